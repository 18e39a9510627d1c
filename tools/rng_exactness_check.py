"""RNG / fp32 accumulation check with NO discretisation bias: Euler on the log-price of a GBM is exact in law, so the
call price through LogGbm + EuroCall(log=True) must equal Black-Scholes within the Monte Carlo error for any number of
steps.  A deviation beyond ~3 sigma would point at the normals (23-bit radius, 16-bit angle) or at fp32 accumulation."""
import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import sde_mc_b200 as sm
exact = sm.bs_call(1, 1, 3, 0.02, 0.3)
sde = sm.LogGbm(0.02, 0.3, torch.tensor([0.0]))
for n, N in ((1, 8e9), (6, 8e9), (60, 8e9), (252, 4e9), (1008, 1e9)):
    solver = sm.EulerSolver(sde, 3, n, device='cuda', seed=99 + n)
    st = sm.mc_simple(int(N), solver, sm.EuroCall(1.0, log=True), sm.ConstantShortRate(0.02), bs=10 ** 6)
    print("n=%5d N=%.0e  estimate %.6f +- %.6f  minus BS %+.2e  = %+.1f sigma" %
          (n, N, st.sample_mean, st.sample_std, st.sample_mean - exact, (st.sample_mean - exact) / st.sample_std))
