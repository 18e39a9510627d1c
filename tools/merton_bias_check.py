"""Weak discretisation bias of the jump-adapted Euler estimator for the Merton call: the estimate at n steps minus
the Merton series should shrink like 1/n (it is a property of the scheme, shared with the reference)."""
import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import sde_mc_b200 as sm
exact = sm.merton_call(1, 1, 3, 0.02, 0.2, -0.05, 0.3, 1)
sde = sm.Merton(0.02, 0.2, 1, -0.05, 0.3, torch.tensor([1.]), 1)
for n, N in ((25, 4e9), (50, 4e9), (100, 4e9), (200, 4e9), (400, 2e9)):
    solver = sm.JumpEulerSolver(sde, 3, n, device='cuda', seed=1234 + n)
    st = sm.mc_simple(int(N), solver, sm.EuroCall(1.0), sm.ConstantShortRate(0.02), bs=10 ** 6, payoff_time='adapted')
    print("n=%4d N=%.0e  estimate %.6f +- %.6f  minus series %+.2e  (x n = %+.4f)" %
          (n, N, st.sample_mean, st.sample_std, st.sample_mean - exact, (st.sample_mean - exact) * n))
