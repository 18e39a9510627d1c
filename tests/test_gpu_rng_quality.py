"""Statistical quality of the in-register noise at the 1e-5 level.  Euler on the LOG-price of a GBM is exact in law
(constant coefficients), so LogGbm + EuroCall(log=True) must reproduce Black-Scholes for ANY step count: a deviation
would have to come from the normals (Philox4x32-10, six per block: 23-bit radius, 16-bit angle) or from fp32
accumulation over the steps.  2e9 paths per case: standard error 1e-5 (relative 4e-5)."""
import pytest
import torch

from common import sm

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("steps", [1, 6, 252])
def test_log_euler_reproduces_black_scholes_for_any_step_count(steps):
    sde = sm.LogGbm(0.02, 0.3, torch.tensor([0.0]))
    solver = sm.EulerSolver(sde, 3, steps, device="cuda", seed=4242 + steps)
    n = 2 * 10 ** 9
    st = sm.mc_simple(n, solver, sm.EuroCall(1.0, log=True), sm.ConstantShortRate(0.02), bs=10 ** 6)
    exact = sm.bs_call(1, 1, 3, 0.02, 0.3)
    print("steps=%d estimate %.6f +- %.6f, Black-Scholes %.6f" % (steps, st.sample_mean, st.sample_std, exact))
    assert st.sample_std < 1.2e-5
    assert abs(st.sample_mean - exact) < 4 * st.sample_std


def test_disjoint_path_ranges_are_uncorrelated_and_reproducible():
    """same seed + same path range -> same moments bit for bit; the two halves of a range behave as independent
    samples (their means differ by an amount consistent with the standard error)."""
    sde = sm.Gbm(0.02, 0.3, torch.tensor([1.0]), 1)
    a = sm.mc_simple(10 ** 7, sm.EulerSolver(sde, 3, 50, device="cuda", seed=5), sm.EuroCall(1.0), sm.ConstantShortRate(0.02), bs=10 ** 6)
    b = sm.mc_simple(10 ** 7, sm.EulerSolver(sde, 3, 50, device="cuda", seed=5), sm.EuroCall(1.0), sm.ConstantShortRate(0.02), bs=10 ** 6)
    assert a.sample_mean == b.sample_mean and a.sample_std == b.sample_std
    s = sm.EulerSolver(sde, 3, 50, device="cuda", seed=5)
    h1 = sm.mc_simple(5 * 10 ** 6, s, sm.EuroCall(1.0), sm.ConstantShortRate(0.02), bs=10 ** 6)   # paths [0, 5e6)
    h2 = sm.mc_simple(5 * 10 ** 6, s, sm.EuroCall(1.0), sm.ConstantShortRate(0.02), bs=10 ** 6)   # paths [5e6, 1e7)
    assert abs(0.5 * (h1.sample_mean + h2.sample_mean) - a.sample_mean) < 1e-9
    assert abs(h1.sample_mean - h2.sample_mean) < 5 * (h1.sample_std ** 2 + h2.sample_std ** 2) ** 0.5


def test_far_out_of_the_money_payoffs_see_the_tail_of_the_normals():
    """The Box-Muller normals stop at 5.65 sigma (23-bit radius uniform) and lie on 65536 directions per pair
    (philox.cuh).  A digital and a call struck FOUR standard deviations out of the money on
    a single exact log-Euler step are priced by the normals' tail alone: 4e9 draws put 1.3e5 of them beyond the
    strike (relative standard error 0.3 %); the mass missing beyond 5.65 sigma is 0.05 % of that.  Both must match the
    Black-Scholes closed forms within 3 standard errors."""
    import math
    r, sigma, T = 0.02, 0.3, 3.0
    strike = math.exp((r - 0.5 * sigma * sigma) * T + 4.0 * sigma * math.sqrt(T))        # 7.41 S0
    n = 4 * 10 ** 9
    for payoff, exact in ((sm.Digital(strike, log=True), sm.bs_digital_call(1, strike, T, r, sigma)),
                          (sm.EuroCall(strike, log=True), sm.bs_call(1, strike, T, r, sigma))):
        solver = sm.EulerSolver(sm.LogGbm(r, sigma, torch.tensor([0.0])), T, 1, device="cuda", seed=99)
        st = sm.mc_simple(n, solver, payoff, sm.ConstantShortRate(r), bs=10 ** 6)
        print("%s: %.4e +- %.1e, closed form %.4e" % (type(payoff).__name__, st.sample_mean, st.sample_std, exact))
        assert st.sample_std < 0.005 * exact
        assert abs(st.sample_mean - exact) < 3 * st.sample_std
