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
