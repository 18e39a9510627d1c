"""User-defined SDEs on the fused kernels.  A subclass of Sde supplies its coefficients twice: as the
Python tensor methods the reference requires (drift / diffusion / jumps, sde.py:63-152) and as CUDA expressions
(kernel_code()).  Parity oracle = the reference's own algorithm: an eager PyTorch Euler loop (schemes.py:5-13,
solvers.py:83-87) driven by the SAME Brownian increments the kernel reports, calling the subclass's tensor methods.
Tolerance 1e-5 relative (fp32)."""
import math

import numpy as np
import pytest
import torch

from common import rel_err, sm

pytestmark = pytest.mark.gpu
DEV = "cuda"


class Cir(sm.DiffusionSde):
    def __init__(self, kappa, theta, xi, x0):
        super().__init__(x0, 1, 1, 'diag')
        self.kappa, self.theta, self.xi = kappa, theta, xi

    def drift(self, t, x):
        return self.kappa * (self.theta - x)

    def diffusion(self, t, x):
        return self.xi * torch.sqrt(torch.clamp(x, min=0))

    def kernel_code(self):
        return dict(drift=["p[0] * (p[1] - x[0])"], diffusion=["p[2] * sqrtf(fmaxf(x[0], 0.f))"],
                    params=[self.kappa, self.theta, self.xi])


class SeasonalPair(sm.DiffusionSde):
    """2-D, time-dependent, coupled: dX0 = a (m(t) - X0) dt + s0 dW0 ;  dX1 = r X1 dt + s1 X1 (1 + 0.5 tanh(X0)) dW1"""

    def __init__(self, x0, corr):
        super().__init__(x0, 2, 2, 'diag', corr)

    def drift(self, t, x):
        m = 0.1 * torch.sin(2.0 * t)
        return torch.stack([1.5 * (m - x[:, 0]), 0.03 * x[:, 1]], dim=1)

    def diffusion(self, t, x):
        return torch.stack([0.2 * torch.ones_like(x[:, 0]), 0.25 * x[:, 1] * (1 + 0.5 * torch.tanh(x[:, 0]))], dim=1)

    def kernel_code(self):
        return dict(drift=["1.5f * (0.1f * sinf(2.f * t) - x[0])", "0.03f * x[1]"],
                    diffusion=["0.2f", "0.25f * x[1] * (1.f + 0.5f * tanhf(x[0]))"], params=[])


class MyMerton(sm.LogNormalJumpsSde):
    """the Merton model written by hand as a user-defined SDE"""

    def __init__(self, mu, sigma, rate, alpha, gamma, x0):
        super().__init__(rate, alpha, gamma, x0, 1, 1, 'diag')
        self.mu, self.sigma = mu, sigma

    def drift(self, t, x):
        return (self.mu - self.rate * self.jump_mean()) * x

    def diffusion(self, t, x):
        return self.sigma * x

    def jumps(self, t, x, jumps):
        return x * jumps

    def kernel_code(self):
        return dict(drift=["p[0] * x[0]"], diffusion=["p[1] * x[0]"], jump=["x[0] * J"],
                    params=[self.mu - float(self.rate) * self.jump_mean(), self.sigma])


def _replay_euler(sde, T, steps, dw):
    """reference algorithm on the GPU in eager PyTorch: x + a(t, x) h + b(t, x) dW with the fp32 clock t += h"""
    bs = dw.shape[0]
    h = torch.tensor(T / steps, device=DEV)
    t = torch.tensor(0.0, device=DEV)
    x = sde.init_value.to(DEV).unsqueeze(0).repeat(bs, 1)
    out = [x]
    for i in range(steps):
        x = x + sde.drift(t, x) * h + sde.diffusion(t, x) * dw[:, i]
        out.append(x)
        t = t + h
    return torch.stack(out, dim=1)


@pytest.mark.parametrize("row_align", [32, 1])
def test_user_cir_paths_match_eager_replay(row_align):
    sde = Cir(2.0, 0.04, 0.3, torch.tensor([0.04]))
    solver = sm.EulerSolver(sde, 3.0, 120, device=DEV, seed=9)
    solver.row_align = row_align
    paths, dw = solver.solve(bs=5000)
    ref = _replay_euler(sde, 3.0, 120, dw)
    assert paths.shape == (5000, 121, 1) and rel_err(paths.cpu().numpy(), ref.cpu().numpy(), 1e-2) < 1e-5
    assert float(dw.std()) == pytest.approx(math.sqrt(3.0 / 120), rel=2e-2)


def test_user_time_dependent_2d_correlated_paths_match_eager_replay():
    sde = SeasonalPair(torch.tensor([0.0, 1.0]), sm.get_corr_matrix([-0.5]))
    solver = sm.EulerSolver(sde, 2.0, 64, device=DEV, seed=4)
    paths, dw = solver.solve(bs=4096)
    ref = _replay_euler(sde, 2.0, 64, dw)
    assert rel_err(paths.cpu().numpy(), ref.cpu().numpy(), 1e-1) < 1e-5
    c = np.corrcoef(dw[:, :, 0].flatten().cpu().numpy(), dw[:, :, 1].flatten().cpu().numpy())[0, 1]
    assert abs(c + 0.5) < 0.01                                       # the increments carry the Cholesky correlation


def test_user_gbm_equals_builtin_and_prices_black_scholes():
    class MyGbm(sm.DiffusionSde):
        def __init__(self):
            super().__init__(torch.tensor([1.0]), 1, 1, 'diag')

        def drift(self, t, x):
            return 0.02 * x

        def diffusion(self, t, x):
            return 0.3 * x

        def kernel_code(self):
            return dict(drift=["p[0] * x[0]"], diffusion=["p[1] * x[0]"], params=[0.02, 0.3])

    user = sm.EulerSolver(MyGbm(), 3.0, 64, device=DEV, seed=2)
    builtin = sm.EulerSolver(sm.Gbm(0.02, 0.3, torch.tensor([1.0]), 1), 3.0, 64, device=DEV, seed=2)
    pu, nu = user.solve(bs=3000)
    pb, nb = builtin.solve(bs=3000)
    assert torch.equal(nu, nb)                                       # same Philox stream
    assert rel_err(pu.cpu().numpy(), pb.cpu().numpy()) < 1e-5
    st = sm.mc_simple(2 * 10 ** 7, user, sm.EuroCall(1.0), sm.ConstantShortRate(0.02), bs=10 ** 6)
    assert abs(st.sample_mean - sm.bs_call(1, 1, 3, 0.02, 0.3)) < 4 * st.sample_std + 5e-4
    # moments kernel == store kernel on the same paths
    fresh = sm.EulerSolver(MyGbm(), 3.0, 64, device=DEV, seed=2)
    one = sm.mc_simple(3000, fresh, sm.EuroCall(1.0), sm.ConstantShortRate(0.02), bs=3000)
    direct = (torch.clamp(pu[:, -1, 0] - 1.0, min=0) * math.exp(-0.06)).double().mean()
    assert abs(one.sample_mean - float(direct)) < 1e-6


def test_user_jump_sde_equals_builtin_merton():
    args = (0.02, 0.2, 1.0, -0.05, 0.3)
    user = sm.JumpEulerSolver(MyMerton(*args, torch.tensor([1.0])), 3.0, 100, device=DEV, seed=6)
    builtin = sm.JumpEulerSolver(sm.Merton(*args, torch.tensor([1.0]), 1), 3.0, 100, device=DEV, seed=6)
    pu, (nu, tu, lu, ku, ju) = user.solve(bs=4096)
    pb, (nb, tb, lb, kb, jb) = builtin.solve(bs=4096)
    assert ku == kb and torch.equal(tu, tb) and torch.equal(ju, jb)  # same jump times and marks (same Philox streams)
    assert rel_err(pu.cpu().numpy(), pb.cpu().numpy()) < 1e-5
    st = sm.mc_simple(2 * 10 ** 7, user, sm.EuroCall(1.0), sm.ConstantShortRate(0.02), bs=10 ** 6, payoff_time='adapted')
    exact = sm.merton_call(1, 1, 3, 0.02, 0.2, -0.05, 0.3, 1)
    assert abs(st.sample_mean - exact) < 4 * st.sample_std + 5e-4


def test_user_sde_unsupported_paths_fail_loudly():
    sde = Cir(2.0, 0.04, 0.3, torch.tensor([0.04]))
    solver = sm.EulerSolver(sde, 3.0, 16, device=DEV)
    with pytest.raises(sm._lib.SdemcError):
        solver.solve(bs=16, inject=dict(z=np.zeros((16, 16, 1, 1), np.float32)))
    with pytest.raises(sm._lib.SdemcError):
        sm.mc_multilevel([1000, 1000], [4, 8], solver, sm.EuroCall(0.0), sm.ConstantShortRate(0.0))


def test_user_defined_option_runs_on_stored_paths():
    """a user Option subclass without kernel coefficients: payoff in PyTorch on the GPU (the reference's own
    evaluation, mc.py:84-93) -- one-shot on trajectories from the storing kernel, batched on the states the fused
    moments kernel reports per path (no trajectory is stored); diffusion and jump solvers"""

    class PowerCall(sm.Option):
        def __init__(self, strike, power):
            super().__init__(log=False)
            self.strike, self.power = strike, power

        def payoff(self, x):
            return torch.clamp(x[:, 0] ** self.power - self.strike, min=0)

    gbm = sm.Gbm(0.02, 0.3, torch.tensor([1.0]), 1)
    csr = sm.ConstantShortRate(0.02)
    one = sm.mc_simple(200000, sm.EulerSolver(gbm, 3.0, 32, device=DEV, seed=3), PowerCall(1.0, 1.0), csr)
    ref = sm.mc_simple(200000, sm.EulerSolver(gbm, 3.0, 32, device=DEV, seed=3), sm.EuroCall(1.0), csr)
    assert one.payoffs.shape == (200000,) and abs(one.sample_mean - ref.sample_mean) < 1e-6   # power 1 == EuroCall
    # batched = the same global path ids through the fused kernel: same estimate as the one-shot call on stored paths
    bat = sm.mc_simple(200000, sm.EulerSolver(gbm, 3.0, 32, device=DEV, seed=3), PowerCall(1.0, 1.0), csr, bs=64000)
    assert abs(bat.sample_mean - float(one.sample_mean)) < 2e-6 and abs(bat.sample_std - float(one.sample_std)) < 2e-6
    big = sm.mc_simple(4 * 10 ** 6, sm.EulerSolver(gbm, 3.0, 32, device=DEV, seed=3), PowerCall(1.0, 1.0), csr, bs=10 ** 6)
    assert abs(big.sample_mean - sm.bs_call(1, 1, 3, 0.02, 0.3)) < 4 * big.sample_std + 1e-3
    merton = sm.Merton(0.02, 0.2, 1, -0.05, 0.3, torch.tensor([1.0]), 1)
    jm = sm.mc_simple(2 * 10 ** 6, sm.JumpEulerSolver(merton, 3.0, 50, device=DEV), PowerCall(1.0, 1.0), csr, bs=5 * 10 ** 5,
                      payoff_time='adapted')
    assert abs(jm.sample_mean - sm.merton_call(1, 1, 3, 0.02, 0.2, -0.05, 0.3, 1)) < 4 * jm.sample_std + 1e-3
    # both payoff indices against the built-in payoff on the same seed (quirk Q1: 'terminal' = array index num_steps)
    for ptime in ('adapted', 'terminal'):
        mk = lambda: sm.JumpEulerSolver(merton, 3.0, 50, device=DEV, seed=9)
        a = sm.mc_simple(300000, mk(), PowerCall(1.0, 1.0), csr, bs=10 ** 5, payoff_time=ptime)
        b = sm.mc_simple(300000, mk(), sm.EuroCall(1.0), csr, bs=10 ** 5, payoff_time=ptime)
        assert abs(a.sample_mean - b.sample_mean) < 2e-6, (ptime, a.sample_mean, b.sample_mean)
