"""GPU tests of the Milstein extension (no reference counterpart): injected-noise parity against the oracle's
Milstein step (1e-5 relative, fp32), strong order 1 of the Philox-driven stored paths against the exact GBM solution
on the same increments, and the Black-Scholes price through mc_simple."""
import numpy as np
import pytest
import torch

from common import oracle, oracle_sde, rel_err, sm

pytestmark = pytest.mark.gpu
DEV = "cuda"
RTOL = 1e-5


def test_milstein_diffusion_inject_vs_oracle():
    rng = np.random.default_rng(1)
    for dim, corr in ((1, None), (3, sm.get_corr_matrix([0.3, -0.2, 0.5]))):
        sde = sm.Gbm(0.02, 0.3, torch.ones(dim), dim, corr)
        solver = sm.MilsteinSolver(sde, 3.0, 64, device=DEV)
        z = rng.standard_normal((4096, 64, dim, 1)).astype(np.float32)
        paths, normals = solver.solve(bs=4096, inject=dict(z=z))
        ref = oracle.diffusion(oracle_sde(solver), z)
        assert rel_err(paths.cpu().numpy(), ref[0]) < RTOL
        # and it is not Euler
        eul, _ = sm.EulerSolver(sde, 3.0, 64, device=DEV).solve(bs=4096, inject=dict(z=z))
        assert float((paths - eul).abs().max()) > 1e-3


def test_milstein_jump_inject_vs_oracle():
    rng = np.random.default_rng(2)
    sde = sm.Merton(0.02, 0.2, 1, -0.05, 0.3, torch.tensor([1.0]), 1)
    solver = sm.JumpMilsteinSolver(sde, 3.0, 50, device=DEV)
    bs, K = 4096, 50 + solver.max_jumps
    z = rng.standard_normal((bs, K, 1)).astype(np.float32)
    jt = np.cumsum(rng.exponential(1.0, (bs, solver.max_jumps)), axis=1).astype(np.float32)
    mk = rng.standard_normal((bs, K)).astype(np.float32)
    paths, aux = solver.solve(bs=bs, inject=dict(z=z, jump_times=jt, marks=mk))
    ref = oracle.jump(oracle_sde(solver), z, None, jt, mk)
    assert aux[3] == ref["total_steps"]
    assert rel_err(paths.cpu().numpy(), ref["paths"][:, :aux[3] + 1]) < RTOL


def test_milstein_strong_order_one_on_philox_paths():
    mu, sigma, T, bs = 0.05, 0.5, 1.0, 200_000
    sde = sm.Gbm(mu, sigma, torch.tensor([1.0]), 1)

    def strong_error(cls, n):
        solver = cls(sde, T, n, device=DEV, seed=11)
        paths, dw = solver.solve(bs=bs)
        w_T = dw[:, :, 0].double().sum(dim=1)
        exact = torch.exp((mu - 0.5 * sigma * sigma) * T + sigma * w_T)
        return float((paths[:, -1, 0].double() - exact).abs().mean())

    mil = [strong_error(sm.MilsteinSolver, n) for n in (16, 32, 64, 128)]
    eul = [strong_error(sm.EulerSolver, n) for n in (16, 32, 64, 128)]
    print("Milstein", mil, "Euler", eul)
    assert all(1.7 < mil[i] / mil[i + 1] < 2.35 for i in range(3))
    assert all(1.25 < eul[i] / eul[i + 1] < 1.6 for i in range(3))
    assert mil[-1] < 0.15 * eul[-1]


def test_milstein_black_scholes_price():
    sde = sm.Gbm(0.02, 0.3, torch.tensor([1.0]), 1)
    solver = sm.MilsteinSolver(sde, 3.0, 64, device=DEV)
    st = sm.mc_simple(2 * 10 ** 7, solver, sm.EuroCall(1.0), sm.ConstantShortRate(0.02), bs=10 ** 6)
    exact = sm.bs_call(1, 1, 3, 0.02, 0.3)
    assert abs(st.sample_mean - exact) < 4 * st.sample_std + 5e-4
    # coupled MLMC pair with Milstein steps runs too (variance of the correction decays like h^2)
    msolver = sm.JumpMilsteinSolver(sm.Merton(0.02, 0.2, 1, -0.05, 0.3, torch.tensor([1.0]), 1), 3.0, 8, device=DEV,
                                    exact_jumps=True)
    est = sm.mc_multilevel([400000, 100000, 50000], [4, 8, 16], msolver, sm.EuroCall(1.0), sm.ConstantShortRate(0.02))
    assert abs(est.sample_mean - sm.merton_call(1, 1, 3, 0.02, 0.2, -0.05, 0.3, 1)) < 4 * est.sample_std + 5e-3
