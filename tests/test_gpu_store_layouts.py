"""Path-storing mode: the three data paths out of the SM must write identical trajectories.
  * TMA tiles (padded 128-byte row pitch)                               csrc/diffusion_tma.cuh, csrc/jump_tma.cuh
  * 16-byte vector flush (padded pitch, TMA disabled)                   csrc/store_tile.cuh
  * 4-byte scalar flush (dense rows = the reference's contiguous layout, row_align = 1)
Checked bit for bit on the same seed, for ragged sizes (1 path, 33 paths, a non-multiple of the CTA size) and for
shapes whose rows are not a multiple of the tile (partial tiles, clipped by the tensor map / scalar tail)."""
import math
import os

import numpy as np
import pytest
import torch

from common import sm

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _solve(solver_factory, bs, row_align, tma=True):
    solver = solver_factory()
    solver.row_align = row_align
    solver.tma_store = tma                    # sdemc_paths_out.flags: SDEMC_OUT_NO_TMA
    out = solver.solve(bs=bs)
    return out


@pytest.mark.parametrize("bs", [1, 33, 1000, 4097])
@pytest.mark.parametrize("steps", [5, 37, 252])
def test_diffusion_store_paths_identical_across_layouts(bs, steps):
    def factory():
        return sm.EulerSolver(sm.Gbm(0.02, 0.3, torch.tensor([1.0]), 1), 3.0, steps, device=DEV, seed=7)

    p_tma, n_tma = _solve(factory, bs, 32)          # padded pitch -> TMA kernel
    p_dense, n_dense = _solve(factory, bs, 1)       # dense rows   -> scalar flush
    assert p_tma.shape == (bs, steps + 1, 1) and n_tma.shape == (bs, steps, 1)
    assert p_dense.is_contiguous() and n_dense.is_contiguous()
    assert torch.equal(p_tma, p_dense) and torch.equal(n_tma, n_dense)
    assert torch.isfinite(p_tma).all() and float(p_tma[:, 0].min()) == 1.0
    p_vec, n_vec = _solve(factory, bs, 32, tma=False)   # padded pitch -> 16-byte vector flush of the LSU kernel
    assert torch.equal(p_vec, p_dense) and torch.equal(n_vec, n_dense)


@pytest.mark.parametrize("dim,corr", [(2, [0.4]), (3, [0.3, -0.2, 0.5]), (4, None)])
def test_diffusion_store_multidim_identical_across_layouts(dim, corr):
    def factory():
        c = sm.get_corr_matrix(corr) if corr else None
        return sm.EulerSolver(sm.Gbm(0.02, 0.3, torch.ones(dim), dim, c), 3.0, 50, device=DEV, seed=5)

    p_tma, n_tma = _solve(factory, 777, 32)
    p_dense, n_dense = _solve(factory, 777, 1)
    assert torch.equal(p_tma, p_dense) and torch.equal(n_tma, n_dense)


def test_heston_and_double_gbm_store_identical_across_layouts():
    def heston():
        return sm.HestonSolver(sm.Heston(0.02, 2.0, 0.04, 0.2, -0.7, torch.tensor([1.0, 0.04])), 3.0, 64, device=DEV)

    def dgbm():
        return sm.EulerSolver(sm.DoubleGbm(0.02, 0.2, 0.1, torch.ones(2), 2, sm.get_corr_matrix([0.3])), 3.0, 40, device=DEV)

    for factory in (heston, dgbm):
        a, an = _solve(factory, 515, 32)
        b, bn = _solve(factory, 515, 1)
        assert torch.equal(a, b) and torch.equal(an, bn)


def _assert_jump_outputs_equal(a, b):
    pa, (na, ta, la, ka, ja) = a
    pb, (nb, tb, lb, kb, jb) = b
    assert ka == kb
    for x, y in ((pa, pb), (na, nb), (ta, tb), (la, lb), (ja, jb)):
        assert x.shape == y.shape and torch.equal(x, y)


@pytest.mark.parametrize("bs", [1, 33, 1000, 4097])
@pytest.mark.parametrize("steps", [1, 7, 24, 100])
def test_jump_store_identical_across_layouts(bs, steps):
    """Merton 1-D (queued jumps): TMA tiles (jump_tma.cuh) == 16-byte vector flush == dense scalar flush, for row
    lengths that are / are not whole super-groups of 12 iterations and whole tiles of 32 elements."""
    def factory():
        sde = sm.Merton(0.02, 0.2, 1, -0.05, 0.3, torch.tensor([1.0]), 1)
        return sm.JumpEulerSolver(sde, 3.0, steps, device=DEV, seed=3)

    a = _solve(factory, bs, 32)
    b = _solve(factory, bs, 1)
    c = _solve(factory, bs, 32, tma=False)
    _assert_jump_outputs_equal(a, b)
    _assert_jump_outputs_equal(c, b)
    pa, (na, ta, la, ka, ja) = a
    # the trajectories end at T and the time grid is non-decreasing
    assert float(ta[:, ka, 0].min()) >= 3.0 - 1e-6
    assert bool((ta[:, 1:ka + 1, 0] >= ta[:, :ka, 0]).all())
    assert float(pa[:, 0].min()) == 1.0 and float(la[:, 0].max()) == 1.0 and float(ja[:, 0].abs().max()) == 0.0


def _jump_models():
    def merton2d():
        sde = sm.Merton(0.02, 0.2, 1, -0.05, 0.3, torch.ones(2), 2, sm.get_corr_matrix([0.4]))
        return sm.JumpEulerSolver(sde, 3.0, 40, device=DEV, seed=11)

    def merton3d():
        sde = sm.Merton(0.02, 0.2, 2, -0.05, 0.3, torch.ones(3), 3)
        return sm.JumpEulerSolver(sde, 1.0, 30, device=DEV, seed=12)

    def explevy2d():   # 'indep' noise: normals (bs, S, 2, 2), dense (inline) jumps
        levy = sm.ExpExampleLevy(1, 1, 0.5, 2, 0.02, 0.3, 0.2, 0.05, dim=2)
        return sm.JumpEulerSolver(sm.LevySde(levy, torch.tensor([1., 1.])), 1.0, 32, device=DEV, seed=13)

    def asian_merton():
        sde = sm.AsianWrapper(sm.Merton(0.02, 0.2, 1, -0.05, 0.3, torch.tensor([1.0]), 1))
        return sm.JumpEulerSolver(sde, 3.0, 50, device=DEV, seed=14)

    def merton_exact():
        sde = sm.Merton(0.02, 0.2, 4, -0.05, 0.3, torch.tensor([1.0]), 1)
        return sm.JumpEulerSolver(sde, 3.0, 20, device=DEV, seed=15, exact_jumps=True)

    return dict(merton2d=merton2d, merton3d=merton3d, explevy2d=explevy2d, asian_merton=asian_merton,
                merton_exact=merton_exact)


@pytest.mark.parametrize("model", sorted(_jump_models()))
def test_jump_store_multidim_identical_across_layouts(model):
    factory = _jump_models()[model]
    b = _solve(factory, 515, 1)
    _assert_jump_outputs_equal(_solve(factory, 515, 32), b)
    _assert_jump_outputs_equal(_solve(factory, 515, 32, tma=False), b)


@pytest.mark.parametrize("bs", [33, 2000])
def test_jump_low_storage_identical_across_layouts(bs):
    """low_storage=True (solvers.py:152-153): only `paths` is produced; TMA kernel with one array."""
    outs = []
    for align, tma in ((32, True), (1, True), (32, False)):
        sde = sm.Merton(0.02, 0.2, 1, -0.05, 0.3, torch.tensor([1.0]), 1)
        solver = sm.JumpEulerSolver(sde, 3.0, 100, device=DEV, seed=3)
        solver.row_align, solver.tma_store = align, tma
        paths, aux = solver.solve(bs=bs, low_storage=True)
        assert aux[0] is None and aux[1] is None and aux[2] is None and aux[4] is None
        outs.append((paths, aux[3]))
    full = sm.JumpEulerSolver(sm.Merton(0.02, 0.2, 1, -0.05, 0.3, torch.tensor([1.0]), 1), 3.0, 100, device=DEV, seed=3)
    ref = full.solve(bs=bs)[0]
    for paths, k in outs:
        assert k == outs[0][1] and torch.equal(paths, outs[0][0])
    assert torch.equal(outs[0][0], ref)


def test_merton_full_storage_at_bench_size_is_self_consistent():
    """The path-storing bench workload (Merton 1-D, 2e6 paths x 100 nominal steps, five arrays through TMA): every
    stored slot is re-derived from its neighbours -- the size-independent form of the step-by-step parity tests.
      times non-decreasing, 0 at index 0, T from the path's last iteration on        solvers.py:190-203
      dt = times[k+1] - times[k] <= h0 and the increments are N(0, dt)                :194-197
      left[k+1] = x[k] (1 + a dt + sigma dW[k])                                       schemes.py:5-9, :203
      paths[k+1] = left[k+1] + x[k] J[k+1]   (exact_jumps = False: the jump acts on the pre-step state)  :214-222
    and the idle tail repeats the final state (dt = 0)."""
    mu, sigma, rate, alpha, gamma_, T, steps, n = 0.02, 0.2, 1.0, -0.05, 0.3, 3.0, 100, 2_000_000
    sde = sm.Merton(mu, sigma, rate, alpha, gamma_, torch.tensor([1.0]), 1)
    solver = sm.JumpEulerSolver(sde, T, steps, device=DEV, seed=21)
    paths, (normals, times, left, total, jumps) = solver.solve(bs=n)
    S = normals.shape[1]
    assert paths.shape == (n, total + 1, 1) and times.shape == (n, S + 1, 1) and left.shape == jumps.shape == (n, S + 1, 1)
    x, t, l, j, dw = paths[:, :, 0], times[:, :total + 1, 0], left[:, :total + 1, 0], jumps[:, :total + 1, 0], normals[:, :total, 0]
    assert float(x[:, 0].min()) == 1.0 == float(x[:, 0].max()) and float(t[:, 0].abs().max()) == 0.0
    dt = t[:, 1:] - t[:, :-1]
    assert float(dt.min()) >= 0.0 and float(dt.max()) <= T / steps + 1e-6       # (differences of fp32 times)
    assert float((t[:, -1] - T).abs().max()) <= 1e-6
    a = mu - rate * (math.exp(alpha + 0.5 * gamma_ ** 2) - 1.0)        # compensated drift of the Merton model
    pred_left = x[:, :-1] * (1.0 + a * dt + sigma * dw)
    assert float(((l[:, 1:] - pred_left).abs() / l[:, 1:].abs().clamp_min(1e-3)).max()) < 2e-6
    pred_x = l[:, 1:] + x[:, :-1] * j[:, 1:]
    assert float(((x[:, 1:] - pred_x).abs() / x[:, 1:].abs().clamp_min(1e-3)).max()) < 2e-6
    # increments: N(0, dt) -- standardised over the active steps
    act = dt > 0
    z = (dw[act] / dt[act].sqrt()).double()
    assert abs(float(z.mean())) < 5.0 / math.sqrt(z.numel()) and abs(float(z.var()) - 1.0) < 5.0 * math.sqrt(2.0 / z.numel()) + 2e-6
    # jumps: rate * T per path on average, marks lognormal - 1
    nj = (j[:, 1:] != 0).sum(dim=1).double()
    assert abs(float(nj.mean()) - rate * T) < 5.0 * math.sqrt(rate * T / n)
    # idle tail: state and time repeat, no increments, no jumps
    it = solver.last_iters.long()
    k = torch.arange(total + 1, device=x.device)[None, :]
    idle = k > it[:, None]
    x_last = x.gather(1, it[:, None])
    assert bool((x[idle] == x_last.expand_as(x)[idle]).all()) and bool(((t[idle] - T).abs() <= 1e-6).all()) and bool((j[idle] == 0).all())
    assert bool((dw[idle[:, 1:]] == 0).all())
    assert int(it.max()) == total


def test_gbm_solve_at_bench_size_is_self_consistent():
    """GBM solve() at the bench size (4e6 x 252, TMA gang of paths + increments): x[k+1] = x[k] (1 + mu h + sigma dW[k])
    for every stored slot (schemes.py:5-9), increments N(0, h)."""
    mu, sigma, T, steps, n = 0.02, 0.3, 3.0, 252, 4_000_000
    solver = sm.EulerSolver(sm.Gbm(mu, sigma, torch.tensor([1.0]), 1), T, steps, device=DEV, seed=22)
    paths, normals = solver.solve(bs=n)
    x, dw = paths[:, :, 0], normals[:, :, 0]
    h = T / steps
    assert float(x[:, 0].min()) == 1.0 == float(x[:, 0].max())
    worst = 0.0
    for lo in range(0, n, 1_000_000):       # in slices: the check itself needs a few temporaries of the slice's size
        xs, ds = x[lo:lo + 1_000_000], dw[lo:lo + 1_000_000]
        pred = xs[:, :-1] * (1.0 + mu * h + sigma * ds)
        worst = max(worst, float(((xs[:, 1:] - pred).abs() / xs[:, 1:].abs().clamp_min(1e-3)).max()))
    assert worst < 2e-6, worst
    z = (dw[:1_000_000] / math.sqrt(h)).double()
    assert abs(float(z.mean())) < 5.0 / math.sqrt(z.numel()) and abs(float(z.var()) - 1.0) < 5.0 * math.sqrt(2.0 / z.numel()) + 2e-6
