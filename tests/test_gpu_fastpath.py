"""Per-path parity of the MOMENTS kernels -- the kernels the benchmark numbers come from.

`sdemc_mc_moments(per_path=...)` makes the fast kernels report what every path contributed (payoff, iteration count,
terminal state).  Each benchmarked kernel is compared
  (a) with the path-storing kernel on the same seed (same Philox counters => same paths), at 1e6 paths, and
  (b) DIRECTLY with the CPU oracle: the noise the kernel draws from Philox is handed to oracle.diffusion /
      oracle.jump, the restatement of the reference loops solvers.py:68-88,164-226.  For the uniform grid the noise
      is re-derived on the CPU (oracle/philox_streams.py, float64 libm).  For the jump-adapted kernels it is written
      by `sdemc_debug_draws` with the device functions the kernels themselves call, because that loop amplifies a
      1e-7 (MUFU-accuracy) shift of a jump time without bound -- a step of length dt -> 0 has sqrt(dt) in its
      increment, and whether t + (T - t) rounds to T decides an extra iteration -- so only bit-identical jump times
      make a per-path comparison meaningful; the CPU restatement pins those draws in turn (last test).
Tolerances: 1e-5 relative against the oracle (BASELINE.json north_star); 2e-6 between two kernels on the same draws.

Iteration counts are asserted EXACTLY everywhere: all kernels evaluate torch.isclose as the reference does,
|tau - t| <= 1e-12 + 1e-5 |t| (solvers.py:212), bit for bit.  (These tests found that the one-FMA form of that test
the 1-D kernels used in round 1 rounded its threshold differently: one path in 1e6 hit a jump an iteration early, or
applied a jump lying 1e-5 T after T that the reference never applies.)
States of short paths (1-4 nominal steps of length up to T) are compared on the O(1) scale of the spot: a single Euler
factor 1 + a h + sigma sqrt(h) z can come close to zero and amplify the relative rounding of a state of 1e-3.
"""
import math

import numpy as np
import pytest
import torch

from common import oracle, oracle_sde, rel_err, sm
from oracle import philox_streams as ps
from sde_mc_b200 import _engine as E
from sde_mc_b200 import _lib as L
from sde_mc_b200 import _spec

pytestmark = pytest.mark.gpu
DEV = "cuda"
CSR = sm.ConstantShortRate(0.02)


def _np(x):
    return x.detach().cpu().numpy()


def _moments_per_path(solver, payoff, n, mode=L.INDEX_ADAPTED):
    pp = {}
    mom = E.run_moments(solver, payoff, CSR, n, mode, per_path=pp).read()
    assert mom["n"] == n
    pay, it, term = _np(pp["payoffs"]), _np(pp["iters"]), _np(pp["terminal"])
    # the per-path outputs ARE what was reduced
    assert abs(mom["sum"] - pay.astype(np.float64).sum()) <= 1e-9 * max(1.0, abs(mom["sum"]))
    assert mom["iters"] == float(it.astype(np.int64).sum())
    return pay, it, term


def _merton(rate=1.0):
    return sm.Merton(0.02, 0.2, rate, -0.05, 0.3, torch.tensor([1.]), 1)


# ---------------------------------------------------------------------------------------------------------------
# FAST1D (diffusion.cuh): the C2 kernel
# ---------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("steps", [252, 50, 5])
def test_fast1d_gbm_per_path_equals_store_kernel(steps):
    """1e6 paths, same seed: the folded-radius, pipelined FAST1D body against the generic storing kernel"""
    n = 1_000_000 if steps != 252 else 400_000
    mk = lambda: sm.EulerSolver(sm.Gbm(0.02, 0.3, torch.tensor([1.]), 1), 3, steps, device=DEV, seed=5)
    pay, it, term = _moments_per_path(mk(), sm.EuroCall(1.0), n)
    po = _spec.payoff_struct(sm.EuroCall(1.0), math.exp(-0.06), L.INDEX_ADAPTED)
    paths, _, payoffs = mk().solve(bs=n, want_payoff=po)
    assert np.all(it == steps)
    # two evaluation orders of the same step (sigma sqrt(h) folded into the Box-Muller radius vs applied to the unit
    # normal): one fp32 rounding of difference per step, accumulating like a random walk -- 4e-7 sqrt(steps) covers
    # the maximum over 1e6 paths (measured 2.9e-6 at 252 steps)
    tol = max(2e-6, 4e-7 * math.sqrt(steps))
    assert rel_err(term[:, 0], _np(paths[:, -1, 0])) < tol
    assert rel_err(pay, _np(payoffs), 1.0) < tol
    # the production launch (no per-path outputs) is the other instantiation of the same kernel source: its fp64 sums
    # must be those of the per-path launch bit for bit
    pp = {}
    a = E.run_moments(mk(), sm.EuroCall(1.0), CSR, n, L.INDEX_ADAPTED).read()
    b = E.run_moments(mk(), sm.EuroCall(1.0), CSR, n, L.INDEX_ADAPTED, per_path=pp).read()
    assert a == b


@pytest.mark.parametrize("family", ["gbm", "loggbm"])
def test_fast1d_per_path_vs_oracle_c2_shape(family):
    """C2 shape (2048 x 252): Philox normals re-derived on the CPU -> oracle.diffusion -> terminal states at 1e-5"""
    n, steps, seed = 2048, 252, 11
    sde = sm.Gbm(0.02, 0.3, torch.tensor([1.]), 1) if family == "gbm" else sm.LogGbm(0.02, 0.2, torch.tensor([0.1]))
    solver = sm.EulerSolver(sde, 3, steps, device=DEV, seed=seed)
    solver._next_path = 12345                       # not the first paths of the stream
    pay, it, term = _moments_per_path(solver, sm.EuroCall(1.0, log=family == "loggbm"), n)
    z = ps.brownian_normals(seed, 12345 + np.arange(n), steps).reshape(n, steps, 1, 1)
    ref_paths, _ = oracle.diffusion(oracle_sde(solver), z)
    assert rel_err(term[:, 0], ref_paths[:, -1, 0], 1.0 if family == "loggbm" else 1e-3) < 1e-5
    kind, strike, aux = sm.EuroCall(1.0).kernel_spec()
    ref_pay = oracle.payoff(oracle.payoff_struct(kind, strike, family == "loggbm", 1.0, aux), ref_paths[:, -1]) * \
        np.float32(math.exp(-0.06))                  # the oracle's payoff leaves the discount factor to the caller
    assert rel_err(pay, ref_pay, 1.0) < 1e-5


# ---------------------------------------------------------------------------------------------------------------
# jump1d_kernel (jump1d.cuh): the north-star kernel
# ---------------------------------------------------------------------------------------------------------------
def _store_reference(solver_factory, payoff, n, mode):
    s = solver_factory()
    po = _spec.payoff_struct(payoff, math.exp(-0.06), mode)
    paths, aux, payoffs = s.solve(bs=n, low_storage=True, want_payoff=po)
    col = s.num_steps if mode == L.INDEX_TERMINAL else None
    iters = _np(s.last_iters)
    P = _np(paths)
    term = P[:, s.num_steps, 0] if col is not None else P[np.arange(n), np.minimum(iters, P.shape[1] - 1), 0]
    return _np(payoffs), iters, term


@pytest.mark.parametrize("case", ["c1", "c1_exact", "c1_terminal", "replay_heavy", "replay_heavy_exact", "steps_13"])
def test_jump1d_per_path_equals_store_kernel(case):
    """Merton, QUEUE strategy: speculative groups, queue refills and replays of jump1d.cuh against the generic
    storing kernel, path by path on the same seed.  `replay_heavy`: rate 3 (9 jumps per path) with the queue depth
    forced to 4, so nearly every path refills and most groups with a refill are replayed (jump1d.cuh:132-146)."""
    heavy = case.startswith("replay_heavy")
    steps = 13 if case == "steps_13" else 100
    n = 1_000_000 if not heavy else 500_000
    mode = L.INDEX_TERMINAL if case == "c1_terminal" else L.INDEX_ADAPTED

    def factory():
        s = sm.JumpEulerSolver(_merton(3.0 if heavy else 1.0), 3, steps, device=DEV, seed=17,
                               exact_jumps=case.endswith("exact"))
        s.jump_strategy = L.JUMPS_QUEUE
        s.queue_depth = 4 if heavy else 0
        return s

    pay, it, term = _moments_per_path(factory(), sm.EuroCall(1.0), n, mode)
    ref_pay, ref_it, ref_term = _store_reference(factory, sm.EuroCall(1.0), n, mode)
    assert np.array_equal(it, ref_it)
    assert it.mean() > steps + (4.0 if heavy else 1.5)   # a jump restarts the grid: about half an iteration each
    assert rel_err(term[:, 0], ref_term, 1e-2) < 3e-6
    if mode == L.INDEX_TERMINAL:
        assert rel_err(pay, ref_pay, 1.0) < 3e-6
    else:
        # the one-shot 'adapted' payoff is read at the LAST column of the batch (mc.py:84-91): a path that reached T
        # keeps stepping with dt = 0 while others run, and a jump lying within 1e-5 T after T is then applied to it
        # (batch-dependent reference quirk, reproduced by the storing kernel); the per-path loop ends at T
        late = np.abs(pay - ref_pay) > 3e-6 * np.maximum(np.abs(ref_pay), 1.0)
        assert late.mean() < 2e-4
    if heavy:  # the branch under test was really taken: more than 4 jumps before T needs at least one refill
        jt = _draws(factory(), L.DRAWS_QUEUE, 0, 20000, 8, 2)[0]
        assert ((jt < 3.0).sum(axis=1) > 4).mean() > 0.9


def _draws(solver, kind, lo, n, count, arrays):
    """sdemc_debug_draws: `arrays` (n, count) float32 arrays of the kernels' own draws for paths lo .. lo + n - 1"""
    lib = L.load()
    dev = torch.device(DEV, 0)
    out = [torch.empty((n, count), device=dev, dtype=torch.float32) for _ in range(arrays)] + [None] * (3 - arrays)
    L.check(lib.sdemc_debug_draws(solver._sde_struct(), L.SdemcRange(int(solver.seed), lo, n), kind, count,
                                  L.ptr(out[0]), L.ptr(out[1]), L.ptr(out[2]), L.stream_ptr(dev)))
    return [_np(t) for t in out[:arrays]]


def _marks_at_hits(osde, z, zc, jt, raw_per_jump):
    """marks are per JUMP in the QUEUE strategy and per hit ITERATION in the reference's contract; the jump-adapted
    clock does not depend on the marks, so a first oracle pass with unit marks locates the hit iterations"""
    n, K = z.shape[0], z.shape[1]
    probe = oracle.jump(osde, z, zc, jt, np.ones((n, K), np.float32))
    hits = probe["jumps"][:, 1:, 0] != 0
    marks = np.zeros((n, K), np.float32)
    order = np.cumsum(hits, axis=1) - 1
    rows, cols = np.nonzero(hits)
    marks[rows, cols] = raw_per_jump[rows, order[rows, cols]]
    return marks


@pytest.mark.parametrize("exact", [False, True])
@pytest.mark.parametrize("rate,qd", [(1.0, 0), (3.0, 4)])
def test_jump1d_per_path_vs_oracle_c1_shape(exact, rate, qd):
    """C1 shape (4096 x 100, K = 100 + max_jumps slots): the kernel's own queue draws and normals -> oracle.jump"""
    n, steps, seed, lo = 4096, 100, 23, 777
    solver = sm.JumpEulerSolver(_merton(rate), 3, steps, device=DEV, seed=seed, exact_jumps=exact)
    solver.jump_strategy, solver.queue_depth, solver._next_path = L.JUMPS_QUEUE, qd, lo
    pay, it, term = _moments_per_path(solver, sm.EuroCall(1.0), n)
    K = -(-(steps + solver.max_jumps) // 6) * 6
    z = _draws(solver, L.DRAWS_BROWNIAN, lo, n, K, 1)[0].reshape(n, K, 1)
    jt, raw = _draws(solver, L.DRAWS_QUEUE, lo, n, -(-solver.max_jumps // 4) * 4, 2)
    jt, raw = jt[:, :solver.max_jumps], raw[:, :solver.max_jumps]
    osde = oracle_sde(solver)
    ref = oracle.jump(osde, z, None, jt, _marks_at_hits(osde, z, None, jt, raw))
    assert np.array_equal(it, ref["iters"])
    ref_term = ref["paths"][np.arange(n), ref["iters"], 0]
    assert rel_err(term[:, 0], ref_term, 1e-2) < 1e-5
    ref_pay = oracle.payoff(oracle.payoff_struct(0, 1.0), ref_term[:, None]) * np.float32(math.exp(-0.06))
    assert np.max(np.abs(pay - ref_pay) / np.maximum(ref_term, 1.0)) < 1e-5      # D (x - K): on the scale of the spot
    # 'terminal' payoff index (quirk Q1): the state at array index num_steps
    solver._next_path = lo
    _, _, term_n = _moments_per_path(solver, sm.EuroCall(1.0), n, L.INDEX_TERMINAL)
    assert rel_err(term_n[:, 0], ref["paths"][:, steps, 0], 1e-2) < 1e-5


# ---------------------------------------------------------------------------------------------------------------
# generic jump_kernel, INLINE strategy (jump.cuh): the C4 kernel
# ---------------------------------------------------------------------------------------------------------------
def test_levy2d_inline_kernel_per_path_vs_oracle():
    """C4 model (2-D exp-Levy, rho = 0.4, dense jumps): the kernel's own normals, gap and mark candidates; the
    candidate-per-iteration strategy is mapped onto the reference's (jump_times, marks) inputs by restating only the
    clock (oracle/philox_streams.py:candidate_jumps); the states come from oracle.jump."""
    n, steps, seed = 512, 32, 31
    levy = sm.ExpExampleLevy(1, 1, 0.5, 2, 0.02, 0.3, 0.2, 0.001, dim=2)
    sde = sm.LevySde(levy, torch.tensor([1., 1.]), corr_matrix=sm.get_corr_matrix([0.4]))
    solver = sm.JumpEulerSolver(sde, 3, steps, device=DEV, seed=seed)
    solver.short_path = L.SHORT_OFF
    pay, it, term = _moments_per_path(solver, sm.Rainbow(1.0), n)
    K = int(it.max()) + 6
    K += K & 1
    zz = _draws(solver, L.DRAWS_BROWNIAN, 0, n, 3 * K, 1)[0].reshape(n, K, 3)
    gap, raw = _draws(solver, L.DRAWS_INLINE, 0, n, K, 2)
    jt, marks, iters = ps.candidate_jumps(3.0 / steps, 3.0, float(sde.jump_rate().sum()), gap, raw, solver.max_jumps)
    ref = oracle.jump(oracle_sde(solver), zz[:, :, :2], zz[:, :, 2], jt, marks)
    assert np.array_equal(it, ref["iters"]) and np.array_equal(it, iters)
    ref_term = ref["paths"][np.arange(n), ref["iters"]]
    assert rel_err(term, ref_term) < 5e-5          # ~400 iterations with |J| up to 10 (as the injected-noise C4 test)


# ---------------------------------------------------------------------------------------------------------------
# short-path kernels (jump_flat.cuh): MLMC level 0
# ---------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("steps,mode", [(1, L.INDEX_ADAPTED), (3, L.INDEX_TERMINAL), (2, L.INDEX_ADAPTED)])
def test_aligned_short_path_kernel_per_path_equals_store_kernel(steps, mode):
    """SDEMC_SHORT_ALIGNED (and plain jump_kernel) run the generic iteration on the counters of the storing kernel:
    payoffs and iteration counts are IDENTICAL, path by path"""
    n = 300_001

    def factory(short):
        s = sm.JumpEulerSolver(_merton(), 3, steps, device=DEV, seed=29)
        s.jump_strategy, s.short_path = L.JUMPS_INLINE, short
        return s

    ref_pay, ref_it, ref_term = _store_reference(lambda: factory(L.SHORT_OFF), sm.EuroCall(1.0), n, mode)
    for short in (L.SHORT_ALIGNED, L.SHORT_OFF, L.SHORT_AUTO):
        pay, it, term = _moments_per_path(factory(short), sm.EuroCall(1.0), n, mode)
        assert np.array_equal(it, ref_it)
        assert np.array_equal(pay, ref_pay)


@pytest.mark.parametrize("steps,exact", [(1, False), (1, True), (4, False)])
def test_packed_short_path_kernels_per_path_vs_oracle(steps, exact):
    """STREAM_PACKED (MLMC level 0): both loop bodies -- the restated one (SDEMC_SHORT_PACKED) and the generic
    jump_iteration (PACKED_GENERIC) -- against oracle.jump on the CPU-derived draws, and against each other."""
    n, seed, lo = 8192, 37, 4242
    ids = lo + np.arange(n)
    outs = {}
    for short in (L.SHORT_PACKED, L.SHORT_PACKED_GENERIC):
        solver = sm.JumpEulerSolver(_merton(), 3, steps, device=DEV, seed=seed, exact_jumps=exact)
        solver.short_path, solver._next_path = short, lo
        outs[short] = _moments_per_path(solver, sm.EuroCall(1.0), n)
    K = int(max(o[1].max() for o in outs.values())) + 2
    K += K & 1
    z, gap, raw = _draws(solver, L.DRAWS_PACKED, lo, n, K, 3)
    jt, marks, iters = ps.candidate_jumps(3.0 / steps, 3.0, 1.0, gap, raw, solver.max_jumps)
    ref = oracle.jump(oracle_sde(solver), z.reshape(n, K, 1), None, jt, marks)
    ref_term = ref["paths"][np.arange(n), ref["iters"], 0]
    for short, (pay, it, term) in outs.items():
        assert np.array_equal(it, ref["iters"]) and np.array_equal(it, iters), short
        assert rel_err(term[:, 0], ref_term, 0.1) < 1e-5, short               # O(1) scale, see the module docstring
    a, b = outs[L.SHORT_PACKED], outs[L.SHORT_PACKED_GENERIC]
    assert rel_err(a[2], b[2], 0.1) < 2e-6


def test_debug_draws_match_the_cpu_restatement_of_the_streams():
    """The draw-reporting hook against oracle/philox_streams.py (Philox counters, key schedule and every bits ->
    normal / gap / mark map restated in numpy with float64 libm): equal to the accuracy of the MUFU units."""
    n, lo, seed = 2000, 987_654_321_012, 41          # path ids beyond 2^32: the high counter word is exercised
    ids = lo + np.arange(n)
    solver = sm.JumpEulerSolver(_merton(2.0), 3, 10, device=DEV, seed=seed)
    # normals: radius (up to 5.6) x the absolute error of MUFU sin / cos and of the fp32 angle (~2e-6)
    z = _draws(solver, L.DRAWS_BROWNIAN, lo, n, 40, 1)[0]
    assert np.max(np.abs(z - ps.brownian_normals(seed, ids, 40))) < 2e-5
    jt, raw = _draws(solver, L.DRAWS_QUEUE, lo, n, 16, 2)
    cjt, craw = ps.queue_jumps(seed, ids, 16, 2.0, "lognormal")
    assert np.max(np.abs(jt - cjt) / (1 + cjt)) < 2e-6 and np.max(np.abs(raw - craw)) < 2e-5   # lg2.approx: abs 2^-22
    gap, raw = _draws(solver, L.DRAWS_INLINE, lo, n, 14, 2)
    cgap, craw = ps.inline_draws(seed, ids, 14, "lognormal")
    assert np.max(np.abs(gap - cgap)) < 2e-6 * (1 + cgap.max()) and np.max(np.abs(raw - craw)) < 2e-5
    z, gap, raw = _draws(solver, L.DRAWS_PACKED, lo, n, 9, 3)
    cz, cgap, craw = ps.packed_draws(seed, ids, 9)
    assert np.max(np.abs(z - cz)) < 2e-5 and np.max(np.abs(raw - craw)) < 2e-5
    assert np.max(np.abs(gap - cgap)) < 2e-6 * (1 + cgap.max())
    levy = sm.ExpExampleLevy(1, 1, 0.5, 2, 0.02, 0.3, 0.2, 0.001, dim=2)
    lsolver = sm.JumpEulerSolver(sm.LevySde(levy, torch.tensor([1., 1.])), 3, 10, device=DEV, seed=seed)
    gap, raw = _draws(lsolver, L.DRAWS_INLINE, lo, n, 14, 2)
    cgap, craw = ps.inline_draws(seed, ids, 14, "icdf")
    assert np.max(np.abs(gap - cgap)) < 2e-6 * (1 + cgap.max()) and np.array_equal(raw, craw)   # uniforms: exact
