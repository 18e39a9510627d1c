"""GPU tests of the fused MLMC pair kernels (H4, H11, E4): injected-noise parity against the golden vectors of the
reference's fp64 run and against the fp32 oracle, then statistical acceptance of mc_multilevel (C5)."""
import math

import numpy as np
import pytest
import torch

from common import MLMC_CASES, golden, golden_json, mlmc_levels, oracle, oracle_sde, rel_err, sm, t

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _np(x):
    return x.detach().cpu().numpy()


@pytest.mark.parametrize("name", sorted(MLMC_CASES))
def test_jump_pair_vs_reference_golden(name):
    """fp32 kernel vs the reference's fp64 coupled pair on the same (float-representable) noise: 2e-5 relative"""
    g = golden(name)
    fine, coarse = mlmc_levels(name)
    solver = sm.JumpEulerSolver(MLMC_CASES[name](g), float(g["T"]), fine, device=DEV,
                                exact_jumps=bool(int(g["exact_jumps"])))
    inject = dict(z=g["z"].astype(np.float32), jump_times=g["jump_times"].astype(np.float32),
                  marks=g["marks"].astype(np.float32))
    if "zc" in g.files:
        inject["zc"] = g["zc"].astype(np.float32)
    (pf, pc), _ = solver.multilevel_solve(g["z"].shape[0], (fine, coarse), inject=inject)
    # the oracle in fp32 on the SAME rounded inputs is the tight check ...
    fl, cl, iters, total = oracle.jump_pair(oracle_sde(solver, fine), fine, coarse, inject["z"], inject.get("zc"),
                                            inject["jump_times"], inject["marks"], np.float32)
    assert total > 0
    assert rel_err(_np(pf)[:, -1], fl) < 1e-5
    assert rel_err(_np(pc)[:, -1], cl) < 1e-5
    # ... and the reference's own fp64 run bounds the fp32 rounding of the whole pair
    assert rel_err(_np(pf)[:, -1], g["fine_last"]) < 5e-5
    assert rel_err(_np(pc)[:, -1], g["coarse_last"]) < 5e-5


@pytest.mark.parametrize("name", sorted(MLMC_CASES))
def test_jump_pair_fp64_vs_reference_golden(name):
    """sdemc_mlmc_pair_f64 (fp64 state, coefficients, marks) on the reference's fp64 noise: the terminal states of the
    reference's own fp64 coupled pair (solvers.py:228-307 under set_default_dtype(float64)) to 1e-12"""
    g = golden(name)
    fine, coarse = mlmc_levels(name)
    torch.set_default_dtype(torch.float64)       # the reference's switch for its jump MLMC; selects the fp64 kernel
    try:
        solver = sm.JumpEulerSolver(MLMC_CASES[name](g), float(g["T"]), fine, device=DEV,
                                    exact_jumps=bool(int(g["exact_jumps"])))
        inject = dict(z=g["z"], jump_times=g["jump_times"], marks=g["marks"])
        if "zc" in g.files:
            inject["zc"] = g["zc"]
        (pf, pc), _ = solver.multilevel_solve(g["z"].shape[0], (fine, coarse), inject=inject)
    finally:
        torch.set_default_dtype(torch.float32)
    assert pf.dtype == torch.float64
    assert rel_err(_np(pf)[:, -1], g["fine_last"]) < 1e-12
    assert rel_err(_np(pc)[:, -1], g["coarse_last"]) < 1e-12


def test_fp64_pair_large_inject_vs_fp64_oracle_and_philox_levels():
    """(1) 4096 pairs of injected fp64 noise against the fp64 oracle at 1e-12; (2) Philox-driven fp64 pairs consume the
    counters of the fp32 pair kernel: per-level moments of P_f - P_c agree to fp32 accuracy of the states, and the
    fp64 run removes the fp32 rounding noise from the finest level's variance"""
    rng = np.random.default_rng(15)
    bs, fine, coarse = 4096, 32, 8
    torch.set_default_dtype(torch.float64)
    try:
        sde = sm.Merton(0.02, 0.2, 1, -0.05, 0.3, torch.tensor([1.]), 1)
        solver = sm.JumpEulerSolver(sde, 3, fine, device=DEV, exact_jumps=True)
        K = coarse + solver.max_jumps
        z = rng.standard_normal((bs, K * 4, 1))
        jt = np.cumsum(rng.exponential(1.0, (bs, solver.max_jumps)), axis=1)
        mk = rng.standard_normal((bs, K))
        fl, cl, iters, total = oracle.jump_pair(oracle_sde(solver, fine), fine, coarse, z, None, jt, mk, np.float64)
        (pf, pc), _ = solver.multilevel_solve(bs, (fine, coarse), inject=dict(z=z, jump_times=jt, marks=mk))
        assert rel_err(_np(pf)[:, -1], fl) < 1e-12 and rel_err(_np(pc)[:, -1], cl) < 1e-12
        from sde_mc_b200.mlmc import _level_moments
        call, csr = sm.EuroCall(1.0), sm.ConstantShortRate(0.02)
        n = 200_000
        s64 = sm.JumpEulerSolver(sde, 3, 1, device=DEV, seed=3, exact_jumps=True)
        m64 = _level_moments(s64, call, csr, n, 128, 64).read()
    finally:
        torch.set_default_dtype(torch.float32)
    sde = sm.Merton(0.02, 0.2, 1, -0.05, 0.3, torch.tensor([1.]), 1)
    s32 = sm.JumpEulerSolver(sde, 3, 1, device=DEV, seed=3, exact_jumps=True)
    m32 = _level_moments(s32, call, csr, n, 128, 64).read()
    assert m64["n"] == m32["n"] == n and m64["iters"] == m32["iters"]          # same draws, same clock
    mean64, mean32 = m64["sum"] / n, m32["sum"] / n
    assert abs(mean64 - mean32) < 2e-6                                          # same pairs up to fp32 rounding of the states
    var64 = m64["sumsq"] / n - mean64 ** 2
    var32 = m32["sumsq"] / n - mean32 ** 2
    assert 0 < var64 <= var32 * 1.02                                            # fp32 adds rounding noise, never removes it


def test_diffusion_pair_vs_reference_golden():
    g = golden("diff_gbm_mlmc_8_2")
    solver = sm.EulerSolver(sm.Gbm(0.02, 0.3, t(g["x0"]), 1), float(g["T"]), 8, device=DEV)
    (pf, pc), _ = solver.multilevel_solve(16, (8, 2), inject=dict(z=g["z"]))
    assert rel_err(_np(pf)[:, -1], g["paths_fine"][:, -1]) < 1e-5
    assert rel_err(_np(pc)[:, -1], g["paths_coarse"][:, -1]) < 1e-5


@pytest.mark.parametrize("fine,coarse", [(8, 2), (16, 8)])
def test_heston_pair_vs_reference_golden(fine, coarse):
    """the coupled pair of the Heston solver (the reference's HestonSolver inherits multilevel_solve): terminal states
    of both paths against the unmodified reference on injected normals"""
    g = golden("mlmc_heston_%d_%d" % (fine, coarse))
    sde = sm.Heston(float(g["r"]), float(g["kappa"]), float(g["theta"]), float(g["xi"]), float(g["rho"]), t(g["x0"]))
    solver = sm.HestonSolver(sde, float(g["T"]), fine, device=DEV)
    (pf, pc), _ = solver.multilevel_solve(g["z"].shape[0], (fine, coarse), inject=dict(z=g["z"]))
    # 5e-5, not the 1e-5 of the other pairs: the variance root (-b - sqrt(b^2 - 4ac)) / 2a of a step cancels when v is
    # small, which amplifies the last-bit differences between evaluation orders (FMA contraction here, none in torch)
    # by 10-100x over steps as long as 0.375 and 1.5; the scalar oracle, ordered like the reference, holds 5e-6
    assert rel_err(_np(pf)[:, -1], g["paths_fine"][:, -1]) < 5e-5
    assert rel_err(_np(pc)[:, -1], g["paths_coarse"][:, -1]) < 5e-5


def test_heston_mc_multilevel_matches_plain_mc():
    """mc_multilevel on the Heston model (levels 4, 8, 16, 32): the telescoping estimator agrees with plain MC on the
    finest grid within the error bars, and the level variances decay"""
    sde = sm.Heston(0.02, 2.0, 0.04, 0.2, -0.7, torch.tensor([1.0, 0.04]))
    call, csr = sm.EuroCall(1.0), sm.ConstantShortRate(0.02)
    levels = [4, 8, 16, 32]
    ml = sm.mc_multilevel([2_000_000, 400_000, 200_000, 100_000], levels, sm.HestonSolver(sde, 3.0, 4, device=DEV), call, csr)
    plain = sm.mc_simple(4_000_000, sm.HestonSolver(sde, 3.0, 32, device=DEV, seed=5), call, csr, bs=10 ** 6)
    assert abs(float(ml.sample_mean) - float(plain.sample_mean)) <= 3.0 * math.hypot(float(ml.sample_std), float(plain.sample_std))


def test_pair_large_inject_vs_oracle():
    rng = np.random.default_rng(5)
    bs, fine, coarse = 4096, 16, 4
    sde = sm.Merton(0.02, 0.2, 1, -0.05, 0.3, torch.tensor([1.]), 1)
    solver = sm.JumpEulerSolver(sde, 3, fine, device=DEV)
    K = coarse + solver.max_jumps
    z = rng.standard_normal((bs, K * 4, 1)).astype(np.float32)
    jt = np.cumsum(rng.exponential(1.0, (bs, solver.max_jumps)), axis=1).astype(np.float32)
    mk = rng.standard_normal((bs, K)).astype(np.float32)
    fl, cl, iters, total = oracle.jump_pair(oracle_sde(solver, fine), fine, coarse, z, None, jt, mk, np.float32)
    (pf, pc), _ = solver.multilevel_solve(bs, (fine, coarse), inject=dict(z=z, jump_times=jt, marks=mk))
    assert rel_err(_np(pf)[:, -1], fl) < 1e-5 and rel_err(_np(pc)[:, -1], cl) < 1e-5


def test_level_variances_match_reference_pilot():
    """per-level Var[P_l - P_{l-1}] against the reference's fp64 pilot (40000 pairs per level)"""
    ref = golden_json("ref_stats")["c5_level_vars_n40000"]
    sde = sm.Merton(0.02, 0.2, 1, -0.05, 0.3, torch.tensor([1.]), 1)
    solver = sm.JumpEulerSolver(sde, 3, 1, device=DEV)
    levels = [1, 2, 4, 8, 16, 32, 64, 128]
    call, csr = sm.EuroCall(1.0), sm.ConstantShortRate(0.02)
    n = 400000
    from sde_mc_b200.mlmc import _level_moments
    for i, lv in enumerate(levels):
        m = _level_moments(solver, call, csr, n, lv, levels[i - 1] if i else 0).read()
        var = (m["sumsq"] - m["sum"] ** 2 / n) / (n - 1)
        # sample variance of a heavy-ish tailed difference at n=40000: allow 15 %
        assert abs(var - ref[i]) < 0.15 * ref[i], (lv, var, ref[i])


def test_mc_multilevel_c5():
    """C5.  With exact_jumps=False the reference's coupled pair applies the jump to the state before the LAST fine
    sub-step (solvers.py:275,296-299), which is the post-step state whenever the jump time is reached early; its
    fine path then has a different law from the same level run as a coarse path, the telescoping sum no longer
    cancels and the estimate sits ~1.4e-3 below the Merton series (the reference's own fp64 run shows it too).
    We reproduce the reference: exact_jumps=False is checked against the reference's CI, exact_jumps=True (the
    LevyRainbowMLMC setting, where the telescoping is exact) against the closed form."""
    sde = sm.Merton(0.02, 0.2, 1, -0.05, 0.3, torch.tensor([1.]), 1)
    levels = [1, 2, 4, 8, 16, 32, 64, 128]
    call, csr = sm.EuroCall(1.0), sm.ConstantShortRate(0.02)
    exact = sm.merton_call(1, 1, 3, 0.02, 0.2, -0.05, 0.3, 1)
    ref = golden_json("ref_stats")["c5_mlmc_eps2e-3"]

    solver = sm.JumpEulerSolver(sde, 3, 1, device=DEV)
    trials = sm.get_optimal_trials(10 ** 5, levels, 2e-4, solver, call, csr)
    assert len(trials) == len(levels) and all(a >= b for a, b in zip(trials, trials[1:]))
    for ours, theirs in zip(trials, ref["trials"]):   # allocation scales like eps^-2: 100x the reference's at 2e-3
        assert 0.6 < ours / (100.0 * theirs) < 1.6, (ours, theirs)
    st = sm.mc_multilevel(trials, levels, solver, call, csr)
    assert st.sample_std * 1.96 < 2.3e-4
    assert abs(st.sample_mean - ref["mean"]) <= 1.96 * math.hypot(ref["se"], st.sample_std)
    assert st.sample_mean < exact - 5e-4              # the reference's telescoping bias is reproduced, not hidden

    solver = sm.JumpEulerSolver(sde, 3, 1, device=DEV, exact_jumps=True)
    trials = sm.get_optimal_trials(10 ** 5, levels, 2e-4, solver, call, csr)
    st = sm.mc_multilevel(trials, levels, solver, call, csr)
    assert st.sample_std * 1.96 < 2.3e-4
    assert abs(st.sample_mean - exact) <= 1.96 * st.sample_std + 3e-4     # + level-7 discretisation bias


def test_mc_multilevel_works_for_diffusions_too():
    """the reference raises TypeError here (no low_storage on DiffusionSolver); ours runs the uniform-grid pair"""
    p = sm.BlackScholesEuroCall.default_params(1, DEV)
    levels = [2, 8, 32, 128]
    st = sm.mc_multilevel([2 * 10 ** 6, 4 * 10 ** 5, 10 ** 5, 3 * 10 ** 4], levels, p.solver, p.payoff, p.discounter)
    assert abs(st.sample_mean - sm.bs_call(1, 1, 3, 0.02, 0.3)) <= 1.96 * st.sample_std + 3e-4


@pytest.mark.parametrize("case", ["merton_1step", "merton_3steps_terminal", "levy2d_2steps", "merton2d_exact"])
def test_short_path_kernel_matches_generic_kernel(case, monkeypatch):
    """jump_flat.cuh (persistent lanes, chosen for short paths such as MLMC level 0) simulates the SAME paths as the
    generic jump kernel: identical path count and iteration total, moments equal up to the fp64 summation order."""
    from sde_mc_b200 import _engine as E
    from sde_mc_b200 import _lib as L
    from sde_mc_b200 import _spec
    if case.startswith("merton_"):
        sde = sm.Merton(0.02, 0.2, 1, -0.05, 0.3, torch.tensor([1.]), 1)
        payoff = sm.EuroCall(1.0)
    elif case == "merton2d_exact":
        sde = sm.Merton(0.02, 0.2, 2, -0.05, 0.3, torch.tensor([1., 1.1]), 2, sm.get_corr_matrix([0.4]))
        payoff = sm.Rainbow(1.0)
    else:
        levy = sm.ExpExampleLevy(1, 1, .5, 2, .02, .3, .2, .05, dim=2)
        sde = sm.LevySde(levy, torch.tensor([1., 1.]))
        payoff = sm.Rainbow(1.0)
    steps = {"merton_1step": 1, "merton_3steps_terminal": 3, "levy2d_2steps": 2, "merton2d_exact": 1}[case]
    solver = sm.JumpEulerSolver(sde, 3, steps, device=DEV, exact_jumps=case == "merton2d_exact")
    mode = L.INDEX_TERMINAL if case.endswith("terminal") else L.INDEX_ADAPTED
    n = 300_001  # ragged: the last warp is partial and lanes run out of paths at different times
    lib = L.load()
    res = {}
    for flat, short in (("0", L.SHORT_OFF), ("1", L.SHORT_ALIGNED)):   # ALIGNED keeps jump_kernel's streams
        solver.short_path = short
        with torch.cuda.device(DEV):
            mom = E.Moments(torch.device(DEV, 0))
            po = _spec.payoff_struct(payoff, math.exp(-0.06), mode)
            L.check(lib.sdemc_mc_moments(solver._sde_struct(steps), po, L.SdemcRange(7, 123, n), None, L.ptr(mom.buf),
                                         L.ptr(L.workspace(torch.device(DEV, 0))), L.stream_ptr(torch.device(DEV, 0))))
            res[flat] = mom.read()
    a, b = res["0"], res["1"]
    assert a["n"] == b["n"] == n
    assert a["iters"] == b["iters"] and a["iters"] > n * steps
    for key in ("sum", "sumsq"):
        assert abs(a[key] - b[key]) <= 1e-11 * abs(a[key]), (key, a[key], b[key])


@pytest.mark.parametrize("case", ["merton_2_1", "merton_8_4_exact", "merton_16_4", "levy2d_4_2_exact"])
def test_persistent_lane_pair_kernel_matches_lockstep_pair_kernel(case, monkeypatch):
    """jump_pair_flat_kernel (chosen for the coarse levels) simulates the SAME coupled pairs as jump_pair_kernel:
    identical pair count and fine sub-step total, moments of P_f - P_c equal up to the fp64 summation order."""
    from sde_mc_b200.mlmc import _level_moments
    if case.startswith("merton"):
        sde = sm.Merton(0.02, 0.2, 1, -0.05, 0.3, torch.tensor([1.]), 1)
        payoff = sm.EuroCall(1.0)
    else:
        levy = sm.ExpExampleLevy(1, 1, .5, 2, .02, .3, .2, .05, dim=2)
        sde = sm.LevySde(levy, torch.tensor([1., 1.]))
        payoff = sm.Rainbow(1.0)
    fine, coarse = {"merton_2_1": (2, 1), "merton_8_4_exact": (8, 4), "merton_16_4": (16, 4),
                    "levy2d_4_2_exact": (4, 2)}[case]
    n = 200_003
    res = {}
    from sde_mc_b200 import _lib as L
    for flat, short in (("0", L.SHORT_OFF), ("1", L.SHORT_ALIGNED)):
        solver = sm.JumpEulerSolver(sde, 3, coarse, device=DEV, exact_jumps=case.endswith("exact"), seed=11)
        solver.short_path = short
        res[flat] = _level_moments(solver, payoff, sm.ConstantShortRate(0.02), n, fine, coarse).read()
    a, b = res["0"], res["1"]
    assert a["n"] == b["n"] == n
    assert a["iters"] == b["iters"] and a["iters"] >= n * fine
    assert a["sumsq"] > 0
    for key in ("sum", "sumsq"):
        assert abs(a[key] - b[key]) <= 1e-10 * max(abs(a[key]), a["sumsq"] ** 0.5), (key, a[key], b[key])


@pytest.mark.parametrize("steps,mode", [(1, "adapted"), (3, "terminal"), (2, "adapted")])
def test_packed_short_path_kernel_same_law_as_generic_kernel(steps, mode, monkeypatch):
    """jump_flat1d_kernel (1-D Merton short paths, MLMC level 0) packs the normals, gaps and marks of two iterations
    into one Philox block: a different stream, the same law.  Against the generic kernel at 2e7 paths each: payoff
    mean within 4 combined standard errors, second moment and mean iteration count within 4 of theirs; path count
    exact."""
    from sde_mc_b200 import _engine as E
    from sde_mc_b200 import _lib as L
    from sde_mc_b200 import _spec
    sde = sm.Merton(0.02, 0.2, 1, -0.05, 0.3, torch.tensor([1.]), 1)
    solver = sm.JumpEulerSolver(sde, 3, steps, device=DEV)
    payoff = sm.EuroCall(1.0)
    idx = L.INDEX_TERMINAL if mode == "terminal" else L.INDEX_ADAPTED
    n = 20_000_001
    lib = L.load()
    res = {}
    for name, short in (("generic", L.SHORT_OFF), ("packed", L.SHORT_PACKED)):
        solver.short_path = short
        with torch.cuda.device(DEV):
            dev = torch.device(DEV, 0)
            mom = E.Moments(dev)
            po = _spec.payoff_struct(payoff, math.exp(-0.06), idx)
            L.check(lib.sdemc_mc_moments(solver._sde_struct(steps), po, L.SdemcRange(5, 1000, n), None, L.ptr(mom.buf),
                                         L.ptr(L.workspace(dev)), L.stream_ptr(dev)))
            res[name] = mom.read()
    a, b = res["generic"], res["packed"]
    assert a["n"] == b["n"] == n
    mean_a, mean_b = a["sum"] / n, b["sum"] / n
    var_a, var_b = a["sumsq"] / n - mean_a ** 2, b["sumsq"] / n - mean_b ** 2
    se = math.sqrt((var_a + var_b) / n)
    assert abs(mean_a - mean_b) < 4 * se, (mean_a, mean_b, se)
    # second moment: standard error from the fourth moment bound payoff^2 <= 30 * payoff here is loose; use 1 %
    assert abs(var_a / var_b - 1.0) < 0.01, (var_a, var_b)
    # iterations per path: Poisson(3)-driven spread, se <= sqrt(3 / n) each
    it_a, it_b = a["iters"] / n, b["iters"] / n
    assert abs(it_a - it_b) < 4 * math.sqrt(2 * 3.0 / n) + 1e-4, (it_a, it_b)
    assert it_b > steps + 1.0                      # every jump before T costs at most one extra iteration
    if steps == 1:
        assert abs(it_b - 4.0) < 2e-3              # one nominal step: exactly 1 + Poisson(rate T = 3) iterations


@pytest.mark.parametrize("steps,exact", [(1, False), (1, True), (4, False)])
def test_packed_kernel_restated_iteration_matches_generic_iteration(steps, exact, monkeypatch):
    """Same packed stream, two forms of the loop body: the generic jump_iteration (SDEMC_SHORT_PACKED_GENERIC) and the
    restated one (stateless mesh, sigma^2 dt folded into the Box-Muller radius, one-FMA hit test; default).  Same
    draws, same mesh, same hits: iteration totals and all five moment sums equal up to fp32 rounding of the states."""
    from sde_mc_b200 import _engine as E
    from sde_mc_b200 import _lib as L
    from sde_mc_b200 import _spec
    sde = sm.Merton(0.02, 0.2, 1, -0.05, 0.3, torch.tensor([1.]), 1)
    solver = sm.JumpEulerSolver(sde, 3, steps, device=DEV, exact_jumps=exact)
    n = 3_000_001
    lib = L.load()
    res = {}
    for mode, short in (("2", L.SHORT_PACKED_GENERIC), ("1", L.SHORT_PACKED)):
        solver.short_path = short
        with torch.cuda.device(DEV):
            dev = torch.device(DEV, 0)
            mom = E.Moments(dev)
            po = _spec.payoff_struct(sm.EuroCall(1.0), math.exp(-0.06), L.INDEX_ADAPTED)
            L.check(lib.sdemc_mc_moments(solver._sde_struct(steps), po, L.SdemcRange(9, 77, n), None, L.ptr(mom.buf),
                                         L.ptr(L.workspace(dev)), L.stream_ptr(dev)))
            res[mode] = mom.read()
    a, b = res["2"], res["1"]
    assert a["n"] == b["n"] == n
    # the two hit tests are the same inequality up to the rounding of its right-hand side: a handful of iterations in 1e7
    assert abs(a["iters"] - b["iters"]) <= 2 + 1e-6 * a["iters"], (a["iters"], b["iters"])
    # per-path states agree to fp32 rounding (~1e-7, partly systematic): sums within 2e-6 per path; a structural
    # difference (a different draw, mesh or hit) moves them by >= 1e-3 per path
    for key in ("sum", "sumsq", "sum_c", "sumsq_c", "sum_pc"):
        assert abs(a[key] - b[key]) <= 2e-6 * n, (key, a[key], b[key])


def test_graph_replay_of_a_pass_equals_queued_launches():
    """opt-in: mc_multilevel / get_optimal_trials replay a captured CUDA graph of the per-level launches (mlmc._GraphedPass):
    same kernels, same global path ids => the (levels, 8) fp64 moments are those of the plain launches bit for bit,
    call after call (every call advances the path ids, which the graph reads from device memory)."""
    from sde_mc_b200 import mlmc as M
    levels, trials = [1, 2, 4, 8, 16], [200_000, 50_000, 30_000, 20_000, 10_001]
    call, csr = sm.EuroCall(1.0), sm.ConstantShortRate(0.02)
    mk = lambda: sm.JumpEulerSolver(sm.Merton(0.02, 0.2, 1, -0.05, 0.3, torch.tensor([1.]), 1), 3, 1, device=DEV,
                                    seed=11, exact_jumps=True)
    a, b = mk(), mk()
    got, want = [], []
    saved = M.GRAPH_LEVELS
    M.GRAPH_LEVELS = True
    try:
        for _ in range(3):
            got.append(M._all_levels(a, call, csr, trials, levels).read())
        s1 = sm.mc_multilevel(trials, levels, mk(), call, csr)
    finally:
        M.GRAPH_LEVELS = saved
    M.GRAPH_LEVELS = False
    try:
        for _ in range(3):
            want.append(M._all_levels(b, call, csr, trials, levels).read())
        s2 = sm.mc_multilevel(trials, levels, mk(), call, csr)
    finally:
        M.GRAPH_LEVELS = saved
    assert got == want
    assert got[0] != got[1]                      # different paths every call
    assert a._next_path == b._next_path == 3 * sum(trials)
    # and the public estimator on top of it
    assert s1.sample_mean == s2.sample_mean and s1.sample_std == s2.sample_std
