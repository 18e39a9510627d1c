"""GPU parity tests: the CUDA path (through the Python API -> ctypes -> C-ABI) against
  (1) the golden vectors of the unmodified reference, (2) the CPU oracle on larger seeded inputs,
  (3) closed forms / the reference's own confidence intervals for the Philox-driven estimators.

Tolerance for deterministic injected-noise parity: 1e-5 relative in fp32 (BASELINE.json north_star), measured
against the O(1) scale of the state for log-price models that cross zero."""
import math

import numpy as np
import pytest
import torch

from common import (DIFFUSION_CASES, JUMP_CASES, golden, golden_json, jump_solver, oracle, oracle_sde, rel_err, sm, t)

pytestmark = pytest.mark.gpu
RTOL = 1e-5
DEV = "cuda"


def _np(x):
    return x.detach().cpu().numpy()


# ---------------------------------------------------------------------------------------------------------------
# (1) golden vectors
# ---------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name", sorted(DIFFUSION_CASES))
def test_diffusion_solve_vs_reference_golden(name):
    g = golden(name)
    build, solver_cls = DIFFUSION_CASES[name]
    solver = solver_cls(build(g), float(g["T"]), int(g["z"].shape[1]), device=DEV)
    paths, normals = solver.solve(bs=g["z"].shape[0], inject=dict(z=g["z"]))
    assert paths.is_cuda and tuple(paths.shape) == g["paths"].shape and tuple(normals.shape) == g["normals"].shape
    floor = 1.0 if name == "diff_loggbm" else 1e-3
    assert rel_err(_np(paths), g["paths"], floor) < RTOL
    assert rel_err(_np(normals), g["normals"], 1e-2) < RTOL


@pytest.mark.parametrize("name", sorted(JUMP_CASES))
def test_jump_solve_vs_reference_golden(name):
    g = golden(name)
    solver = jump_solver(name, g, DEV)
    inject = dict(z=g["z"], jump_times=g["jump_times"], marks=g["marks"])
    if "zc" in g.files:
        inject["zc"] = g["zc"]
    paths, (normals, times, left, total_steps, jumps) = solver.solve(bs=g["z"].shape[0], inject=inject)
    ts = int(g["total_steps"])
    assert total_steps == ts
    assert tuple(paths.shape) == g["paths"].shape
    floor = 1.0 if name in ("jump_addlevy_1d", "jump_levy2d") else 1e-3
    tol = RTOL if "levy" not in name else 3e-5  # several hundred iterations with |J| up to ~10 compound
    assert rel_err(_np(paths), g["paths"], floor) < tol
    assert rel_err(_np(left)[:, :ts + 1], g["left_paths"], floor) < tol
    assert rel_err(_np(times)[:, :ts + 1, 0], g["time_paths"]) < 2e-6
    assert rel_err(_np(jumps)[:, :ts + 1], g["jump_paths"], 1e-2) < 2e-5
    assert rel_err(_np(normals)[:, :ts], g["normals"], 1e-2) < RTOL


def test_low_storage_solve_returns_only_paths():
    g = golden("jump_merton_1d_ex0")
    solver = jump_solver("jump_merton_1d_ex0", g, DEV)
    paths, aux = solver.solve(bs=g["z"].shape[0], low_storage=True,
                              inject=dict(z=g["z"], jump_times=g["jump_times"], marks=g["marks"]))
    assert aux[0] is None and aux[1] is None and aux[2] is None and aux[4] is None
    assert rel_err(_np(paths)[:, -1], g["paths"][:, -1]) < RTOL


def test_estimators_on_injected_noise_vs_reference():
    """mc_simple / mc_terminal_cv one-shot values computed by the reference on the same increments"""
    g = golden("est_gbm_1d")
    solver = sm.EulerSolver(sm.Gbm(0.02, 0.3, t(g["x0"]), 1), 3.0, 16, device=DEV)
    po = sm._spec.payoff_struct(sm.EuroCall(1.0), math.exp(-0.06), 1)
    paths, normals, payoffs = solver.solve(bs=256, inject=dict(z=g["z"]), want_payoff=po)
    # payoff = D (x - K): the subtraction cancels, so measure against the O(1) scale of the spot
    assert rel_err(_np(payoffs), g["payoffs"], 1.0) < RTOL
    assert abs(float(payoffs.mean()) - float(g["mean"])) < 1e-6
    assert abs(float(payoffs.std()) / 16.0 - float(g["std"])) < 1e-6


# ---------------------------------------------------------------------------------------------------------------
# (2) larger seeded inputs against the oracle
# ---------------------------------------------------------------------------------------------------------------
def test_merton_c1_shape_inject_vs_oracle():
    """C1 acceptance: Merton 1-D, 100 steps, bs=4096, K=133 iterations of injected noise"""
    rng = np.random.default_rng(11)
    bs, steps = 4096, 100
    for exact in (False, True):
        sde = sm.Merton(0.02, 0.2, 1, -0.05, 0.3, torch.tensor([1.]), 1)
        solver = sm.JumpEulerSolver(sde, 3, steps, device=DEV, exact_jumps=exact)
        K = steps + solver.max_jumps
        z = rng.standard_normal((bs, K, 1)).astype(np.float32)
        jt = np.cumsum(rng.exponential(1.0, (bs, solver.max_jumps)), axis=1).astype(np.float32)
        mk = rng.standard_normal((bs, K)).astype(np.float32)
        ref = oracle.jump(oracle_sde(solver), z, None, jt, mk)
        paths, (normals, times, left, total, jumps) = solver.solve(bs=bs, inject=dict(z=z, jump_times=jt, marks=mk))
        assert total == ref["total_steps"]
        assert np.array_equal(_np(solver.last_iters), ref["iters"])
        assert rel_err(_np(paths)[:, -1], ref["paths"][:, total]) < RTOL
        assert rel_err(_np(paths), ref["paths"][:, :total + 1]) < RTOL
        assert rel_err(_np(times)[:, :total + 1, 0], ref["times"][:, :total + 1]) < 2e-6


def test_levy_c4_shape_inject_vs_oracle():
    """C4 model (2-D exp-Levy, rho = 0.4), 32 nominal steps, ~400 executed iterations per path"""
    rng = np.random.default_rng(12)
    bs, steps = 256, 32
    levy = sm.ExpExampleLevy(1, 1, 0.5, 2, 0.02, 0.3, 0.2, 0.001, dim=2)
    sde = sm.LevySde(levy, torch.tensor([1., 1.]), corr_matrix=sm.get_corr_matrix([0.4]))
    solver = sm.JumpEulerSolver(sde, 3, steps, device=DEV)
    K = steps + solver.max_jumps
    z = rng.standard_normal((bs, K, 2)).astype(np.float32)
    zc = rng.standard_normal((bs, K)).astype(np.float32)
    jt = np.cumsum(rng.exponential(1.0 / float(sde.jump_rate()), (bs, solver.max_jumps)), axis=1).astype(np.float32)
    mk = rng.random((bs, K)).astype(np.float32)
    ref = oracle.jump(oracle_sde(solver), z, zc, jt, mk)
    paths, (normals, times, left, total, jumps) = solver.solve(bs=bs, inject=dict(z=z, zc=zc, jump_times=jt, marks=mk))
    assert total == ref["total_steps"]
    assert np.array_equal(_np(solver.last_iters), ref["iters"])
    assert rel_err(_np(paths)[:, -1], ref["paths"][:, total]) < 5e-5


def test_gbm_c2_shape_inject_vs_oracle():
    rng = np.random.default_rng(13)
    bs, steps = 2048, 252
    solver = sm.EulerSolver(sm.Gbm(0.02, 0.3, torch.tensor([1.]), 1), 3, steps, device=DEV)
    z = rng.standard_normal((bs, steps, 1, 1)).astype(np.float32)
    ref_paths, _ = oracle.diffusion(oracle_sde(solver), z)
    paths, _ = solver.solve(bs=bs, inject=dict(z=z))
    assert rel_err(_np(paths)[:, -1], ref_paths[:, -1]) < RTOL


# ---------------------------------------------------------------------------------------------------------------
# (3) Philox-driven kernels: replay, moments-vs-store consistency, statistical acceptance
# ---------------------------------------------------------------------------------------------------------------
def test_philox_store_replayed_through_oracle():
    """solve() with Philox noise returns its increments; feeding them back through the oracle must reproduce the
    stored trajectories (validates the Philox-driven kernel end to end, including jump times and marks)."""
    sde = sm.Merton(0.02, 0.2, 1, -0.05, 0.3, torch.tensor([1.]), 1)
    solver = sm.JumpEulerSolver(sde, 3, 50, device=DEV, seed=3)
    paths, (normals, times, left, total, jumps) = solver.solve(bs=512)
    P, Tm, Lf, J, N = _np(paths), _np(times)[..., 0], _np(left), _np(jumps), _np(normals)
    assert total >= 50 and np.all(np.isfinite(P))
    # reconstruct each step: x_left = x_prev (1 + a dt + b dW), x = x_left + x_base J
    a = sde.kernel_spec().a[0]
    dt = np.diff(Tm[:, :total + 1], axis=1)
    x_prev = P[:, :total, 0]
    left_pred = x_prev * (1.0 + a * dt + 0.2 * N[:, :total, 0])
    assert rel_err(left_pred, Lf[:, 1:total + 1, 0]) < RTOL
    post_pred = Lf[:, 1:total + 1, 0] + x_prev * J[:, 1:total + 1, 0]
    assert rel_err(post_pred, P[:, 1:total + 1, 0]) < RTOL
    # increments have the right scale: dW / sqrt(dt) ~ N(0,1) on the active steps
    act = dt > 1e-6
    zhat = N[:, :total, 0][act] / np.sqrt(dt[act])
    assert abs(zhat.mean()) < 0.02 and abs(zhat.std() - 1.0) < 0.02
    # jump count ~ Poisson(rate T = 3)
    nj = (J[:, :total + 1, 0] != 0).sum(1)
    assert abs(nj.mean() - 3.0) < 0.35


@pytest.mark.parametrize("strategy", [1, 2])
def test_moments_kernel_equals_store_kernel_same_seed(strategy):
    """the fused moments kernel and the path-storing kernel consume identical Philox streams"""
    sde = sm.Merton(0.02, 0.2, 1, -0.05, 0.3, torch.tensor([1.]), 1)
    n = 20000
    s1 = sm.JumpEulerSolver(sde, 3, 40, device=DEV, seed=9)
    s2 = sm.JumpEulerSolver(sde, 3, 40, device=DEV, seed=9)
    s1.jump_strategy = s2.jump_strategy = strategy
    call, csr = sm.EuroCall(1.0), sm.ConstantShortRate(0.02)
    for mode in ("adapted", "terminal"):
        a = sm.mc_simple(n, s1, call, csr, bs=n, payoff_time=mode)
        b = sm.mc_simple(n, s2, call, csr, payoff_time=mode)
        assert abs(a.sample_mean - b.sample_mean) < 2e-6
        assert abs(a.sample_std - b.sample_std) < 2e-6


def test_gbm_moments_equals_store_same_seed():
    n = 30000
    s1 = sm.EulerSolver(sm.Gbm(0.02, 0.3, torch.tensor([1.]), 1), 3, 50, device=DEV, seed=4)
    s2 = sm.EulerSolver(sm.Gbm(0.02, 0.3, torch.tensor([1.]), 1), 3, 50, device=DEV, seed=4)
    call, csr = sm.EuroCall(1.0), sm.ConstantShortRate(0.02)
    a = sm.mc_simple(n, s1, call, csr, bs=n)
    b = sm.mc_simple(n, s2, call, csr)
    assert abs(a.sample_mean - b.sample_mean) < 2e-6 and abs(a.sample_std - b.sample_std) < 2e-6


def test_path_ranges_are_split_invariant():
    """one call over N paths == two calls over N/2 paths each (global path-id counters)"""
    call, csr = sm.EuroCall(1.0), sm.ConstantShortRate(0.02)
    sde = sm.Merton(0.02, 0.2, 1, -0.05, 0.3, torch.tensor([1.]), 1)
    one = sm.JumpEulerSolver(sde, 3, 30, device=DEV, seed=21)
    two = sm.JumpEulerSolver(sde, 3, 30, device=DEV, seed=21)
    a = sm.mc_simple(40000, one, call, csr, bs=40000, payoff_time='adapted')
    b1 = sm.mc_simple(20000, two, call, csr, bs=20000, payoff_time='adapted')
    b2 = sm.mc_simple(20000, two, call, csr, bs=20000, payoff_time='adapted')
    assert abs(a.sample_mean - 0.5 * (b1.sample_mean + b2.sample_mean)) < 1e-9


def test_c2_gbm_price_vs_black_scholes():
    """C2 statistical acceptance at 2e8 paths: |est - BS| <= 1.96 se + Euler weak bias"""
    p = sm.BlackScholesEuroCall.default_params(252, DEV)
    st = sm.mc_simple(2 * 10 ** 8, p.solver, p.payoff, p.discounter, bs=10 ** 6)
    bs_price = sm.bs_call(1, 1, 3, 0.02, 0.3)
    assert abs(bs_price - 0.22943206) < 1e-7
    assert abs(st.sample_mean - bs_price) <= 1.96 * st.sample_std + 7.6e-5
    ref = golden_json("ref_stats")["c2_gbm_2e5x252"]
    assert abs(st.sample_mean - ref["mean"]) <= 1.96 * math.hypot(ref["se"], st.sample_std)


@pytest.mark.parametrize("strategy", [0, 1, 2])
def test_c1_merton_price_vs_series_and_reference_ci(strategy):
    sde = sm.Merton(0.02, 0.2, 1, -0.05, 0.3, torch.tensor([1.]), 1)
    solver = sm.JumpEulerSolver(sde, 3, 100, device=DEV)
    solver.jump_strategy = strategy
    call, csr = sm.EuroCall(1.0), sm.ConstantShortRate(0.02)
    st = sm.mc_simple(4 * 10 ** 7, solver, call, csr, bs=10 ** 5, payoff_time='adapted')
    exact = sm.merton_call(1, 1, 3, 0.02, 0.2, -0.05, 0.3, 1)
    assert abs(exact - 0.26298121) < 1e-7
    assert abs(st.sample_mean - exact) <= 1.96 * st.sample_std + 2e-4  # + Euler weak bias at h = 0.03
    ref = golden_json("ref_stats")
    for mode in ("adapted", "terminal"):
        r = ref["c1_merton_1e5x100_" + mode]
        s = sm.mc_simple(4 * 10 ** 6, solver, call, csr, bs=10 ** 5, payoff_time=mode)
        assert abs(s.sample_mean - r["mean"]) <= 1.96 * math.hypot(r["se"], s.sample_std), mode
    # the 'terminal' index (quirk Q1) is visibly biased low, as in the reference
    lo = sm.mc_simple(4 * 10 ** 6, solver, call, csr, bs=10 ** 5, payoff_time='terminal')
    assert lo.sample_mean < exact - 2.5 * lo.sample_std


@pytest.mark.parametrize("rho", [None, 0.4])
def test_c4_levy_rainbow_inside_reference_ci(rho):
    corr = None if rho is None else sm.get_corr_matrix([rho])
    levy = sm.ExpExampleLevy(1, 1, 0.5, 2, 0.02, 0.3, 0.2, 0.001, dim=2)
    solver = sm.JumpEulerSolver(sm.LevySde(levy, torch.tensor([1., 1.]), corr_matrix=corr), 3, 256, device=DEV)
    st = sm.mc_simple(10 ** 6, solver, sm.Rainbow(1.0), sm.ConstantShortRate(0.02), bs=10 ** 5, payoff_time='adapted')
    r = golden_json("ref_stats")["c4_levy_6e4x256_" + ("rho0" if rho is None else "rho04")]
    assert abs(st.sample_mean - r["mean"]) <= 1.96 * math.hypot(r["se"], st.sample_std)


def test_heston_inside_reference_ci():
    p = sm.HestonEuroCall.default_params(100, DEV)
    st = sm.mc_simple(4 * 10 ** 6, p.solver, p.payoff, p.discounter, bs=10 ** 5)
    r = golden_json("ref_stats")["heston_1e5x100"]
    assert abs(st.sample_mean - r["mean"]) <= 1.96 * math.hypot(r["se"], st.sample_std)


def test_terminal_cv_reduces_variance_and_keeps_mean():
    p = sm.BlackScholesEuroCall.default_params(64, DEV)
    plain = sm.mc_simple(4 * 10 ** 6, p.solver, p.payoff, p.discounter, bs=10 ** 5)
    cv = sm.mc_terminal_cv(4 * 10 ** 6, p.solver, p.payoff, p.discounter, bs=10 ** 5)
    assert cv.sample_std < 0.6 * plain.sample_std
    assert abs(cv.sample_mean - sm.bs_call(1, 1, 3, 0.02, 0.3)) <= 1.96 * cv.sample_std + 3e-4


def test_large_path_count_properties():
    """BASELINE-size property checks (1e9 paths x 252 steps would take ~0.3 s; use 2.5e8): mean of the terminal
    control D(T) S_T - S_0 is zero (martingale), n and the iteration count are exact."""
    p = sm.BlackScholesEuroCall.default_params(252, DEV)
    n = 250_000_000
    mom = sm._engine.run_moments(p.solver, p.payoff, p.discounter, n, 1).read()
    assert mom["n"] == float(n) and mom["iters"] == float(n) * 252
    se_c = math.sqrt(max(mom["sumsq_c"] / n - (mom["sum_c"] / n) ** 2, 0.0) / n)
    assert abs(mom["sum_c"] / n) <= 3.0 * se_c + 1e-4  # Euler mean is exact for GBM up to fp32 rounding drift
