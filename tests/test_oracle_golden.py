"""Pins the CPU oracle (oracle/sde_oracle.c) against the golden vectors produced by the unmodified reference
(tests/golden/make_golden.py).  CPU only.  Tolerances: fp32 cases 2e-6 relative per stored value (different libm
vs torch's vectorised exp/log/pow, FMA-free both sides); fp64 cases 1e-12."""
import numpy as np
import pytest

from common import (DIFFUSION_CASES, JUMP_CASES, MLMC_CASES, golden, golden_json, jump_solver, mlmc_levels, oracle,
                    oracle_sde, rel_err, sm, t)

F32_RTOL = 2e-6


@pytest.mark.parametrize("name", sorted(DIFFUSION_CASES))
def test_diffusion_solver_matches_reference(name):
    g = golden(name)
    build, solver_cls = DIFFUSION_CASES[name]
    solver = solver_cls(build(g), float(g["T"]), int(g["z"].shape[1]))
    paths, normals = oracle.diffusion(oracle_sde(solver), g["z"])
    assert paths.shape == g["paths"].shape
    assert rel_err(paths, g["paths"]) < F32_RTOL
    assert rel_err(normals, g["normals"], floor=1e-2) < F32_RTOL


def test_diffusion_pair_matches_reference():
    g = golden("diff_gbm_mlmc_8_2")
    solver = sm.EulerSolver(sm.Gbm(0.02, 0.3, t(g["x0"]), 1), float(g["T"]), 8)
    pf, pc = oracle.diffusion_pair(oracle_sde(solver, 8), 8, 2, g["z"])
    assert rel_err(pf, g["paths_fine"]) < F32_RTOL
    assert rel_err(pc, g["paths_coarse"]) < F32_RTOL


@pytest.mark.parametrize("fine,coarse", [(8, 2), (16, 8)])
def test_heston_pair_matches_reference(fine, coarse):
    """HestonSolver.multilevel_solve (solvers.py:90-119 on HestonScheme steps, schemes.py:16-22), fp32"""
    g = golden("mlmc_heston_%d_%d" % (fine, coarse))
    sde = sm.Heston(float(g["r"]), float(g["kappa"]), float(g["theta"]), float(g["xi"]), float(g["rho"]), t(g["x0"]))
    solver = sm.HestonSolver(sde, float(g["T"]), fine)
    pf, pc = oracle.diffusion_pair(oracle_sde(solver, fine), fine, coarse, g["z"])
    assert rel_err(pf, g["paths_fine"]) < 5e-6
    assert rel_err(pc, g["paths_coarse"]) < 5e-6


@pytest.mark.parametrize("name", sorted(JUMP_CASES))
def test_jump_solver_matches_reference(name):
    g = golden(name)
    solver = jump_solver(name, g)
    assert solver.max_jumps == int(g["max_jumps"])
    zc = g["zc"] if "zc" in g.files else None
    res = oracle.jump(oracle_sde(solver), g["z"], zc, g["jump_times"], g["marks"])
    ts = int(g["total_steps"])
    assert res["total_steps"] == ts
    # Levy marks reach |J| ~ 10 and several hundred iterations compound: allow a little more there
    tol = F32_RTOL if "merton" in name else 2e-5
    # log-price (arithmetic) models cross zero: measure their error against the O(1) scale of the state
    floor = 1.0 if name in ("jump_addlevy_1d", "jump_levy2d") else 1e-3
    assert rel_err(res["paths"][:, :ts + 1], g["paths"], floor) < tol
    assert rel_err(res["left"][:, :ts + 1], g["left_paths"], floor) < tol
    assert rel_err(res["times"][:, :ts + 1], g["time_paths"]) < 1e-6
    assert rel_err(res["jumps"][:, :ts + 1], g["jump_paths"], floor=1e-2) < 1e-5
    assert rel_err(res["normals"][:, :ts], g["normals"], floor=1e-2) < F32_RTOL


@pytest.mark.parametrize("name", sorted(MLMC_CASES))
def test_jump_pair_matches_reference_fp64(name):
    g = golden(name)
    fine, coarse = mlmc_levels(name)
    sde = MLMC_CASES[name](g)
    solver = sm.JumpEulerSolver(sde, float(g["T"]), fine, exact_jumps=bool(int(g["exact_jumps"])))
    zc = g["zc"] if "zc" in g.files else None
    fl, cl, iters, total = oracle.jump_pair(oracle_sde(solver, fine), fine, coarse, g["z"], zc, g["jump_times"],
                                            g["marks"], np.float64)
    assert total > 0
    assert rel_err(fl, g["fine_last"]) < 1e-12
    assert rel_err(cl, g["coarse_last"]) < 1e-12
    if "paths_fine" in g.files:
        assert total == g["paths_fine"].shape[1] - 1


@pytest.mark.parametrize("name", ["mlmc_merton_8_2_ex0", "mlmc_merton_16_8_ex1"])
def test_jump_pair_fp32_oracle_close_to_fp64(name):
    """the fp32 pair (with the dt >= 0 clamp) tracks the fp64 reference run to fp32 accuracy"""
    g = golden(name)
    fine, coarse = mlmc_levels(name)
    solver = sm.JumpEulerSolver(MLMC_CASES[name](g), float(g["T"]), fine, exact_jumps=bool(int(g["exact_jumps"])))
    fl, cl, _, total = oracle.jump_pair(oracle_sde(solver, fine), fine, coarse, g["z"], None, g["jump_times"],
                                        g["marks"], np.float32)
    assert total > 0
    assert rel_err(fl, g["fine_last"]) < 2e-5
    assert rel_err(cl, g["coarse_last"]) < 2e-5


def test_payoffs_match_reference():
    g = golden("payoffs")
    kinds = {"euro_call": (0, 1.0, False, 1.0, 1.0), "euro_put": (1, 1.0, False, 1.0, 1.0),
             "binary_aon": (2, 1.0, False, 1.0, 1.0), "basket_arith": (3, 1.0, False, 1.0, 1.0),
             "basket_geom": (4, 1.0, False, 1.0, 1.0), "rainbow": (5, 1.0, False, 1.0, 1.0),
             "digital": (6, 1.0, False, 1.0, 1.0), "asian_call": (7, 0.3, False, 1.0, 3.0),
             "heston_rainbow": (8, 1.0, False, 1.0, 1.0), "best_of": (9, 1.0, False, 1.0, 1.0),
             "euro_call_disc": (0, 0.9, False, 0.94, 1.0), "euro_call_log": (0, 1.0, True, 0.94, 1.0),
             "rainbow_log": (5, 1.0, True, 0.94, 1.0), "asian_call_log": (7, 1.0, True, 1.0, 3.0)}
    checked = 0
    for dim in (1, 2, 3, 4):
        x = g["x%d" % dim]
        for key, (kind, strike, log, disc, aux) in kinds.items():
            gk = "%s_%d" % (key, dim)
            if gk not in g.files:
                continue
            xin = np.log(x) if log else x
            out = oracle.payoff(oracle.payoff_struct(kind, strike, log, disc, aux), xin)
            assert rel_err(out, g[gk]) < 2e-6, gk
            checked += 1
    assert checked == 54


def test_icdf_matches_reference():
    g = golden("icdf")
    ic = sm.InverseCdf(1, 1, 2, 0.5, 0.01)
    cf = golden_json("closed_forms")["icdf_params"]
    assert abs(ic.lda - cf["lda"]) < 1e-12 and abs(ic.y2 - cf["y2"]) < 1e-15
    out = oracle.icdf(ic.mark_params(), g["u"])
    assert rel_err(out, g["x"], floor=1e-2) < 5e-6


def _mlps(g, prefix):
    return oracle.mlp_struct([g["%s_w%d" % (prefix, i)] for i in range(4)], [g["%s_b%d" % (prefix, i)] for i in range(4)])


def test_cv_gamma_jump_matches_reference():
    """apply_adapted_control_variates (varred.py:98-131) per path, on reference-trained nets"""
    g = golden("cv_merton_1d")
    sde = sm.Merton(0.02, 0.2, 1, -0.05, 0.3, t(g["x0"]), 1)
    solver = sm.JumpEulerSolver(sde, 3.0, int(g["z"].shape[1]) - int(g["max_jumps"]))
    osde = oracle_sde(solver)
    res = oracle.jump(osde, g["z"], None, g["jump_times"], g["marks"])
    assert res["total_steps"] == int(g["total_steps"])
    last = res["paths"][np.arange(len(res["iters"])), res["iters"]]
    pay = oracle.payoff(oracle.payoff_struct(0, 1.0), last) * np.float32(np.exp(-0.02 * 3.0))
    assert rel_err(pay, g["payoffs"], floor=1e-2) < 5e-6
    gam = oracle.cv_gamma_jump(osde, res, pay, 0.02, float(g["jump_mean"]), _mlps(g, "f"), _mlps(g, "g"))
    assert np.max(np.abs(gam - g["cv_gamma"])) < 2e-5
    assert abs(float(gam.astype(np.float64).sum()) - float(g["sum_gamma"])) < 2e-4


def test_cv_gamma_levy_2d_matches_reference():
    """the 2-D 'indep' exp-Levy control variates of levy_rainbow_cv_experiment.py:39-40 (f: 4 outputs, one per driver
    of every component; g: 2): per-path gamma of the unmodified reference on injected noise"""
    g = golden("cv_levy_2d")
    levy = sm.ExpExampleLevy(1, 1, 0.5, 2, 0.02, 0.3, 0.2, float(g["eps"]), dim=2)
    sde = sm.LevySde(levy, t(g["x0"]))
    solver = sm.JumpEulerSolver(sde, 3.0, int(g["z"].shape[1]) - int(g["max_jumps"]))
    osde = oracle_sde(solver)
    res = oracle.jump(osde, g["z"], g["zc"], g["jump_times"], g["marks"])
    assert res["total_steps"] == int(g["total_steps"])
    last = res["paths"][np.arange(len(res["iters"])), res["iters"]]
    pay = oracle.payoff(oracle.payoff_struct(5, 1.0), last) * np.float32(np.exp(-0.02 * 3.0))
    assert rel_err(pay, g["payoffs"], floor=1e-2) < 2e-5
    gam = oracle.cv_gamma_jump(osde, res, pay, 0.02, float(g["jump_mean"]), _mlps(g, "f"), _mlps(g, "g"))
    assert np.max(np.abs(gam - g["cv_gamma"])) < 5e-5


def test_cv_gamma_diffusion_matches_reference():
    g = golden("cv_gbm_1d")
    solver = sm.EulerSolver(sm.Gbm(0.02, 0.3, t(g["x0"]), 1), 3.0, 16)
    osde = oracle_sde(solver)
    paths, normals = oracle.diffusion(osde, g["z"])
    pay = oracle.payoff(oracle.payoff_struct(0, 1.0), paths[:, -1]) * np.float32(np.exp(-0.02 * 3.0))
    assert rel_err(pay, g["payoffs"], floor=1e-2) < 5e-6
    gam = oracle.cv_gamma_diffusion(osde, paths, normals, pay, 0.02, _mlps(g, "f"))
    assert np.max(np.abs(gam - g["cv_gamma"])) < 2e-5


def test_philox_known_answer():
    """Random123 known-answer vectors for philox4x32-10 (kat_vectors: zero and all-ones / pi inputs)"""
    assert oracle.philox(0, 0, 0, 0, 0, 0) == [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]
    assert oracle.philox(0xffffffff, 0xffffffff, 0xffffffff, 0xffffffff, 0xffffffff, 0xffffffff) == \
        [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]
    assert oracle.philox(0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344, 0xa4093822, 0x299f31d0) == \
        [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]


def test_cv_gamma_with_tol_matches_reference():
    """integrate_cv's `tol` (varred.py:202-209, remove_steps helpers.py:71-74): only the Brownian sum is cut, at an
    index of num_steps (diffusion) resp. of the batch's total_steps (jump solver).  Golden: the unmodified reference
    with tol > 0 on the nets and noise of cv_gbm_1d / cv_merton_1d (tests/golden/make_golden_cv_tol.py)."""
    gt = golden("cv_tol")
    g = golden("cv_gbm_1d")
    solver = sm.EulerSolver(sm.Gbm(0.02, 0.3, t(g["x0"]), 1), 3.0, 16)
    osde = oracle_sde(solver)
    paths, normals = oracle.diffusion(osde, g["z"])
    pay = oracle.payoff(oracle.payoff_struct(0, 1.0), paths[:, -1]) * np.float32(np.exp(-0.02 * 3.0))
    for tol in (0.5, 1.0):
        keep = oracle.remove_steps(tol, 16, 3.0)
        assert keep == int(gt["gbm_keep%g" % tol])
        gam = oracle.cv_gamma_diffusion(osde, paths, normals, pay, 0.02, _mlps(g, "f"), keep)
        assert np.max(np.abs(gam - gt["gbm_tol%g" % tol])) < 2e-5
        assert np.max(np.abs(gam - g["cv_gamma"])) > 1e-2          # the cut matters on this fixture
    g = golden("cv_merton_1d")
    sde = sm.Merton(0.02, 0.2, 1, -0.05, 0.3, t(g["x0"]), 1)
    solver = sm.JumpEulerSolver(sde, 3.0, int(g["z"].shape[1]) - int(g["max_jumps"]))
    osde = oracle_sde(solver)
    res = oracle.jump(osde, g["z"], None, g["jump_times"], g["marks"])
    last = res["paths"][np.arange(len(res["iters"])), res["iters"]]
    pay = oracle.payoff(oracle.payoff_struct(0, 1.0), last) * np.float32(np.exp(-0.02 * 3.0))
    keep = oracle.remove_steps(0.5, int(res["total_steps"]), 3.0)
    assert keep == int(gt["merton_keep0.5"])
    gam = oracle.cv_gamma_jump(osde, res, pay, 0.02, float(g["jump_mean"]), _mlps(g, "f"), _mlps(g, "g"), keep)
    assert np.max(np.abs(gam - gt["merton_tol0.5"])) < 2e-5
