"""GPU tests of the fused control-variate kernel (E5/E6/E7, tcgen05 tensor cores).

Deterministic parity: per-path gamma against the reference's apply_adapted_control_variates /
apply_diffusion_control_variate on the same injected noise and the same (reference-trained) weights.
Tolerance: the payoff part is fp32-exact (1e-5); the control-variate part evaluates the MLPs with bf16 weights and
hidden activations (fp32 inputs via hi/lo split, fp32 accumulation), so gamma is compared at 5e-3 absolute
(|gamma| is O(0.1 - 1); SURVEY.md C3 asks for ~1e-2 relative on the CV term).  Statistical: unbiasedness against the
Merton series and variance reduction equal to the fp32 PyTorch application of the same nets."""
import math

import numpy as np
import pytest
import torch

from common import golden, rel_err, sm, t

pytestmark = pytest.mark.gpu
DEV = "cuda"
CV_ATOL = 5e-3


def _net_from_golden(g, prefix, in_dim=2, out_dim=1):
    net = sm.Mlp(in_dim, [50, 50, 50], out_dim, batch_norm=False, batch_norm_init=False, device=DEV)
    lin = [m for m in net.net if isinstance(m, torch.nn.Linear)]
    with torch.no_grad():
        for i, l in enumerate(lin):
            l.weight.copy_(torch.as_tensor(g["%s_w%d" % (prefix, i)]))
            l.bias.copy_(torch.as_tensor(g["%s_b%d" % (prefix, i)]))
    return net.eval()


def test_fused_cv_jump_gamma_vs_reference_golden():
    g = golden("cv_merton_1d")
    sde = sm.Merton(0.02, 0.2, 1, -0.05, 0.3, t(g["x0"]), 1)
    solver = sm.JumpEulerSolver(sde, 3.0, int(g["z"].shape[1]) - int(g["max_jumps"]), device=DEV)
    f, gnet = _net_from_golden(g, "f"), _net_from_golden(g, "g")
    assert sm.fused_cv_supported([f, gnet], solver)
    n = g["z"].shape[0]
    mom, gam = sm.mc_cv_fused([f, gnet], solver, n, sm.EuroCall(1.0), sm.ConstantShortRate(0.02),
                              inject=dict(z=g["z"], jump_times=g["jump_times"], marks=g["marks"],
                                          total_steps=int(g["total_steps"])), gamma_out=True)
    gam = gam.cpu().numpy()
    m = mom.read()
    err = np.max(np.abs(gam - g["cv_gamma"]))
    print("fused CV (jump) max |gamma - ref| = %.3e over %d paths" % (err, n))
    assert err < CV_ATOL
    assert abs(m["sum"] - float(g["sum_gamma"])) < CV_ATOL * n
    # the plain payoff rides along in the control slot and is fp32-exact
    assert abs(m["sum_c"] - float(np.sum(g["payoffs"].astype(np.float64)))) < 1e-4


def _bf16_emulated_gamma(g, solver, nets, payoff_kind):
    """apply_adapted_control_variates (varred.py:98-131) on the oracle's trajectories of the golden's injected noise
    with the precision of the fused kernel: weights, biases and hidden activations rounded to bf16, inputs as bf16
    hi + lo pairs, fp32 accumulation"""
    from common import oracle, oracle_sde
    zc = g["zc"] if "zc" in g.files else None
    osde = oracle_sde(solver)
    res = oracle.jump(osde, g["z"], zc, g["jump_times"], g["marks"])
    S = int(res["total_steps"])
    n = res["paths"].shape[0]
    bf = lambda v: v.to(torch.bfloat16).to(torch.float32)

    def hilo(v):
        hi = bf(v)
        return hi + bf(v - hi)

    def forward(net, inp):
        lin = [m for m in net.net if isinstance(m, torch.nn.Linear)]
        h = hilo(inp)
        for i, l in enumerate(lin):
            h = h @ bf(l.weight.detach().cpu()).t() + bf(l.bias.detach().cpu())
            if i < 3:
                h = bf(torch.relu(h))
        return h

    T = torch.as_tensor(res["times"][:, :S])
    D = torch.exp(-T * 0.02)
    P, Lf = torch.as_tensor(res["paths"][:, :S]), torch.as_tensor(res["left"][:, :S])
    J = torch.as_tensor(res["jumps"][:, :S])
    N = torch.as_tensor(res["normals"][:, :S]).reshape(n, S, -1)
    fo = forward(nets[0], torch.cat([T.unsqueeze(-1), P], -1).reshape(n * S, -1)).reshape(n, S, -1)
    go = forward(nets[1], torch.cat([T.unsqueeze(-1), Lf], -1).reshape(n * S, -1)).reshape(n, S, -1)
    bcv = ((N * fo).sum(-1) * D).sum(-1)
    jcv = (go * D.unsqueeze(-1) * J).sum(-1).sum(-1)
    h = torch.diff(torch.as_tensor(res["times"][:, :S + 1]), dim=1)[:, :S - 1]
    comp = (-float(solver.sde.jump_rate().sum()) * float(solver.sde.jump_mean()) * go[:, :S - 1] *
            D[:, :S - 1].unsqueeze(-1) * h.unsqueeze(-1)).sum(-1).sum(-1)
    last = res["paths"][np.arange(n), res["iters"]]
    pay = oracle.payoff(oracle.payoff_struct(payoff_kind, 1.0), last) * np.float32(np.exp(-0.06))
    return pay + (bcv + jcv + comp).numpy()


def _levy_solver(eps, steps, **kw):
    levy = sm.ExpExampleLevy(1, 1, 0.5, 2, 0.02, 0.3, 0.2, eps, dim=2)
    return sm.JumpEulerSolver(sm.LevySde(levy, torch.tensor([1., 1.])), 3.0, steps, device=DEV, **kw)


def test_fused_cv_levy_2d_gamma_vs_reference_golden():
    """the configuration of levy_rainbow_cv_experiment.py:39-40 in the fused kernel: 2-D 'indep' exp-Levy SDE, rainbow
    payoff, f = Mlp(3, [50, 50, 50], 4), g = Mlp(3, [50, 50, 50], 2); per-path gamma of the unmodified reference"""
    g = golden("cv_levy_2d")
    solver = _levy_solver(float(g["eps"]), int(g["z"].shape[1]) - int(g["max_jumps"]))
    f, gnet = _net_from_golden(g, "f", 3, 4), _net_from_golden(g, "g", 3, 2)
    assert sm.fused_cv_supported([f, gnet], solver)
    n = g["z"].shape[0]
    mom, gam = sm.mc_cv_fused([f, gnet], solver, n, sm.Rainbow(1.0), sm.ConstantShortRate(0.02),
                              inject=dict(z=g["z"], zc=g["zc"], jump_times=g["jump_times"], marks=g["marks"],
                                          total_steps=int(g["total_steps"])), gamma_out=True)
    gam = gam.cpu().numpy()
    m = mom.read()
    err = np.max(np.abs(gam - g["cv_gamma"]))
    # what the kernel's arithmetic should give: the reference's formula with bf16 weights / hidden activations and
    # fp32 accumulation (the tensor-core precision), evaluated in PyTorch on the oracle's trajectories
    emu = _bf16_emulated_gamma(g, solver, [f, gnet], 5)
    err_emu = np.max(np.abs(gam - emu))
    print("fused CV (Levy 2-D) over %d paths: max |gamma - ref fp32| = %.3e, max |gamma - bf16 emulation| = %.3e, "
          "max |emulation - ref| = %.3e" % (n, err, err_emu, np.max(np.abs(emu - g["cv_gamma"]))))
    assert err_emu < 2e-3                      # the kernel computes the reference's formula at tensor-core precision
    assert err < 2e-2                          # ... whose distance to the fp32 reference is bf16 rounding: six outputs
    assert abs(np.mean(gam - g["cv_gamma"])) < 4e-3     # x 57 iterations x marks up to |J| ~ 10; zero-mean, so unbiased
    assert abs(m["sum"] - float(np.sum(gam.astype(np.float64)))) < 1e-4
    assert abs(m["sum_c"] - float(np.sum(g["payoffs"].astype(np.float64)))) < 2e-4


def test_fused_cv_levy_2d_unbiased_and_equal_to_the_stored_path_application():
    """Philox-driven: (1) mc_apply_cvs takes the fused kernel for the Levy-rainbow nets, (2) its estimate agrees with
    plain Monte Carlo (any adapted f, g is an unbiased control variate), (3) the variance of gamma equals that of
    the PyTorch application of the same nets on stored trajectories within 10 % (bf16 weights / activations)"""
    g = golden("cv_levy_2d")
    f, gnet = _net_from_golden(g, "f", 3, 4), _net_from_golden(g, "g", 3, 2)
    opt, csr = sm.Rainbow(1.0), sm.ConstantShortRate(0.02)
    n = 400_000
    solver = _levy_solver(0.05, 40, seed=5)
    fused = sm.mc_apply_cvs([f, gnet], solver, n, opt, csr, sim_bs=10 ** 5, bs=2000)
    plain = sm.mc_simple(4 * n, _levy_solver(0.05, 40, seed=6), opt, csr, bs=10 ** 5, payoff_time='adapted')
    assert abs(fused.sample_mean - plain.sample_mean) < 4 * math.hypot(fused.sample_std, plain.sample_std)
    from sde_mc_b200 import varred
    varred.FUSED_CV_ENABLED = False
    try:
        stored = sm.mc_apply_cvs([f, gnet], _levy_solver(0.05, 40, seed=7), 40_000, opt, csr, sim_bs=20_000, bs=2000)
    finally:
        varred.FUSED_CV_ENABLED = True
    var_fused, var_stored = fused.sample_std ** 2 * n, stored.sample_std ** 2 * 40_000
    assert abs(var_fused / var_stored - 1.0) < 0.10, (var_fused, var_stored)


def test_fused_cv_diffusion_gamma_vs_reference_golden():
    g = golden("cv_gbm_1d")
    solver = sm.EulerSolver(sm.Gbm(0.02, 0.3, t(g["x0"]), 1), 3.0, 16, device=DEV)
    f = _net_from_golden(g, "f")
    assert sm.fused_cv_supported(f, solver)
    n = g["z"].shape[0]
    mom, gam = sm.mc_cv_fused(f, solver, n, sm.EuroCall(1.0), sm.ConstantShortRate(0.02), inject=dict(z=g["z"]),
                              gamma_out=True)
    err = np.max(np.abs(gam.cpu().numpy() - g["cv_gamma"]))
    print("fused CV (diffusion) max |gamma - ref| = %.3e over %d paths" % (err, n))
    assert err < CV_ATOL


def test_fused_cv_unbiased_and_matches_torch_application():
    """C3 shape at reduced N: Merton 1-D, 200 steps, reference-trained nets"""
    g = golden("cv_merton_1d")
    sde = sm.Merton(0.02, 0.2, 1, -0.05, 0.3, torch.tensor([1.]), 1)
    f, gnet = _net_from_golden(g, "f"), _net_from_golden(g, "g")
    call, csr = sm.EuroCall(1.0), sm.ConstantShortRate(0.02)
    exact = sm.merton_call(1, 1, 3, 0.02, 0.2, -0.05, 0.3, 1)

    solver = sm.JumpEulerSolver(sde, 3, 200, device=DEV)
    n = 4 * 10 ** 6
    fused = sm.mc_apply_cvs([f, gnet], solver, n, call, csr, sim_bs=10 ** 5, bs=2000)
    plain = sm.mc_simple(n, solver, call, csr, bs=10 ** 5, payoff_time='adapted')
    assert abs(fused.sample_mean - exact) <= 1.96 * fused.sample_std + 2e-4
    assert fused.sample_std < 0.8 * plain.sample_std          # the nets do reduce variance

    # same nets applied by PyTorch (fp32) on trajectories stored by the path-storing kernel
    sm.varred.FUSED_CV_ENABLED = False
    try:
        solver2 = sm.JumpEulerSolver(sde, 3, 200, device=DEV, seed=7)
        n2 = 2 * 10 ** 5
        torch_path = sm.mc_apply_cvs([f, gnet], solver2, n2, call, csr, sim_bs=5 * 10 ** 4, bs=5000)
    finally:
        sm.varred.FUSED_CV_ENABLED = True
    var_fused = fused.sample_std ** 2 * n
    var_torch = torch_path.sample_std ** 2 * n2
    assert abs(var_fused / var_torch - 1.0) < 0.10, (var_fused, var_torch)
    assert abs(torch_path.sample_mean - fused.sample_mean) <= 1.96 * math.hypot(torch_path.sample_std, fused.sample_std)


@pytest.mark.parametrize("hidden", [8, 56, 60])
def test_fused_cv_other_widths_vs_torch_application(hidden):
    """Widths other than the experiments' 50: 56 is the widest net whose epilogues read 56 accumulator columns and
    write the constant-one unit themselves, 60 takes the full-width path (constant carried by W[63][63] = 1).
    Oracle = the reference's algorithm itself: the same (random-init) nets applied by PyTorch in fp32
    (`_adapted_gammas`, varred.py:98-131) to trajectories stored by the path-storing kernel on the same injected noise."""
    from sde_mc_b200.nets import AdaptedPathData
    from sde_mc_b200.varred import _adapted_gammas
    rng = np.random.default_rng(hidden)
    torch.manual_seed(hidden)
    sde = sm.Merton(0.02, 0.2, 1, -0.05, 0.3, torch.tensor([1.]), 1)
    steps, bs = 24, 1000
    solver = sm.JumpEulerSolver(sde, 3.0, steps, device=DEV)
    K = steps + solver.max_jumps
    z = rng.standard_normal((bs, K, 1)).astype(np.float32)
    jt = np.cumsum(rng.exponential(1.0, (bs, solver.max_jumps)), axis=1).astype(np.float32)
    mk = rng.standard_normal((bs, K)).astype(np.float32)
    f = sm.Mlp(2, [hidden] * 3, 1, batch_norm=False, batch_norm_init=False, device=DEV).eval()
    g = sm.Mlp(2, [hidden] * 3, 1, batch_norm=False, batch_norm_init=False, device=DEV).eval()
    assert sm.fused_cv_supported([f, g], solver)
    call, csr = sm.EuroCall(1.0), sm.ConstantShortRate(0.02)

    paths, (normals, times, left, total, jumps) = solver.solve(bs=bs, inject=dict(z=z, jump_times=jt, marks=mk))
    n = total + 1
    payoffs = call(paths[:, total]) * float(csr(3.0))   # D(T) payoff(S_T), mc.py:111-112
    data = AdaptedPathData(paths[:, :n], payoffs, normals[:, :total], left[:, :n], times[:, :n], jumps[:, :n], total)
    batch = ((data.paths, data.normals, data.left_paths, data.time_paths, data.jump_paths), data.payoffs)
    with torch.inference_mode():
        ref = _adapted_gammas([f, g], batch, solver, csr, total, 1, bs, 0).cpu().numpy()

    mom, gam = sm.mc_cv_fused([f, g], solver, bs, call, csr,
                              inject=dict(z=z, jump_times=jt, marks=mk, total_steps=int(total)), gamma_out=True)
    err = np.max(np.abs(gam.cpu().numpy() - ref))
    cv_scale = float(np.abs(ref - payoffs.cpu().numpy()).max())   # random-init nets: control-variate terms are O(1-5)
    print("fused CV, hidden %d: max |gamma - torch fp32| = %.3e, CV term up to %.2f" % (hidden, err, cv_scale))
    # bf16 weights and hidden activations: 1.5e-2 relative to the control-variate term (~1e-2), the
    # payoff part is fp32-exact; a structural error (bias unit, column order, a missing layer) is O(1) relative
    tol = 1.5e-2 * cv_scale + 1e-3
    assert err < tol
    assert abs(mom.read()["sum"] - float(ref.astype(np.float64).sum())) < tol * bs


def test_fused_cv_tol_vs_reference_golden():
    """tol > 0 (integrate_cv varred.py:202-209): the fused kernel drops the f dW terms from index
    remove_steps(tol, steps, T) on (sdemc_mlp.cv_steps); jump and compensator sums are untouched.  Golden: the
    unmodified reference with tol > 0 (tests/golden/make_golden_cv_tol.py)."""
    gt = golden("cv_tol")
    g = golden("cv_gbm_1d")
    solver = sm.EulerSolver(sm.Gbm(0.02, 0.3, t(g["x0"]), 1), 3.0, 16, device=DEV)
    f = _net_from_golden(g, "f")
    n = g["z"].shape[0]
    for tol in (0.5, 1.0):
        assert sm.fused_cv_supported(f, solver, tol)
        _, gam = sm.mc_cv_fused(f, solver, n, sm.EuroCall(1.0), sm.ConstantShortRate(0.02), inject=dict(z=g["z"]),
                                gamma_out=True, tol=tol)
        err = np.max(np.abs(gam.cpu().numpy() - gt["gbm_tol%g" % tol]))
        assert err < CV_ATOL, (tol, err)
    g = golden("cv_merton_1d")
    sde = sm.Merton(0.02, 0.2, 1, -0.05, 0.3, t(g["x0"]), 1)
    solver = sm.JumpEulerSolver(sde, 3.0, int(g["z"].shape[1]) - int(g["max_jumps"]), device=DEV)
    f, gnet = _net_from_golden(g, "f"), _net_from_golden(g, "g")
    assert not sm.fused_cv_supported([f, gnet], solver, 0.5)     # production: the cut index needs the batch
    _, gam = sm.mc_cv_fused([f, gnet], solver, g["z"].shape[0], sm.EuroCall(1.0), sm.ConstantShortRate(0.02),
                            inject=dict(z=g["z"], jump_times=g["jump_times"], marks=g["marks"],
                                        total_steps=int(g["total_steps"])), gamma_out=True, tol=0.5)
    err = np.max(np.abs(gam.cpu().numpy() - gt["merton_tol0.5"]))
    assert err < CV_ATOL, err


def test_mc_apply_cvs_with_tol_fused_equals_stored_route():
    """mc_apply_cvs(..., tol) on a diffusion: the fused kernel against the reference-style application (PyTorch on
    trajectories stored by the path-storing kernel) -- same estimate within the error bars, same variance."""
    from sde_mc_b200 import varred
    g = golden("cv_gbm_1d")
    f = _net_from_golden(g, "f")
    call, csr = sm.EuroCall(1.0), sm.ConstantShortRate(0.02)
    mk = lambda seed: sm.EulerSolver(sm.Gbm(0.02, 0.3, torch.tensor([1.]), 1), 3.0, 16, device=DEV, seed=seed)
    n = 400_000
    fused = sm.mc_apply_cvs(f, mk(1), n, call, csr, sim_bs=10 ** 5, bs=2000, tol=0.5)
    varred.FUSED_CV_ENABLED = False
    try:
        stored = sm.mc_apply_cvs(f, mk(7), n, call, csr, sim_bs=10 ** 5, bs=2000, tol=0.5)
    finally:
        varred.FUSED_CV_ENABLED = True
    assert abs(float(fused.sample_mean) - float(stored.sample_mean)) <= 3 * math.hypot(float(fused.sample_std), float(stored.sample_std))
    assert abs(float(fused.sample_std) ** 2 / float(stored.sample_std) ** 2 - 1.0) < 0.10
