#!/usr/bin/env python
"""Generate the golden fixtures in tests/golden/ by running the UNMODIFIED reference
(piers-hinds/sde_mc at /root/reference) with injected noise.

Run only in the build container (the reference is not present on the GPU box):

    python tests/golden/make_golden.py

Nothing in tests/, bench.py or __graft_entry__.py imports this file; they read the .npz/.json it wrote.

How the injection works: three sampling hooks of a reference solver instance are
replaced so that the loop consumes OUR arrays instead of torch's global RNG --
  sample_corr_normals(size, h, corr)  /root/reference/sde_mc/solvers.py:51-56
  sample_jump_times(size)             /root/reference/sde_mc/solvers.py:143-144
  sample_one_jump(size)               /root/reference/sde_mc/solvers.py:146-148
everything else (step loop, schemes, coefficient functions, payoffs, estimators) is the reference's own code.
"""
import json
import math
import os
import sys
import types

import numpy
import numpy as np

numpy.math = math  # NumPy-2 shim for /root/reference/sde_mc/options.py:96
sys.path.insert(0, "/root/reference")
import torch  # noqa: E402
import sde_mc as ref  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
RNG = np.random.default_rng(20261017)


def save(name, **arrays):
    out = {}
    for k, v in arrays.items():
        if torch.is_tensor(v):
            v = v.detach().cpu().numpy()
        out[k] = np.asarray(v)
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    print("wrote", name, {k: v.shape for k, v in out.items()})


# ----------------------------------------------------------------------------------------------
# injection adaptors
# ----------------------------------------------------------------------------------------------
def inject_diffusion(solver, z):
    """z: (bs, steps, dim, m) unit normals. DiffusionSolver draws everything in one call
    (/root/reference/sde_mc/solvers.py:77-81)."""
    zt = torch.as_tensor(z)

    def sample_corr_normals(self, size, h, corr=True):
        assert tuple(size) == tuple(zt.shape), (size, zt.shape)
        normals = zt * torch.sqrt(h)
        return torch.matmul(self.lower_cholesky, normals).squeeze(-1)

    solver.sample_corr_normals = types.MethodType(sample_corr_normals, solver)


class JumpInjector:
    """Feeds z[:, k] / zc[:, k] / marks[:, k] on loop iteration k, and the given jump times."""

    def __init__(self, solver, z, zc, jump_times, marks, mark_kind):
        self.k = 0          # index of sample_corr_normals(corr=True) calls  == diffusion sub-step index
        self.j = 0          # index of sample_one_jump calls                  == outer iteration index
        self.z = torch.as_tensor(z)                      # (bs, K, dim)
        self.zc = None if zc is None else torch.as_tensor(zc)   # (bs, K)
        self.jt = torch.as_tensor(jump_times)            # (bs, max_jumps)
        self.marks = torch.as_tensor(marks)              # (bs, K)
        inj = self

        def sample_corr_normals(self, size, h, corr=True):
            if corr:
                n = inj.z[:, inj.k].unsqueeze(-1) * torch.sqrt(h)
                out = torch.matmul(self.lower_cholesky, n).squeeze(-1)
                if inj.zc is None:
                    inj.k += 1
                return out
            n = inj.zc[:, inj.k].reshape(-1, 1, 1) * torch.sqrt(h)
            inj.k += 1
            return n.squeeze(-1)

        def sample_jump_times(self, size):
            assert size[1] == inj.jt.shape[1]
            return inj.jt.unsqueeze(-1).clone()

        def sample_one_jump(self, size):
            raw = inj.marks[:, inj.j].unsqueeze(-1)
            inj.j += 1
            if mark_kind == "lognormal":   # /root/reference/sde_mc/sde.py:325-326
                sde = self.sde.base_sde if hasattr(self.sde, "base_sde") else self.sde
                jumps = (raw * sde.gamma + sde.alpha).exp() - 1
            else:                          # /root/reference/sde_mc/levy.py:85-87
                jumps = self.sde.levy.icdf(raw + ref.levy.UNIFORM_TOL / 3)
            return jumps.repeat(1, self.sde.dim)

        solver.sample_corr_normals = types.MethodType(sample_corr_normals, solver)
        solver.sample_jump_times = types.MethodType(sample_jump_times, solver)
        solver.sample_one_jump = types.MethodType(sample_one_jump, solver)


def jump_noise(bs, K, dim, max_jumps, rate, two_drivers, mark_kind, dtype=np.float32):
    z = RNG.standard_normal((bs, K, dim)).astype(dtype)
    zc = RNG.standard_normal((bs, K)).astype(dtype) if two_drivers else None
    gaps = RNG.exponential(1.0 / rate, (bs, max_jumps))
    jt = np.cumsum(gaps, axis=1).astype(dtype)
    if mark_kind == "lognormal":
        marks = RNG.standard_normal((bs, K)).astype(dtype)
    else:
        marks = RNG.random((bs, K)).astype(dtype)
    return z, zc, jt, marks


# ----------------------------------------------------------------------------------------------
# diffusion solver cases  (H3, S1, S2, M1, M2)
# ----------------------------------------------------------------------------------------------
def run_diffusion(name, sde, solver_cls, T, steps, bs, meta):
    solver = solver_cls(sde, T, steps)
    m = sde.brown_dim // sde.dim
    z = RNG.standard_normal((bs, steps, sde.dim, m)).astype(np.float32)
    inject_diffusion(solver, z)
    paths, normals = solver.solve(bs=bs)
    save(name, z=z, paths=paths, normals=normals, chol=solver.lower_cholesky, **meta)


def diffusion_cases():
    run_diffusion("diff_gbm_1d", ref.Gbm(0.02, 0.3, torch.tensor([1.]), 1), ref.EulerSolver, 3, 16, 32,
                  dict(mu=[0.02], sigma=[0.3], x0=[1.0], T=3.0))
    corr = ref.get_corr_matrix([0.7, 0.2, -0.3])
    run_diffusion("diff_gbm_3d_corr", ref.Gbm(0.02, 0.3, torch.ones(3), 3, corr), ref.EulerSolver, 3, 12, 16,
                  dict(mu=[0.02] * 3, sigma=[0.3] * 3, x0=[1.0] * 3, T=3.0, corr=corr))
    run_diffusion("diff_gbm_2d_vec",
                  ref.Gbm(torch.tensor([0.02, 0.05]), torch.tensor([0.2, 0.4]), torch.tensor([1., 2.]), 2),
                  ref.EulerSolver, 2.0, 10, 16, dict(mu=[0.02, 0.05], sigma=[0.2, 0.4], x0=[1.0, 2.0], T=2.0))
    run_diffusion("diff_loggbm", ref.LogGbm(0.02, 0.2, torch.tensor([0.])), ref.EulerSolver, 3, 10, 16,
                  dict(mu=[0.02], sigma=[0.2], x0=[0.0], T=3.0))
    corr2 = ref.get_corr_matrix([0.5])
    run_diffusion("diff_double_gbm_2d", ref.DoubleGbm(0.02, 0.2, 0.1, torch.tensor([1., 1.]), 2, corr2),
                  ref.EulerSolver, 3, 10, 16,
                  dict(mu=[0.02] * 2, sigma1=[0.2] * 2, sigma2=[0.1] * 2, x0=[1.0, 1.0], T=3.0, corr=corr2))
    run_diffusion("diff_heston", ref.Heston(0.02, 0.25, 0.5, 0.3, -0.3, torch.tensor([1., 0.15])),
                  ref.HestonSolver, 3, 20, 32,
                  dict(r=0.02, kappa=0.25, theta=0.5, xi=0.3, rho=-0.3, x0=[1.0, 0.15], T=3.0))
    run_diffusion("diff_asian_gbm", ref.AsianWrapper(ref.Gbm(0.02, 0.3, torch.tensor([1.]), 1)), ref.EulerSolver,
                  3, 12, 16, dict(mu=[0.02], sigma=[0.3], x0=[1.0, 0.0], T=3.0))

    # H4: coupled fine/coarse uniform-grid pair
    gbm = ref.Gbm(0.02, 0.3, torch.tensor([1.]), 1)
    solver = ref.EulerSolver(gbm, 3, 8)
    z = RNG.standard_normal((16, 8, 1, 1)).astype(np.float32)
    inject_diffusion(solver, z)
    (pf, pc), normals = solver.multilevel_solve(16, (8, 2))
    save("diff_gbm_mlmc_8_2", z=z, paths_fine=pf, paths_coarse=pc, mu=[0.02], sigma=[0.3], x0=[1.0], T=3.0)


# ----------------------------------------------------------------------------------------------
# jump-adapted solver cases  (H10, M3, M4, M5)
# ----------------------------------------------------------------------------------------------
def run_jump(name, sde, T, steps, bs, mark_kind, exact_jumps, meta):
    solver = ref.JumpEulerSolver(sde, T, steps, exact_jumps=exact_jumps)
    K = steps + solver.max_jumps
    two = sde.diffusion_struct != 'diag'
    z, zc, jt, marks = jump_noise(bs, K, sde.dim, solver.max_jumps, float(sde.jump_rate().sum()), two, mark_kind)
    JumpInjector(solver, z, zc, jt, marks, mark_kind)
    paths, (normals, time_paths, left_paths, total_steps, jump_paths) = solver.solve(bs=bs)
    extra = {} if zc is None else {"zc": zc}
    save(name, z=z, jump_times=jt, marks=marks, paths=paths, normals=normals[:, :total_steps],
         time_paths=time_paths[:, :total_steps + 1, 0], left_paths=left_paths[:, :total_steps + 1],
         jump_paths=jump_paths[:, :total_steps + 1], total_steps=total_steps, max_jumps=solver.max_jumps,
         exact_jumps=int(exact_jumps), chol=solver.lower_cholesky, **extra, **meta)
    return solver


def merton_meta(mu, sigma, rate, alpha, gamma, x0, T):
    return dict(mu=mu, sigma=sigma, rate=rate, alpha=alpha, gamma=gamma, x0=x0, T=T,
                jump_mean=np.exp(alpha + 0.5 * gamma * gamma) - 1)


def levy_meta(levy, sde, x0, T):
    ic = levy.icdf
    return dict(cm=ic.cm, cp=ic.cp, lmu=ic.mu, lalpha=ic.alpha, eps=ic.eps, lda=ic.lda, y1=ic.y1, y2=ic.y2, y3=ic.y3,
                gamma_eps=levy.gamma(), beta_eps=levy.beta(), x0=x0, T=T)


def jump_cases():
    for ex in (False, True):
        run_jump("jump_merton_1d_ex%d" % ex, ref.Merton(0.02, 0.2, 1, -0.05, 0.3, torch.tensor([1.]), 1), 3, 20, 64,
                 "lognormal", ex, merton_meta(0.02, 0.2, 1.0, -0.05, 0.3, [1.0], 3.0))
    corr = ref.get_corr_matrix([0.4])
    run_jump("jump_merton_2d_corr", ref.Merton(0.02, 0.3, 2, -0.05, 0.3, torch.tensor([1., 1.]), 2, corr), 3, 10, 32,
             "lognormal", False, merton_meta(0.02, 0.3, 2.0, -0.05, 0.3, [1.0, 1.0], 3.0))
    run_jump("jump_asian_merton", ref.AsianWrapper(ref.Merton(0.02, 0.2, 1, -0.05, 0.3, torch.tensor([1.]), 1)), 3, 12,
             32, "lognormal", False, merton_meta(0.02, 0.2, 1.0, -0.05, 0.3, [1.0, 0.0], 3.0))

    # C4 model: LevySde(ExpExampleLevy), rho = 0 (shipped example) and rho = 0.4 (problem.py:156)
    for tag, corr in (("rho0", None), ("rho04", ref.get_corr_matrix([0.4]))):
        levy = ref.ExpExampleLevy(1, 1, 0.5, 2, 0.02, 0.3, 0.2, 0.001, dim=2)
        sde = ref.LevySde(levy, torch.tensor([1., 1.]), corr_matrix=corr)
        meta = levy_meta(levy, sde, [1.0, 1.0], 3.0)
        meta.update(r=0.02, sigma=0.3, f=0.2)
        run_jump("jump_explevy_2d_" + tag, sde, 3, 8, 8, "icdf", False, meta)
    # exact_jumps variant with a coarser epsilon (fewer jumps)
    levy = ref.ExpExampleLevy(1, 1, 0.5, 2, 0.02, 0.3, 0.2, 0.05, dim=2)
    sde = ref.LevySde(levy, torch.tensor([1., 1.]))
    meta = levy_meta(levy, sde, [1.0, 1.0], 3.0)
    meta.update(r=0.02, sigma=0.3, f=0.2)
    run_jump("jump_explevy_2d_eps05_ex1", sde, 3, 16, 16, "icdf", True, meta)

    # additive (log-price) Levy: LevyCall preset problem.py:109-121 and the Levy2d class
    chol = torch.tensor([[1.]])
    levy = ref.ExampleLevy(1, 1, 0.5, 2, 0.02, torch.tensor([0.2]), torch.tensor([0.2]), chol, 0.01, 1)
    sde = ref.LevySde(levy, torch.tensor([0.]))
    meta = levy_meta(levy, sde, [0.0], 3.0)
    meta.update(drift=(levy.drift(0, torch.zeros(1, 1)) - levy.jumps(0, torch.zeros(1, 1), 1) * levy.gamma())[0],
                sigma=[0.2], f=[0.2])
    run_jump("jump_addlevy_1d", sde, 3, 10, 16, "icdf", False, meta)
    levy = ref.Levy2d(1.2, 0.8, 0.5, 2, 0.15, 0.02)
    sde = ref.LevySde(levy, torch.tensor([0., 0.]))
    meta = levy_meta(levy, sde, [0.0, 0.0], 3.0)
    meta.update(drift=(levy.drift(0, torch.zeros(1, 2)) - levy.jumps(0, torch.zeros(1, 2), 1) * levy.gamma())[0],
                sigma=[1.0, 1.0], f=[0.15, 0.15])
    run_jump("jump_levy2d", sde, 3, 10, 8, "icdf", False, meta)


# ----------------------------------------------------------------------------------------------
# coupled multilevel pairs in fp64 (H11); the reference's fp32 version asserts
# ----------------------------------------------------------------------------------------------
def mlmc_cases():
    torch.set_default_dtype(torch.float64)
    try:
        for ex in (False, True):
            for fine, coarse in ((4, 2), (8, 2), (16, 8)):
                sde = ref.Merton(0.02, 0.2, 1, -0.05, 0.3, torch.tensor([1.]), 1)
                solver = ref.JumpEulerSolver(sde, 3, fine, exact_jumps=ex)
                bs = 48
                factor = fine // coarse
                K = coarse + solver.max_jumps          # outer iterations available
                z, _, jt, marks = jump_noise(bs, K * factor, 1, solver.max_jumps, 1.0, False, "lognormal", np.float64)
                marks = marks[:, :K]
                JumpInjector(solver, z, None, jt, marks, "lognormal")
                (pf, pc), _ = solver.multilevel_solve(bs, (fine, coarse))
                save("mlmc_merton_%d_%d_ex%d" % (fine, coarse, ex), z=z, jump_times=jt, marks=marks,
                     fine_last=pf[:, -1], coarse_last=pc[:, -1], paths_fine=pf, paths_coarse=pc,
                     max_jumps=solver.max_jumps, exact_jumps=int(ex),
                     **merton_meta(0.02, 0.2, 1.0, -0.05, 0.3, [1.0], 3.0))
        levy = ref.ExpExampleLevy(1, 1, 0.5, 2, 0.02, 0.3, 0.2, 0.05, dim=2)
        sde = ref.LevySde(levy, torch.tensor([1., 1.]))
        solver = ref.JumpEulerSolver(sde, 3, 8, exact_jumps=True)
        bs, fine, coarse = 16, 8, 4
        K = coarse + solver.max_jumps
        z, zc, jt, marks = jump_noise(bs, K * 2, 2, solver.max_jumps, float(sde.jump_rate()), True, "icdf", np.float64)
        marks = marks[:, :K]
        JumpInjector(solver, z, zc, jt, marks, "icdf")
        (pf, pc), _ = solver.multilevel_solve(bs, (fine, coarse))
        meta = levy_meta(levy, sde, [1.0, 1.0], 3.0)
        meta.update(r=0.02, sigma=0.3, f=0.2)
        save("mlmc_explevy_2d_8_4_ex1", z=z, zc=zc, jump_times=jt, marks=marks, fine_last=pf[:, -1],
             coarse_last=pc[:, -1], max_jumps=solver.max_jumps, exact_jumps=1, **meta)
    finally:
        torch.set_default_dtype(torch.float32)


# ----------------------------------------------------------------------------------------------
# payoffs, discounter, closed forms, helpers (P1-P4, E2, known-answer tests of the reference suite)
# ----------------------------------------------------------------------------------------------
def payoff_cases():
    out = {}
    for dim in (1, 2, 3, 4):
        x = (RNG.random((24, dim)) * 2.0 + 0.05).astype(np.float32)
        out["x%d" % dim] = x
        xt = torch.as_tensor(x)
        lx = torch.log(xt)
        specs = {
            "euro_call": ref.EuroCall(1.0), "euro_put": ref.EuroPut(1.0), "binary_aon": ref.BinaryAoN(1.0),
            "basket_arith": ref.Basket(1.0), "basket_geom": ref.Basket(1.0, 'geometric'), "rainbow": ref.Rainbow(1.0),
            "digital": ref.Digital(1.0), "heston_rainbow": ref.HestonRainbow(1.0), "best_of": ref.BestOf(1.0),
            "euro_call_disc": ref.EuroCall(0.9, discount=0.94),
        }
        if dim >= 2:
            specs["asian_call"] = ref.AsianCall(3.0, 0.3)
        for k, opt in specs.items():
            out["%s_%d" % (k, dim)] = opt(xt).numpy()
        logspecs = {"euro_call_log": ref.EuroCall(1.0, log=True, discount=0.94),
                    "rainbow_log": ref.Rainbow(1.0, log=True, discount=0.94)}
        if dim >= 2:
            logspecs["asian_call_log"] = ref.AsianCall(3.0, 1.0, log=True)
        for k, opt in logspecs.items():
            out["%s_%d" % (k, dim)] = opt(lx).numpy()
    save("payoffs", **out)

    cf = {
        "bs_call_1_1_3_.02_.2": ref.bs_call(1, 1, 3, 0.02, 0.2),
        "bs_call_1_1_3_.02_.3": ref.bs_call(1, 1, 3, 0.02, 0.3),
        "bs_binary_aon_1_1_3_.02_.2": ref.bs_binary_aon(1, 1, 3, 0.02, 0.2),
        "bs_digital_call_1_1_3_.02_.2": float(ref.bs_digital_call(1, 1, 3, 0.02, 0.2)),
        "bs_asian_call_1_1_3_.02_.2": ref.bs_asian_call(1, 1, 3, 0.02, 0.2),
        "merton_call_1_1_3_.02_.3_-.05_.3_2": ref.merton_call(1, 1, 3, 0.02, 0.3, -0.05, 0.3, 2),
        "merton_call_1_1_3_.02_.2_-.05_.3_1": ref.merton_call(1, 1, 3, 0.02, 0.2, -0.05, 0.3, 1),
        "remove_steps_0.1_1000_3": ref.remove_steps(0.1, 1000, 3),
        "ceil_mult_10.5_4": ref.ceil_mult(10.5, 4),
        "get_jump_comp_1_1_.5_2_.2": ref.get_jump_comp(1, 1, 0.5, 2, 0.2),
        "max_jumps_merton_rate1_T3": ref.JumpEulerSolver(
            ref.Merton(0.02, 0.2, 1, -0.05, 0.3, torch.tensor([1.]), 1), 3, 10).max_jumps,
        "max_jumps_explevy_eps001_T3": ref.JumpEulerSolver(
            ref.LevySde(ref.ExpExampleLevy(1, 1, 0.5, 2, 0.02, 0.3, 0.2, 0.001, dim=2), torch.tensor([1., 1.])),
            3, 10).max_jumps,
        "corr_matrix_.7_.2_-.3": ref.get_corr_matrix([0.7, 0.2, -0.3]).tolist(),
        "partition_3_4_right": ref.partition(3, 4).tolist(),
        "partition_3_4_left": ref.partition(3, 4, ends='left').tolist(),
        "mlmc_bs_from_trials": ref.mlmc_bs_from_trials(torch.tensor([10 ** 8, 10 ** 5, 10 ** 3]), [1, 4, 16],
                                                       dim=1, max_jumps=33).tolist(),
    }
    ic = ref.InverseCdf(1, 1, 2, 0.5, 0.01)
    u = np.linspace(1e-6, 1 - 1e-6, 97).astype(np.float32)
    cf["icdf_params"] = dict(lda=ic.lda, y1=ic.y1, y2=ic.y2, y3=ic.y3)
    with open(os.path.join(HERE, "closed_forms.json"), "w") as fh:
        json.dump(cf, fh, indent=1, default=float)
    save("icdf", u=u, x=ic(torch.as_tensor(u)))
    a = torch.tensor([1., 2., -1.])
    b = torch.tensor([0., 5., 1.])
    c = torch.tensor([-4., 2., 2.])
    save("solve_quadratic", a=a, b=b, c=c, root=ref.solve_quadratic((a, b, c)))
    print("wrote closed_forms.json")


# ----------------------------------------------------------------------------------------------
# control variates (E5-E7): per-path gamma of apply_adapted / apply_diffusion with fixed MLP weights
# ----------------------------------------------------------------------------------------------
def mlp_weights(net):
    lin = [l for l in net.net if isinstance(l, torch.nn.Linear)]
    w = {}
    for i, l in enumerate(lin):
        w["w%d" % i] = l.weight.detach().numpy().copy()
        w["b%d" % i] = l.bias.detach().numpy().copy()
    return w


def cv_cases():
    torch.manual_seed(7)
    f = ref.Mlp(2, [50, 50, 50], 1, batch_norm=False, batch_norm_init=False)
    g = ref.Mlp(2, [50, 50, 50], 1, batch_norm=False, batch_norm_init=False)
    # make the nets non-trivial: a few Adam steps of the reference's own training on reference-simulated data
    sde = ref.Merton(0.02, 0.2, 1, -0.05, 0.3, torch.tensor([1.]), 1)
    solver = ref.JumpEulerSolver(sde, 3, 20)
    csr = ref.ConstantShortRate(0.02)
    call = ref.EuroCall(1.0)
    adam = torch.optim.Adam(list(f.parameters()) + list(g.parameters()))
    dl = ref.simulate_adapted_data(2000, solver, call, csr, bs=200)
    ref.train_adapted_control_variates([f, g], adam, dl, solver, csr, 5, False)
    f.eval(), g.eval()

    # injected inference batch; batch_size 1 loaders give per-path gamma
    solver = ref.JumpEulerSolver(sde, 3, 20)
    bs = 24
    K = 20 + solver.max_jumps
    z, _, jt, marks = jump_noise(bs, K, 1, solver.max_jumps, 1.0, False, "lognormal")
    JumpInjector(solver, z, None, jt, marks, "lognormal")
    dl = ref.simulate_adapted_data(bs, solver, call, csr, bs=1, inference=True)
    gammas = []
    with torch.inference_mode():
        for i in range(bs):
            s, _ = ref.apply_adapted_control_variates([f, g], _ShapeProxy(dl.dataset, i), solver, csr)
            gammas.append(float(s))
    s_all, ss_all = ref.apply_adapted_control_variates([f, g], dl, solver, csr)
    wf = {"f_" + k: v for k, v in mlp_weights(f).items()}
    wg = {"g_" + k: v for k, v in mlp_weights(g).items()}
    save("cv_merton_1d", z=z, jump_times=jt, marks=marks, cv_gamma=np.array(gammas, np.float32),
         payoffs=dl.dataset.payoffs, total_steps=dl.dataset.total_steps, sum_gamma=float(s_all),
         sumsq_gamma=float(ss_all), max_jumps=solver.max_jumps, disc_rate=0.02,
         **merton_meta(0.02, 0.2, 1.0, -0.05, 0.3, [1.0], 3.0), **wf, **wg)

    # pure diffusion CV (varred.py:75-95) on GBM
    torch.manual_seed(11)
    fd = ref.Mlp(2, [50, 50, 50], 1, batch_norm=False, batch_norm_init=False)
    gbm = ref.Gbm(0.02, 0.3, torch.tensor([1.]), 1)
    solver = ref.EulerSolver(gbm, 3, 16)
    adam = torch.optim.Adam(fd.parameters())
    dl = ref.simulate_data(2000, solver, call, csr, bs=200)
    ref.train_diffusion_control_variate(fd, adam, dl, solver, csr, 5, False)
    fd.eval()
    solver = ref.EulerSolver(gbm, 3, 16)
    zd = RNG.standard_normal((bs, 16, 1, 1)).astype(np.float32)
    inject_diffusion(solver, zd)
    dl = ref.simulate_data(bs, solver, call, csr, bs=1, inference=True)
    gammas = []
    for i in range(bs):
        s, _ = ref.apply_diffusion_control_variate(fd, _ShapeProxy(dl.dataset, i), solver, csr)
        gammas.append(float(s))
    save("cv_gbm_1d", z=zd, cv_gamma=np.array(gammas, np.float32), payoffs=dl.dataset.payoffs, disc_rate=0.02,
         mu=[0.02], sigma=[0.3], x0=[1.0], T=3.0, **{"f_" + k: v for k, v in mlp_weights(fd).items()})


def cv_levy_cases():
    """per-path gamma of apply_adapted_control_variates (varred.py:98-131) for the 2-D 'indep' exp-Levy SDE with the
    nets of levy_rainbow_cv_experiment.py:39-40: f = Mlp(3, [50, 50, 50], 4), g = Mlp(3, [50, 50, 50], 2).  Own RNG,
    so the section can be regenerated on its own."""
    global RNG
    keep, RNG = RNG, np.random.default_rng(20261018)
    try:
        torch.manual_seed(13)
        f = ref.Mlp(3, [50, 50, 50], 4, batch_norm=False, batch_norm_init=False)
        g = ref.Mlp(3, [50, 50, 50], 2, batch_norm=False, batch_norm_init=False)
        eps = 0.05
        levy = ref.ExpExampleLevy(1, 1, 0.5, 2, 0.02, 0.3, 0.2, eps, dim=2)
        sde = ref.LevySde(levy, torch.tensor([1., 1.]))
        csr = ref.ConstantShortRate(0.02)
        opt = ref.Rainbow(1.0)
        solver = ref.JumpEulerSolver(sde, 3, 10)
        adam = torch.optim.Adam(list(f.parameters()) + list(g.parameters()))
        dl = ref.simulate_adapted_data(1000, solver, opt, csr, bs=200)
        ref.train_adapted_control_variates([f, g], adam, dl, solver, csr, 5, False)
        f.eval(), g.eval()

        solver = ref.JumpEulerSolver(sde, 3, 10)
        bs = 24
        K = 10 + solver.max_jumps
        z, zc, jt, marks = jump_noise(bs, K, 2, solver.max_jumps, float(sde.jump_rate()), True, "icdf")
        JumpInjector(solver, z, zc, jt, marks, "icdf")
        dl = ref.simulate_adapted_data(bs, solver, opt, csr, bs=1, inference=True)
        gammas = []
        with torch.inference_mode():
            for i in range(bs):
                s, _ = ref.apply_adapted_control_variates([f, g], _ShapeProxy(dl.dataset, i), solver, csr)
                gammas.append(float(s))
        s_all, ss_all = ref.apply_adapted_control_variates([f, g], dl, solver, csr)
        wf = {"f_" + k: v for k, v in mlp_weights(f).items()}
        wg = {"g_" + k: v for k, v in mlp_weights(g).items()}
        save("cv_levy_2d", z=z, zc=zc, jump_times=jt, marks=marks, cv_gamma=np.array(gammas, np.float32),
             payoffs=dl.dataset.payoffs, total_steps=dl.dataset.total_steps, sum_gamma=float(s_all),
             sumsq_gamma=float(ss_all), max_jumps=solver.max_jumps, disc_rate=0.02, eps=eps, x0=[1.0, 1.0], T=3.0,
             jump_mean=float(sde.jump_mean()), rate=float(sde.jump_rate()), **wf, **wg)
    finally:
        RNG = keep


class _ShapeProxy:
    """A one-sample 'DataLoader' over sample i of a reference dataset: iterates one batch of size 1 and exposes
    .dataset.paths / .batch_size the way varred.py:76,99 reads them."""

    def __init__(self, dataset, i):
        self.dataset = dataset
        self.batch_size = 1
        self.i = i

    def __iter__(self):
        item = self.dataset[self.i]

        def b(v):
            return v.unsqueeze(0)
        xs, y = item
        yield tuple(b(v) for v in xs), b(y)


# ----------------------------------------------------------------------------------------------
# estimator-level known answers on injected noise + reference statistics for the acceptance CIs
# ----------------------------------------------------------------------------------------------
def estimator_cases():
    gbm = ref.Gbm(0.02, 0.3, torch.tensor([1.]), 1)
    solver = ref.EulerSolver(gbm, 3, 16)
    z = RNG.standard_normal((256, 16, 1, 1)).astype(np.float32)
    inject_diffusion(solver, z)
    st = ref.mc_simple(256, solver, ref.EuroCall(1.0), ref.ConstantShortRate(0.02))
    st_cv = ref.mc_terminal_cv(256, solver, ref.EuroCall(1.0), ref.ConstantShortRate(0.02))
    save("est_gbm_1d", z=z, payoffs=st.payoffs, mean=st.sample_mean, std=st.sample_std,
         cv_mean=st_cv.sample_mean, cv_std=st_cv.sample_std, mu=[0.02], sigma=[0.3], x0=[1.0], T=3.0)

    stats = {}
    sde = ref.Merton(0.02, 0.2, 1, -0.05, 0.3, torch.tensor([1.]), 1)
    solver = ref.JumpEulerSolver(sde, 3, 100)
    for mode in ("adapted", "terminal"):
        s = ref.mc_simple(10 ** 5, solver, ref.EuroCall(1.0), ref.ConstantShortRate(0.02), bs=10 ** 5 // 4,
                          payoff_time=mode)
        stats["c1_merton_1e5x100_" + mode] = dict(mean=s.sample_mean, se=s.sample_std, n=10 ** 5)
    bsz = ref.BlackScholesEuroCall.default_params(252, 'cpu')
    s = ref.mc_simple(2 * 10 ** 5, bsz.solver, bsz.payoff, bsz.discounter, bs=10 ** 5)
    stats["c2_gbm_2e5x252"] = dict(mean=s.sample_mean, se=s.sample_std, n=2 * 10 ** 5)
    for tag, corr in (("rho0", None), ("rho04", ref.get_corr_matrix([0.4]))):
        levy = ref.ExpExampleLevy(1, 1, 0.5, 2, 0.02, 0.3, 0.2, 0.001, dim=2)
        lsde = ref.LevySde(levy, torch.tensor([1., 1.]), corr_matrix=corr)
        solver = ref.JumpEulerSolver(lsde, 3, 256)
        s = ref.mc_simple(60000, solver, ref.Rainbow(1.0), ref.ConstantShortRate(0.02), bs=20000,
                          payoff_time='adapted')
        stats["c4_levy_6e4x256_" + tag] = dict(mean=s.sample_mean, se=s.sample_std, n=60000)
    hp = ref.HestonEuroCall.default_params(100, 'cpu')
    s = ref.mc_simple(10 ** 5, hp.solver, hp.payoff, hp.discounter, bs=10 ** 5)
    stats["heston_1e5x100"] = dict(mean=s.sample_mean, se=s.sample_std, n=10 ** 5)
    # C5: fp64 MLMC pilot (level variances) and estimate at eps = 2e-3
    torch.set_default_dtype(torch.float64)
    try:
        sde = ref.Merton(0.02, 0.2, 1, -0.05, 0.3, torch.tensor([1.]), 1)
        solver = ref.JumpEulerSolver(sde, 3, 1)
        levels = [1, 2, 4, 8, 16, 32, 64, 128]
        call, csr = ref.EuroCall(1.0), ref.ConstantShortRate(0.02)
        trials = ref.get_optimal_trials(20000, levels, 2e-3, solver, call, csr)
        s = ref.mc_multilevel(trials, levels, solver, call, csr)
        stats["c5_mlmc_eps2e-3"] = dict(mean=s.sample_mean, se=s.sample_std, trials=trials, levels=levels)
        # per-level variance pilot for CI checks
        lv = []
        solver.num_steps = 1
        p, _ = solver.solve(bs=40000, low_storage=True)
        lv.append(float((call(p[:, -1, :]) * csr(3)).var()))
        for i in range(len(levels) - 1):
            (pf, pc), _ = solver.multilevel_solve(40000, (levels[i + 1], levels[i]))
            d = csr(3) * (call(pf[:, -1]) - call(pc[:, -1]))
            lv.append(float(d.var()))
        stats["c5_level_vars_n40000"] = lv
    finally:
        torch.set_default_dtype(torch.float32)
    with open(os.path.join(HERE, "ref_stats.json"), "w") as fh:
        json.dump(stats, fh, indent=1, default=float)
    print("wrote ref_stats.json", json.dumps(stats, default=float)[:400])


# ----------------------------------------------------------------------------------------------
# seeded runs of the reference with ITS OWN rng (pins oracle/torch_port.py bit-for-bit: same draws, same order)
# ----------------------------------------------------------------------------------------------
def seeded_cases():
    out = {}
    solver = ref.EulerSolver(ref.Gbm(0.02, 0.3, torch.tensor([1.]), 1), 3, 16, seed=5)
    p, n = solver.solve(bs=64)
    out["gbm_paths"], out["gbm_normals"] = p, n
    solver = ref.JumpEulerSolver(ref.Merton(0.02, 0.2, 1, -0.05, 0.3, torch.tensor([1.]), 1), 3, 20, seed=5)
    p, (n, tp, lp, ts, jp) = solver.solve(bs=64)
    out["merton_paths"], out["merton_total_steps"] = p, ts
    out["merton_times"] = tp[:, :ts + 1, 0]
    solver = ref.JumpEulerSolver(ref.Merton(0.02, 0.2, 1, -0.05, 0.3, torch.tensor([1.]), 1), 3, 20, seed=5)
    p, _ = solver.solve(bs=64, low_storage=True)
    out["merton_low_last"] = p[:, -1]
    levy = ref.ExpExampleLevy(1, 1, 0.5, 2, 0.02, 0.3, 0.2, 0.01, dim=2)
    solver = ref.JumpEulerSolver(ref.LevySde(levy, torch.tensor([1., 1.])), 3, 8, seed=5)
    p, _ = solver.solve(bs=32, low_storage=True)
    out["levy_low_last"] = p[:, -1]
    solver = ref.HestonSolver(ref.Heston(0.02, 0.25, 0.5, 0.3, -0.3, torch.tensor([1., 0.15])), 3, 20, seed=5)
    p, _ = solver.solve(bs=32)
    out["heston_last"] = p[:, -1]
    save("seeded", **out)


if __name__ == "__main__":
    which = [a for a in sys.argv[1:] if a != "api"] or ([] if "api" in sys.argv[1:] else
                                                          ["diffusion", "jump", "mlmc", "payoff", "cv", "estimator", "seeded"])
    torch.manual_seed(0)
    for w in which:
        {"diffusion": diffusion_cases, "jump": jump_cases, "mlmc": mlmc_cases, "payoff": payoff_cases,
         "cv": cv_cases, "cv_levy": cv_levy_cases, "estimator": estimator_cases, "seeded": seeded_cases}[w]()


def api_names():
    """public names of the reference's star-exported namespace (the drop-in must offer every one of them)"""
    names = sorted(n for n in dir(ref) if not n.startswith("_"))
    sigs = {}
    import inspect
    for n in names:
        obj = getattr(ref, n)
        if inspect.isclass(obj) and obj.__module__.startswith("sde_mc"):
            try:
                sigs[n] = list(inspect.signature(obj.__init__).parameters)
            except (TypeError, ValueError):
                pass
        elif inspect.isfunction(obj) and obj.__module__.startswith("sde_mc"):
            sigs[n] = list(inspect.signature(obj).parameters)
    with open(os.path.join(HERE, "api_names.json"), "w") as fh:
        json.dump({"names": names, "signatures": sigs}, fh, indent=1)
    print("wrote api_names.json", len(names), len(sigs))


if __name__ == "__main__" and "api" in sys.argv[1:]:
    api_names()
