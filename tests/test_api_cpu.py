"""CPU tests of the host layer: the drop-in API surface, payoffs / closed forms / helpers against the reference's
known answers (tests/test_options.py, test_helpers.py, test_sde.py, test_levy.py, test_nets.py of the reference and
the committed goldens), the C-ABI library (loads, exports every declared symbol, struct layouts), and the
no-CPU-fallback rule."""
import ctypes
import inspect
import os
import re
import subprocess
import tempfile

import numpy as np
import pytest
import torch

from common import ROOT, golden, golden_json, rel_err, sm, t
from sde_mc_b200 import _lib as L
from sde_mc_b200 import _spec


# ---- API surface -------------------------------------------------------------------------------------------------
def test_every_public_name_of_the_reference_exists():
    api = golden_json("api_names")
    missing = [n for n in api["names"] if not hasattr(sm, n)]
    # submodule names leak through the reference's star imports as well; they must resolve too
    assert missing == [], missing


def test_signatures_match_the_reference():
    api = golden_json("api_names")
    extra_ok = {"inject", "want_payoff"}   # keyword-only extensions (deterministic parity mode)
    bad = []
    for name, ref_sig in api["signatures"].items():
        obj = getattr(sm, name)
        sig = inspect.signature(obj.__init__ if inspect.isclass(obj) else obj)
        ours = [p for p in sig.parameters if p not in extra_ok]
        theirs = list(ref_sig)
        if ours[:len(theirs)] != theirs:
            bad.append((name, ours, theirs))
    assert bad == [], bad[:5]


# ---- known answers of the reference's own test-suite ---------------------------------------------------------------
def test_closed_forms():
    cf = golden_json("closed_forms")
    assert abs(sm.bs_binary_aon(1, 1, 3, 0.02, 0.2) - 0.63548275523) < 1e-9          # tests/test_options.py:5-8
    assert abs(sm.bs_call(1, 1, 3, 0.02, 0.2) - 0.1646004265) < 1e-6                  # :11-14 (torch.isclose there)
    assert abs(sm.bs_call(1, 1, 3, 0.02, 0.2) - cf["bs_call_1_1_3_.02_.2"]) < 1e-14
    assert abs(sm.merton_call(1, 1, 3, 0.02, 0.3, -0.05, 0.3, 2) - 0.36328189504657027) < 1e-12   # :17-20
    assert abs(sm.bs_call(1, 1, 3, 0.02, 0.3) - cf["bs_call_1_1_3_.02_.3"]) < 1e-14
    assert abs(sm.bs_digital_call(1, 1, 3, 0.02, 0.2) - cf["bs_digital_call_1_1_3_.02_.2"]) < 1e-12
    assert abs(sm.bs_asian_call(1, 1, 3, 0.02, 0.2) - cf["bs_asian_call_1_1_3_.02_.2"]) < 1e-14
    assert abs(sm.merton_call(1, 1, 3, 0.02, 0.2, -0.05, 0.3, 1) - cf["merton_call_1_1_3_.02_.2_-.05_.3_1"]) < 1e-14


def test_payoffs_known_answers():
    x1 = torch.tensor([[1.], [3.], [0.5], [0.]])                                       # tests/test_options.py:23-50
    x2 = torch.tensor([[1., 2.], [0., 0.5], [1.2, 0.9]])
    assert torch.allclose(sm.EuroCall(strike=1)(x1), torch.tensor([0., 2., 0., 0.]))
    assert torch.allclose(sm.BinaryAoN(strike=1)(x1), torch.tensor([1., 3., 0., 0.]))
    assert torch.allclose(sm.Digital(1.)(x1), torch.tensor([0., 1., 0., 0.]))
    assert torch.allclose(sm.Basket(strike=1)(x2), torch.tensor([0.5, 0, 0.05]))
    assert torch.allclose(sm.Rainbow(strike=1)(x2), torch.tensor([1., 0., 0.2]))
    assert torch.allclose(sm.BestOf(strike=1)(x2), torch.tensor([2., 1., 1.2]))
    tp = torch.tensor([0., 1., 2.])
    assert torch.allclose(sm.ConstantShortRate(r=0.02)(tp), (tp * -0.02).exp())


def test_payoffs_match_reference_golden():
    g = golden("payoffs")
    for dim in (1, 2, 3, 4):
        x = t(g["x%d" % dim])
        lx = torch.log(x)
        specs = {"euro_call": sm.EuroCall(1.0), "euro_put": sm.EuroPut(1.0), "binary_aon": sm.BinaryAoN(1.0),
                 "basket_arith": sm.Basket(1.0), "basket_geom": sm.Basket(1.0, 'geometric'), "rainbow": sm.Rainbow(1.0),
                 "digital": sm.Digital(1.0), "heston_rainbow": sm.HestonRainbow(1.0), "best_of": sm.BestOf(1.0),
                 "euro_call_disc": sm.EuroCall(0.9, discount=0.94)}
        if dim >= 2:
            specs["asian_call"] = sm.AsianCall(3.0, 0.3)
        for k, opt in specs.items():
            assert rel_err(opt(x).numpy(), g["%s_%d" % (k, dim)]) < 1e-6, (k, dim)
        logs = {"euro_call_log": sm.EuroCall(1.0, log=True, discount=0.94),
                "rainbow_log": sm.Rainbow(1.0, log=True, discount=0.94)}
        if dim >= 2:
            logs["asian_call_log"] = sm.AsianCall(3.0, 1.0, log=True)
        for k, opt in logs.items():
            assert rel_err(opt(lx).numpy(), g["%s_%d" % (k, dim)]) < 2e-6, (k, dim)


def test_helpers_known_answers():
    cf = golden_json("closed_forms")
    g = golden("solve_quadratic")
    assert torch.allclose(sm.solve_quadratic((t(g["a"]), t(g["b"]), t(g["c"]))), t(g["root"]))
    x = torch.randn(50, dtype=torch.float64)
    mean, var = sm.mc_estimates(x.sum(), (x * x).sum(), 50)
    assert torch.isclose(mean, x.mean()) and torch.isclose(var, x.var())
    assert sm.remove_steps(0.1, 1000, 3) == 966 == cf["remove_steps_0.1_1000_3"]          # tests/test_helpers.py:32-34
    assert sm.ceil_mult(10.5, 4) == cf["ceil_mult_10.5_4"]
    assert torch.allclose(sm.get_corr_matrix([0.7, 0.2, -0.3]), torch.tensor(cf["corr_matrix_.7_.2_-.3"]))
    with pytest.raises(AssertionError):
        sm.get_corr_matrix([0.1, 0.2])
    with pytest.raises(RuntimeError):
        sm.get_corr_matrix([0.99, 0.99, -0.99])
    assert sm.partition(3, 4).tolist() == cf["partition_3_4_right"]
    assert sm.partition(3, 4, ends='left').tolist() == cf["partition_3_4_left"]
    assert abs(sm.get_jump_comp(1, 1, 0.5, 2, 0.2) - cf["get_jump_comp_1_1_.5_2_.2"]) < 1e-12
    got = sm.mlmc_bs_from_trials(torch.tensor([10 ** 8, 10 ** 5, 10 ** 3]), [1, 4, 16], dim=1, max_jumps=33).tolist()
    assert got == cf["mlmc_bs_from_trials"]
    a, b = torch.randn(20), torch.randn(20)
    assert torch.isclose(sm.sample_cov(a, b), torch.cov(torch.stack([a, b]))[0, 1])


def test_sde_coefficient_algebra():
    x = torch.tensor([[1., 2.], [3., 4.]])                                              # tests/test_sde.py:5-41
    gbm = sm.Gbm(0.02, 0.2, torch.tensor([1., 2.]), dim=2)
    assert torch.allclose(gbm.drift(0, x), 0.02 * x) and torch.allclose(gbm.diffusion(0, x), 0.2 * x)
    assert gbm.jump_rate() == 0 and gbm.jumps(0, x, x) is None
    lg = sm.LogGbm(0.02, 0.2, torch.tensor([1.]))
    assert torch.allclose(lg.drift(0, x[:, :1]), torch.full((2, 1), 0.02 - 0.5 * 0.04))
    dg = sm.DoubleGbm(0.02, 0.2, 0.1, torch.tensor([1., 1.]), 2)
    assert dg.diffusion(0, x).shape == (2, 2, 2) and dg.brown_dim == 4 and dg.diffusion_struct == 'indep'
    h = sm.Heston(0.02, 0.5, 0.1, 0.15, -0.5, torch.tensor([1., 2.]))
    assert torch.allclose(h.drift(0, x), torch.tensor([[0.02, 0.], [0.06, 0.]]))
    assert torch.allclose(h.diffusion(0, x)[:, 0], x[:, 1].sqrt() * x[:, 0])
    with pytest.raises(AssertionError):
        sm.Heston(0.02, 0.1, 0.1, 0.5, 0., torch.tensor([1., 1.]))
    m = sm.Merton(0.02, 0.3, 2, -0.05, 0.3, torch.tensor([1., 1.]), dim=2)
    jm = np.exp(-0.05 + 0.045) - 1
    assert abs(m.jump_mean() - jm) < 1e-15 and float(m.jump_rate()) == 2
    assert torch.allclose(m.drift(0, x), (0.02 - 2 * jm) * x) and torch.allclose(m.jumps(0, x, x), x * x)
    assert m.sample_jumps([5, 1], 'cpu').shape == (5, 1)
    aw = sm.AsianWrapper(sm.Gbm(0.02, 0.2, torch.tensor([1.]), 1))
    assert aw.dim == 2 and torch.allclose(aw.drift(0, x)[:, 1], x[:, 0])
    assert list(sm.UniformGrid(0., 3., 10))[-1] == pytest.approx(2.7)                    # tests/test_sde.py:44-46


def test_kernel_specs_of_builtin_models():
    m = sm.Merton(0.02, 0.2, 1, -0.05, 0.3, torch.tensor([1.]), 1).kernel_spec()
    assert (m.family, m.marks, m.dim, m.m) == (L.FAMILY_GEOMETRIC, L.MARKS_LOGNORMAL, 1, 1)
    assert abs(m.a[0] - (0.02 - (np.exp(-0.05 + 0.045) - 1))) < 1e-15 and m.c[0] == 1.0
    levy = sm.ExpExampleLevy(1, 1, 0.5, 2, 0.02, 0.3, 0.2, 0.001, dim=2)
    sde = sm.LevySde(levy, torch.tensor([1., 1.]), corr_matrix=sm.get_corr_matrix([0.4]))
    s = sde.kernel_spec()
    assert (s.family, s.marks, s.dim, s.m) == (L.FAMILY_GEOMETRIC, L.MARKS_ICDF, 2, 2)
    assert abs(s.rate - 123.49) < 0.01 and abs(s.b2[0] - 0.2 * levy.beta()) < 1e-15
    assert abs(s.chol[4] - 0.4) < 1e-7 and abs(s.chol[5] - np.sqrt(1 - 0.16)) < 1e-7
    g = golden("jump_levy2d")
    l2 = sm.LevySde(sm.Levy2d(1.2, 0.8, 0.5, 2, 0.15, 0.02), torch.tensor([0., 0.])).kernel_spec()
    assert l2.family == L.FAMILY_ARITHMETIC and np.allclose(l2.a[:2], g["drift"], rtol=1e-6)
    g = golden("jump_addlevy_1d")
    ex = sm.LevySde(sm.ExampleLevy(1, 1, 0.5, 2, 0.02, torch.tensor([0.2]), torch.tensor([0.2]),
                                   torch.tensor([[1.]]), 0.01, 1), torch.tensor([0.])).kernel_spec()
    assert np.allclose(ex.a[:1], g["drift"], rtol=1e-6)

    class Custom(sm.DiffusionSde):
        def drift(self, t_, x):
            return x

        def diffusion(self, t_, x):
            return x

    solver = sm.EulerSolver(Custom(torch.tensor([1.]), 1, 1, 'diag'), 1, 4)
    with pytest.raises(L.SdemcError):           # user-defined SDEs have no kernel and there is no CPU fallback
        solver._sde_struct()


def test_levy_classes():
    g = golden("icdf")
    ic = sm.InverseCdf(1, 1, 2, 0.5, 0.01)                                               # tests/test_levy.py:8-14
    assert rel_err(ic(t(g["u"])).numpy(), g["x"], 1e-2) < 1e-6
    lv = sm.ExpExampleLevy(1, 1, 0.5, 2, 0.02, 0.2, 0.1, 0.01)
    x = torch.tensor([[1.], [2.]])
    assert torch.allclose(lv.drift(0, x), 0.02 * x) and torch.allclose(lv.jumps(0, x, 2.), 0.1 * x * 2.)
    assert lv.gamma() == 0 and lv.beta() > 0
    assert sm.UNIFORM_TOL == 5.960464477539063e-08


def test_nets():
    torch.manual_seed(0)
    net = sm.Mlp(2, [5, 5], 1)                                                           # tests/test_nets.py
    net.eval()
    assert net(torch.randn(7, 2)).shape == (7, 1) and net.mlp_layers() is None           # has BatchNorm
    plain = sm.Mlp(2, [50, 50, 50], 1, batch_norm=False, batch_norm_init=False)
    layers = plain.mlp_layers()
    assert [tuple(w.shape) for w, _ in layers] == [(50, 2), (50, 50), (50, 50), (1, 50)]
    assert sm.ZeroFunction(3)(torch.randn(4, 2)).shape == (4, 3)
    assert sm.Lstm(2, 4, 1)(torch.randn(3, 6, 2)).shape == (3, 6, 1)
    assert sm.Gru(2, 4, 1)(torch.randn(3, 6, 2)).shape == (3, 6, 1)
    st = sm.MCStatistics(torch.tensor(1.0), torch.tensor(0.5), 2.0, 100)
    assert str(st).startswith("Mean: 1.000000  +/- 0.980000")


def test_solver_construction_and_max_jumps():
    cf = golden_json("closed_forms")
    s1 = sm.JumpEulerSolver(sm.Merton(0.02, 0.2, 1, -0.05, 0.3, torch.tensor([1.]), 1), 3, 10)
    assert s1.max_jumps == cf["max_jumps_merton_rate1_T3"] == 33 and bool(s1.has_jumps)
    lv = sm.LevySde(sm.ExpExampleLevy(1, 1, 0.5, 2, 0.02, 0.3, 0.2, 0.001, dim=2), torch.tensor([1., 1.]))
    assert sm.JumpEulerSolver(lv, 3, 10).max_jumps == cf["max_jumps_explevy_eps001_T3"] == 588
    e = sm.EulerSolver(sm.Gbm(0.02, 0.2, torch.tensor([1., 2.]), 2, sm.get_corr_matrix([0.5])), 3, 10)
    assert not bool(e.has_jumps) and e.lower_cholesky.shape == (2, 2)
    assert sm.HestonEuroCall.default_params(100, 'cpu').solver.sde.simulation_method == 'heston'
    assert isinstance(sm.LevyCallOnMax.default_params(3, 10, 'cpu'), str)
    p = sm.LevyCallOnMax.default_params(2, 10, 'cpu')
    assert p.payoff.log and p.solver.sde.kernel_spec().family == L.FAMILY_ARITHMETIC


# ---- the C-ABI library ------------------------------------------------------------------------------------------------
def test_library_exports_every_declared_symbol():
    header = open(os.path.join(ROOT, "include", "sdemc_b200.h")).read()
    declared = set(re.findall(r"\b(sdemc_[a-z0-9_]+)\s*\(", header))
    assert declared == set(L.EXPORTED_SYMBOLS), declared ^ set(L.EXPORTED_SYMBOLS)
    lib = L.load()
    for name in declared:
        assert getattr(lib, name) is not None
    assert lib.sdemc_version() == L.ABI_VERSION == 4
    assert lib.sdemc_workspace_bytes() >= 64 + 8 * 8
    assert b"bad argument" in lib.sdemc_strerror(-1) and lib.sdemc_strerror(0) == b"ok"


_C_NAMES = {"SdemcSde": "sdemc_sde", "SdemcPayoff": "sdemc_payoff", "SdemcRange": "sdemc_range",
            "SdemcInject": "sdemc_inject", "SdemcPathsOut": "sdemc_paths_out", "SdemcMlp": "sdemc_mlp",
            "SdemcCoeffsF64": "sdemc_coeffs_f64", "SdemcInjectF64": "sdemc_inject_f64"}


def test_struct_layouts_match_the_header():
    """A C program that includes include/sdemc_b200.h prints sizeof of every struct and offsetof of every field; both
    must equal what the ctypes mirror in _lib.py lays out, and what the built library reports (sdemc_abi_layout)."""
    classes = [getattr(L, n) for n in _C_NAMES]
    lines = ['#include <stdio.h>', '#include <stddef.h>', '#include "sdemc_b200.h"', 'int main(){']
    for cls in classes:
        c = _C_NAMES[cls.__name__]
        lines.append('printf("%s %%zu\\n", sizeof(%s));' % (c, c))
        for name, _ in cls._fields_:
            lines.append('printf("%s.%s %%zu\\n", offsetof(%s, %s));' % (c, name, c, name))
    lines.append('printf("sdemc_moments %zu\\n", sizeof(sdemc_moments));')
    lines.append('printf("version %d\\n", SDEMC_ABI_VERSION); return 0;}')
    with tempfile.TemporaryDirectory() as d:
        open(os.path.join(d, "s.c"), "w").write("\n".join(lines))
        subprocess.run(["gcc", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), "-o", os.path.join(d, "s"),
                        os.path.join(d, "s.c")], check=True)
        out = subprocess.run([os.path.join(d, "s")], capture_output=True, text=True, check=True).stdout
    theirs = dict((ln.split()[0], int(ln.split()[1])) for ln in out.splitlines())
    for cls in classes:
        c = _C_NAMES[cls.__name__]
        assert theirs[c] == ctypes.sizeof(cls), c
        for name, _ in cls._fields_:
            assert theirs["%s.%s" % (c, name)] == getattr(cls, name).offset, (c, name)
        assert cls().struct_size == ctypes.sizeof(cls)        # the constructor fills the handshake field
    assert theirs["sdemc_moments"] == 8 * L.NUM_MOMENTS
    assert theirs["version"] == L.ABI_VERSION
    lib = L.load()
    sizes = (ctypes.c_uint32 * 9)()
    assert lib.sdemc_abi_layout(sizes, 9) == 9
    assert list(sizes) == L.struct_sizes() == [theirs[c] for c in ("sdemc_sde", "sdemc_payoff", "sdemc_range",
                                                                    "sdemc_inject", "sdemc_moments",
                                                                    "sdemc_paths_out", "sdemc_mlp", "sdemc_coeffs_f64",
                                                                    "sdemc_inject_f64")]


def test_integration_md_binding_snippet_lays_out_the_same_structs():
    """INTEGRATION.md shows the ctypes classes a maintainer of the reference would paste; they must be the header's."""
    import re
    text = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    block = re.search(r"```python\n(import ctypes as C, torch\n.*?)```", text, re.S).group(1)
    ns = {}
    code = block.split("# ---- call site")[0].replace('C.CDLL("libsdemc_b200.so")', 'C.CDLL(%r)' % L.LIB_PATH)
    exec(compile(code, "INTEGRATION.md", "exec"), ns)
    ns["_check_layout"]()                                   # the snippet's own handshake against the built library
    for name in ("SdemcSde", "SdemcRange", "SdemcPathsOut"):
        assert ctypes.sizeof(ns[name]) == ctypes.sizeof(getattr(L, name)), name
        assert [f[0] for f in ns[name]._fields_] == [f[0] for f in getattr(L, name)._fields_], name


def test_bad_arguments_are_rejected_without_touching_a_gpu():
    lib = L.load()
    assert lib.sdemc_mc_moments(None, None, None, None, None, None, None) == -1
    s = L.SdemcSde()
    s.dim, s.m, s.num_steps, s.T = 9, 1, 10, 1.0
    assert lib.sdemc_solve_paths(s, None, L.SdemcRange(1, 0, 4), None, L.SdemcPathsOut(), None, None) == -1


def test_struct_size_handshake_rejects_foreign_layouts():
    """a struct laid out by another header version (wrong struct_size) is refused with BAD_ARG, never read past"""
    lib = L.load()
    good = sm._spec.sde_struct(sm.Gbm(0.02, 0.3, torch.tensor([1.0]), 1).kernel_spec(), 3.0, 10)
    po = sm._spec.payoff_struct(sm.EuroCall(1.0), 1.0, L.INDEX_ADAPTED)
    fake = ctypes.c_void_p(16)       # non-NULL placeholders: validation happens before any dereference on the device
    rng = L.SdemcRange(1, 0, 0)      # zero paths: a fully valid call returns OK without touching a GPU
    assert lib.sdemc_mc_moments(good, po, rng, None, fake, fake, None) == 0
    short = sm._spec.sde_struct(sm.Gbm(0.02, 0.3, torch.tensor([1.0]), 1).kernel_spec(), 3.0, 10)
    short.struct_size = 256          # the layout INTEGRATION.md showed in round 1
    assert lib.sdemc_mc_moments(short, po, rng, None, fake, fake, None) == -1
    bad_po = sm._spec.payoff_struct(sm.EuroCall(1.0), 1.0, L.INDEX_ADAPTED)
    bad_po.struct_size = 28
    assert lib.sdemc_mc_moments(good, bad_po, rng, None, fake, fake, None) == -1
    bad_rng = L.SdemcRange(1, 0, 0)
    bad_rng.struct_size = 24
    assert lib.sdemc_mc_moments(good, po, bad_rng, None, fake, fake, None) == -1
    out = L.SdemcPathsOut()
    out.struct_size = 88
    assert lib.sdemc_solve_paths(good, po, rng, None, out, fake, None) == -1
    # trajectories are not a per-path output of the moments kernels
    traj = L.SdemcPathsOut(d_paths=fake)
    assert lib.sdemc_mc_moments(good, po, rng, traj, fake, fake, None) == -1
    assert lib.sdemc_eval_payoff(bad_po, 1, fake, 4, fake, None) == -1
    # kernel choices are struct fields, validated like the rest
    good.queue_depth = 6
    assert lib.sdemc_mc_moments(good, po, rng, None, fake, fake, None) == -1
    good.queue_depth, good.short_path = 0, 9
    assert lib.sdemc_mc_moments(good, po, rng, None, fake, fake, None) == -1


def test_library_reads_no_environment_variables():
    """kernel selection is a function of the structs only (no hidden dispatch): no translation unit of the library
    calls getenv (the statically linked CUDA runtime does, for its own CUDA_* variables)"""
    csrc = os.path.join(ROOT, "sde_mc_b200", "csrc")
    for name in sorted(os.listdir(csrc)):
        if name.endswith((".cu", ".cuh", ".in")):
            assert "getenv" not in open(os.path.join(csrc, name)).read(), name


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU behaviour")
def test_no_cpu_fallback():
    solver = sm.EulerSolver(sm.Gbm(0.02, 0.2, torch.tensor([1.]), 1), 3, 10)
    with pytest.raises(L.SdemcError):
        solver.solve(bs=4)
    with pytest.raises(L.SdemcError):
        sm.mc_simple(100, solver, sm.EuroCall(1.0), sm.ConstantShortRate(0.02), bs=10)


def test_bench_reference_arm_prints_exactly_one_json_line():
    """bench.py contract: rank 0 prints ONE JSON line on stdout (everything else -- progress, native libraries such as
    NCCL's version banner -- goes to stderr).  The reference arm runs without a GPU: the reference's CPU algorithm
    (the unmodified package in oracle/_ref, else oracle/torch_port.py) on a bounded sample of the default workload."""
    import json
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    res = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "1",
                          "--warmup", "0"], capture_output=True, text=True, timeout=600, cwd=root)
    assert res.returncode == 0, res.stderr[-2000:]
    lines = [ln for ln in res.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, res.stdout
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "path_steps_per_sec" and d["unit"] == "path-steps/s"
    assert d["value"] > 0 and d["higher_is_better"] is True and d["gpu_launches"] == 0
    # the unmodified reference from oracle/_ref where `make -C oracle ref` installed it, else the pinned port
    have_ref = os.path.isdir(os.path.join(root, "oracle", "_ref", "sde_mc"))
    assert d["cpu_baseline"]["kind"] == ("reference" if have_ref else "port") and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    assert d["config"]["workload"] == "merton_1d_eurocall_jump_adapted_euler_1e9x100"     # the north-star config


def test_subclass_overrides_are_not_silently_replaced_by_the_parents_kernels():
    """The reference always calls the Python methods, so a subclass that overrides one changes the model.  The
    parent's kernel_spec() must then not be used: payoffs fall back to Python-on-stored-paths, SDEs need their own
    kernel_code()."""
    class Capped(sm.EuroCall):
        def payoff(self, x):
            return torch.clamp(super().payoff(x), max=0.5)

    class Shifted(sm.EuroCall):                       # overrides nothing the loop calls: still the built-in
        def describe(self):
            return "call"

    assert _spec.payoff_kernel_spec(sm.EuroCall(1.0)) is not None
    assert _spec.payoff_kernel_spec(Shifted(1.0)) is not None
    assert _spec.payoff_kernel_spec(Capped(1.0)) is None
    with pytest.raises(L.SdemcError):
        _spec.payoff_struct(Capped(1.0), 1.0, L.INDEX_ADAPTED)

    class Cev(sm.Gbm):
        def diffusion(self, t, x):
            return self.sigma * torch.sqrt(torch.clamp(x, min=0))

    class CevWithCode(Cev):
        def kernel_code(self):
            return dict(drift=["p[0] * x[0]"], diffusion=["p[1] * sqrtf(fmaxf(x[0], 0.f))"], params=[0.02, 0.3])

    class CodeThenOverride(CevWithCode):              # kernel_code describes CevWithCode, not this class
        def drift(self, t, x):
            return 0 * x

    x0 = torch.tensor([1.0])
    assert _spec.spec_of(sm.Gbm(0.02, 0.3, x0, 1)).family == L.FAMILY_GEOMETRIC
    with pytest.raises(L.SdemcError) as e:
        _spec.spec_of(Cev(0.02, 0.3, x0, 1))
    assert "diffusion" in str(e.value) and "Cev" in str(e.value)
    assert _spec.spec_of(CevWithCode(0.02, 0.3, x0, 1)).family == L.FAMILY_USER
    with pytest.raises(L.SdemcError):
        _spec.spec_of(CodeThenOverride(0.02, 0.3, x0, 1))
    with pytest.raises(L.SdemcError):                 # wrappers look through to the wrapped model
        _spec.spec_of(sm.AsianWrapper(Cev(0.02, 0.3, x0, 1)))
