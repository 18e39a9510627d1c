"""Shared test helpers: golden-fixture loading and the (golden case -> product objects) registry."""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import sde_mc_b200 as sm  # noqa: E402
from oracle import oracle  # noqa: E402  (test infrastructure)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def golden(name):
    return np.load(os.path.join(GOLDEN, name + ".npz"))


def golden_json(name):
    with open(os.path.join(GOLDEN, name + ".json")) as fh:
        return json.load(fh)


def t(a):
    return torch.as_tensor(np.asarray(a, dtype=np.float32))


def _merton(g, dim, corr=None):
    return sm.Merton(float(g["mu"]), float(g["sigma"]), float(g["rate"]), float(g["alpha"]), float(g["gamma"]),
                     t(g["x0"])[:dim], dim, corr)


def _explevy(g, corr=None):
    levy = sm.ExpExampleLevy(float(g["cm"]), float(g["cp"]), float(g["lalpha"]), float(g["lmu"]), float(g["r"]),
                             float(g["sigma"]), float(g["f"]), float(g["eps"]), dim=2)
    return sm.LevySde(levy, t(g["x0"]), corr_matrix=corr)


# name -> (builder(golden) -> sde, solver class, steps)
DIFFUSION_CASES = {
    "diff_gbm_1d": (lambda g: sm.Gbm(0.02, 0.3, t(g["x0"]), 1), sm.EulerSolver),
    "diff_gbm_3d_corr": (lambda g: sm.Gbm(0.02, 0.3, t(g["x0"]), 3, t(g["corr"])), sm.EulerSolver),
    "diff_gbm_2d_vec": (lambda g: sm.Gbm(t(g["mu"]), t(g["sigma"]), t(g["x0"]), 2), sm.EulerSolver),
    "diff_loggbm": (lambda g: sm.LogGbm(0.02, 0.2, t(g["x0"])), sm.EulerSolver),
    "diff_double_gbm_2d": (lambda g: sm.DoubleGbm(0.02, 0.2, 0.1, t(g["x0"]), 2, t(g["corr"])), sm.EulerSolver),
    "diff_heston": (lambda g: sm.Heston(float(g["r"]), float(g["kappa"]), float(g["theta"]), float(g["xi"]),
                                        float(g["rho"]), t(g["x0"])), sm.HestonSolver),
    "diff_asian_gbm": (lambda g: sm.AsianWrapper(sm.Gbm(0.02, 0.3, t(g["x0"])[:1], 1)), sm.EulerSolver),
}

JUMP_CASES = {
    "jump_merton_1d_ex0": lambda g: _merton(g, 1),
    "jump_merton_1d_ex1": lambda g: _merton(g, 1),
    "jump_merton_2d_corr": lambda g: _merton(g, 2, sm.get_corr_matrix([0.4])),
    "jump_asian_merton": lambda g: sm.AsianWrapper(_merton(g, 1)),
    "jump_explevy_2d_rho0": lambda g: _explevy(g),
    "jump_explevy_2d_rho04": lambda g: _explevy(g, sm.get_corr_matrix([0.4])),
    "jump_explevy_2d_eps05_ex1": lambda g: _explevy(g),
    "jump_addlevy_1d": lambda g: sm.LevySde(
        sm.ExampleLevy(1, 1, 0.5, 2, 0.02, torch.tensor([0.2]), torch.tensor([0.2]), torch.tensor([[1.]]), 0.01, 1),
        t(g["x0"])),
    "jump_levy2d": lambda g: sm.LevySde(sm.Levy2d(1.2, 0.8, 0.5, 2, 0.15, 0.02), t(g["x0"])),
}

MLMC_CASES = {
    **{"mlmc_merton_%d_%d_ex%d" % (f, c, e): (lambda g: _merton(g, 1)) for e in (0, 1)
       for f, c in ((4, 2), (8, 2), (16, 8))},
    "mlmc_explevy_2d_8_4_ex1": lambda g: _explevy(g),
}


def mlmc_levels(name):
    parts = name.split("_")
    return int(parts[-3]), int(parts[-2])


def steps_of_diffusion(g):
    return int(g["z"].shape[1])


def jump_solver(name, g, device="cpu"):
    sde = JUMP_CASES[name](g)
    steps = int(g["z"].shape[1]) - int(g["max_jumps"])
    return sm.JumpEulerSolver(sde, float(g["T"]), steps, device=device, exact_jumps=bool(int(g["exact_jumps"])))


def oracle_sde(solver, num_steps=None):
    spec = solver.sde.kernel_spec()
    if isinstance(solver, sm.MilsteinScheme):
        import dataclasses
        spec = dataclasses.replace(spec, scheme=2)
    return oracle.sde_struct(spec, solver.time_interval, solver.num_steps if num_steps is None else num_steps,
                             solver._max_jumps(), solver._exact_jumps())


def rel_err(a, b, floor=1e-3):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return float(np.max(np.abs(a - b) / np.maximum(np.abs(b), floor))) if a.size else 0.0
