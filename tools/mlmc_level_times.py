"""Per-level kernel times of the C5 MLMC pass (Merton 1-D, levels 1..128, eps = 1e-4 allocation).  GPU only."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import sde_mc_b200 as sm  # noqa: E402
from sde_mc_b200 import mlmc as M  # noqa: E402

levels = [1, 2, 4, 8, 16, 32, 64, 128]
sde = sm.Merton(0.02, 0.2, 1, -0.05, 0.3, torch.tensor([1.]), 1)
solver = sm.JumpEulerSolver(sde, 3, 1, device="cuda", exact_jumps=True)
call, csr = sm.EuroCall(1.0), sm.ConstantShortRate(0.02)
trials = sm.get_optimal_trials(10 ** 5, levels, 1e-4, solver, call, csr)
for rep in range(2):
    tot = 0.0
    for i, lv in enumerate(levels):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        m = M._level_moments(solver, call, csr, trials[i], lv, levels[i - 1] if i else 0)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        tot += ms
        if rep:
            r = m.read()
            print("level %3d: %10d pairs  %.3f ms  iterations/pair %.2f" % (lv, trials[i], ms, r["iters"] / r["n"]))
    if rep:
        print("total %.3f ms" % tot)
