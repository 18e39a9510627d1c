"""C5 MLMC pass timed three ways on one GPU: queued launches with host ranges, queued launches with device-resident
ranges (the run_mlmc mechanism), captured graph (device ranges).  python tools/mlmc_modes.py"""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import sde_mc_b200 as sm
from sde_mc_b200 import mlmc as M, _engine as E

levels = [1, 2, 4, 8, 16, 32, 64, 128]
solver = sm.JumpEulerSolver(sm.Merton(0.02, 0.2, 1, -0.05, 0.3, torch.tensor([1.]), 1), 3, 1, device="cuda", exact_jumps=True)
call, csr = sm.EuroCall(1.0), sm.ConstantShortRate(0.02)
trials = sm.get_optimal_trials(10 ** 5, levels, 1e-4, solver, call, csr)

def timed(fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n

M.GRAPH_LEVELS = False
print("host ranges, streams   : %.3f ms" % timed(lambda: M._all_levels(solver, call, csr, trials, levels)))
plan = E.DeviceRange("cuda", len(levels))
lo, rows = 10 ** 9, []
for n in trials:
    rows.append((lo, int(n))); lo += int(n)
plan.ranges.copy_(torch.tensor(rows, dtype=torch.int64))
print("device ranges, streams : %.3f ms" % timed(lambda: M._all_levels(solver, call, csr, [int(n) for n in trials], levels, plan=plan)))
print("device ranges, streams, full grids : %.3f ms" % timed(lambda: M._all_levels(solver, call, csr, [0] * len(levels), levels, plan=plan)))
M.GRAPH_LEVELS = True
print("captured graph         : %.3f ms" % timed(lambda: M._all_levels(solver, call, csr, trials, levels)))
for i, lv in enumerate(levels[:2]):
    for dr in (None, (plan, i)):
        t = timed(lambda: M._level_moments(solver, call, csr, int(trials[i]), lv, levels[i - 1] if i else 0, dev_range=dr, reduce=False))
        print("level %d alone, %s range: %.3f ms" % (lv, "device" if dr else "host", t))
