"""Time solver.solve() (path-storing kernels) with CUDA events: python tools/store_variants.py gbm|merton [paths]"""
import sys, os, statistics, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import sde_mc_b200 as sm
which = sys.argv[1]
if which == "gbm":
    solver = sm.BlackScholesEuroCall.default_params(252, 'cuda').solver
    bs = int(float(sys.argv[2])) if len(sys.argv) > 2 else 4_000_000
    nbytes = bs * (253 + 252) * 4
else:
    sde = sm.Merton(0.02, 0.2, 1, -0.05, 0.3, torch.tensor([1.]), 1)
    solver = sm.JumpEulerSolver(sde, 3, 100, device='cuda')
    bs = int(float(sys.argv[2])) if len(sys.argv) > 2 else 2_000_000
    S = 100 + solver.max_jumps
    nbytes = bs * (4 * (S + 1) + S) * 4
ts = []
for i in range(12):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    out = None
    e0.record()
    out = solver.solve(bs=bs)
    e1.record()
    torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1))
ts = ts[3:]
med = statistics.median(ts)
print("%s lib=%s paths=%d median %.3f ms min %.3f ms -> %.0f GB/s" % (which, os.path.basename(os.environ.get("SDEMC_B200_LIB", "default")), bs, med, min(ts), nbytes / med / 1e6))
