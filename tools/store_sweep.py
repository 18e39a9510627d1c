"""GBM solve() (path-storing kernel) timed with CUDA events over a sweep of step counts and path counts:
python tools/store_sweep.py [steps,steps,...] [paths,paths,...]"""
import sys, os, statistics, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import sde_mc_b200 as sm
steps_list = [int(s) for s in (sys.argv[1] if len(sys.argv) > 1 else "252,255,256,128,512").split(",")]
paths_list = [int(float(s)) for s in (sys.argv[2] if len(sys.argv) > 2 else "4e6").split(",")]
for steps in steps_list:
    for bs in paths_list:
        solver = sm.EulerSolver(sm.Gbm(0.02, 0.3, torch.tensor([1.0]), 1), 3.0, steps, device='cuda')
        nbytes = bs * (2 * steps + 1) * 4
        ts = []
        for i in range(10):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            out = None
            e0.record()
            out = solver.solve(bs=bs)
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        med = statistics.median(ts[3:])
        print("lib=%s steps=%d paths=%d median %.3f ms -> %.0f GB/s" % (os.path.basename(os.environ.get("SDEMC_B200_LIB", "default")), steps, bs, med, nbytes / med / 1e6), flush=True)
