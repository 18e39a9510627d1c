"""Probe: does solve() run slower when consecutive calls write to different buffers (previous result kept alive)?"""
import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import sde_mc_b200 as sm
p = sm.BlackScholesEuroCall.default_params(252, 'cuda')
solver = p.solver
bs = 4_000_000
def timed(keep):
    prev = None
    ts = []
    for i in range(8):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        if not keep:
            prev = None
        e0.record()
        out = solver.solve(bs=bs)
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
        prev = out
    return ts
print("free-before", ["%.2f" % t for t in timed(False)])
print("keep-prev  ", ["%.2f" % t for t in timed(True)])
print("free-before", ["%.2f" % t for t in timed(False)])
print(torch.cuda.memory_stats()["num_device_alloc"], torch.cuda.memory_reserved() / 1e9)
