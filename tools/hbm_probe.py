"""Probe HBM bandwidth on this GPU: pure write (fill_), pure read (sum), copy."""
import torch, statistics
n = 2_000_000_000  # floats = 8 GB
a = torch.empty(n, device='cuda', dtype=torch.float32)
b = torch.empty(n, device='cuda', dtype=torch.float32)
def t(f, reps=6):
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); f(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return statistics.median(ts[1:])
ms = t(lambda: a.fill_(1.0)); print("fill_  8 GB: %.3f ms -> %.0f GB/s (write only)" % (ms, 8e9 / ms / 1e6))
ms = t(lambda: a.zero_()); print("zero_  8 GB: %.3f ms -> %.0f GB/s (memset)" % (ms, 8e9 / ms / 1e6))
ms = t(lambda: b.copy_(a)); print("copy_  8 GB: %.3f ms -> %.0f GB/s (read+write counted)" % (ms, 16e9 / ms / 1e6))
ms = t(lambda: a.sum()); print("sum    8 GB: %.3f ms -> %.0f GB/s (read only)" % (ms, 8e9 / ms / 1e6))
# strided-row write pattern like the store kernels: rows of 1024 B, each warp writes 128 B per row
