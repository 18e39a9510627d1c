"""Estimator-level plumbing shared by mc.py / mlmc.py / varred.py: launch the fused moments kernels on this rank's
share of the paths and combine the fp64 moments across ranks.

Multi-GPU: paths are i.i.d., so rank g of G simulates the global path ids
[lo + g*N/G, lo + (g+1)*N/G) -- disjoint Philox streams by construction -- and the only exchange is ONE
all-reduce of the 8 fp64 moments (64 bytes) over NCCL.  Single process = no collective at all.
"""
import ctypes as C

import torch
import torch.distributed as dist

from . import _lib as L
from . import _spec


def world():
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def shard(n_total, rank, size):
    """contiguous, balanced split of n_total paths: returns (offset, count) of this rank."""
    base, rem = divmod(int(n_total), size)
    count = base + (1 if rank < rem else 0)
    offset = rank * base + min(rank, rem)
    return offset, count


class Moments:
    """fp64 (sum, sumsq, ...) of one estimator, resident on the GPU until read."""

    def __init__(self, device):
        self.buf = torch.zeros(L.NUM_MOMENTS, dtype=torch.float64, device=device)

    def all_reduce(self):
        if world()[1] > 1:
            dist.all_reduce(self.buf, op=dist.ReduceOp.SUM)
        return self

    def read(self):
        """device -> host read of the 64-byte result (synchronises the stream)."""
        vals = self.buf.tolist()
        return dict(zip(L.MOMENT_FIELDS, vals))


def mean_and_stderr(total, total_sq, n):
    """(sample mean, standard error) from fp64 sums; unbiased variance like mc.py:119-120."""
    mean = total / n
    var = max(total_sq / n - mean * mean, 0.0) * (n / (n - 1)) if n > 1 else float('nan')
    return mean, (var / n) ** 0.5


def run_moments(solver, payoff, discounter, num_trials, index_mode, moments=None, num_steps=None, reduce=True,
                per_path=None):
    """Simulate `num_trials` paths (split over the ranks of the default process group) through the fused
    step-loop + payoff + reduction kernel.  Returns the Moments holder (all-reduced unless reduce=False).
    per_path: optional dict; filled with this rank's per-path 'payoffs', 'iters' and 'terminal' device tensors
    (what each path contributed -- the hook the parity tests compare with the path-storing kernel and the oracle)."""
    num_trials = int(num_trials)
    dev = solver._compute_device()
    lib = solver._engine_lib()
    rank, size = world()
    lo = solver._take_paths(num_trials)           # every rank advances the global path counter identically
    off, cnt = shard(num_trials, rank, size)
    df = float(discounter(solver.time_interval))
    po = _spec.payoff_struct(payoff, df, index_mode)
    sde = solver._sde_struct(num_steps)
    with torch.cuda.device(dev):
        if moments is None:
            moments = Moments(dev)
        rng = L.SdemcRange(int(solver.seed), lo + off, cnt)
        pp = None
        if per_path is not None:
            per_path['payoffs'] = torch.empty((cnt,), device=dev, dtype=torch.float32)
            per_path['iters'] = torch.empty((cnt,), device=dev, dtype=torch.int32)
            per_path['terminal'] = torch.empty((cnt, solver.sde.dim), device=dev, dtype=torch.float32)
            pp = L.SdemcPathsOut(d_payoffs=L.ptr(per_path['payoffs']), d_iters=L.ptr(per_path['iters']),
                                 d_terminal=L.ptr(per_path['terminal']))
        getattr(lib, 'check', L.check)(lib.sdemc_mc_moments(sde, po, rng, pp, L.ptr(moments.buf),
                                                            L.ptr(L.workspace(dev)), L.stream_ptr(dev)))
        if reduce:
            moments.all_reduce()
    return moments
