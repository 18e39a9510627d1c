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
    """fp64 (sum, sumsq, ...) of one estimator, resident on the GPU until read.  `buf` may be a row of a larger
    (levels, 8) tensor so that all levels of an MLMC estimator are reduced and read together."""

    def __init__(self, device, buf=None):
        self.buf = torch.zeros(L.NUM_MOMENTS, dtype=torch.float64, device=device) if buf is None else buf

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


class DeviceRange:
    """A path range that only exists in device memory (written by plan_mc / plan_mlmc): rows of (path_lo, n_paths)
    uint64 pairs of THIS rank's share, plus the global trial counts.  torch has no uint64 arithmetic, so the bits
    travel as int64."""

    def __init__(self, device, rows=1):
        self.ranges = torch.zeros((rows, 2), dtype=torch.int64, device=device)
        self.trials = torch.zeros((rows,), dtype=torch.int64, device=device)

    def row_ptr(self, row=0):
        return C.c_void_p(self.ranges.data_ptr() + 16 * row)


def plan_mc(solver, pilot, pilot_trials, eps, multiple_of=1, max_trials=0):
    """Size the main run from the pilot's (all-reduced) moments on the device (sdemc_plan_mc; find_num_trials
    mc.py:418-427).  The main run's global path ids start at the solver's next id; its length is known to the host
    only after the final read -- call `solver._take_paths(n)` then."""
    dev = solver._compute_device()
    rank, size = world()
    with torch.cuda.device(dev):
        plan = DeviceRange(dev)
        L.check(L.load().sdemc_plan_mc(L.ptr(pilot.buf), int(pilot_trials), float(eps), int(multiple_of), int(max_trials),
                                       int(solver._next_path), rank, size, plan.row_ptr(), L.ptr(plan.trials),
                                       L.stream_ptr(dev)))
    return plan


def run_moments(solver, payoff, discounter, num_trials, index_mode, moments=None, num_steps=None, reduce=True,
                per_path=None, dev_range=None):
    """Simulate `num_trials` paths (split over the ranks of the default process group) through the fused
    step-loop + payoff + reduction kernel.  Returns the Moments holder (all-reduced unless reduce=False).
    per_path: optional dict; filled with this rank's per-path 'payoffs', 'iters' and 'terminal' device tensors
    (what each path contributed -- the hook the parity tests compare with the path-storing kernel and the oracle)."""
    dev = solver._compute_device()
    lib = solver._engine_lib()
    rank, size = world()
    if dev_range is not None:
        # the range is read by the kernel from device memory (DeviceRange row 0); num_trials only bounds the grid
        lo, off, cnt = 0, 0, int(num_trials or 0)
    else:
        num_trials = int(num_trials)
        lo = solver._take_paths(num_trials)           # every rank advances the global path counter identically
        off, cnt = shard(num_trials, rank, size)
    df = float(discounter(solver.time_interval))
    po = _spec.payoff_struct(payoff, df, index_mode)
    sde = solver._sde_struct(num_steps)
    with torch.cuda.device(dev):
        if moments is None:
            moments = Moments(dev)
        rng = L.SdemcRange(int(solver.seed), lo + off, cnt, dev_range.row_ptr() if dev_range is not None else None)
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


def queue_to_tolerance(solver, eps, init_trials, launch, multiple_of=1, max_trials=0):
    """queue pilot -> plan -> main run on the current stream without synchronising; returns (main Moments, plan)"""
    init_trials = int(init_trials)
    pilot = launch(init_trials)
    plan = plan_mc(solver, pilot, init_trials, eps, multiple_of, max_trials)
    return launch(0, dev_range=plan), plan


def run_to_tolerance(solver, payoff, discounter, eps, init_trials, index_mode, launch=None, multiple_of=1, max_trials=0):
    """Pilot of `init_trials` paths -> trial count for a 95% half-width of eps -> main run, queued as ONE submission:
    the pilot's moments are all-reduced and turned into the main run's path range on the device (plan_mc), the main
    kernels read that range when they start, and the host reads once at the end -- (moments of the main run, N).
    Replaces find_num_trials + the second estimator call of run_mc / run_cv_mc (mc.py:418-467).
    launch(num_trials, dev_range) -> Moments runs one estimator pass (default: the fused plain-MC kernel)."""
    if launch is None:
        def launch(num_trials, dev_range=None):
            return run_moments(solver, payoff, discounter, num_trials, index_mode, dev_range=dev_range)
    main, plan = queue_to_tolerance(solver, eps, init_trials, launch, multiple_of, max_trials)
    dev = solver._compute_device()
    with torch.cuda.device(dev):
        packed = torch.cat([main.buf, plan.trials.double()]).tolist()      # the one device -> host read
    mom = dict(zip(L.MOMENT_FIELDS, packed[:L.NUM_MOMENTS]))
    trials = int(packed[L.NUM_MOMENTS])
    solver._take_paths(trials)
    return mom, trials
