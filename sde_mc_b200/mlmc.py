"""Multilevel Monte Carlo (API of /root/reference/sde_mc/mlmc.py).

Each level is one launch of the fused pair kernel (`sdemc_mlmc_pair`): a thread carries the fine and the coarse
path together, sharing Brownian increments and jumps in registers, and only D(T)(P_fine - P_coarse) is reduced.
Unlike the reference -- whose `mc_multilevel` only works with jump solvers (its diffusion solver has no
`low_storage` argument) and whose fp32 jump pair asserts (solvers.py:264) -- both solver kinds work, in fp32.
"""
import time

import torch

from . import _engine as E
from . import _lib as L
from . import _spec
from .helpers import mc_estimates
from .mc import MCStatistics


def _pair_lib(solver):
    """the library with the coupled-pair entry point for this solver's model, or a clear error: JIT-built user models
    and the Asian wrapper have no pair kernels"""
    spec = _spec.spec_of(solver.sde)
    if spec.family == L.FAMILY_USER or spec.asian:
        raise L.SdemcError("MLMC pair kernels (mc_multilevel, get_optimal_trials, multilevel_solve) exist for the "
                           "built-in geometric / arithmetic / Heston models only, not for %s" % type(solver.sde).__name__)
    return L.load()


def _wants_fp64(solver):
    """The reference's jump MLMC runs under torch.set_default_dtype(torch.float64) (its fp32 pair asserts,
    solvers.py:264).  The same switch selects the fp64-state pair kernel here (sdemc_mlmc_pair_f64); the fp32 pair
    kernel -- dt clamped at 0 instead of the assert -- serves the default dtype."""
    return torch.get_default_dtype() == torch.float64 and bool(solver.has_jumps)


def _coeffs_f64(solver, payoff=None, df=1.0):
    """the model's coefficients as doubles (KernelSpec holds python floats; sdemc_sde narrows them to fp32)"""
    spec = _spec.spec_of(solver.sde)
    c = L.SdemcCoeffsF64()
    c.T = float(solver.time_interval)
    for i in range(L.MAX_DIM):
        c.x0[i], c.a[i], c.b1[i], c.b2[i], c.c[i] = spec.x0[i], spec.a[i], spec.b1[i], spec.b2[i], spec.c[i]
    for i in range(L.MAX_DIM * L.MAX_DIM):
        c.chol[i] = spec.chol[i]
    c.rate = spec.rate
    for i in range(12):
        c.mark_p[i] = spec.mark_p[i]
    c.strike, c.transform_discount, c.aux, c.df = 0.0, 1.0, 1.0, float(df)
    if payoff is not None:
        _, strike, aux = _spec.payoff_kernel_spec(payoff)()
        c.strike, c.transform_discount, c.aux = float(strike), float(payoff.discount), float(aux)
    return c


class _LevelContext:
    """what every level launch of one estimator call shares: the library, the payoff / coefficient structs and a
    template of the SDE struct (only num_steps changes from level to level).  Built once per call: at 8 GPUs a whole
    MLMC pass is ~1 ms of device time, and rebuilding these structs per level (kernel_spec, tensor -> float
    conversions) cost about as much on the host."""

    def __init__(self, solver, payoff, discounter):
        self.dev = solver._compute_device()
        self.lib = _pair_lib(solver)
        self.df = float(discounter(solver.time_interval))
        self.po = _spec.payoff_struct(payoff, self.df, L.INDEX_ADAPTED)
        self.sde0 = solver._sde_struct(1)
        self.use64 = _wants_fp64(solver)
        self.coeffs64 = _coeffs_f64(solver, payoff, self.df) if self.use64 else None
        self.seed = int(solver.seed)
        self.rank, self.size = E.world()

    def sde(self, fine):
        s = L.SdemcSde.from_buffer_copy(self.sde0)
        s.num_steps = int(fine)
        return s


def _level_moments(solver, payoff, discounter, trials, fine, coarse, buf=None, dev_range=None, reduce=True, ctx=None,
                   count_on_host=False):
    """moments of D(T) (P(fine) - P(coarse)) over `trials` coupled pairs (coarse == 0: single level), accumulated
    into `buf` (a row of the estimator's (levels, 8) tensor) or a fresh Moments.  dev_range = (DeviceRange, row): the
    kernel reads this level's path range from device memory (run_mlmc)."""
    if ctx is None:
        ctx = _LevelContext(solver, payoff, discounter)
    dev, lib, rank, size = ctx.dev, ctx.lib, ctx.rank, ctx.size
    if dev_range is not None:
        lo, off, cnt = 0, 0, int(trials or 0)
        d_range = dev_range[0].row_ptr(dev_range[1])
    else:
        trials = int(trials)
        lo = solver._take_paths(trials)
        off, cnt = E.shard(trials, rank, size)
        d_range = None
    sde, po = ctx.sde(fine), ctx.po
    with torch.cuda.device(dev):
        mom = E.Moments(dev, buf)
        # count_on_host: `cnt` is exact and only the first path id is read from device memory (graph replays)
        rng = L.SdemcRange(ctx.seed, lo + off, cnt, d_range,
                           L.RANGE_COUNT_ON_HOST if (count_on_host and d_range is not None and cnt > 0) else 0)
        ws = L.ptr(L.workspace(dev))          # one workspace per stream: the levels run on streams of their own
        if ctx.use64 and coarse > 0:
            L.check(lib.sdemc_mlmc_pair_f64(sde, ctx.coeffs64, po, int(fine), int(coarse), rng, None,
                                            L.ptr(mom.buf), None, ws, L.stream_ptr(dev)))
        else:
            L.check(lib.sdemc_mlmc_pair(sde, po, int(fine), int(coarse), rng, None, L.ptr(mom.buf), None,
                                        ws, L.stream_ptr(dev)))
        if reduce:
            mom.all_reduce()
    return mom


_level_streams = {}


class LevelMoments:
    """The fp64 moments of all levels of one MLMC estimator: ONE (levels, 8) device tensor, so the estimator costs
    one all-reduce (8 x 64 bytes) and one device -> host read, however many levels it has."""

    def __init__(self, device, n_levels):
        self.buf = torch.zeros((n_levels, L.NUM_MOMENTS), dtype=torch.float64, device=device)

    def all_reduce(self):
        if E.world()[1] > 1:
            torch.distributed.all_reduce(self.buf, op=torch.distributed.ReduceOp.SUM)
        return self

    def read(self):
        return [dict(zip(L.MOMENT_FIELDS, row)) for row in self.buf.tolist()]


class _GraphedPass:
    """One MLMC pass (zero the moments, fork, one launch per level on its own stream, join) captured ONCE as a CUDA
    graph and replayed (opt-in, see GRAPH_LEVELS: measured no faster than queueing the launches).  The kernels read their
    path ranges from device memory (sdemc_range.d_range, the run_mlmc mechanism), so a replay only needs the 16 bytes
    per level that say which global path ids this call simulates."""

    def __init__(self, solver, payoff, discounter, counts, levels, ctx):
        dev = ctx.dev
        self.buf = torch.zeros((len(levels), L.NUM_MOMENTS), dtype=torch.float64, device=dev)
        self.plan = E.DeviceRange(dev, len(levels))
        warm = torch.cuda.Stream(device=dev)           # capture must not start on the legacy default stream
        warm.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(warm):
            _queue_levels(solver, payoff, discounter, counts, levels, ctx, self.buf, self.plan)   # allocates workspaces
        torch.cuda.current_stream(dev).wait_stream(warm)
        torch.cuda.synchronize(dev)
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.buf.zero_()
            _queue_levels(solver, payoff, discounter, counts, levels, ctx, self.buf, self.plan)

    def run(self, ranges):
        """ranges: [(first global path id of this rank's share, count)] per level"""
        host = torch.tensor(ranges, dtype=torch.int64).pin_memory()
        self.plan.ranges.copy_(host, non_blocking=True)
        self.graph.replay()
        out = LevelMoments.__new__(LevelMoments)
        out.buf = self.buf.clone()                     # the static tensor is rewritten by the next replay
        return out


_graphs = {}
import os as _os
# Opt-in (SDEMC_MLMC_GRAPH=1 or mlmc.GRAPH_LEVELS = True): mc_multilevel / get_optimal_trials replay a captured pass.
# Measured on B200 (C5 pass, ms): 1 GPU 5.84 graph / 5.56 queued launches (the graph's branches overlap less than
# eight streams do), 2 GPUs 3.01 / 2.89, 4 GPUs 1.58 / 1.54, 8 GPUs 0.86 / 0.86 -- queueing from Python is not what
# bounds a pass once the per-call context is built once (_LevelContext), so the default stays with the streams.
GRAPH_LEVELS = _os.environ.get("SDEMC_MLMC_GRAPH", "0") == "1"


def _queue_levels(solver, payoff, discounter, counts, levels, ctx, buf, plan):
    """one launch per level into the rows of `buf`, each on a stream of its own forked from and joined back into the
    current stream; every kernel takes its path range from row l of `plan` (counts only bound the grids)"""
    dev = ctx.dev
    coarse = [0] + list(levels[:-1])
    pool = _level_streams.setdefault(dev, [])
    while len(pool) < len(levels):
        pool.append(torch.cuda.Stream(device=dev))
    cur = torch.cuda.current_stream(dev)
    fork = torch.cuda.Event()
    fork.record(cur)
    for i, (side, n, f, c) in enumerate(zip(pool, counts, levels, coarse)):
        side.wait_event(fork)
        with torch.cuda.stream(side):
            _level_moments(solver, payoff, discounter, n, f, c, buf[i], (plan, i), reduce=False, ctx=ctx,
                           count_on_host=True)
        join = torch.cuda.Event()
        join.record(side)
        cur.wait_event(join)


def _all_levels(solver, payoff, discounter, trials, levels, plan=None):
    """Queue one launch per level into the rows of one LevelMoments and all-reduce it ONCE.  The levels are
    independent, so each goes to a stream of its own (forked from and joined back into the current stream): level 0
    fills the GPU first, the small fine levels -- a wave or two of long serial paths each -- then overlap instead of
    running their tails one after the other.  SDEMC_MLMC_STREAMS=0 keeps everything on the current stream.
    plan: an E.DeviceRange with one row per level (run_mlmc) -- the kernels read their ranges from device memory.
    Host-known trial counts (mc_multilevel, get_optimal_trials) replay a captured CUDA graph of the same launches."""
    import os
    trials = [trials] * len(levels) if not isinstance(trials, (list, tuple)) else trials
    dev = solver._compute_device()
    coarse = [0] + list(levels[:-1])
    ctx = _LevelContext(solver, payoff, discounter)
    streams = os.environ.get("SDEMC_MLMC_STREAMS", "1") != "0" and len(levels) >= 2
    with torch.cuda.device(dev):
        if plan is None and streams and GRAPH_LEVELS:
            ranges, counts = [], []
            for n in trials:
                lo = solver._take_paths(int(n))
                off, cnt = E.shard(int(n), ctx.rank, ctx.size)
                ranges.append((lo + off, cnt))
                counts.append(cnt)
            key = (dev, ctx.seed, ctx.rank, ctx.size, ctx.use64, bytes(ctx.sde0), bytes(ctx.po),
                   bytes(ctx.coeffs64) if ctx.use64 else b"", tuple(int(l) for l in levels), tuple(counts))
            g = _graphs.get(key)
            if g is None:
                if len(_graphs) >= 8:
                    _graphs.pop(next(iter(_graphs)))
                g = _graphs[key] = _GraphedPass(solver, payoff, discounter, counts, levels, ctx)
            return g.run(ranges).all_reduce()
        out = LevelMoments(dev, len(levels))
        rows = [out.buf[i] for i in range(len(levels))]
        dr = [(plan, i) if plan is not None else None for i in range(len(levels))]
        if not streams:
            for n, f, c, row, d in zip(trials, levels, coarse, rows, dr):
                _level_moments(solver, payoff, discounter, n, f, c, row, d, reduce=False, ctx=ctx)
            return out.all_reduce()
        pool = _level_streams.setdefault(dev, [])
        while len(pool) < len(levels):
            pool.append(torch.cuda.Stream(device=dev))
        cur = torch.cuda.current_stream(dev)
        fork = torch.cuda.Event()
        fork.record(cur)
        for side, n, f, c, row, d in zip(pool, trials, levels, coarse, rows, dr):
            side.wait_event(fork)
            with torch.cuda.stream(side):
                _level_moments(solver, payoff, discounter, n, f, c, row, d, reduce=False, ctx=ctx)
            join = torch.cuda.Event()
            join.record(side)
            cur.wait_event(join)
        out.buf.record_stream(cur)
        return out.all_reduce()


def _combine(levels_read, trials):
    total_mean, total_var = 0.0, 0.0
    for n, m in zip(trials, levels_read):
        mean, var = mc_estimates(m['sum'], m['sumsq'], int(n))
        total_mean += mean
        total_var += var / int(n)
    return total_mean, total_var ** 0.5


def mc_multilevel(trials, levels, solver, payoff, discounter, bs=None):
    """MLMC estimate  sum_l E[P_l - P_{l-1}]  with trials[l] coupled pairs on level l (mlmc.py:7-74).
    `bs` is accepted for compatibility; nothing is stored so no batching is needed.  One launch per level, ONE
    all-reduce and ONE host read for the whole estimator."""
    start = time.time()
    pending = _all_levels(solver, payoff, discounter, [int(n) for n in trials], levels)
    mean, stderr = _combine(pending.read(), trials)
    return MCStatistics(mean, stderr, time.time() - start, trials[-1])


def get_optimal_trials(trials, levels, epsilon, solver, payoff, discounter):
    """Pilot of `trials` pairs per level -> N_l = ceil(1.96^2/eps^2 sqrt(V_l h_l) sum_k sqrt(V_k / h_k))
    (mlmc.py:77-97; eps is a 95% half-width)."""
    pending = _all_levels(solver, payoff, discounter, [int(trials)] * len(levels), levels)
    variances = [mc_estimates(m['sum'], m['sumsq'], int(trials))[1] for m in pending.read()]
    variances = torch.tensor(variances, dtype=torch.float64)
    step_sizes = solver.time_interval / torch.tensor(levels, dtype=torch.float64)
    solver.num_steps = levels[0]                  # the reference leaves the solver on the coarsest level (:83)
    total = (variances / step_sizes).sqrt().sum()
    optimal = (1.96 ** 2 / (epsilon * epsilon)) * (variances * step_sizes).sqrt() * total
    return optimal.ceil().long().tolist()


def run_mlmc(levels, epsilon, solver, payoff, discounter, pilot_trials=10 ** 5, max_trials=0):
    """MLMC to tolerance as ONE submission (extension; the reference offers the two halves get_optimal_trials and
    mc_multilevel, mlmc.py:7-97, with a host round trip between them): pilot of `pilot_trials` pairs per level -> one
    all-reduce -> the allocation formula on the device (sdemc_plan_mlmc) -> the main run of every level, whose
    kernels read their path ranges from device memory -> one all-reduce -> one host read of (moments, N_l).
    Returns (MCStatistics, trials per level)."""
    start = time.time()
    dev = solver._compute_device()
    rank, size = E.world()
    pilot_trials = int(pilot_trials)
    pilot = _all_levels(solver, payoff, discounter, [pilot_trials] * len(levels), levels)
    with torch.cuda.device(dev):
        plan = E.DeviceRange(dev, len(levels))
        d_levels = torch.tensor([int(l) for l in levels], dtype=torch.int32, device=dev)
        L.check(L.load().sdemc_plan_mlmc(L.ptr(pilot.buf), len(levels), L.ptr(d_levels), pilot_trials,
                                         float(solver.time_interval), float(epsilon), int(max_trials),
                                         int(solver._next_path), rank, size, L.ptr(plan.ranges), L.ptr(plan.trials),
                                         L.stream_ptr(dev)))
        main = _all_levels(solver, payoff, discounter, [0] * len(levels), levels, plan=plan)
        packed = torch.cat([main.buf.reshape(-1), plan.trials.double()]).tolist()        # the one host read
    n_lv = len(levels)
    rows = [dict(zip(L.MOMENT_FIELDS, packed[i * L.NUM_MOMENTS:(i + 1) * L.NUM_MOMENTS])) for i in range(n_lv)]
    trials = [int(v) for v in packed[n_lv * L.NUM_MOMENTS:]]
    solver._take_paths(sum(trials))
    solver.num_steps = levels[0]
    mean, stderr = _combine(rows, trials)
    return MCStatistics(mean, stderr, time.time() - start, trials[-1]), trials


def mlmc_bs_from_trials(trials, levels, max_mem=5 * 10 ** 8, dim=1, max_jumps=0):
    """Batch sizes that keep the reference's stored paths under max_mem floats (mlmc.py:100-101)."""
    return torch.minimum(max_mem / (dim * (torch.tensor(levels) + max_jumps)), trials).ceil().long()


# ---- solver.multilevel_solve backends (terminal states of coupled pairs) --------------------------------------------
def _pair_terminals(solver, bs, levels, inject):
    fine, coarse = int(levels[0]), int(levels[1])
    dev = solver._compute_device()
    lib = _pair_lib(solver)
    d = solver.sde.dim
    sde = solver._sde_struct(fine)
    use64 = _wants_fp64(solver)
    dt = torch.float64 if use64 else torch.float32
    keep = []

    def as_dev(a):
        return None if a is None else torch.as_tensor(a).to(device=dev, dtype=dt).contiguous()

    with torch.cuda.device(dev):
        out = torch.empty((bs, 2, d), device=dev, dtype=dt)
        inj = None
        if inject is not None:
            z, zc, jt, mk = (as_dev(inject.get(k)) for k in ('z', 'zc', 'jump_times', 'marks'))
            keep += [z, zc, jt, mk]
            K = int(mk.shape[1]) if mk is not None else coarse
            if use64:
                inj = L.SdemcInjectF64(K, L.ptr(z), L.ptr(zc), L.ptr(jt), L.ptr(mk))
            else:
                inj = L.SdemcInject(L.ptr(z), L.ptr(zc), L.ptr(jt), L.ptr(mk), K)
        mom = E.Moments(dev)
        rng = L.SdemcRange(int(solver.seed), solver._take_paths(bs), bs)
        if use64:
            L.check(lib.sdemc_mlmc_pair_f64(sde, _coeffs_f64(solver), None, fine, coarse, rng, inj, L.ptr(mom.buf),
                                            L.ptr(out), L.ptr(L.workspace(dev)), L.stream_ptr(dev)))
        else:
            L.check(lib.sdemc_mlmc_pair(sde, None, fine, coarse, rng, inj, L.ptr(mom.buf), L.ptr(out),
                                        L.ptr(L.workspace(dev)), L.stream_ptr(dev)))
    return solver._to_user_device(out)


def _pair_paths_jump(solver, bs, levels, inject):
    """((paths_fine, paths_coarse), None) where only the LAST index is meaningful -- the estimators read
    paths[:, -1] only (mlmc.py:64-65); shape (bs, 2, dim): index 0 = initial value, -1 = terminal state."""
    term = _pair_terminals(solver, bs, levels, inject)
    x0 = solver.sde.init_value.to(device=term.device, dtype=term.dtype).unsqueeze(0).repeat(bs, 1)
    pf = torch.stack([x0, term[:, 0]], dim=1)
    pc = torch.stack([x0, term[:, 1]], dim=1)
    return (pf, pc), None


def _pair_paths_diffusion(solver, bs, levels, inject):
    (pf, pc), _ = _pair_paths_jump(solver, bs, levels, inject)
    return (pf, pc), None
