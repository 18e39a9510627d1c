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
    and the Heston scheme have no pair kernels (the reference's Heston solver could run multilevel_solve)"""
    spec = _spec.spec_of(solver.sde)
    if spec.family in (L.FAMILY_USER, L.FAMILY_HESTON) or spec.asian:
        raise L.SdemcError("MLMC pair kernels (mc_multilevel, get_optimal_trials, multilevel_solve) exist for the "
                           "built-in geometric / arithmetic models only, not for %s" % type(solver.sde).__name__)
    return L.load()


def _level_moments(solver, payoff, discounter, trials, fine, coarse):
    """moments of D(T) (P(fine) - P(coarse)) over `trials` coupled pairs (coarse == 0: single level)."""
    trials = int(trials)
    dev = solver._compute_device()
    lib = _pair_lib(solver)
    rank, size = E.world()
    lo = solver._take_paths(trials)
    off, cnt = E.shard(trials, rank, size)
    po = _spec.payoff_struct(payoff, float(discounter(solver.time_interval)), L.INDEX_ADAPTED)
    sde = solver._sde_struct(fine)
    with torch.cuda.device(dev):
        mom = E.Moments(dev)
        rng = L.SdemcRange(int(solver.seed), lo + off, cnt)
        L.check(lib.sdemc_mlmc_pair(sde, po, int(fine), int(coarse), 0, rng, None, L.ptr(mom.buf), None,
                                    L.ptr(L.workspace(dev)), L.stream_ptr(dev)))
        mom.all_reduce()
    return mom


_level_streams = {}


def _all_levels(solver, payoff, discounter, trials, levels):
    """Queue one launch per level and return their Moments.  The levels are independent, so each goes to a stream of
    its own (forked from and joined back into the current stream): level 0 fills the GPU first, the small fine levels
    -- a wave or two of long serial paths each -- then overlap instead of running their tails one after the other.
    SDEMC_MLMC_STREAMS=0 keeps everything on the current stream."""
    import os
    trials = [trials] * len(levels) if not isinstance(trials, (list, tuple)) else trials
    dev = solver._compute_device()
    coarse = [0] + list(levels[:-1])
    if os.environ.get("SDEMC_MLMC_STREAMS", "1") == "0" or len(levels) < 2:
        return [_level_moments(solver, payoff, discounter, n, f, c) for n, f, c in zip(trials, levels, coarse)]
    with torch.cuda.device(dev):
        pool = _level_streams.setdefault(dev, [])
        while len(pool) < len(levels):
            pool.append(torch.cuda.Stream(device=dev))
        cur = torch.cuda.current_stream(dev)
        fork = torch.cuda.Event()
        fork.record(cur)
        pending = []
        for side, n, f, c in zip(pool, trials, levels, coarse):
            side.wait_event(fork)
            with torch.cuda.stream(side):
                mom = _level_moments(solver, payoff, discounter, n, f, c)
                mom.buf.record_stream(cur)       # read (and possibly freed) on the caller's stream
            join = torch.cuda.Event()
            join.record(side)
            cur.wait_event(join)
            pending.append(mom)
    return pending


def mc_multilevel(trials, levels, solver, payoff, discounter, bs=None):
    """MLMC estimate  sum_l E[P_l - P_{l-1}]  with trials[l] coupled pairs on level l (mlmc.py:7-74).
    `bs` is accepted for compatibility; nothing is stored so no batching is needed."""
    start = time.time()
    pending = _all_levels(solver, payoff, discounter, [int(n) for n in trials], levels)
    total_mean, total_var = 0.0, 0.0
    for n, mom in zip(trials, pending):          # one host read per level, after all launches are queued
        m = mom.read()
        mean, var = mc_estimates(m['sum'], m['sumsq'], int(n))
        total_mean += mean
        total_var += var / int(n)
    return MCStatistics(total_mean, total_var ** 0.5, time.time() - start, trials[-1])


def get_optimal_trials(trials, levels, epsilon, solver, payoff, discounter):
    """Pilot of `trials` pairs per level -> N_l = ceil(1.96^2/eps^2 sqrt(V_l h_l) sum_k sqrt(V_k / h_k))
    (mlmc.py:77-97; eps is a 95% half-width)."""
    pending = _all_levels(solver, payoff, discounter, [int(trials)] * len(levels), levels)
    variances = []
    for mom in pending:
        m = mom.read()
        variances.append(mc_estimates(m['sum'], m['sumsq'], int(trials))[1])
    variances = torch.tensor(variances, dtype=torch.float64)
    step_sizes = solver.time_interval / torch.tensor(levels, dtype=torch.float64)
    solver.num_steps = levels[0]                  # the reference leaves the solver on the coarsest level (:83)
    total = (variances / step_sizes).sqrt().sum()
    optimal = (1.96 ** 2 / (epsilon * epsilon)) * (variances * step_sizes).sqrt() * total
    return optimal.ceil().long().tolist()


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
    keep = []
    with torch.cuda.device(dev):
        out = torch.empty((bs, 2, d), device=dev, dtype=torch.float32)
        inj = None
        if inject is not None:
            from .solvers import _as_dev_f32
            z = _as_dev_f32(inject['z'], dev)
            zc = _as_dev_f32(inject.get('zc'), dev)
            jt = _as_dev_f32(inject.get('jump_times'), dev)
            mk = _as_dev_f32(inject.get('marks'), dev)
            keep += [z, zc, jt, mk]
            K = int(mk.shape[1]) if mk is not None else coarse
            inj = L.SdemcInject(L.ptr(z), L.ptr(zc), L.ptr(jt), L.ptr(mk), K)
        mom = E.Moments(dev)
        rng = L.SdemcRange(int(solver.seed), solver._take_paths(bs), bs)
        L.check(lib.sdemc_mlmc_pair(sde, None, fine, coarse, 0, rng, inj, L.ptr(mom.buf), L.ptr(out),
                                    L.ptr(L.workspace(dev)), L.stream_ptr(dev)))
    return solver._to_user_device(out)


def _pair_paths_jump(solver, bs, levels, inject):
    """((paths_fine, paths_coarse), None) where only the LAST index is meaningful -- the estimators read
    paths[:, -1] only (mlmc.py:64-65); shape (bs, 2, dim): index 0 = initial value, -1 = terminal state."""
    term = _pair_terminals(solver, bs, levels, inject)
    x0 = solver.sde.init_value.to(term.device).unsqueeze(0).repeat(bs, 1)
    pf = torch.stack([x0, term[:, 0]], dim=1)
    pc = torch.stack([x0, term[:, 1]], dim=1)
    return (pf, pc), None


def _pair_paths_diffusion(solver, bs, levels, inject):
    (pf, pc), _ = _pair_paths_jump(solver, bs, levels, inject)
    return (pf, pc), None
