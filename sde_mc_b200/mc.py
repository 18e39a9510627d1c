"""Monte Carlo estimators (API of /root/reference/sde_mc/mc.py).

Entry points, argument meaning and the returned MCStatistics are the reference's; the bodies dispatch to the fused
kernels.  Whenever per-path outputs are not requested (batched `mc_simple`, `mc_terminal_cv`, `mc_apply_cvs`) one
launch simulates, applies the payoff and reduces (sum, sum^2) in fp64 -- no trajectory or increment tensor is
materialised, so `bs` no longer bounds memory and is only kept for signature compatibility.
"""
import gc
import time

import numpy as np
import torch
from torch.utils.data import DataLoader

from . import _engine as E
from . import _lib as L
from . import _spec
from .helpers import ceil_mult, mc_estimates, partition, sample_cov
from .nets import AdaptedPathData, Mlp, NormalJumpsPathData, NormalPathData
from .options import ConstantShortRate, EuroCall
from .varred import (EarlyStopping, apply_adapted_control_variates, apply_diffusion_control_variate,
                     fused_cv_supported, mc_cv_fused, train_adapted_control_variates,
                     train_diffusion_control_variate)


def _scalar(v):
    return v.item() if torch.is_tensor(v) else float(v)


class MCStatistics:
    """Result record (mc.py:13-50): sample_mean, sample_std (= standard error), time_elapsed, num_trials and, for
    one-shot runs, the simulated paths / payoffs / increments."""

    def __init__(self, sample_mean, sample_std, time_elapsed, num_trials, paths=None, payoffs=None, normals=None):
        self.sample_mean = _scalar(sample_mean)
        self.sample_std = _scalar(sample_std)
        self.time_elapsed = time_elapsed
        self.num_trials = num_trials
        self.paths = paths
        self.payoffs = payoffs
        self.normals = normals

    def __str__(self):
        return 'Mean: {:.6f}  +/- {:.6f}    Time taken (s): {:.2f}    N: {:.2E}'.format(
            self.sample_mean, self.sample_std * 1.96, self.time_elapsed, self.num_trials)


def _index_mode(payoff_time):
    return L.INDEX_ADAPTED if payoff_time == 'adapted' else L.INDEX_TERMINAL


def _sync(solver):
    torch.cuda.synchronize(solver._compute_device())


def mc_simple(num_trials, sde_solver, payoff, discounter=None, bs=None, return_normals=False, payoff_time='terminal'):
    """Plain Monte Carlo of a payoff of the solution (mc.py:53-123).

    bs falsy : one shot -- paths, payoffs and increments are stored and returned in the statistics.
    bs given : moments only -- fused kernel, nothing stored (the reference loops over batches of bs paths).
    payoff_time: 'terminal' evaluates the payoff at array index num_steps, 'adapted' at the last state (for
    jump-adapted paths these differ: SURVEY.md quirk Q1)."""
    if discounter is None:
        discounter = ConstantShortRate(r=0.0)
    num_trials = int(num_trials)
    start = time.time()
    if _spec.payoff_kernel_spec(payoff) is None:
        return _mc_simple_python_payoff(num_trials, sde_solver, payoff, discounter, bs, payoff_time, start)
    if not bs:
        po = _spec.payoff_struct(payoff, float(discounter(sde_solver.time_interval)), _index_mode(payoff_time))
        out, normals, payoffs = sde_solver.solve(bs=num_trials, return_normals=return_normals, want_payoff=po)
        mean = payoffs.mean()
        stderr = payoffs.std() / np.sqrt(num_trials)
        _sync(sde_solver)
        return MCStatistics(mean, stderr, time.time() - start, num_trials, out, payoffs, normals)
    mom = E.run_moments(sde_solver, payoff, discounter, num_trials, _index_mode(payoff_time)).read()
    mean, stderr = E.mean_and_stderr(mom['sum'], mom['sumsq'], num_trials)
    return MCStatistics(mean, stderr, time.time() - start, num_trials)


def _mc_simple_python_payoff(num_trials, sde_solver, payoff, discounter, bs, payoff_time, start):
    """mc_simple for a USER-DEFINED Option subclass (any callable on (bs, dim) tensors, options.py:156-176): the fused
    kernels cannot evaluate Python code, so the payoff is applied with PyTorch on the GPU as the reference does
    (mc.py:84-93) and the sums are accumulated in fp64.
      one shot (bs falsy): trajectories from the path-storing kernel, returned in the statistics;
      batched            : nothing but the state the payoff reads is needed, so each batch runs the fused moments
                           kernel with its per-path hook (`sdemc_mc_moments(per_path=...)`: 4 dim bytes per path
                           instead of a stored trajectory), sharded over the ranks like every other estimator."""
    df = discounter(sde_solver.time_interval)
    jumps = bool(sde_solver.has_jumps)
    if not bs:
        if jumps:
            out, aux = sde_solver.solve(bs=num_trials)
            idx = aux[3] if (payoff_time == 'adapted' and aux[3] is not None) else sde_solver.num_steps
        else:
            out, aux = sde_solver.solve(bs=num_trials)
            idx = sde_solver.num_steps
        payoffs = payoff(out[:, idx]) * df
        mean, stderr = payoffs.mean(), payoffs.std() / np.sqrt(num_trials)
        _sync(sde_solver)
        return MCStatistics(mean, stderr, time.time() - start, num_trials, out, payoffs, aux)
    dev = sde_solver._compute_device()
    index_mode = _index_mode(payoff_time) if jumps else L.INDEX_TERMINAL
    stand_in = EuroCall(1.0)      # the kernel's own payoff slot; its moments are not read
    with torch.cuda.device(dev):
        totals = torch.zeros(2, dtype=torch.float64, device=dev)
        remaining = num_trials
        while remaining > 0:
            n = int(min(bs, remaining))
            remaining -= n
            pp = {}
            E.run_moments(sde_solver, stand_in, discounter, n, index_mode, reduce=False, per_path=pp)
            p64 = (payoff(sde_solver._to_user_device(pp['terminal'])) * df).double().to(dev)
            totals[0] += p64.sum()
            totals[1] += (p64 * p64).sum()
        if E.world()[1] > 1:
            E.dist.all_reduce(totals, op=E.dist.ReduceOp.SUM)
        total, total_sq = totals.tolist()
    mean, stderr = E.mean_and_stderr(total, total_sq, num_trials)
    return MCStatistics(mean, stderr, time.time() - start, num_trials)


def mc_terminal_cv(num_trials, sde_solver, payoff, discounter=None, bs=None, return_normals=False):
    """Terminal spot D(T) X_T - X_0 as control variate (mc.py:329-375).  Batched mode: beta comes from the first
    `bs` paths (as in the reference) and all five moments are accumulated by the fused kernel."""
    if discounter is None:
        discounter = ConstantShortRate(r=0.0)
    num_trials = int(num_trials)
    start = time.time()
    if not bs:
        df = discounter(sde_solver.time_interval)
        po = _spec.payoff_struct(payoff, float(df), L.INDEX_ADAPTED)
        out, normals, payoffs = sde_solver.solve(bs=num_trials, return_normals=return_normals, want_payoff=po)
        spots = out[:, -1]
        control = float(df) * spots[:, 0] - float(sde_solver.sde.init_value[0])
        beta = sample_cov(control, payoffs) / control.var()
        adjusted = payoffs - beta * control
        mean, stderr = adjusted.mean(), adjusted.std() / np.sqrt(num_trials)
        _sync(sde_solver)
        return MCStatistics(mean, stderr, time.time() - start, num_trials, out, payoffs, normals)
    first = min(int(bs), num_trials)
    if first < 2:
        raise ValueError("mc_terminal_cv estimates beta from the first batch: bs and num_trials must be at least 2")
    a = E.run_moments(sde_solver, payoff, discounter, first, L.INDEX_ADAPTED).read()
    cov = (a['sum_pc'] - a['sum'] * a['sum_c'] / first) / (first - 1)
    var_c = (a['sumsq_c'] - a['sum_c'] ** 2 / first) / (first - 1)
    beta = cov / var_c
    if num_trials > first:
        rest = E.run_moments(sde_solver, payoff, discounter, num_trials - first, L.INDEX_ADAPTED).read()
        a = {k: a[k] + rest[k] for k in a}
    total = a['sum'] - beta * a['sum_c']
    total_sq = a['sumsq'] - 2 * beta * a['sum_pc'] + beta * beta * a['sumsq_c']
    mean, stderr = E.mean_and_stderr(total, total_sq, num_trials)
    return MCStatistics(mean, stderr, time.time() - start, num_trials)


def mc_apply_cvs(models, solver, trials, payoff, discounter, sim_bs=1e5, bs=1000, tol=0):
    """Monte Carlo with already-trained control variates (mc.py:195-242).  When the nets are the BN-free MLPs of
    the experiments the whole thing -- simulation, both MLPs on tensor cores, the three CV sums, payoff and
    (sum, sum^2) -- is one fused kernel; otherwise trajectories are stored by the path-storing kernel and the nets
    are applied with PyTorch on the GPU."""
    start = time.time()
    trials = int(trials)
    if (fused_cv_supported(models, solver, tol) and isinstance(discounter, ConstantShortRate)
            and _spec.payoff_kernel_spec(payoff) is not None):
        mom = mc_cv_fused(models, solver, trials, payoff, discounter, tol=tol).read()
        mean, stderr = E.mean_and_stderr(mom['sum'], mom['sumsq'], trials)
        return MCStatistics(mean, stderr, time.time() - start, trials)
    run_sum, run_sum_sq = 0, 0
    remaining = trials
    while remaining > 0:
        batch = int(min(sim_bs, remaining))
        remaining -= batch
        if solver.has_jumps:
            dl = simulate_adapted_data(batch, solver, payoff, discounter, bs=bs, inference=True)
            s, ss = apply_adapted_control_variates(models, dl, solver, discounter, tol)
        else:
            dl = simulate_data(batch, solver, payoff, discounter, bs=bs, inference=True)
            s, ss = apply_diffusion_control_variate(models, dl, solver, discounter, tol)
        run_sum += s
        run_sum_sq += ss
    mean, var = mc_estimates(run_sum, run_sum_sq, trials)
    stderr = var.sqrt() / torch.tensor(trials).sqrt()
    return MCStatistics(mean, stderr, time.time() - start, trials)


def mc_control_variates(models, opt, solver, trials, steps, payoff, discounter, sim_bs=(1e5, 1e5), bs=(1000, 1000),
                        epochs=10, print_losses=True, tol=0, early_stopping=None):
    """Train a Brownian control variate on a coarse grid, then apply it on a fine one (mc.py:126-192)."""
    assert not solver.has_jumps
    (train_trials, test_trials), (train_steps, test_steps) = trials, steps
    (train_bs, test_bs), (train_sim_bs, test_sim_bs) = bs, sim_bs
    if early_stopping is not None:
        early_stopping.batch_size = bs[1]
    solver.num_steps = train_steps
    t0 = time.time()
    sim_train_control_variates(models, opt, solver, train_trials, payoff, discounter, train_sim_bs, train_bs, epochs,
                               print_losses, tol, early_stopping)
    train_time = time.time() - t0
    solver.num_steps = test_steps
    stats = mc_apply_cvs(models, solver, test_trials, payoff, discounter, test_sim_bs, test_bs, tol)
    stats.time_elapsed += train_time
    return stats


def mc_adaptive_cv(models, opt, solver, trials, steps, payoff, discounter, sim_bs=(1e4, 1e4), bs=(1000, 1000),
                   epochs=10, print_losses=True, pre_trained=False, tol=0, early_stopping=None):
    """Train (f, g) on jump-adapted coarse paths, then apply them on a fine grid (mc.py:245-326)."""
    (train_trials, test_trials), (train_steps, test_steps) = trials, steps
    (train_bs, test_bs), (_, test_sim_bs) = bs, sim_bs
    if early_stopping is not None:
        early_stopping.batch_size = bs[1]
    solver.num_steps = train_steps
    t0 = time.time()
    if not pre_trained:
        dl = simulate_adapted_data(train_trials, solver, payoff, discounter, bs=train_bs)
        train_adapted_control_variates(models, opt, dl, solver, discounter, epochs, print_losses, tol,
                                       early_stopping=early_stopping)
    train_time = time.time() - t0
    solver.num_steps = test_steps
    stats = mc_apply_cvs(models, solver, test_trials, payoff, discounter, test_sim_bs, test_bs, tol)
    stats.time_elapsed += train_time
    return stats


def simulate_data(trials, solver, payoff, discounter, bs=1000, inference=False):
    """Simulate and wrap (paths, increments, payoffs) in a DataLoader (mc.py:378-388).  The tensors stay on the
    solver's device; they come straight out of the path-storing kernel."""
    if inference:
        assert not trials % bs, 'Batch size should partition total trials evenly'
    stats = mc_simple(trials, solver, payoff, discounter, return_normals=True)
    normals = stats.normals[0] if isinstance(stats.normals, tuple) else stats.normals
    paths = stats.paths
    if isinstance(stats.normals, tuple):  # jump solver with the 'terminal' index, as in the reference
        normals = normals[:, :paths.shape[1] - 1]
    dset = NormalPathData(paths, stats.payoffs, normals)
    return DataLoader(dset, batch_size=int(bs), shuffle=not inference, drop_last=not inference)


def simulate_adapted_data(trials, solver, payoff, discounter, bs=1000, inference=False):
    """Jump-adapted trajectories for training / applying (f, g) (mc.py:391-398)."""
    stats = mc_simple(trials, solver, payoff, discounter, return_normals=True, payoff_time='adapted')
    if solver.has_jumps:
        normals, time_paths, left_paths, total_steps, jump_paths = stats.normals
        n = total_steps + 1
        dset = AdaptedPathData(stats.paths[:, :n], stats.payoffs, normals[:, :total_steps], left_paths[:, :n],
                               time_paths[:, :n], jump_paths[:, :n], total_steps)
    return DataLoader(dset, batch_size=int(bs), shuffle=not inference, drop_last=not inference)


def sim_train_control_variates(models, opt, solver, trials, payoff, discounter, sim_bs, bs, epochs=10,
                               print_losses=True, tol=0, early_stopping=None):
    if solver.has_jumps:
        dl = simulate_adapted_data(trials, solver, payoff, discounter, bs=bs)
        train_adapted_control_variates(models, opt, dl, solver, discounter, epochs, print_losses, tol, early_stopping)
    else:
        dl = simulate_data(trials, solver, payoff, discounter, bs=bs)
        train_diffusion_control_variate(models, opt, dl, solver, discounter, epochs, print_losses, tol, early_stopping)


def sample_batch_cost(solver, option, discounter, models, trials, bs, nn_bs):
    out = mc_apply_cvs(models, solver, trials, option, discounter, bs, nn_bs)
    return out.time_elapsed / (trials / nn_bs)


def _trials_for_tolerance(stats, eps, init_trials):
    return int(np.ceil((stats.sample_std * 1.96 / eps) ** 2 * init_trials))


def find_num_trials(problem, eps, models=None, init_trials=1e5, bs=1e5):
    """Pilot run -> number of trials for a 95% half-width of eps (mc.py:418-427)."""
    payoff_time = 'adapted' if problem.solver.has_jumps else 'terminal'
    if models is None:
        stats = mc_simple(init_trials, problem.solver, problem.payoff, problem.discounter, bs, payoff_time=payoff_time)
    else:
        stats = mc_apply_cvs(models, problem.solver, init_trials, problem.payoff, problem.discounter, bs)
    return _trials_for_tolerance(stats, eps, init_trials)


def find_num_trials_terminal_cv(problem, eps, init_trials, bs):
    stats = mc_terminal_cv(init_trials, problem.solver, problem.payoff, problem.discounter, bs)
    return _trials_for_tolerance(stats, eps, init_trials)


def run_mc(problem, eps, bs=1e5, init_trials=1e5):
    """Plain MC to tolerance (mc.py:437-440): pilot -> trial count -> main run.  With a built-in payoff the three
    stages are one submission (E.run_to_tolerance): the count is computed on the device from the pilot's moments and
    the main kernel reads its path range from device memory; the host reads (moments, N) once at the end."""
    payoff_time = 'adapted' if problem.solver.has_jumps else 'terminal'
    if _spec.payoff_kernel_spec(problem.payoff) is None:
        trials = find_num_trials(problem, eps, None, init_trials, bs)
        return mc_simple(trials, problem.solver, problem.payoff, problem.discounter, bs=bs, payoff_time=payoff_time)
    start = time.time()
    discounter = problem.discounter if problem.discounter is not None else ConstantShortRate(r=0.0)
    mom, trials = E.run_to_tolerance(problem.solver, problem.payoff, discounter, eps, init_trials,
                                     _index_mode(payoff_time))
    mean, stderr = E.mean_and_stderr(mom['sum'], mom['sumsq'], trials)
    return MCStatistics(mean, stderr, time.time() - start, trials)


def run_cv_mc(problem, models, opt, eps, train_size, step_factor=30, sim_bs=1e5, train_bs=1e3, nn_bs=1e3, epochs=10,
              early_stopping=False, print_losses=True, init_trials=1e5):
    """Train control variates on a coarse grid, size the run by a pilot, run to tolerance (mc.py:443-467)."""
    es = None
    if early_stopping:
        cost = sample_batch_cost(problem.solver, problem.payoff, problem.discounter, models, sim_bs, sim_bs, nn_bs)
        es = EarlyStopping(eps, 1.96, cost, 1)
        es.batch_size = nn_bs
    steps = problem.solver.num_steps
    problem.solver.num_steps = int(np.ceil(steps / step_factor))
    t0 = time.time()
    sim_train_control_variates(models, opt, problem.solver, train_size, problem.payoff, problem.discounter, sim_bs,
                               train_bs, epochs, print_losses, 0, es)
    train_time = time.time() - t0
    gc.collect()
    problem.solver.num_steps = steps
    if (fused_cv_supported(models, problem.solver, 0) and isinstance(problem.discounter, ConstantShortRate)
            and _spec.payoff_kernel_spec(problem.payoff) is not None):
        # pilot with the control variates -> ceil_mult(N, nn_bs) on the device -> main run: one submission
        t1 = time.time()
        mom, trials = E.run_to_tolerance(
            problem.solver, problem.payoff, problem.discounter, eps, init_trials, L.INDEX_ADAPTED, multiple_of=int(nn_bs),
            launch=lambda n, dev_range=None: mc_cv_fused(models, problem.solver, n, problem.payoff, problem.discounter,
                                                         dev_range=dev_range))
        mean, stderr = E.mean_and_stderr(mom['sum'], mom['sumsq'], trials)
        stats = MCStatistics(mean, stderr, time.time() - t1, trials)
    else:
        trials = ceil_mult(find_num_trials(problem, eps, models, init_trials, sim_bs), nn_bs)
        stats = mc_apply_cvs(models, problem.solver, trials, problem.payoff, problem.discounter, sim_bs, nn_bs)
    test_time = stats.time_elapsed
    stats.time_elapsed += train_time
    return stats, train_time, test_time


def run_mc_terminal_cv(problem, eps, bs=1e5, init_trials=1e5):
    trials = find_num_trials_terminal_cv(problem, eps, init_trials, bs)
    return mc_terminal_cv(trials, problem.solver, problem.payoff, problem.discounter, bs=bs)
