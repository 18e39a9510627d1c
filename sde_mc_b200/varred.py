"""Neural control variates (API of /root/reference/sde_mc/varred.py).

Training is one-off autograd work on stored trajectories and stays in PyTorch (it consumes the path-storing
kernel's output).  Applying trained nets is the hot part: `mc_cv_fused` runs simulation + both MLPs + the control
variate sums in one kernel; the tensor-valued `apply_*` functions remain for nets the kernel cannot evaluate
(batch norm, recurrent nets) and as the API the reference exposes.
"""
import time

import numpy as np
import torch

from . import _engine as E
from . import _lib as L
from . import _spec
from .helpers import partition, remove_steps


class EarlyStopping:
    """Stop training once an epoch's variance reduction no longer pays for its cost (varred.py:7-20)."""

    def __init__(self, eps, quantile, cost_batch, alpha):
        self.eps = eps
        self.quantile = quantile
        self.cost_batch = cost_batch
        self.alpha = alpha
        self.cost_epoch = None
        self.batch_size = None

    def threshold(self):
        return self.alpha * (self.cost_epoch * self.eps ** 2 * self.batch_size) / (self.cost_batch * self.quantile ** 2)

    def stop(self, delta_gamma):
        return delta_gamma < self.threshold()


def integrate_cv(normals, f_out, discounts, diffusion_struct, tol=0, time_interval=None):
    """sum_k f_k . dW_k D_k per path; `tol` trims the steps closest to maturity (varred.py:202-214)."""
    if tol != 0:
        assert time_interval is not None
        keep = remove_steps(tol, normals.shape[1], time_interval)
        normals, f_out, discounts = normals[:, :keep], f_out[:, :keep], discounts[:, :keep]
    if diffusion_struct == 'diag':
        return (normals * f_out * discounts).sum(-1).sum(-1)
    return ((normals * f_out).sum(-1) * discounts).sum(-1).sum(-1)


def _net_inputs(net, times, states, batch, steps, dim):
    """(t, x) rows for a pointwise net, (batch, steps, 1+dim) sequences for a recurrent one."""
    if net.sequential:
        return torch.cat([times, states], dim=-1)
    return torch.cat([times.reshape(batch * steps, 1), states.reshape(batch * steps, dim)], dim=-1)


def _diffusion_gammas(model, batch, time_points, discounts, solver, steps, dim, bs, tol):
    (paths, normals), payoffs = batch
    if model.sequential:
        times = time_points.unsqueeze(-1).repeat(bs, 1, 1)
    else:
        times = time_points.repeat(bs).unsqueeze(-1)
    f_out = model(_net_inputs(model, times, paths, bs, steps, dim)).view(normals.shape)
    return payoffs + integrate_cv(normals, f_out, discounts, solver.sde.diffusion_struct, tol=tol,
                                  time_interval=solver.time_interval)


def _adapted_gammas(models, batch, solver, discounter, steps, dim, bs, tol):
    """gamma = payoff + sum f dW D + sum g D J - sum rate E[J] g D h, the last sum without the final interval
    (varred.py:103-128)."""
    f, g = models
    (paths, normals, left_paths, time_paths, jump_paths), payoffs = batch
    h = torch.diff(time_paths, dim=1)
    discounts = discounter(time_paths)
    f_out = f(_net_inputs(f, time_paths, paths, bs, steps, dim)).view(normals.shape)
    brownian_cv = integrate_cv(normals, f_out, discounts, solver.sde.diffusion_struct, tol=tol,
                               time_interval=solver.time_interval)
    g_out = g(_net_inputs(g, time_paths, left_paths, bs, steps, dim)).view(bs, steps, dim)
    jump_cv = (g_out * discounts * jump_paths).sum(-1).sum(-1)
    comp = (-solver.sde.jump_rate() * solver.sde.jump_mean() * g_out[:, :-1] * discounts[:, :-1] * h).sum(-1).sum(-1)
    return payoffs + brownian_cv + jump_cv + comp


def _train(nets, opt, dl, epochs, print_losses, early_stopping, gammas_of):
    trials = dl.dataset.paths.shape[0]
    losses = []
    spent = 0.0
    for epoch in range(epochs):
        t0 = time.time()
        for net in nets:
            net.train()
        total = 0.0
        for batch in dl:
            opt.zero_grad()
            loss = gammas_of(batch).var()
            total += loss.item()
            loss.backward()
            opt.step()
        losses.append(total / len(dl))
        if print_losses:
            print('{}: Train loss: {:.5f}     95% confidence interval: {:.5f}'.format(
                epoch, losses[epoch], np.sqrt(losses[epoch]) * 2 / np.sqrt(trials)))
        for net in nets:
            net.eval()
        spent += time.time() - t0
        if early_stopping is not None and epoch > 0:
            early_stopping.cost_epoch = spent / (epoch + 1)
            if early_stopping.stop(losses[epoch - 1] - losses[epoch]):
                break
    return losses


def train_diffusion_control_variate(model, opt, dl, solver, discounter, epochs, print_losses=True, tol=0,
                                    early_stopping=None):
    """Minimise the batch variance of payoff + sum f dW D (varred.py:23-72).  Returns (seconds, losses)."""
    _, steps, dim = dl.dataset.paths.shape
    time_points = partition(solver.time_interval, solver.num_steps, ends='left', device=dl.dataset.paths.device)
    discounts = discounter(time_points).view(1, len(time_points), 1)
    t0 = time.time()
    losses = _train([model], opt, dl, epochs, print_losses, early_stopping,
                    lambda b: _diffusion_gammas(model, b, time_points, discounts, solver, steps, dim, dl.batch_size, tol))
    return time.time() - t0, losses


def apply_diffusion_control_variate(model, dl, solver, discounter, tol=0):
    """(sum gamma, sum gamma^2) over a DataLoader of stored paths (varred.py:75-95)."""
    _, steps, dim = dl.dataset.paths.shape
    time_points = partition(solver.time_interval, solver.num_steps, ends='left', device=dl.dataset.paths.device)
    discounts = discounter(time_points).view(1, len(time_points), 1)
    run_sum, run_sum_sq = 0, 0
    with torch.inference_mode():
        for batch in dl:
            gammas = _diffusion_gammas(model, batch, time_points, discounts, solver, steps, dim, dl.batch_size, tol)
            run_sum += gammas.sum()
            run_sum_sq += (gammas * gammas).sum()
    return run_sum, run_sum_sq


def train_adapted_control_variates(models, opt, dl, solver, discounter, epochs=10, print_losses=True, tol=0,
                                   early_stopping=None):
    """Minimise the batch variance of the jump-adapted gamma over (f, g) (varred.py:134-199).  Returns the losses."""
    _, steps, dim = dl.dataset.paths.shape
    return _train(list(models), opt, dl, epochs, print_losses, early_stopping,
                  lambda b: _adapted_gammas(models, b, solver, discounter, steps, dim, dl.batch_size, tol))


def apply_adapted_control_variates(models, dl, solver, discounter, tol=0):
    """(sum gamma, sum gamma^2) over a DataLoader of stored jump-adapted paths (varred.py:98-131)."""
    _, steps, dim = dl.dataset.paths.shape
    run_sum, run_sum_sq = 0, 0
    with torch.inference_mode():
        for batch in dl:
            gammas = _adapted_gammas(models, batch, solver, discounter, steps, dim, dl.batch_size, tol)
            run_sum += gammas.sum()
            run_sum_sq += (gammas * gammas).sum()
    return run_sum, run_sum_sq


# ---- fused path: simulation + MLPs + CV sums in one kernel --------------------------------------------------------
FUSED_CV_ENABLED = True   # set to False to force the stored-trajectory + PyTorch path for every net


def _export_mlp(net, dev, keep, in_dim=2, out_dim=1):
    layers = net.mlp_layers() if hasattr(net, 'mlp_layers') else None
    if layers is None or len(layers) != 4:
        return None
    m = L.SdemcMlp()
    for i, (w, b) in enumerate(layers):
        wt = w.detach().to(device=dev, dtype=torch.float32).contiguous()
        bt = b.detach().to(device=dev, dtype=torch.float32).contiguous()
        keep += [wt, bt]
        m.d_w[i] = wt.data_ptr()
        m.d_b[i] = bt.data_ptr()
    m.in_dim = layers[0][0].shape[1]
    m.hidden = layers[0][0].shape[0]
    m.out_dim = layers[3][0].shape[0]
    m.n_hidden_layers = 3
    if layers[1][0].shape != (m.hidden, m.hidden) or layers[2][0].shape != (m.hidden, m.hidden) or m.hidden > 63:
        return None
    if m.in_dim != in_dim or m.out_dim != out_dim:
        return None
    return m


def fused_cv_supported(models, solver, tol=0):
    """True when `sdemc_mc_cv` can evaluate these nets: BN-free Linear/ReLU stacks with three equal hidden layers of
    width <= 63 (the architecture of the experiments) on a model shape the kernel is built for -- 1-D geometric
    'diag' SDEs (Gbm, Merton: merton_cv_experiment.py:37-38, Mlp(2, .., 1)) and the 2-D geometric 'indep' Levy SDE
    (levy_rainbow_cv_experiment.py:39-40, f = Mlp(3, .., 4), g = Mlp(3, .., 2)) -- with a constant short rate.
    tol > 0 (integrate_cv varred.py:202-209) is fused for diffusions, where the cut index depends on num_steps only;
    for the jump solver the reference cuts at an index derived from the BATCH's total_steps (the stored arrays are
    trimmed to it, mc.py:394-396): a fused kernel has no batch, so that case keeps the stored-trajectory route
    (the kernel itself takes the cut index: `mc_cv_fused(..., inject=dict(total_steps=...), tol=...)`)."""
    if tol != 0 and solver.has_jumps:
        return False
    try:
        spec = _spec.spec_of(solver.sde)
    except L.SdemcError:
        return False
    if spec.family != L.FAMILY_GEOMETRIC or spec.asian:
        return False
    shape = (spec.dim, spec.m, spec.marks)
    if shape not in ((1, 1, L.MARKS_NONE), (1, 1, L.MARKS_LOGNORMAL), (2, 2, L.MARKS_ICDF)):
        return False
    if not FUSED_CV_ENABLED:
        return False
    nets = list(models) if isinstance(models, (list, tuple)) else [models]
    if len(nets) != (2 if solver.has_jumps else 1):
        return False
    keep = []
    outs = [spec.dim * spec.m, spec.dim]
    return all(_export_mlp(n, 'cpu', keep, spec.dim + 1, o) is not None for n, o in zip(nets, outs))


def mc_cv_fused(models, solver, trials, payoff, discounter, inject=None, gamma_out=False, dev_range=None, tol=0):
    """One launch: jump-adapted (or uniform) Euler + f/g MLPs on tensor cores + gamma + moments (sdemc_mc_cv).
    dev_range: an _engine.DeviceRange -- the kernel reads its path range from device memory (run_cv_mc).
    tol: the Brownian control-variate sum keeps the first remove_steps(tol, steps, T) steps (varred.py:202-209);
    steps = num_steps for diffusions, inject['total_steps'] for the jump solver (parity mode only)."""
    trials = int(trials)
    dev = solver._compute_device()
    lib = L.load()
    nets = list(models) if isinstance(models, (list, tuple)) else [models]
    keep = []
    with torch.cuda.device(dev):
        d, m = solver.sde.dim, solver.sde.brown_dim // solver.sde.dim
        f = _export_mlp(nets[0], dev, keep, d + 1, d * m)
        g = _export_mlp(nets[1], dev, keep, d + 1, d) if solver.has_jumps else None
        if tol != 0:
            if solver.has_jumps and not (inject and inject.get('total_steps')):
                raise L.SdemcError("tol > 0 with the jump solver needs the batch's total_steps (inject=dict(total_steps=..))")
            steps = int(inject['total_steps']) if solver.has_jumps else int(solver.num_steps)
            kept = remove_steps(tol, steps, solver.time_interval)
            if kept <= 0:
                raise L.SdemcError("tol = %r removes every step of the Brownian control variate" % (tol,))
            f.cv_steps = kept
        rank, size = E.world()
        if dev_range is not None:
            lo, off, cnt = 0, 0, trials
        else:
            lo = solver._take_paths(trials)
            off, cnt = E.shard(trials, rank, size) if inject is None else (0, trials)
        df = float(discounter(solver.time_interval))
        po = _spec.payoff_struct(payoff, df, L.INDEX_ADAPTED)
        sde = solver._sde_struct()
        mom = E.Moments(dev)
        gam = torch.empty((cnt,), device=dev, dtype=torch.float32) if gamma_out else None
        jm = float(solver.sde.jump_mean()) if solver.has_jumps else 0.0
        inj = None
        if inject is not None:
            from .solvers import _as_dev_f32
            z = _as_dev_f32(inject['z'], dev)
            K = int(z.shape[1])                  # (n, K[, dim]) unit normals of the correlated driver
            z = z.reshape(trials, K, -1)
            zc = _as_dev_f32(inject.get('zc'), dev)
            jt = _as_dev_f32(inject.get('jump_times'), dev)
            mk = _as_dev_f32(inject.get('marks'), dev)
            keep += [z, zc, jt, mk]
            inj = L.SdemcInject(L.ptr(z), L.ptr(zc), L.ptr(jt), L.ptr(mk), K, int(inject.get('total_steps', 0)))
        rng = L.SdemcRange(int(solver.seed), lo + off, cnt, dev_range.row_ptr() if dev_range is not None else None)
        L.check(lib.sdemc_mc_cv(sde, po, float(discounter.r), jm, f, g, rng, inj, L.ptr(mom.buf), L.ptr(gam),
                                L.ptr(L.workspace(dev)), L.stream_ptr(dev)))
        if inject is None:
            mom.all_reduce()
    if gamma_out:
        return mom, gam
    return mom
