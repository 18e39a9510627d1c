"""TEST INFRASTRUCTURE -- batched PyTorch-CPU restatement of the reference's hot path, with the reference's own
draw order from torch's global RNG.  This is how the reference itself computes (eager ATen ops on (bs, dim)
tensors in a Python loop), so it is what `bench.py --impl reference` and the `cpu_baseline` leg time on the GPU
box's host cores ("kind": "port"), and it is pinned BIT-FOR-BIT against seeded runs of the unmodified reference
(tests/golden/seeded.npz, tests/test_torch_port.py).

Only tests/, __graft_entry__.smoke() and bench.py's CPU legs import this module; sde_mc_b200 never does.
Every function cites the reference lines (/root/reference/sde_mc/...) it restates.
"""
import math
import time

import torch
from scipy.stats import poisson


def _chol(spec):
    d = spec.dim
    L = torch.tensor(spec.chol, dtype=torch.get_default_dtype()).reshape(4, 4)[:d, :d]
    # solvers.py:33-36: a 1x1 correlation matrix becomes [[1.]]
    return L.contiguous()


class PortModel:
    """Coefficient callbacks of one built-in SDE on (bs, dim) tensors, built from a KernelSpec-like object.
    Restates the drift/diffusion/jumps methods of sde.py:201-205,251-255,369-375 and levy.py:74-83,148-155 for the
    geometric/arithmetic families, plus sample_jumps sde.py:325-326 / levy.py:85-87 with icdf levy.py:19-30."""

    def __init__(self, spec):
        self.spec = spec
        self.dim, self.m = spec.dim, spec.m
        dt = torch.get_default_dtype()
        d = spec.dim
        self.a = torch.tensor(spec.a[:d], dtype=dt)
        self.b1 = torch.tensor(spec.b1[:d], dtype=dt)
        self.b2 = torch.tensor(spec.b2[:d], dtype=dt)
        self.c = torch.tensor(spec.c[:d], dtype=dt)
        self.x0 = torch.tensor(spec.x0[:d], dtype=dt)
        self.geometric = spec.family == 0
        self.heston = spec.family == 2
        self.rate = torch.tensor(float(spec.rate), dtype=dt)

    def drift(self, x):
        return self.a * x if self.geometric else self.a * torch.ones_like(x)

    def diffusion(self, x):
        one = x if self.geometric else torch.ones_like(x)
        if self.m == 1:
            return self.b1 * one
        return torch.stack([self.b1 * one, self.b2 * one], dim=-1)

    def jumps(self, x, marks):
        return (self.c * x if self.geometric else self.c * torch.ones_like(x)) * marks

    def sample_marks(self, bs):
        s = self.spec
        if s.marks == 1:
            return (torch.randn(size=[bs, 1]) * s.mark_p[1] + s.mark_p[0]).exp() - 1
        cm, cp, mu, al, eps, lda, y1, y2, y3 = s.mark_p[:9]
        y = torch.rand([bs, 1]) + 5.960464477539063e-08 / 3
        x1 = torch.log((mu * lda * y) / cm) / mu - 1
        x2 = -(al * ((lda * y / cm) - (1 / mu)) + 1) ** (-1 / al)
        x3 = ((-al / cp) * (lda * y - cm / mu - cm * ((eps ** (-al) - 1) / al)) + eps ** (-al)) ** (-1 / al)
        x4 = 1 - (1 / mu) * torch.log(mu * lda * (1 - y) / cp)
        return torch.where(y <= y1, x1, torch.where(y < y2, x2, torch.where(y < y3, x3, x4)))


def _euler(model, x, h, dW):
    """EulerScheme.step schemes.py:5-13"""
    if model.m == 1:
        return x + model.drift(x) * h + model.diffusion(x) * dW
    return x + model.drift(x) * h + (model.diffusion(x) * dW).sum(dim=-1)


def _corr_normals(L, size, h, corr=True):
    """SdeSolver.sample_corr_normals solvers.py:51-56"""
    normals = torch.randn(size=size) * torch.sqrt(h)
    return torch.matmul(L, normals).squeeze(-1) if corr else normals.squeeze(-1)


def diffusion_solve(spec, T, num_steps, bs, store=True):
    """DiffusionSolver.solve solvers.py:68-88 (Euler; all increments drawn up front, paths stored every step)."""
    model, L = PortModel(spec), _chol(spec)
    h = torch.tensor(T / num_steps)
    x = model.x0.unsqueeze(0).repeat(bs, 1)
    paths = torch.empty(size=(bs, num_steps + 1, model.dim))
    paths[:, 0] = x
    dW_all = _corr_normals(L, (bs, num_steps, model.dim, model.m), h)
    for i in range(num_steps):
        x = _euler(model, x, h, dW_all[:, i])
        paths[:, i + 1] = x
    return paths, dW_all


def max_jumps(T, total_rate):
    """JumpDiffusionSolver.__init__ solvers.py:133"""
    return max(int(T * poisson.ppf(1 - 1 / 1e9, total_rate)), 5)


def jump_solve(spec, T, num_steps, bs, exact_jumps=False, low_storage=True):
    """JumpDiffusionSolver.solve solvers.py:164-226 with low_storage semantics (paths only) or full storage."""
    model, L = PortModel(spec), _chol(spec)
    mj = max_jumps(T, float(spec.rate))
    d = model.dim
    h = torch.tensor(T / num_steps)
    x = model.x0.unsqueeze(0).repeat(bs, 1)
    t = torch.zeros((bs, 1))
    S = num_steps + mj
    paths = torch.zeros(size=(bs, S + 1, d))
    if not low_storage:                                                   # init_storage :150-162
        left = torch.zeros_like(paths)
        jump_paths = torch.zeros_like(paths)
        time_paths = torch.zeros(size=(bs, S + 1, 1)) + T
        normals = torch.zeros(size=(bs, S, d) if model.m == 1 else (bs, S, d, model.m))
        left[:, 0] = x
        time_paths[:, 0] = t
    paths[:, 0] = x
    jump_times = torch.empty((bs, mj, 1)).exponential_(model.rate).cumsum(dim=1)   # :143-144
    jump_idx = torch.zeros_like(jump_times[:, 0, :]).long()
    rows = torch.arange(bs)
    k = 0
    while torch.any(t < T):                                               # :182
        k += 1
        tau = jump_times[rows, jump_idx.squeeze(-1), :]
        h = torch.minimum(h, torch.maximum(T - t, torch.tensor(0.)))     # :190
        dt = torch.minimum(h, tau - t)                                    # :191
        assert (tau >= t).all()
        if model.m == 1:
            dW = _corr_normals(L, x.shape + torch.Size([1]), dt.unsqueeze(-1))
        else:                                                             # :198-201
            dW = torch.stack([_corr_normals(L, x.shape + torch.Size([1]), dt.unsqueeze(-1)),
                              _corr_normals(L, [x.shape[0], 1, 1], dt.unsqueeze(-1), corr=False).repeat(1, x.shape[1])],
                             dim=-1)
        old_x = x
        x = _euler(model, x, dt, dW)
        t += dt
        if not low_storage:
            normals[:, k - 1] = dW
            left[:, k] = x
            time_paths[:, k] = t
        marks = model.sample_marks(bs).repeat(1, d)                       # :146-148, drawn every iteration
        hit = torch.isclose(tau, t, atol=1e-12)                           # :212 (default rtol = 1e-5)
        now = torch.where(hit, marks, torch.zeros_like(marks))
        x = x + model.jumps(x if exact_jumps else old_x, now)            # :214-217
        paths[:, k] = x
        if not low_storage:
            jump_paths[:, k] = now
        jump_idx = torch.where(hit, jump_idx + 1, jump_idx)
    if low_storage:
        return paths[:, :k + 1], (None, None, None, k, None)
    return paths[:, :k + 1], (normals, time_paths, left, k, jump_paths)


def payoff_call_on(kind, strike):
    """EuroCall options.py:196-199 / Rainbow options.py:269-272 (log=False, discount=1)"""
    zero = torch.tensor(0.)
    if kind == "euro_call":
        return lambda x: torch.where(x[:, 0] > strike, x[:, 0] - strike, zero)
    if kind == "rainbow":
        return lambda x: torch.where(x.max(1).values > strike, x.max(1).values - strike, zero)
    raise ValueError(kind)


def mc_simple_batched(spec, T, num_steps, num_trials, bs, payoff, rate_r, jumps, payoff_time="adapted"):
    """mc_simple batched branch mc.py:101-123: loop of solve() + payoff + running (sum, sum^2) in fp32 tensors."""
    df = torch.exp(-torch.tensor(float(T)) * rate_r)                      # ConstantShortRate options.py:334-337
    remaining = int(num_trials)
    s1, s2 = 0.0, 0.0
    start = time.time()
    while remaining:
        bs = min(bs, remaining)
        remaining -= bs
        if jumps:
            out, aux = jump_solve(spec, T, num_steps, bs, low_storage=False)   # mc_simple always stores (mc.py:110)
            idx = aux[3] if payoff_time == "adapted" else num_steps
        else:
            out, _ = diffusion_solve(spec, T, num_steps, bs)
            idx = num_steps
        pay = payoff(out[:, idx]) * df
        s1 += pay.sum()
        s2 += (pay ** 2).sum()
    mean = s1 / num_trials
    sd = torch.sqrt((s2 / num_trials - mean ** 2) * (num_trials / (num_trials - 1))) / math.sqrt(num_trials)
    return float(mean), float(sd), time.time() - start


def apply_adapted_cvs(f, g, out, aux, payoffs, rate, jump_mean, r, bs):
    """apply_adapted_control_variates varred.py:98-131 over stored jump-adapted paths, in NN batches of `bs` rows
    like the reference's DataLoader (AdaptedPathData drops the last time index, nets.py:192-200)."""
    normals, time_paths, left, total, jump_paths = aux
    n = out.shape[0]
    paths, left, times, jumps, normals = out[:, :total], left[:, :total], time_paths[:, :total], \
        jump_paths[:, :total], normals[:, :total]
    s1, s2 = 0.0, 0.0
    with torch.inference_mode():
        for lo in range(0, n, bs):
            sl = slice(lo, min(lo + bs, n))
            b, S, d = paths[sl].shape
            tt = times[sl]
            h = torch.diff(tt, dim=1)
            disc = torch.exp(-tt * r)
            f_out = f(torch.cat([tt.reshape(b * S, 1), paths[sl].reshape(b * S, d)], dim=-1)).view(normals[sl].shape)
            bcv = (normals[sl] * f_out * disc).sum(-1).sum(-1)
            g_out = g(torch.cat([tt.reshape(b * S, 1), left[sl].reshape(b * S, d)], dim=-1)).view(b, S, d)
            jcv = (g_out * disc * jumps[sl]).sum(-1).sum(-1)
            comp = (-rate * jump_mean * g_out[:, :-1] * disc[:, :-1] * h).sum(-1).sum(-1)
            gam = payoffs[sl] + bcv + jcv + comp
            s1 += gam.sum()
            s2 += (gam * gam).sum()
    return s1, s2


def mc_apply_cvs_batched(spec, T, num_steps, num_trials, nets, rate_r, payoff, nn_bs, sim_bs=10 ** 5):
    """mc_apply_cvs mc.py:195-242 for a jump SDE: simulate with full storage, apply (f, g), running sums."""
    f, g = nets
    df = torch.exp(-torch.tensor(float(T)) * rate_r)
    remaining, s1, s2 = int(num_trials), 0.0, 0.0
    start = time.time()
    while remaining > 0:
        batch = min(int(sim_bs), remaining)
        remaining -= batch
        out, aux = jump_solve(spec, T, num_steps, batch, low_storage=False)
        pay = payoff(out[:, aux[3]]) * df
        a, b = apply_adapted_cvs(f, g, out, aux, pay, float(spec.rate), float(spec.jump_mean), rate_r, nn_bs)
        s1 += a
        s2 += b
    mean = s1 / num_trials
    var = (s2 - s1 * s1 / num_trials) / (num_trials - 1)
    return float(mean), float(var.sqrt() / math.sqrt(num_trials)), time.time() - start
