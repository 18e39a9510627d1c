"""Solvers: the step loops of /root/reference/sde_mc/solvers.py, executed as fused sm_100a kernels.

`solve()` / `multilevel_solve()` keep the reference's signatures and return layouts; the Python `for`/`while`
loops over time steps (solvers.py:83-87, :182-225, :254-306) are replaced by one kernel launch in which every path
lives in the registers of one CUDA thread.  Noise is Philox4x32-10 keyed by `seed` and countered by a global path
id that keeps advancing across calls, so successive `solve()` calls draw fresh, reproducible paths; pass
`inject=` to drive the kernels with explicit noise arrays instead (the deterministic parity mode).
"""
from abc import ABC, abstractmethod

import numpy as np
import torch
import torch.nn.functional as F  # noqa: F401  (star-exported by the reference's solvers.py:2)
from scipy.stats import poisson

from . import _lib as L
from . import _spec
from .schemes import EulerScheme, HestonScheme, MilsteinScheme


def _alloc_rows(bs, inner_shape, dev, align):
    """(bs, *inner_shape) fp32 output whose rows (one per path) start `align` floats apart-aligned: the storing
    kernels write 16-byte vectors / whole 128-byte lines when the row pitch allows it.  Returns (tensor, pitch);
    with padding the tensor is a strided view (same shape and values as the reference's dense allocation)."""
    length = int(np.prod(inner_shape))
    pitch = -(-length // align) * align if align and align > 1 else length
    buf = torch.empty((bs, pitch), device=dev, dtype=torch.float32)
    if pitch == length:
        return buf.view((bs,) + tuple(inner_shape)), pitch
    return buf[:, :length].unflatten(1, tuple(inner_shape)), pitch


def _as_dev_f32(a, dev):
    if a is None:
        return None
    t = torch.as_tensor(a)
    return t.to(device=dev, dtype=torch.float32).contiguous()


class SdeSolver(ABC):
    """Common state of all solvers (solvers.py:9-56).

    `device` is where returned tensors live.  The computation always runs on a CUDA device: `device` itself if it
    is one, else the current CUDA device (results are then copied back).  Without a GPU every solve raises --
    there is no CPU implementation in this package.
    """

    def __init__(self, sde, time_interval, num_steps, device='cpu', seed=1):
        self.sde = sde
        self.time_interval = time_interval
        self.num_steps = num_steps
        self.device = device
        self.seed = seed
        self.has_jumps = self.sde.jump_rate().any()
        if len(self.sde.corr_matrix) > 1:
            self.lower_cholesky = torch.linalg.cholesky(self.sde.corr_matrix.to(device))
        else:
            self.lower_cholesky = torch.tensor([[1.]], device=device)
        torch.manual_seed(seed)       # the reference reseeds torch's global RNG here (solvers.py:37)
        self._next_path = 0           # global Philox path id of the next path to simulate
        self.jump_strategy = L.JUMPS_AUTO    # sdemc_jump_strategy: how the kernels draw compound-Poisson jumps
        self.queue_depth = 0                 # QUEUE strategy: pre-drawn jumps per refill (0 = sized from rate * T)
        self.short_path = L.SHORT_AUTO       # sdemc_short_path: persistent-lane kernels for short paths
        self.tma_store = True                # solve(): TMA tiles where the layout allows them
        # Row pitch of stored trajectories in floats: rows are padded to a multiple of this (32 floats = one
        # 128-byte line) so the path-storing kernels can use 16-byte vector stores.  Set to 1 for the reference's
        # dense (contiguous) allocations -- same values, 4-byte store path.
        self.row_align = 32
        # A list here makes solve() append a (start, end) pair of CUDA events recorded around its kernel launch on the
        # current stream (bench.py: the launch duration behind roofline.achieved, without allocation and host gaps).
        self.kernel_events = None

    # ---- engine plumbing -----------------------------------------------------------------------------------
    def _record_kernel_event(self, start):
        """kernel_events hook: called before (start=None) and after the launch of solve()'s kernel."""
        if self.kernel_events is None:
            return None
        ev = torch.cuda.Event(enable_timing=True)
        ev.record()
        if start is not None:
            self.kernel_events.append((start, ev))
        return ev

    def _compute_device(self):
        dev = torch.device(self.device)
        if dev.type == 'cuda':
            return L.require_cuda(dev)
        if not torch.cuda.is_available():
            raise L.SdemcError("sde_mc_b200 needs a B200 GPU: no CUDA device is available and there is no CPU path")
        return torch.device('cuda', torch.cuda.current_device())

    def _take_paths(self, n):
        lo = self._next_path
        self._next_path += int(n)
        return lo

    def _max_jumps(self):
        return 0

    def _exact_jumps(self):
        return False

    def _sde_struct(self, num_steps=None):
        spec = _spec.spec_of(self.sde)
        if isinstance(self, MilsteinScheme):
            if spec.m != 1 or spec.family == L.FAMILY_HESTON:
                raise L.SdemcError("the Milstein scheme is available for 'diag' geometric / arithmetic SDEs only")
            import dataclasses
            spec = dataclasses.replace(spec, scheme=L.SCHEME_MILSTEIN)
        return _spec.sde_struct(spec, self.time_interval, self.num_steps if num_steps is None else num_steps,
                                self._max_jumps(), self._exact_jumps(), self.jump_strategy, self.queue_depth,
                                self.short_path)

    def _engine_lib(self):
        """the library serving this solver's model: the stock engine or the JIT-built one of a user-defined SDE"""
        return _spec.engine_lib(_spec.spec_of(self.sde))

    def _to_user_device(self, t):
        return t if t is None or torch.device(self.device) == t.device else t.to(self.device)

    # ---- reference API ----------------------------------------------------------------------------------------
    @abstractmethod
    def solve(self, bs=1, return_normals=False):
        pass

    @abstractmethod
    def init_storage(self, bs, steps):
        pass

    @abstractmethod
    def step(self, t, x, h, corr_normals):
        pass

    def sample_corr_normals(self, size, h, corr=True):
        """torch-RNG Brownian increments with the reference's shape contract (solvers.py:51-56).  Kept for API
        compatibility; the kernels draw their own Philox normals and never call this."""
        normals = torch.randn(size=size, device=self.device) * torch.sqrt(h)
        if not corr:
            return normals.squeeze(-1)
        return torch.matmul(self.lower_cholesky, normals).squeeze(-1)


class DiffusionSolver(SdeSolver):
    """Uniform-grid solver for SDEs without jumps (solvers.py:59-119)."""

    def init_storage(self, bs, steps):
        return torch.empty(size=(bs, steps + 1, self.sde.dim), device=self.device)

    def solve(self, bs=1, return_normals=False, inject=None, want_payoff=None):
        """Returns (paths (bs, steps+1, dim), normals (bs, steps, dim[, m])) like solvers.py:68-88
        (`return_normals` is ignored there too: the increments are always returned).

        inject: optional dict(z=(bs, steps, dim, m) unit normals) -- deterministic parity mode.
        want_payoff: optional (payoff_struct) to also get per-path discounted payoffs as a third return value."""
        bs = int(bs)
        dev = self._compute_device()
        lib = self._engine_lib()
        S, d = int(self.num_steps), self.sde.dim
        m = self.sde.brown_dim // self.sde.dim
        sde = self._sde_struct()
        with torch.cuda.device(dev):
            paths, p_state = _alloc_rows(bs, (S + 1, d), dev, self.row_align)
            normals, p_norm = _alloc_rows(bs, (S, d) if m == 1 else (S, d, m), dev, self.row_align)
            payoffs = torch.empty((bs,), device=dev, dtype=torch.float32) if want_payoff is not None else None
            out = L.SdemcPathsOut(L.ptr(paths), None, None, None, L.ptr(normals), L.ptr(payoffs), None, None,
                                  p_state, 0, p_norm, flags=0 if self.tma_store else L.OUT_NO_TMA)
            inj = None
            keep = []
            if inject is not None:
                z = _as_dev_f32(inject['z'], dev).reshape(bs, S, d, m)
                keep.append(z)
                inj = L.SdemcInject(L.ptr(z), None, None, None, S)
            rng = L.SdemcRange(int(self.seed), self._take_paths(bs), bs)
            ev = self._record_kernel_event(None)
            getattr(lib, 'check', L.check)(lib.sdemc_solve_paths(sde, want_payoff, rng, inj, out, L.ptr(L.workspace(dev)),
                                                                  L.stream_ptr(dev)))
            self._record_kernel_event(ev)
        paths, normals = self._to_user_device(paths), self._to_user_device(normals)
        if want_payoff is not None:
            return paths, normals, self._to_user_device(payoffs)
        return paths, normals

    def multilevel_solve(self, bs, levels, return_normals=False, inject=None):
        """Coupled fine/coarse pair on shared increments (solvers.py:90-119; Euler or Heston steps).
        Returns ((paths_fine, paths_coarse), None) with paths of shape (bs, 2, dim): index 0 = the initial value,
        index -1 = the terminal state -- what the estimators read (mlmc.py:64-65).  The reference's intermediate grid
        points and its `corr_normals` are not materialised: the pair lives in the registers of one thread."""
        from .mlmc import _pair_paths_diffusion
        return _pair_paths_diffusion(self, int(bs), levels, inject)


class EulerSolver(EulerScheme, DiffusionSolver):
    pass


class HestonSolver(HestonScheme, DiffusionSolver):
    pass


class MilsteinSolver(MilsteinScheme, DiffusionSolver):
    """Uniform-grid Milstein solver -- extension, not in the reference."""
    pass


class JumpDiffusionSolver(SdeSolver):
    """Jump-adapted solver: steps of the mesh size, shortened to land on every jump time (solvers.py:130-307)."""

    def __init__(self, sde, time_interval, num_steps, device='cpu', seed=1, exact_jumps=False):
        super().__init__(sde, time_interval, num_steps, device, seed)
        total_rate = float(self.sde.jump_rate().sum())
        self.max_jumps = max(int(self.time_interval * poisson.ppf(1 - 1 / 1e9, total_rate)), 5)
        self.exact_jumps = exact_jumps

    def _max_jumps(self):
        return self.max_jumps

    def _exact_jumps(self):
        return self.exact_jumps

    def add_jumps(self, t, old_x, x, jumps):
        return x + self.sde.jumps(t, old_x, jumps)

    def sample_jump_times(self, size):
        """torch-RNG cumulative jump times (solvers.py:143-144); API compatibility only."""
        return torch.empty(size, device=self.device).exponential_(self.sde.jump_rate().sum()).cumsum(dim=1)

    def sample_one_jump(self, size):
        """one common mark per path, repeated over the components (solvers.py:146-148); API compatibility only."""
        return self.sde.sample_jumps([size, 1], self.device).repeat(1, self.sde.dim)

    def init_storage(self, bs, steps, low_storage=False):
        """Allocate the reference's storage layout (solvers.py:150-162)."""
        d = self.sde.dim
        paths = torch.zeros((bs, steps + 1, d), device=self.device)
        if low_storage:
            return paths, None, None, None, None
        m = self.sde.brown_dim // d
        nshape = (bs, steps, d) if self.sde.diffusion_struct == 'diag' else (bs, steps, d, m)
        return (paths, torch.zeros_like(paths), torch.zeros((bs, steps + 1, 1), device=self.device) + self.time_interval,
                torch.zeros_like(paths), torch.zeros(nshape, device=self.device))

    def solve(self, bs=1, return_normals=False, low_storage=False, inject=None, want_payoff=None):
        """Returns (paths[:, :total_steps+1], (normals, time_paths, left_paths, total_steps, jump_paths)) like
        solvers.py:164-226; with low_storage only `paths` is produced and the tuple holds Nones.

        inject: optional dict(z=(bs, K, dim), zc=(bs, K) [indep only], jump_times=(bs, max_jumps), marks=(bs, K))."""
        bs = int(bs)
        dev = self._compute_device()
        lib = self._engine_lib()
        d = self.sde.dim
        m = self.sde.brown_dim // d
        sde = self._sde_struct()
        S = int(self.num_steps) + int(self.max_jumps)
        keep = []
        inj = None
        with torch.cuda.device(dev):
            if inject is not None:
                z = _as_dev_f32(inject['z'], dev)
                S = int(z.shape[1])
                zc = _as_dev_f32(inject.get('zc'), dev)
                jt = _as_dev_f32(inject['jump_times'], dev)
                mk = _as_dev_f32(inject['marks'], dev)
                assert z.shape == (bs, S, d) and jt.shape == (bs, self.max_jumps) and mk.shape == (bs, S)
                keep += [z, zc, jt, mk]
                inj = L.SdemcInject(L.ptr(z), L.ptr(zc), L.ptr(jt), L.ptr(mk), S)
            paths, p_state = _alloc_rows(bs, (S + 1, d), dev, self.row_align)
            p_times = p_norm = 0
            total = torch.zeros((1,), device=dev, dtype=torch.int32)
            iters = torch.empty((bs,), device=dev, dtype=torch.int32)
            payoffs = torch.empty((bs,), device=dev, dtype=torch.float32) if want_payoff is not None else None
            left = times = jumps = normals = None
            if not low_storage:
                left, _ = _alloc_rows(bs, (S + 1, d), dev, self.row_align)
                jumps, _ = _alloc_rows(bs, (S + 1, d), dev, self.row_align)
                times, p_times = _alloc_rows(bs, (S + 1,), dev, self.row_align)
                normals, p_norm = _alloc_rows(bs, (S, d) if m == 1 else (S, d, m), dev, self.row_align)
            out = L.SdemcPathsOut(L.ptr(paths), L.ptr(left), L.ptr(times), L.ptr(jumps), L.ptr(normals),
                                  L.ptr(payoffs), L.ptr(iters), L.ptr(total), p_state, p_times, p_norm,
                                  flags=0 if self.tma_store else L.OUT_NO_TMA)
            rng = L.SdemcRange(int(self.seed), self._take_paths(bs), bs)
            ev = self._record_kernel_event(None)
            getattr(lib, 'check', L.check)(lib.sdemc_solve_paths(sde, want_payoff, rng, inj, out, L.ptr(L.workspace(dev)),
                                                                  L.stream_ptr(dev)))
            self._record_kernel_event(ev)
            total_steps = int(total.item())
        self.last_iters = iters
        u = self._to_user_device
        aux = (u(normals), None if times is None else u(times.unsqueeze(-1)), u(left), total_steps, u(jumps))
        paths = u(paths[:, :total_steps + 1])
        if want_payoff is not None:
            return paths, aux, u(payoffs)
        return paths, aux

    def multilevel_solve(self, bs, levels, return_normals=False, inject=None):
        """Coupled fine/coarse jump-adapted pair sharing increments and jumps (solvers.py:228-307).
        Returns ((paths_fine, paths_coarse), None) with paths of shape (bs, 2, dim): index 0 = the initial value,
        index -1 = the terminal state -- what the estimators read (mlmc.py:64-65); the reference's states at the
        coarse iterations in between are not materialised."""
        from .mlmc import _pair_paths_jump
        return _pair_paths_jump(self, int(bs), levels, inject)


class JumpEulerSolver(EulerScheme, JumpDiffusionSolver):
    pass


class JumpMilsteinSolver(MilsteinScheme, JumpDiffusionSolver):
    """Jump-adapted solver with Milstein steps between the jumps -- extension, not in the reference."""
    pass


class Grid(ABC):
    """Iterator over time points (solvers.py:314-327)."""

    def __init__(self, start, end):
        self.start = start
        self.end = end
        self.time_interval = end - start
        self.t = start

    def __iter__(self):
        return self

    @abstractmethod
    def __next__(self):
        pass


class UniformGrid(Grid):
    def __init__(self, start, end, num_steps):
        super().__init__(start, end)
        self.num_steps = num_steps
        self.h = (end - start) / num_steps
        assert self.h > 1e-8

    def __next__(self):
        if self.t > self.end - 1e-8:
            raise StopIteration
        now = self.t
        self.t += self.h
        return now
