"""Parameter extraction: turns (Sde, solver, Option, discounter) objects into the POD structs of the C-ABI.

The reference evaluates Python coefficient callbacks on (bs, dim) tensors every step (sde.py:63-152); here each
built-in model publishes its coefficients once, as doubles, and the fused kernels evaluate them in registers.
User-defined Sde subclasses have no kernel: they raise (no CPU fallback).
"""
from dataclasses import dataclass, field
from typing import List

import numpy as np
import torch

from . import _lib as L


def _vec(v, dim, name):
    """scalar / tensor / sequence -> list of `dim` python floats (padded to MAX_DIM with zeros)."""
    if torch.is_tensor(v):
        v = v.detach().cpu().double().reshape(-1).tolist()
    elif isinstance(v, np.ndarray):
        v = v.astype(np.float64).reshape(-1).tolist()
    elif isinstance(v, (list, tuple)):
        v = [float(e) for e in v]
    else:
        v = [float(v)]
    if len(v) == 1:
        v = v * dim
    if len(v) != dim:
        raise ValueError("%s has %d entries for a %d-dimensional SDE" % (name, len(v), dim))
    return v + [0.0] * (L.MAX_DIM - dim)


@dataclass
class KernelSpec:
    """Coefficients of one SDE in the form the kernels consume (all python floats = doubles)."""
    family: int
    dim: int
    m: int = 1
    marks: int = L.MARKS_NONE
    scheme: int = L.SCHEME_EULER
    asian: int = 0
    x0: List[float] = field(default_factory=lambda: [0.0] * L.MAX_DIM)
    chol: List[float] = field(default_factory=lambda: [0.0] * (L.MAX_DIM * L.MAX_DIM))
    a: List[float] = field(default_factory=lambda: [0.0] * L.MAX_DIM)
    b1: List[float] = field(default_factory=lambda: [0.0] * L.MAX_DIM)
    b2: List[float] = field(default_factory=lambda: [0.0] * L.MAX_DIM)
    c: List[float] = field(default_factory=lambda: [0.0] * L.MAX_DIM)
    rate: float = 0.0
    mark_p: List[float] = field(default_factory=lambda: [0.0] * 12)
    heston: List[float] = field(default_factory=lambda: [0.0] * 4)
    jump_mean: float = 0.0
    user_p: List[float] = field(default_factory=lambda: [0.0] * 16)   # USER family: parameters p[]
    user_code: object = None   # USER family: dict(drift=[...], diffusion=[...], jump=[...]) of CUDA expressions


def cholesky_rows(corr_matrix, dim):
    """Lower Cholesky factor of the correlation matrix as a flat row-major list with stride MAX_DIM
    (SdeSolver.__init__ solvers.py:33-36: identity-like [[1.]] when the matrix is 1x1)."""
    out = [0.0] * (L.MAX_DIM * L.MAX_DIM)
    if corr_matrix is None or len(corr_matrix) <= 1:
        for i in range(dim):
            out[i * L.MAX_DIM + i] = 1.0
        return out
    # in the dtype the reference would factor it in (solvers.py:33-36: the default dtype; fp64 only for the fp64 pair)
    chol = torch.linalg.cholesky(torch.as_tensor(corr_matrix).detach().cpu().to(torch.get_default_dtype())).double()
    n = chol.shape[0]
    if n != dim:
        raise ValueError("correlation matrix is %dx%d for a %d-dimensional SDE" % (n, n, dim))
    for i in range(n):
        for j in range(n):
            out[i * L.MAX_DIM + j] = float(chol[i, j])
    return out


def sde_struct(spec, T, num_steps, max_jumps=0, exact_jumps=False, jump_strategy=L.JUMPS_AUTO, queue_depth=0,
               short_path=L.SHORT_AUTO):
    s = L.SdemcSde()
    s.family, s.scheme, s.dim, s.m, s.marks = spec.family, spec.scheme, spec.dim, spec.m, spec.marks
    s.num_steps, s.max_jumps, s.exact_jumps, s.asian = int(num_steps), int(max_jumps), int(bool(exact_jumps)), spec.asian
    s.jump_strategy, s.queue_depth, s.short_path = int(jump_strategy), int(queue_depth), int(short_path)
    s.T = float(T)
    for i in range(L.MAX_DIM):
        s.x0[i], s.a[i], s.b1[i], s.b2[i], s.c[i] = spec.x0[i], spec.a[i], spec.b1[i], spec.b2[i], spec.c[i]
    for i in range(L.MAX_DIM * L.MAX_DIM):
        s.chol[i] = spec.chol[i]
    s.rate = spec.rate
    for i in range(12):
        s.mark_p[i] = spec.mark_p[i]
    for i in range(4):
        s.heston[i] = spec.heston[i]
    for i in range(16):
        s.user_p[i] = spec.user_p[i]
    return s


# ---- whose coefficients are these? -----------------------------------------------------------------------------------
# The reference always calls the Python methods of an Sde / Option, so a user subclass that overrides one of them
# changes the model (class Cev(Gbm) with a new diffusion(), class MyCall(EuroCall) with a new payoff()).  The
# kernel_spec() of a built-in class describes the built-in model only: it may not be used when a user class -- any
# class outside this package, more derived than the built-in -- overrides a method the reference's loop would call,
# unless that user class restates the coefficients itself (kernel_spec / kernel_code at the same level or below).
_SDE_METHODS = ("drift", "diffusion", "jumps", "sample_jumps", "jump_mean", "jump_rate")
_OPTION_METHODS = ("payoff", "transform", "__call__")
_PACKAGE = __name__.rsplit(".", 1)[0]


def _user_classes(obj):
    """the classes of the object's MRO that are more derived than the first class of this package"""
    out = []
    for cls in type(obj).__mro__:
        if cls.__module__ == _PACKAGE or cls.__module__.startswith(_PACKAGE + "."):
            break
        out.append(cls)
    return out


def user_overrides(obj, names, restated_by):
    """(overridden, restated): the methods among `names` that user classes override WITHOUT restating the kernel
    coefficients (an attribute of `restated_by`) at that level or further down the hierarchy, and whether any user
    class restates them at all"""
    over = []
    for cls in _user_classes(obj):
        if any(a in cls.__dict__ for a in restated_by):
            return over, True
        over += [n for n in names if n in cls.__dict__]
    return over, False


def payoff_kernel_spec(payoff):
    """the payoff's kernel_spec method when the fused kernels evaluate exactly what its Python payoff computes, else
    None (user-defined Option, or a subclass of a built-in that overrides payoff / transform: those run as Python
    on stored trajectories, mc._mc_simple_python_payoff)"""
    fn = getattr(payoff, "kernel_spec", None)
    if fn is None or user_overrides(payoff, _OPTION_METHODS, ("kernel_spec",))[0]:
        return None
    return fn


def payoff_struct(payoff, discount_factor, index_mode):
    spec = payoff_kernel_spec(payoff)
    if spec is None:
        raise L.SdemcError("payoff %r is not one of the built-in Option classes; the fused kernels cannot evaluate "
                           "arbitrary Python payoffs (no CPU fallback)" % type(payoff).__name__)
    kind, strike, aux = spec()
    p = L.SdemcPayoff()
    p.kind, p.log, p.index_mode = int(kind), int(bool(payoff.log)), int(index_mode)
    p.strike = float(strike)
    p.transform_discount = float(payoff.discount)
    p.aux = float(aux)
    p.df = float(discount_factor)
    return p


def engine_lib(spec):
    """the library whose entry points serve this model: the stock engine, or the JIT-built one of a user model"""
    if spec.family == L.FAMILY_USER:
        from . import _jit
        return _jit.library_for(spec)
    return L.load()


def spec_of(sde):
    fn = getattr(sde, "kernel_spec", None)
    if fn is None:
        raise L.SdemcError("%s does not publish kernel coefficients; only the built-in SDE classes run on the B200 "
                           "engine (no CPU fallback)" % type(sde).__name__)
    over, restated = user_overrides(sde, _SDE_METHODS, ("kernel_spec", "kernel_code"))
    if over:
        raise L.SdemcError("%s overrides %s but inherits the kernel coefficients of a parent class; the fused kernels "
                           "would simulate the parent model.  Define kernel_code() on the subclass (see "
                           "Sde.kernel_spec); there is no CPU fallback" % (type(sde).__name__, ", ".join(over)))
    if restated and not any("kernel_spec" in c.__dict__ for c in _user_classes(sde)):
        from .sde import Sde                       # the user's kernel_code() describes the model: USER family
        return Sde.kernel_spec(sde)
    return fn()
