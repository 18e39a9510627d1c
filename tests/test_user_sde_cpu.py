"""User-defined SDEs (the reference's plugin point is subclassing Sde): CPU-side checks -- the
kernel_spec() contract of a subclass with kernel_code(), the generated translation unit, and that it compiles for
sm_100a and exports the two entry points (no compute without a GPU)."""
import ctypes
import os

import pytest
import torch

from common import sm
from sde_mc_b200 import _jit, _lib as L, _spec


class Cir(sm.DiffusionSde):
    """dX = kappa (theta - X) dt + xi sqrt(max(X, 0)) dW -- not one of the built-in models"""

    def __init__(self, kappa, theta, xi, x0):
        super().__init__(x0, 1, 1, 'diag')
        self.kappa, self.theta, self.xi = kappa, theta, xi

    def drift(self, t, x):
        return self.kappa * (self.theta - x)

    def diffusion(self, t, x):
        return self.xi * torch.sqrt(torch.clamp(x, min=0))

    def kernel_code(self):
        return dict(drift=["p[0] * (p[1] - x[0])"], diffusion=["p[2] * sqrtf(fmaxf(x[0], 0.f))"],
                    params=[self.kappa, self.theta, self.xi])


class NoCode(sm.DiffusionSde):
    def __init__(self):
        super().__init__(torch.tensor([1.0]), 1, 1, 'diag')

    def drift(self, t, x):
        return x

    def diffusion(self, t, x):
        return x


def test_user_spec_and_source():
    spec = _spec.spec_of(Cir(2.0, 0.04, 0.2, torch.tensor([0.04])))
    assert spec.family == L.FAMILY_USER and spec.dim == 1 and spec.m == 1 and spec.marks == L.MARKS_NONE
    assert spec.user_p[:3] == [2.0, 0.04, 0.2] and len(spec.user_p) == 16
    src = _jit.source_for(spec.dim, spec.marks, spec.user_code)
    assert "case 0: return (float)(p[0] * (p[1] - x[0]));" in src and "@DIM@" not in src and "@JUMP_CASES@" not in src
    s = _spec.sde_struct(spec, 3.0, 10)
    assert s.family == 3 and abs(s.user_p[2] - 0.2) < 1e-7


def test_subclass_without_kernel_code_raises():
    with pytest.raises(L.SdemcError):
        _spec.spec_of(NoCode())


def test_user_library_builds_and_exports_entry_points():
    spec = _spec.spec_of(Cir(2.0, 0.04, 0.2, torch.tensor([0.04])))
    path = _jit.build(spec.dim, spec.marks, spec.user_code)
    assert os.path.exists(path)
    lib = ctypes.CDLL(path)
    for name in ("sdemc_user_version", "sdemc_user_mc_moments", "sdemc_user_solve_paths", "sdemc_user_last_cuda_error"):
        getattr(lib, name)
    assert lib.sdemc_user_version() == L.ABI_VERSION == L.load().sdemc_version()
    # wrong model shape is refused before any CUDA call
    other = _spec.sde_struct(_spec.spec_of(sm.Gbm(0.02, 0.3, torch.tensor([1.0]), 1)), 3.0, 10)
    lib.sdemc_user_mc_moments.argtypes = [ctypes.POINTER(L.SdemcSde), ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                                          ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]
    rng = L.SdemcRange(1, 0, 4)
    assert lib.sdemc_user_mc_moments(other, None, ctypes.cast(ctypes.pointer(rng), ctypes.c_void_p), None, None, None,
                                     None) == -2


def test_bad_expression_reports_the_compiler_error():
    class Broken(Cir):
        def kernel_code(self):
            return dict(drift=["p[0] * (p[1] - y[0])"], diffusion=["0.f"], params=[1.0, 1.0])

    spec = _spec.spec_of(Broken(2.0, 0.04, 0.2, torch.tensor([0.04])))
    with pytest.raises(L.SdemcError) as e:
        _jit.build(spec.dim, spec.marks, spec.user_code)
    assert "nvcc failed" in str(e.value) and "y" in str(e.value)
