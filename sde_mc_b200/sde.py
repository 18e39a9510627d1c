"""SDE model definitions -- the 'plugin API' of the package (mirrors /root/reference/sde_mc/sde.py).

Each class keeps the reference's constructor signature, attributes and tensor-valued coefficient methods (so code
written against piers-hinds/sde_mc keeps working), and additionally publishes `kernel_spec()`: the constant
coefficients the fused sm_100a kernels evaluate in registers.  The solvers never call the tensor methods.
"""
from abc import ABC, abstractmethod

import numpy as np
import torch

from . import _lib as L
from ._spec import KernelSpec, _vec, cholesky_rows


class Sde(ABC):
    """Base class: an SDE driven by Brownian motion and a compound Poisson process (sde.py:6-152).

    Attributes: init_value (dim,), dim, brown_dim, diffusion_struct in {'diag', 'indep', 'general'},
    corr_matrix (identity when not given), simulation_method.
    """

    def __init__(self, init_value, dim, brown_dim, diffusion_struct, corr_matrix=None, method='euler'):
        self.init_value = init_value
        self.dim = dim
        self.brown_dim = brown_dim
        self.diffusion_struct = diffusion_struct
        self.simulation_method = method
        self.corr_matrix = torch.eye(dim) if corr_matrix is None else corr_matrix

    @abstractmethod
    def drift(self, t, x):
        """(bs, dim) drift vector at (t, x)."""

    @abstractmethod
    def diffusion(self, t, x):
        """diffusion coefficients: (bs, dim) for 'diag', (bs, dim, brown_dim / dim) for 'indep'."""

    @abstractmethod
    def jumps(self, t, x, jumps):
        """(bs, dim) jump coefficient for jump marks `jumps`."""

    @abstractmethod
    def sample_jumps(self, size, device):
        """draw jump marks of shape `size`."""

    @abstractmethod
    def jump_mean(self):
        """expected jump mark."""

    @abstractmethod
    def jump_rate(self):
        """intensity of the Poisson process (tensor)."""

    # ---- engine interface -------------------------------------------------------------------------------
    def kernel_spec(self):
        """Coefficients for the fused kernels.  Built-in models override this with closed-form coefficient
        families; a USER-DEFINED subclass (the reference's plugin point: subclass Sde, write drift / diffusion /
        jumps) gets its own JIT-built kernels by also defining

            def kernel_code(self):
                return dict(drift=[...], diffusion=[...], jump=[...] or None, params=[...])

        with one CUDA float expression per component in terms of i, t, x[], p[] (and J for the jump coefficient),
        mirroring its Python tensor methods.  'diag' noise only; jump marks must be the log-normal ones of
        LogNormalJumpsSde.  See sde_mc_b200/_jit.py and csrc/user_model.cu.in."""
        code_fn = getattr(self, "kernel_code", None)
        if code_fn is None:
            raise L.SdemcError("%s does not publish kernel coefficients: built-in SDE classes do, user-defined ones "
                               "must define kernel_code() (no CPU fallback)" % type(self).__name__)
        if self.diffusion_struct != 'diag' or self.brown_dim != self.dim:
            raise L.SdemcError("user-defined SDEs run on the fused kernels with the 'diag' diffusion structure only")
        code = dict(code_fn())
        params = [float(v) for v in code.pop("params", [])]
        if len(params) > 16:
            raise L.SdemcError("at most 16 parameters p[] per user-defined SDE")
        rate = self.jump_rate()
        has_jumps = bool(torch.as_tensor(rate).any())
        spec = self._base_spec(L.FAMILY_USER, marks=L.MARKS_LOGNORMAL if has_jumps else L.MARKS_NONE)
        if has_jumps:
            if not (hasattr(self, "alpha") and hasattr(self, "gamma")) or code.get("jump") is None:
                raise L.SdemcError("user-defined jump SDEs need log-normal marks (alpha, gamma as in "
                                   "LogNormalJumpsSde) and a 'jump' expression in kernel_code()")
            spec.rate = float(torch.as_tensor(rate).double().sum())
            spec.mark_p[0], spec.mark_p[1] = float(self.alpha), float(self.gamma)
            spec.jump_mean = float(self.jump_mean())
        spec.user_p = params + [0.0] * (16 - len(params))
        spec.user_code = dict(drift=code["drift"], diffusion=code["diffusion"], jump=code.get("jump"))
        return spec

    def _base_spec(self, family, m=1, marks=L.MARKS_NONE):
        spec = KernelSpec(family=family, dim=self.dim, m=m, marks=marks)
        spec.x0 = _vec(self.init_value, self.dim, 'init_value')
        spec.chol = cholesky_rows(self.corr_matrix, self.dim)
        return spec


class DiffusionSde(Sde):
    """An SDE without jumps (sde.py:155-176)."""

    def jumps(self, t, x, jumps):
        return None

    def sample_jumps(self, size, device):
        return None

    def jump_mean(self):
        return None

    def jump_rate(self):
        return torch.tensor(0)


class Gbm(DiffusionSde):
    """Correlated multi-dimensional geometric Brownian motion dX = mu X dt + sigma X dW (sde.py:179-207)."""

    def __init__(self, mu, sigma, init_value, dim, corr_matrix=None, method='euler'):
        super().__init__(init_value, dim, dim, 'diag', corr_matrix, method)
        self.mu = mu
        self.sigma = sigma

    def drift(self, t, x):
        return self.mu * x

    def diffusion(self, t, x):
        return self.sigma * x

    def kernel_spec(self):
        spec = self._base_spec(L.FAMILY_GEOMETRIC)
        spec.a = _vec(self.mu, self.dim, 'mu')
        spec.b1 = _vec(self.sigma, self.dim, 'sigma')
        return spec


class LogGbm(Gbm):
    """One-dimensional log-price of a GBM: constant coefficients (sde.py:210-219)."""

    def __init__(self, mu, sigma, init_value):
        super().__init__(mu, sigma, init_value, 1)

    def drift(self, t, x):
        return torch.ones_like(x) * (self.mu - 0.5 * self.sigma * self.sigma)

    def diffusion(self, t, x):
        return torch.ones_like(x) * self.sigma

    def kernel_spec(self):
        spec = self._base_spec(L.FAMILY_ARITHMETIC)
        mu, sigma = _vec(self.mu, 1, 'mu')[0], _vec(self.sigma, 1, 'sigma')[0]
        spec.a[0] = mu - 0.5 * sigma * sigma
        spec.b1[0] = sigma
        return spec


class DoubleGbm(DiffusionSde):
    """GBM driven by two independent d-dimensional Brownian motions -- the 'indep' structure (sde.py:222-255)."""

    def __init__(self, mu, sigma1, sigma2, init_value, dim, corr_matrix=None, method='euler'):
        super().__init__(init_value, dim, 2 * dim, 'indep', corr_matrix, method)
        self.mu = mu
        self.sigma1 = sigma1
        self.sigma2 = sigma2

    def drift(self, t, x):
        return self.mu * x

    def diffusion(self, t, x):
        return torch.stack((self.sigma1 * x, self.sigma2 * x), dim=-1)

    def kernel_spec(self):
        spec = self._base_spec(L.FAMILY_GEOMETRIC, m=2)
        spec.a = _vec(self.mu, self.dim, 'mu')
        spec.b1 = _vec(self.sigma1, self.dim, 'sigma1')
        spec.b2 = _vec(self.sigma2, self.dim, 'sigma2')
        return spec


class Heston(DiffusionSde):
    """Heston stochastic volatility, state (S, v), simulated with the drift-implicit square-root scheme
    (sde.py:258-279, schemes.py:16-22).  Requires the Feller-type condition 2 kappa theta > xi^2."""

    def __init__(self, r, kappa, theta, xi, rho, init_value):
        assert 2 * kappa * theta > xi ** 2
        super().__init__(init_value, 2, 2, 'diag', torch.tensor([[1., rho], [rho, 1.]]))
        self.simulation_method = 'heston'
        self.r = r
        self.kappa = kappa
        self.theta = theta
        self.xi = xi

    def drift(self, t, x):
        return torch.stack((self.r * x[:, 0], torch.zeros_like(x[:, 1])), dim=1)

    def diffusion(self, t, x):
        return torch.stack((x[:, 1].sqrt() * x[:, 0], torch.zeros_like(x[:, 1])), dim=1)

    def quadratic_parameters(self, x, h, normals):
        """coefficients (a, b, c) of the quadratic solved for sqrt(v_next)."""
        a = -torch.ones_like(x) - self.kappa * h
        b = self.xi * normals
        c = x + self.kappa * self.theta * h - 0.5 * h * self.xi ** 2
        return a, b, c

    def kernel_spec(self):
        spec = self._base_spec(L.FAMILY_HESTON)
        spec.scheme = L.SCHEME_HESTON
        spec.heston = [float(self.r), float(self.kappa), float(self.theta), float(self.xi)]
        return spec


class LogNormalJumpsSde(Sde):
    """Base for SDEs whose jump marks are shifted log-normals exp(gamma Z + alpha) - 1 (sde.py:282-332)."""

    def __init__(self, rate, alpha, gamma, init_value, dim, brown_dim, diffusion_struct, corr_matrix=None,
                 method='euler'):
        super().__init__(init_value, dim, brown_dim, diffusion_struct, corr_matrix, method)
        self.rate = rate if torch.is_tensor(rate) else torch.tensor(rate)
        self.alpha = alpha
        self.gamma = gamma

    def sample_jumps(self, size, device):
        return torch.exp(torch.randn(size=size, device=device) * self.gamma + self.alpha) - 1

    def jump_mean(self):
        return np.exp(self.alpha + 0.5 * self.gamma * self.gamma) - 1

    def jump_rate(self):
        return self.rate


class Merton(LogNormalJumpsSde):
    """Merton jump-diffusion: GBM plus compound-Poisson log-normal jumps, compensated drift (sde.py:335-375)."""

    def __init__(self, mu, sigma, rate, alpha, gamma, init_value, dim, corr_matrix=None, method='euler'):
        self.mu = mu
        self.sigma = sigma
        super().__init__(rate, alpha, gamma, init_value, dim, dim, 'diag', corr_matrix, method)

    def drift(self, t, x):
        return (self.mu - self.rate * self.jump_mean()) * x

    def diffusion(self, t, x):
        return self.sigma * x

    def jumps(self, t, x, jumps):
        return x * jumps

    def kernel_spec(self):
        if self.rate.numel() != 1:
            raise L.SdemcError("the Merton kernel takes a scalar jump rate")
        spec = self._base_spec(L.FAMILY_GEOMETRIC, marks=L.MARKS_LOGNORMAL)
        rate = float(self.rate)
        jm = float(self.jump_mean())
        spec.a = [mu - rate * jm if i < self.dim else 0.0 for i, mu in enumerate(_vec(self.mu, self.dim, 'mu'))]
        spec.b1 = _vec(self.sigma, self.dim, 'sigma')
        spec.c = [1.0] * self.dim + [0.0] * (L.MAX_DIM - self.dim)
        spec.rate = rate
        spec.mark_p[0], spec.mark_p[1] = float(self.alpha), float(self.gamma)
        spec.jump_mean = jm
        return spec


class AsianWrapper(Sde):
    """Augments a 1-D SDE X with its running integral: state (X_t, int_0^t X_s ds) (sde.py:378-406)."""

    def __init__(self, base_sde):
        super().__init__(torch.cat([base_sde.init_value, torch.tensor([0.])]), 2, 2, 'diag', None,
                         base_sde.simulation_method)
        self.base_sde = base_sde

    def drift(self, t, x):
        return torch.stack((self.base_sde.drift(t, x[:, 0]), x[:, 0]), dim=1)

    def diffusion(self, t, x):
        return torch.stack((self.base_sde.diffusion(t, x[:, 0]), torch.zeros_like(x[:, 0])), dim=1)

    def jumps(self, t, x, jumps):
        return torch.stack((self.base_sde.jumps(t, x[:, 0], jumps[:, 0]), torch.zeros_like(x[:, 0])), dim=1)

    def sample_jumps(self, size, device):
        return self.base_sde.sample_jumps(size, device)

    def jump_mean(self):
        return self.base_sde.jump_mean()

    def jump_rate(self):
        return self.base_sde.jump_rate()

    def kernel_spec(self):
        from ._spec import spec_of
        base = spec_of(self.base_sde)
        if base.dim != 1 or base.m != 1 or base.family == L.FAMILY_HESTON:
            raise L.SdemcError("AsianWrapper kernels exist for 1-D 'diag' base SDEs only")
        spec = KernelSpec(family=base.family, dim=2, m=1, marks=base.marks, asian=1)
        spec.x0 = [base.x0[0], 0.0, 0.0, 0.0]
        spec.chol = cholesky_rows(None, 2)
        spec.a, spec.b1, spec.c = list(base.a), list(base.b1), list(base.c)
        spec.rate, spec.mark_p, spec.jump_mean = base.rate, list(base.mark_p), base.jump_mean
        return spec
