"""Infinite-activity Levy models (API of /root/reference/sde_mc/levy.py): jumps smaller than epsilon are replaced by
an extra Brownian term, larger ones are sampled by inverting the CDF of the Levy measure."""
from abc import abstractmethod

import numpy as np
import torch

from . import _lib as L
from ._spec import _vec
from .helpers import get_jump_comp
from .sde import Sde

UNIFORM_TOL = 5.960464477539063e-08  # 2^-24: keeps torch.rand's 0 away from the icdf singularity (levy.py:7)


class InverseCdf:
    """Inverse CDF of the normalised Levy measure restricted to |x| > epsilon (levy.py:10-30):
    density  c_- e^{-mu(|x|-1)} on x<-1,  c_- |x|^{-1-alpha} on (-1,-eps),  c_+ x^{-1-alpha} on (eps,1),
    c_+ e^{-mu(x-1)} on x>1;  `lda` is its total mass (the jump intensity)."""

    def __init__(self, c_minus, c_plus, mu, alpha, epsilon):
        self.cm, self.cp, self.mu, self.alpha, self.eps = c_minus, c_plus, mu, alpha, epsilon
        power_mass = (epsilon ** (-alpha) - 1) / alpha
        self.lda = (c_minus + c_plus) * (1 / mu + power_mass)
        self.y1 = c_minus / (mu * self.lda)
        self.y2 = (c_minus / mu + c_minus * power_mass) / self.lda
        self.y3 = 1 - c_plus / (mu * self.lda)

    def __call__(self, y):
        cm, cp, mu, al, lda = self.cm, self.cp, self.mu, self.alpha, self.lda
        eps_pow = self.eps ** (-al)
        left_tail = torch.log((mu * lda * y) / cm) / mu - 1
        left_core = -(al * ((lda * y / cm) - (1 / mu)) + 1) ** (-1 / al)
        right_core = ((-al / cp) * (lda * y - cm / mu - cm * ((eps_pow - 1) / al)) + eps_pow) ** (-1 / al)
        right_tail = 1 - (1 / mu) * torch.log(mu * lda * (1 - y) / cp)
        return torch.where(y <= self.y1, left_tail,
                           torch.where(y < self.y2, left_core, torch.where(y < self.y3, right_core, right_tail)))

    def mark_params(self):
        return [float(v) for v in (self.cm, self.cp, self.mu, self.alpha, self.eps, self.lda, self.y1, self.y2,
                                   self.y3)] + [0.0] * 3


class Levy:
    """A Levy-driven SDE with infinite activity (levy.py:33-62)."""

    def __init__(self, dim, icdf):
        self.dim = dim
        self.icdf = icdf

    @abstractmethod
    def drift(self, t, x):
        pass

    @abstractmethod
    def diffusion(self, t, x):
        pass

    @abstractmethod
    def jumps(self, t, x, jumps):
        pass

    @abstractmethod
    def gamma(self):
        """first moment of the removed small jumps (drift correction)."""

    @abstractmethod
    def beta(self):
        """standard deviation of the removed small jumps (extra diffusion)."""

    @abstractmethod
    def jump_mean(self):
        pass


class LevySde(Sde):
    """Sde view of a Levy model: second Brownian driver for the small jumps, compound Poisson for the rest
    (levy.py:65-96)."""

    def __init__(self, levy, init_value, corr_matrix=None, scale_jump_rate=False, device='cpu', seed=1):
        super().__init__(init_value, levy.dim, levy.dim * 2, 'indep', corr_matrix)
        self.levy = levy
        self.scale_rate = scale_jump_rate

    def drift(self, t, x):
        return self.levy.drift(t, x) - self.levy.jumps(t, x, 1) * self.levy.gamma()

    def diffusion(self, t, x):
        return torch.stack((self.levy.diffusion(t, x), self.levy.jumps(t, x, 1) * self.levy.beta()), dim=-1)

    def jumps(self, t, x, jumps):
        return self.levy.jumps(t, x, jumps)

    def sample_jumps(self, size, device):
        return self.levy.icdf(torch.rand(size, device=device) + UNIFORM_TOL / 3)

    def jump_rate(self):
        scale = self.levy.dim if self.scale_rate else 1
        return torch.tensor(self.levy.icdf.lda * scale)

    def jump_mean(self):
        return self.levy.jump_mean()

    def kernel_spec(self):
        fn = getattr(self.levy, "kernel_coefficients", None)
        if fn is None:
            raise L.SdemcError("%s has no kernel coefficients (no CPU fallback)" % type(self.levy).__name__)
        family, a, b1, f = fn()
        spec = self._base_spec(family, m=2, marks=L.MARKS_ICDF)
        gamma, beta = float(self.levy.gamma()), float(self.levy.beta())
        d = self.dim
        spec.a = [a[i] - f[i] * gamma if i < d else 0.0 for i in range(L.MAX_DIM)]
        spec.b1 = list(b1)
        spec.b2 = [f[i] * beta if i < d else 0.0 for i in range(L.MAX_DIM)]
        spec.c = list(f)
        spec.rate = float(self.jump_rate())
        spec.mark_p = self.levy.icdf.mark_params()
        spec.jump_mean = float(self.levy.jump_mean())
        return spec


class _SmallJumpMoments:
    def gamma(self):
        return (self.cp - self.cm) * (1 - self.epsilon ** (1 - self.alpha)) / (1 - self.alpha)

    def beta(self):
        return np.sqrt((self.cp + self.cm) * (self.epsilon ** (2 - self.alpha)) / (2 - self.alpha))


class ExampleLevy(_SmallJumpMoments, Levy):
    """Log-price model with constant coefficients: dX = -(1/2 |sigma_i|^2 + comp_i) dt + sigma dW + f dL
    (levy.py:99-129)."""

    def __init__(self, c_plus, c_minus, alpha, mu, r, sigma, f, chol_corr, epsilon, dim):
        super().__init__(dim, InverseCdf(c_minus, c_plus, mu, alpha, epsilon))
        self.cm, self.cp, self.alpha, self.mu = c_minus, c_plus, alpha, mu
        self.f = f
        self.sigma = sigma
        self.epsilon = epsilon
        self.jump_comp = torch.tensor([get_jump_comp(c_plus, c_minus, alpha, mu, f[i].item()) for i in range(dim)],
                                      device=f.device)
        self.sigma_matrix = torch.matmul(torch.diag(sigma), chol_corr)
        self.row_sum_sq = (self.sigma_matrix ** 2).sum(-1)

    def drift(self, t, x):
        return -0.5 * self.row_sum_sq - self.jump_comp

    def diffusion(self, t, x):
        return torch.ones_like(x) * self.sigma

    def jumps(self, t, x, jumps):
        return torch.ones_like(x) * self.f * jumps

    def jump_mean(self):
        return 0

    def kernel_coefficients(self):
        d = self.dim
        drift = (-0.5 * self.row_sum_sq - self.jump_comp).detach().cpu().double()
        return (L.FAMILY_ARITHMETIC, _vec(drift, d, 'drift'), _vec(self.sigma, d, 'sigma'), _vec(self.f, d, 'f'))


class ExpExampleLevy(_SmallJumpMoments, Levy):
    """Price-level model dS = r S dt + sigma S dW + f S dL (levy.py:132-160)."""

    def __init__(self, c_minus, c_plus, alpha, mu, r, sigma, f, epsilon, dim=1):
        super().__init__(dim, InverseCdf(c_minus, c_plus, mu, alpha, epsilon))
        self.cm, self.cp, self.alpha, self.mu = c_minus, c_plus, alpha, mu
        self.r = r
        self.sigma = sigma
        self.f = f
        self.epsilon = epsilon

    def drift(self, t, x):
        return self.r * x

    def diffusion(self, t, x):
        return x * self.sigma

    def jumps(self, t, x, jumps):
        return self.f * x * jumps

    def jump_mean(self):
        return 0

    def kernel_coefficients(self):
        d = self.dim
        return (L.FAMILY_GEOMETRIC, _vec(self.r, d, 'r'), _vec(self.sigma, d, 'sigma'), _vec(self.f, d, 'f'))


class Levy2d(_SmallJumpMoments, Levy):
    """Two-dimensional additive model with unit Brownian coefficient and a common scaled jump (levy.py:163-192)."""

    def __init__(self, c_plus, c_minus, alpha, mu, f, epsilon):
        super().__init__(2, InverseCdf(c_minus, c_plus, mu, alpha, epsilon))
        self.cp, self.cm, self.alpha, self.mu = c_plus, c_minus, alpha, mu
        self.f = f
        self.epsilon = epsilon

    def _drift_constant(self):
        return -self.f * (self.cp - self.cm) * (1 / self.mu + 1 / (self.mu ** 2))

    def drift(self, t, x):
        return torch.ones_like(x) * self._drift_constant()

    def diffusion(self, t, x):
        return torch.ones_like(x)

    def jumps(self, t, x, jumps):
        return torch.ones_like(x) * self.f * jumps

    def jump_mean(self):
        if self.cp == self.cm:
            return 0
        mass = (self.cp + self.cm) * (1 / self.mu + (self.epsilon ** (-self.alpha) - 1) / self.alpha)
        first = (1 / self.mu + 1 / (self.mu ** 2)) + (1 - self.epsilon ** (1 - self.alpha)) / (1 - self.alpha)
        return (self.cp - self.cm) * first / mass

    def kernel_coefficients(self):
        return (L.FAMILY_ARITHMETIC, _vec(self._drift_constant(), 2, 'drift'), _vec(1.0, 2, 'sigma'),
                _vec(self.f, 2, 'f'))
