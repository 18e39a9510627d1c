"""One-step update rules as tensor functions (API of /root/reference/sde_mc/schemes.py).

The fused kernels implement the same rules in registers (csrc/engine.cuh: euler_step, heston_step_uniform); these
mixins only keep `solver.step(t, x, h, dW)` available to user code on tensors."""
import torch

from .helpers import solve_quadratic


class EulerScheme:
    def step(self, t, x, h, corr_normals):
        """Euler-Maruyama: x + a h + b dW; 'indep' sums the contributions of the drivers (schemes.py:5-13)."""
        increment = self.sde.diffusion(t, x) * corr_normals
        if self.sde.diffusion_struct == 'indep':
            increment = increment.sum(dim=-1)
        elif self.sde.diffusion_struct != 'diag':
            raise NotImplementedError("'general' diffusion structure is not supported (it is broken upstream too)")
        return x + self.sde.drift(t, x) * h + increment


class HestonScheme:
    def step(self, t, x, h, corr_normals):
        """Euler for the price, drift-implicit square-root step for the variance (schemes.py:16-22)."""
        out = x + self.sde.drift(t, x) * h + self.sde.diffusion(t, x) * corr_normals
        root = solve_quadratic(self.sde.quadratic_parameters(x[:, 1], h, corr_normals[:, 1]))
        out[:, 1] = root * root
        return out


class MilsteinScheme:
    """Milstein: Euler + 1/2 b b' (dW^2 - h) for 'diag' noise.  EXTENSION -- the reference has no Milstein scheme
    (SURVEY.md S3); validated by strong order 1 against the exact GBM solution.  The derivative b' is known for the
    built-in families: geometric b = sigma x -> b' = sigma; arithmetic (constant b) -> b' = 0, i.e. Euler."""

    def step(self, t, x, h, corr_normals):
        if self.sde.diffusion_struct != 'diag':
            raise NotImplementedError("Milstein is implemented for the 'diag' diffusion structure only")
        b = self.sde.diffusion(t, x)
        out = x + self.sde.drift(t, x) * h + b * corr_normals
        spec = self.sde.kernel_spec()
        from . import _lib as L
        if spec.family == L.FAMILY_GEOMETRIC:
            sigma = torch.tensor(spec.b1[:self.sde.dim], dtype=x.dtype, device=x.device)
            if spec.asian:
                sigma[-1] = 0.0
            out = out + 0.5 * b * sigma * (corr_normals * corr_normals - h)
        return out
