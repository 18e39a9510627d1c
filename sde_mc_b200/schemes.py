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
