"""Named presets bundling (solver, discounter, payoff) -- API of /root/reference/sde_mc/problem.py."""
from abc import ABC

import torch

from .helpers import get_corr_matrix
from .levy import ExampleLevy, ExpExampleLevy, LevySde
from .options import BestOf, ConstantShortRate, EuroCall, Rainbow
from .sde import Gbm, Heston, Merton
from .solvers import EulerSolver, HestonSolver, JumpEulerSolver


class Problem(ABC):
    def __init__(self, solver, discounter, payoff):
        self.solver = solver
        self.discounter = discounter
        self.payoff = payoff

    def dim(self):
        return self.solver.sde.dim

    def set_steps(self, steps):
        self.solver.num_steps = steps


def _spots(spot, dim):
    return spot if torch.is_tensor(spot) else torch.ones(dim) * spot


class BlackScholesEuroCall(Problem):
    """problem.py:23-33"""

    def __init__(self, r, sigma, spot, strike, maturity, steps, device):
        solver = EulerSolver(Gbm(r, sigma, torch.tensor([spot]), 1), maturity, steps, device)
        super().__init__(solver, ConstantShortRate(r), EuroCall(strike))

    @classmethod
    def default_params(cls, steps, device):
        return BlackScholesEuroCall(0.02, 0.3, 1, 1, 3, steps, device)


class BlackScholesRainbow(Problem):
    """problem.py:36-48"""

    def __init__(self, r, sigma, spot, strike, maturity, dim, corr_matrix, steps, device):
        solver = EulerSolver(Gbm(r, sigma, torch.ones(dim) * spot, dim, corr_matrix), maturity, steps, device)
        super().__init__(solver, ConstantShortRate(r), Rainbow(strike))

    @classmethod
    def default_params(cls, steps, device):
        return BlackScholesRainbow(0.02, 0.3, 1, 1, 3, 3, get_corr_matrix([0.7, 0.2, -0.3]), steps, device)


class HestonEuroCall(Problem):
    """problem.py:51-61"""

    def __init__(self, r, kappa, theta, xi, rho, spot, v0, strike, maturity, steps, device):
        solver = HestonSolver(Heston(r, kappa, theta, xi, rho, torch.tensor([spot, v0])), maturity, steps, device)
        super().__init__(solver, ConstantShortRate(r), EuroCall(strike))

    @classmethod
    def default_params(cls, steps, device):
        return HestonEuroCall(0.02, 0.25, 0.5, 0.3, -0.3, 1, 0.15, 1, 3, steps, device)


class MertonEuroCall(Problem):
    """problem.py:64-74"""

    def __init__(self, mu, sigma, rate, alpha, gamma, spot, strike, maturity, steps, device):
        solver = JumpEulerSolver(Merton(mu, sigma, rate, alpha, gamma, torch.tensor([spot]), 1), maturity, steps, device)
        super().__init__(solver, ConstantShortRate(mu), EuroCall(strike))

    @classmethod
    def default_params(cls, steps, device):
        return MertonEuroCall(0.02, 0.2, 1, -0.05, 0.3, 1, 1, 3, steps, device)


def _exp_levy_solver(c_minus, c_plus, alpha, mu, r, sigma, f, epsilon, dim, spot, maturity, steps, device, **kw):
    sde = LevySde(ExpExampleLevy(c_minus, c_plus, alpha, mu, r, sigma, f, epsilon, dim), _spots(spot, dim), device=device)
    return JumpEulerSolver(sde, maturity, steps, device=device, **kw)


class LevyRainbow(Problem):
    """problem.py:77-90"""

    def __init__(self, c_minus, c_plus, alpha, mu, r, sigma, f, epsilon, dim, spot, strike, maturity, steps, device):
        solver = _exp_levy_solver(c_minus, c_plus, alpha, mu, r, sigma, f, epsilon, dim, spot, maturity, steps, device)
        super().__init__(solver, ConstantShortRate(r), Rainbow(strike))

    @classmethod
    def default_params(cls, steps, device):
        return LevyRainbow(1, 1, 0.5, 2, 0.02, 0.3, 0.2, 0.001, 2, 1, 1, 3, steps, device)


class LevyRainbowMLMC(Problem):
    """problem.py:93-106 (exact_jumps=True)"""

    def __init__(self, c_minus, c_plus, alpha, mu, r, sigma, f, epsilon, dim, spot, strike, maturity, steps, device):
        solver = _exp_levy_solver(c_minus, c_plus, alpha, mu, r, sigma, f, epsilon, dim, spot, maturity, steps, device,
                                  exact_jumps=True)
        super().__init__(solver, ConstantShortRate(r), Rainbow(strike))

    @classmethod
    def default_params(cls, steps, device):
        return LevyRainbow(1, 1, 0.5, 2, 0.02, 0.3, 0.2, 0.001, 2, 1, 1, 3, steps, device)


class LevyCall(Problem):
    """problem.py:109-121: call on exp(log-price), payoff un-discounts the spot"""

    def __init__(self, c_minus, c_plus, alpha, mu, r, sigma, f, epsilon, spot, strike, maturity, steps, device):
        levy = ExampleLevy(c_minus, c_plus, alpha, mu, r, torch.tensor([sigma], device=device),
                           torch.tensor([f], device=device), torch.tensor([[1.]], device=device), epsilon, 1)
        solver = JumpEulerSolver(LevySde(levy, torch.tensor([spot]), device=device), maturity, steps, device=device)
        csr = ConstantShortRate(r)
        super().__init__(solver, csr, EuroCall(strike, log=True, discount=csr(-maturity)))

    @classmethod
    def default_params(cls, steps, device):
        return LevyCall(1, 1, 0.5, 2, 0.02, 0.2, 0.2, 0.001, 0, 1, 3, steps, device)


class LevyBestOf(Problem):
    """problem.py:124-137"""

    def __init__(self, c_minus, c_plus, alpha, mu, r, sigma, f, epsilon, dim, spot, strike, maturity, steps, device):
        solver = _exp_levy_solver(c_minus, c_plus, alpha, mu, r, sigma, f, epsilon, dim, spot, maturity, steps, device)
        super().__init__(solver, ConstantShortRate(r), BestOf(strike))

    @classmethod
    def default_params(cls, steps, device):
        return LevyBestOf(1, 1, 0.2, 2, 0.02, 0.3, 0.2, 0.001, 4, 1, 1, 3, steps, device)


class LevyCallOnMax(Problem):
    """problem.py:140-165"""

    def __init__(self, c_minus, c_plus, alpha, mu, r, sigma, f, chol_corr, epsilon, dim, spot, strike, maturity, steps,
                 device):
        levy = ExampleLevy(c_plus, c_minus, alpha, mu, r, sigma, f, chol_corr, epsilon, dim)
        solver = JumpEulerSolver(LevySde(levy, _spots(spot, dim), device=device), maturity, steps, device=device)
        csr = ConstantShortRate(r)
        super().__init__(solver, csr, Rainbow(strike, log=True, discount=csr(-maturity)))

    @classmethod
    def default_params(cls, dim, steps, device):
        presets = {
            2: ([0.4], [0.2, 0.2], [0.15, 0.15]),
            4: ([0.87, 0.94, 0.86, 0.87, 0.93, 0.96], [0.2, 0.15, 0.15, 0.1], [0.1, 0.1, 0.1, 0.1]),
        }
        if dim not in presets:
            return 'No default parameters for dimension {:}'.format(dim)
        rhos, fs, sigmas = presets[dim]
        chol = torch.linalg.cholesky(get_corr_matrix(rhos)).to(device)
        return LevyCallOnMax(1, 1, 0.5, 2, 0.02, torch.tensor(sigmas, device=device), torch.tensor(fs, device=device),
                             chol, 0.001, dim, 0, 1, 3, steps, device)
