"""Small host-side utilities (API of /root/reference/sde_mc/helpers.py).  Nothing here is on the GPU hot path."""
import math

import numpy as np
import torch
from scipy.integrate import quad


def partition(interval, steps, ends='right', device='cpu'):
    """Uniform grid on [0, interval] with `steps` cells; `ends` chooses which endpoints are kept
    ('right' -> t_1..t_n, 'left' -> t_0..t_{n-1}, 'both', 'none').  helpers.py:6-33"""
    assert ends in ['right', 'left', 'both', 'none']
    first = 0 if ends in ('left', 'both') else 1
    last = steps if ends in ('right', 'both') else steps - 1
    return torch.tensor([interval * k / steps for k in range(first, last + 1)], device=device)


def solve_quadratic(coefficients):
    """Largest root of a x^2 + b x + c per batch element.  helpers.py:36-48"""
    a, b, c = coefficients
    root = torch.sqrt(b * b - 4 * a * c)
    return torch.maximum((-b + root) / (2 * a), (-b - root) / (2 * a))


def mc_estimates(run_sum, run_sum_squares, n):
    """(sample mean, unbiased sample variance) from running sums.  helpers.py:51-68"""
    mean = run_sum / n
    var = (run_sum_squares - run_sum * run_sum / n) / (n - 1)
    return mean, var


def remove_steps(time_tol, steps, time_interval):
    """Index of the last step left after trimming `time_tol` off the end of the interval.  helpers.py:71-74"""
    return int(np.floor(steps - time_tol / (time_interval / steps)))


def get_corr_matrix(rhos):
    """Correlation matrix from its strict upper triangle listed row by row.  helpers.py:77-100"""
    k = len(rhos)
    n = (1 + math.isqrt(1 + 8 * k)) // 2
    assert n * (n - 1) // 2 == k, "Length of correlation vector is not triangular"
    corr = torch.eye(n)
    iu = torch.triu_indices(row=n, col=n, offset=1)
    corr[iu[0], iu[1]] = torch.tensor(rhos)
    corr = corr + torch.triu(corr, 1).t()
    try:
        torch.linalg.cholesky(corr)
    except Exception:
        raise RuntimeError('Matrix is not positive semidefinite')
    return corr


def ceil_mult(x, n):
    """Smallest multiple of n that is >= x.  helpers.py:103-113"""
    return int(np.ceil(float(x) / n) * n)


def get_jump_comp(c_plus, c_minus, alpha, mu, f):
    """Exponential-moment compensator  int (e^{f x} - 1 - f x 1_{|x|<1}) nu(dx)  of the tempered-stable-like Levy
    measure (power law inside (-1, 1), exponential tails outside).  helpers.py:116-133"""
    def tail(x):
        return (np.exp(f * x) - 1) * np.exp(-mu * (abs(x) - 1))

    def core(x):
        return (np.exp(f * x) - 1 - f * x) * abs(x) ** (-alpha - 1)

    total = c_minus * quad(tail, -np.inf, -1)[0]
    total += c_minus * quad(core, -1, 0)[0]
    total += c_plus * quad(core, 0, 1)[0]
    total += c_plus * quad(tail, 1, np.inf)[0]
    return total


def sample_cov(x, y):
    """Unbiased sample covariance of two 1-D tensors.  helpers.py:136-137"""
    return ((x - x.mean()) * (y - y.mean())).sum() / (len(x) - 1)
