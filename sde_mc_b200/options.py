"""Payoffs, discounting and closed-form prices (API of /root/reference/sde_mc/options.py).

The Option classes are callable on (bs, dim) tensors like the reference's, and each publishes `kernel_spec()` so
the fused kernels can evaluate it at the end of every path without the terminal states ever touching HBM.
The closed forms are host-side scalar math used as statistical oracles.
"""
import math
from abc import ABC, abstractmethod

import numpy as np
import torch
from scipy.integrate import quad
from scipy.stats import lognorm, norm

from . import _lib as L


# ---- closed forms ------------------------------------------------------------------------------------------
def bs_binary_aon(spot, strike, expiry, r, sigma):
    """Black-Scholes value of a binary asset-or-nothing option, as the reference defines it: Phi(d1) evaluated by
    quadrature of the standard normal density (options.py:8-32)."""
    d1 = (np.log(spot / strike) + (r + 0.5 * sigma * sigma) * expiry) / (sigma * np.sqrt(expiry))
    area, _ = quad(lambda z: np.exp(-0.5 * z * z), -np.inf, d1)
    return area / np.sqrt(2 * np.pi)


def bs_call(spot, strike, expiry, r, sigma):
    """Black-Scholes European call (options.py:35-58)."""
    vol = sigma * np.sqrt(expiry)
    d1 = (np.log(spot / strike) + (r + 0.5 * sigma ** 2) * expiry) / vol
    return spot * norm.cdf(d1) - strike * np.exp(-r * expiry) * norm.cdf(d1 - vol)


def merton_call(spot, strike, expiry, r, sigma, alpha, gamma, rate):
    """European call under Merton's jump-diffusion: 40-term Poisson mixture of Black-Scholes prices
    (options.py:61-99)."""
    jm = np.exp(alpha + 0.5 * gamma * gamma) - 1
    lam = (jm + 1) * rate * expiry
    price = 0.0
    for k in range(40):
        weight = np.exp(-lam) * lam ** k / math.factorial(k)
        r_k = r - rate * jm + k * np.log(jm + 1) / expiry
        sigma_k = np.sqrt(sigma ** 2 + k * gamma ** 2 / expiry)
        price += weight * bs_call(spot, strike, expiry, r_k, sigma_k)
    return price


def bs_digital_call(spot, strike, expiry, r, sigma):
    """Black-Scholes cash-or-nothing digital call (options.py:102-125)."""
    log_mean = np.log(spot) + (r - 0.5 * sigma * sigma) * expiry
    tail = 1 - lognorm.cdf(strike, s=sigma * np.sqrt(expiry), scale=np.exp(log_mean))
    return tail * np.exp(-r * expiry)


def bs_asian_call(spot, strike, expiry, r, sigma):
    """Geometric-average Asian call under Black-Scholes (options.py:128-153)."""
    sig_g = sigma / np.sqrt(3)
    b = 0.5 * (r - 0.5 * sig_g ** 2)
    vol = sig_g * np.sqrt(expiry)
    d1 = (np.log(spot / strike) + (b + 0.5 * sig_g ** 2) * expiry) / vol
    return spot * np.exp((b - r) * expiry) * norm.cdf(d1) - strike * np.exp(-r * expiry) * norm.cdf(d1 - vol)


# ---- payoffs -----------------------------------------------------------------------------------------------
class Option(ABC):
    """payoff(transform(x)) with transform(x) = discount * (exp(x) if log else x)  (options.py:156-176)."""

    def __init__(self, log=False, discount=1):
        self.log = log
        self.discount = discount

    def transform(self, x):
        return self.discount * (torch.exp(x) if self.log else x)

    @abstractmethod
    def payoff(self, x):
        pass

    def __call__(self, x):
        return self.payoff(self.transform(x))


class _StrikeOption(Option):
    KIND = None

    def __init__(self, strike, log=False, discount=1):
        super().__init__(log, discount)
        self.strike = strike

    def kernel_spec(self):
        """(payoff kind, strike, aux) for the fused kernels."""
        return self.KIND, float(self.strike), 1.0


class EuroCall(_StrikeOption):
    """max(S_0 - K, 0) on the first component (options.py:179-199)."""
    KIND = L.PAYOFF_EURO_CALL

    def payoff(self, x):
        s = x[:, 0]
        return torch.where(s > self.strike, s - self.strike, torch.zeros((), dtype=s.dtype, device=s.device))


class EuroPut(_StrikeOption):
    """max(K - S_0, 0) (options.py:202-222)."""
    KIND = L.PAYOFF_EURO_PUT

    def payoff(self, x):
        s = x[:, 0]
        return torch.where(s < self.strike, self.strike - s, torch.zeros((), dtype=s.dtype, device=s.device))


class BinaryAoN(_StrikeOption):
    """asset-or-nothing: S_0 if S_0 >= K (options.py:225-244)."""
    KIND = L.PAYOFF_BINARY_AON

    def payoff(self, x):
        s = x[:, 0]
        return torch.where(s >= self.strike, s, torch.zeros((), dtype=s.dtype, device=s.device))


class Basket(Option):
    """call on the arithmetic or geometric average of the components (options.py:247-259)."""

    def __init__(self, strike, average_type='arithmetic', log=False, discount=1):
        assert average_type in ['arithmetic', 'geometric']
        super().__init__(log, discount)
        self.strike = strike
        self.average_type = average_type

    def payoff(self, x):
        avg = x.mean(1) if self.average_type == 'arithmetic' else torch.log(x).mean(1).exp()
        return torch.where(avg > self.strike, avg - self.strike, torch.zeros((), dtype=avg.dtype, device=avg.device))

    def kernel_spec(self):
        kind = L.PAYOFF_BASKET_ARITH if self.average_type == 'arithmetic' else L.PAYOFF_BASKET_GEOM
        return kind, float(self.strike), 1.0


class Rainbow(_StrikeOption):
    """call on the maximum component (options.py:262-272)."""
    KIND = L.PAYOFF_RAINBOW

    def payoff(self, x):
        best = x.max(1).values
        return torch.where(best > self.strike, best - self.strike,
                           torch.zeros((), dtype=best.dtype, device=best.device))


class Digital(_StrikeOption):
    """1 if S_0 > K (options.py:275-284)."""
    KIND = L.PAYOFF_DIGITAL

    def payoff(self, x):
        s = x[:, 0]
        return (s > self.strike).to(s.dtype)


class AsianCall(Option):
    """call on the time average stored in component 1 of an AsianWrapper state (options.py:287-298)."""

    def __init__(self, time_interval, strike, log=False, discount=1):
        super().__init__(log, discount)
        self.time_interval = time_interval
        self.strike = strike

    def payoff(self, x):
        avg = x[:, 1] / self.time_interval
        if self.log:
            avg = torch.exp(avg)
        return torch.where(avg > self.strike, avg - self.strike, torch.zeros((), dtype=avg.dtype, device=avg.device))

    def kernel_spec(self):
        return L.PAYOFF_ASIAN_CALL, float(self.strike), float(self.time_interval)


class HestonRainbow(_StrikeOption):
    """call on the maximum over the even-indexed (price) components of stacked Heston states (options.py:301-311)."""
    KIND = L.PAYOFF_HESTON_RAINBOW

    def payoff(self, x):
        best = x[:, 0::2].max(1).values
        return torch.where(best > self.strike, best - self.strike,
                           torch.zeros((), dtype=best.dtype, device=best.device))


class BestOf(_StrikeOption):
    """max(max_i S_i, K) (options.py:314-321)."""
    KIND = L.PAYOFF_BEST_OF

    def payoff(self, x):
        best = torch.max(x, dim=1).values
        return torch.maximum(best, torch.full_like(best, self.strike))


class ConstantShortRate:
    """discount factor exp(-r t) for float or tensor t (options.py:324-337)."""

    def __init__(self, r):
        self.r = r

    def __call__(self, t):
        t = t if torch.is_tensor(t) else torch.tensor(t)
        return torch.exp(-t * self.r)
