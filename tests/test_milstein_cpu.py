"""Milstein scheme (EXTENSION: absent from the reference, SURVEY.md S3 -- parity unpinned).  The oracle's Milstein
step is validated here on the CPU by (1) an independent numpy restatement, (2) strong order 1 against the exact GBM
solution driven by the same Brownian path (Euler shows order 1/2 on the same paths)."""
import dataclasses

import numpy as np
import torch

from common import oracle, sm


def _gbm_spec(mu, sigma, scheme):
    return dataclasses.replace(sm.Gbm(mu, sigma, torch.tensor([1.0]), 1).kernel_spec(), scheme=scheme)


def _strong_errors(scheme, steps_list, n_paths=4000, mu=0.05, sigma=0.5, T=1.0, seed=3):
    rng = np.random.default_rng(seed)
    nmax = max(steps_list)
    z_fine = rng.standard_normal((n_paths, nmax, 1, 1))
    w_T = z_fine.sum(axis=1)[:, 0, 0] * np.sqrt(T / nmax)
    exact = np.exp((mu - 0.5 * sigma * sigma) * T + sigma * w_T)
    errs = []
    for n in steps_list:
        k = nmax // n
        z = z_fine.reshape(n_paths, n, k, 1, 1).sum(axis=2) / np.sqrt(k)   # unit normals of the coarser grid
        osde = oracle.sde_struct(_gbm_spec(mu, sigma, scheme), T, n)
        paths = oracle.diffusion(osde, z, dtype=np.float64)[0]
        errs.append(float(np.mean(np.abs(paths[:, -1, 0] - exact))))
    return errs


def test_oracle_milstein_matches_numpy_restatement():
    rng = np.random.default_rng(0)
    n, steps, mu, sigma, T = 64, 20, 0.03, 0.4, 2.0
    z = rng.standard_normal((n, steps, 1, 1))
    osde = oracle.sde_struct(_gbm_spec(mu, sigma, 2), T, steps)
    got = oracle.diffusion(osde, z, dtype=np.float64)[0][:, :, 0]
    h = T / steps
    x = np.ones(n)
    ref = [x.copy()]
    for k in range(steps):
        dw = z[:, k, 0, 0] * np.sqrt(h)
        x = x + mu * x * h + sigma * x * dw + 0.5 * sigma * sigma * x * (dw * dw - h)
        ref.append(x.copy())
    assert np.max(np.abs(got - np.stack(ref, axis=1))) < 1e-12


def test_oracle_milstein_strong_order_one_euler_half():
    steps = [16, 32, 64, 128]
    mil = _strong_errors(2, steps)
    eul = _strong_errors(0, steps)
    mil_ratios = [mil[i] / mil[i + 1] for i in range(len(steps) - 1)]
    eul_ratios = [eul[i] / eul[i + 1] for i in range(len(steps) - 1)]
    print("Milstein strong errors", mil, "ratios", mil_ratios)
    print("Euler    strong errors", eul, "ratios", eul_ratios)
    assert all(1.75 < r < 2.3 for r in mil_ratios)            # order 1: halving h halves the error
    assert all(1.25 < r < 1.6 for r in eul_ratios)            # order 1/2: factor sqrt(2)
    assert mil[-1] < 0.15 * eul[-1]


def test_milstein_tensor_step_matches_formula_and_arithmetic_is_euler():
    x = torch.tensor([[1.0, 2.0], [0.5, 1.5]])
    dw = torch.tensor([[0.1, -0.2], [0.05, 0.3]])
    h = 0.01

    class _S(sm.MilsteinScheme):
        def __init__(self, sde):
            self.sde = sde

    gbm = sm.Gbm(0.02, 0.3, torch.tensor([1.0, 1.0]), 2)
    got = _S(gbm).step(0.0, x, h, dw)
    want = x + 0.02 * x * h + 0.3 * x * dw + 0.5 * 0.3 * 0.3 * x * (dw * dw - h)
    assert torch.allclose(got, want, atol=1e-7)
    lg = sm.LogGbm(0.02, 0.2, torch.tensor([0.0]))
    x1, dw1 = x[:, :1], dw[:, :1]
    assert torch.allclose(_S(lg).step(0.0, x1, h, dw1), sm.EulerScheme.step(_S(lg), 0.0, x1, h, dw1))
