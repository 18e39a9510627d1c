"""Host logic of the MLMC estimators (mlmc.py:7-101 of the reference) on CPU: the kernels are replaced by fake
per-level moments, so what is pinned here is the allocation formula, the telescoping sum and its standard error.
No kernel runs."""
import math

import torch

from common import sm  # noqa: F401  (puts the repo on sys.path)
from sde_mc_b200 import mlmc as M


class _FakeLevels:
    """stands in for mlmc.LevelMoments: read() -> one moments dict per level"""

    def __init__(self, rows):
        self._rows = rows

    def read(self):
        return self._rows


def _FakeMoments(total, total_sq, n):
    return {"sum": total, "sumsq": total_sq, "n": float(n), "iters": 0.0}


class _FakeSolver:
    time_interval = 3.0
    num_steps = 7


def _moments_with(mean, var, n):
    # sums that give exactly this sample mean and unbiased variance under helpers.mc_estimates (helpers.py:51-68)
    total = mean * n
    total_sq = var * (n - 1) + total * total / n
    return _FakeMoments(total, total_sq, n)


def test_get_optimal_trials_is_the_reference_allocation(monkeypatch):
    levels = [1, 2, 4, 8]
    variances = [2.97e-1, 7.8e-3, 6.0e-3, 3.8e-3]          # level variances of C5
    pilot = 10 ** 5
    monkeypatch.setattr(M, "_all_levels", lambda solver, payoff, disc, trials, lv: _FakeLevels([
        _moments_with(0.1, v, n) for v, n in zip(variances, trials)]))
    solver = _FakeSolver()
    eps = 1e-3
    got = M.get_optimal_trials(pilot, levels, eps, solver, None, None)
    # mlmc.py:92-96: N_l = ceil(1.96^2 / eps^2 * sqrt(V_l h_l) * sum_k sqrt(V_k / h_k)),  h_l = T / levels[l]
    h = [3.0 / lv for lv in levels]
    total = sum(math.sqrt(v / hl) for v, hl in zip(variances, h))
    want = [math.ceil(1.96 ** 2 / eps ** 2 * math.sqrt(v * hl) * total) for v, hl in zip(variances, h)]
    assert all(abs(g - w) <= 1 for g, w in zip(got, want)), (got, want)   # fp64 tensor vs python float rounding
    assert solver.num_steps == levels[0]                                  # the reference leaves the solver there (:83)


def test_mc_multilevel_telescopes_means_and_adds_variances(monkeypatch):
    levels = [1, 2, 4]
    trials = [1000, 400, 100]
    means = [0.25, 0.01, 0.003]
    variances = [0.3, 0.008, 0.006]
    monkeypatch.setattr(M, "_all_levels", lambda solver, payoff, disc, tr, lv: _FakeLevels([
        _moments_with(m, v, n) for m, v, n in zip(means, variances, tr)]))
    st = M.mc_multilevel(trials, levels, _FakeSolver(), None, None)
    assert abs(st.sample_mean - sum(means)) < 1e-12
    assert abs(st.sample_std - math.sqrt(sum(v / n for v, n in zip(variances, trials)))) < 1e-12
    assert st.num_trials == trials[-1]


def test_mlmc_bs_from_trials_keeps_the_reference_memory_budget():
    levels = [1, 2, 4, 128]
    trials = torch.tensor([10 ** 9, 10 ** 8, 10 ** 7, 10 ** 3])
    bs = M.mlmc_bs_from_trials(trials, levels, max_mem=5 * 10 ** 8, dim=1, max_jumps=33)
    # values of the unmodified reference (mlmc.py:100-101, evaluated in this container): the division runs in fp32
    # tensors, so 5e8 / 34 = 14705882.35 rounds to 14705882 before the ceil
    assert [int(b) for b in bs] == [14705882, 14285714, 10000000, 1000]
