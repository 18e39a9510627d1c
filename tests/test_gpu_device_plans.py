"""N2: the run-to-tolerance drivers as ONE submission -- pilot, trial count and main run without a host
read in between.  The count is computed on the device from the pilot's moments (sdemc_plan_mc / sdemc_plan_mlmc:
find_num_trials mc.py:418-427, get_optimal_trials mlmc.py:77-97) and the main kernels read their path range from
device memory (sdemc_range.d_range).  Checked against the host-sized two-call sequence of the reference API on
identically seeded solvers: same trial counts, same estimates (same global path ids => same paths)."""
import math

import numpy as np
import pytest
import torch

from common import sm
from sde_mc_b200 import _engine as E
from sde_mc_b200 import _lib as L
from sde_mc_b200 import mlmc as M

pytestmark = pytest.mark.gpu
DEV = "cuda"


class _P:
    def __init__(self, solver, payoff, discounter):
        self.solver, self.payoff, self.discounter = solver, payoff, discounter


def _problem(kind, seed=3):
    if kind == "merton":
        sde = sm.Merton(0.02, 0.2, 1, -0.05, 0.3, torch.tensor([1.]), 1)
        return _P(sm.JumpEulerSolver(sde, 3, 40, device=DEV, seed=seed), sm.EuroCall(1.0), sm.ConstantShortRate(0.02))
    if kind == "gbm":
        return _P(sm.EulerSolver(sm.Gbm(0.02, 0.3, torch.tensor([1.]), 1), 3, 30, device=DEV, seed=seed), sm.EuroPut(1.0),
                  sm.ConstantShortRate(0.02))
    levy = sm.ExpExampleLevy(1, 1, 0.5, 2, 0.02, 0.3, 0.2, 0.05, dim=2)
    return _P(sm.JumpEulerSolver(sm.LevySde(levy, torch.tensor([1., 1.])), 3, 16, device=DEV, seed=seed), sm.Rainbow(1.0),
              sm.ConstantShortRate(0.02))


@pytest.mark.parametrize("kind", ["merton", "gbm", "levy2d"])
def test_run_mc_device_sized_equals_host_sized(kind):
    eps, init = 2e-3, 20000
    dev_stats = sm.run_mc(_problem(kind), eps, bs=10 ** 5, init_trials=init)
    # the reference's sequence, host-sized: find_num_trials (a batched mc_simple pilot) then mc_simple(trials)
    p = _problem(kind)
    payoff_time = 'adapted' if p.solver.has_jumps else 'terminal'
    trials = sm.find_num_trials(p, eps, None, init, 10 ** 5)
    host_stats = sm.mc_simple(trials, p.solver, p.payoff, p.discounter, bs=10 ** 5, payoff_time=payoff_time)
    assert abs(dev_stats.num_trials - trials) <= 1                      # device fp64 vs numpy: the ceil may differ by one
    assert dev_stats.num_trials > init                                  # the main run is a real run
    if dev_stats.num_trials == trials:
        assert abs(dev_stats.sample_mean - host_stats.sample_mean) < 1e-12
        assert abs(dev_stats.sample_std - host_stats.sample_std) < 1e-12
    assert 1.96 * dev_stats.sample_std <= eps * 1.15                    # the tolerance was met (pilot variance: +-few %)


def test_pilot_plan_and_main_run_are_queued_without_a_host_sync():
    """torch's sync debug mode turns every synchronising call into an error: queueing the three stages must not
    trip it (the ONE read happens afterwards, in run_to_tolerance)"""
    p = _problem("merton")
    idx = L.INDEX_ADAPTED
    launch = lambda n, dev_range=None: E.run_moments(p.solver, p.payoff, p.discounter, n, idx, dev_range=dev_range)
    E.run_moments(p.solver, p.payoff, p.discounter, 1000, idx)     # warm up allocations / workspace
    torch.cuda.synchronize()
    torch.cuda.set_sync_debug_mode("error")
    try:
        main, plan = E.queue_to_tolerance(p.solver, 2e-3, 20000, launch)
    finally:
        torch.cuda.set_sync_debug_mode("default")
    n = int(plan.trials.item())
    mom = main.read()
    assert mom["n"] == n and n > 100000


def test_plan_mc_matches_the_host_formula_and_shards_the_range():
    lib = L.load()
    dev = torch.device(DEV, 0)
    rng = np.random.default_rng(5)
    for _ in range(20):
        n_pilot = int(rng.integers(1000, 10 ** 6))
        mean, var = float(rng.uniform(0.05, 2.0)), float(rng.uniform(1e-3, 3.0))
        total = mean * n_pilot
        total_sq = var * (n_pilot - 1) + total * total / n_pilot
        eps = float(rng.uniform(1e-4, 1e-2))
        mult = int(rng.choice([1, 1000]))
        pilot = torch.tensor([total, total_sq, 0, 0, 0, n_pilot, 0, 0], dtype=torch.float64, device=dev)
        m, se = E.mean_and_stderr(total, total_sq, n_pilot)
        want = int(np.ceil((se * 1.96 / eps) ** 2 * n_pilot))          # mc.py:424-427
        want = sm.ceil_mult(want, mult) if mult > 1 else want          # mc.py:459
        world = int(rng.integers(1, 9))
        base = int(rng.integers(0, 2 ** 40))
        got_lo, got_n = [], []
        for rank in range(world):
            plan = E.DeviceRange(dev)
            L.check(lib.sdemc_plan_mc(L.ptr(pilot), n_pilot, eps, mult, 0, base, rank, world, plan.row_ptr(),
                                      L.ptr(plan.trials), L.stream_ptr(dev)))
            lo, cnt = plan.ranges[0].tolist()
            assert abs(int(plan.trials.item()) - want) <= mult
            n_total = int(plan.trials.item())
            assert (lo - base, cnt) == E.shard(n_total, rank, world)
            got_lo.append(lo)
            got_n.append(cnt)
        assert sum(got_n) == n_total and got_lo[0] == base
        assert all(got_lo[i] + got_n[i] == got_lo[i + 1] for i in range(world - 1))    # contiguous, disjoint
    # cap
    plan = E.DeviceRange(dev)
    L.check(lib.sdemc_plan_mc(L.ptr(pilot), n_pilot, 1e-9, 1, 12345, 0, 0, 1, plan.row_ptr(), L.ptr(plan.trials),
                              L.stream_ptr(dev)))
    assert int(plan.trials.item()) == 12345


def test_run_mlmc_equals_get_optimal_trials_then_mc_multilevel():
    levels, eps, pilot = [1, 2, 4, 8, 16], 2e-3, 20000
    call, csr = sm.EuroCall(1.0), sm.ConstantShortRate(0.02)
    sde = sm.Merton(0.02, 0.2, 1, -0.05, 0.3, torch.tensor([1.]), 1)
    s1 = sm.JumpEulerSolver(sde, 3, 1, device=DEV, seed=7, exact_jumps=True)
    stats, trials = sm.run_mlmc(levels, eps, s1, call, csr, pilot_trials=pilot)
    s2 = sm.JumpEulerSolver(sde, 3, 1, device=DEV, seed=7, exact_jumps=True)
    want = sm.get_optimal_trials(pilot, levels, eps, s2, call, csr)
    assert all(abs(a - b) <= 1 for a, b in zip(trials, want)), (trials, want)
    ref = sm.mc_multilevel(want, levels, s2, call, csr)
    if trials == want:
        assert abs(stats.sample_mean - ref.sample_mean) < 1e-12 and abs(stats.sample_std - ref.sample_std) < 1e-12
    assert s1._next_path == s2._next_path or trials != want
    exact = sm.merton_call(1, 1, 3, 0.02, 0.2, -0.05, 0.3, 1)
    assert abs(stats.sample_mean - exact) < 1.96 * stats.sample_std + 1.5e-3      # + bias of the finest level (h = 3/16)
    assert 1.96 * stats.sample_std <= eps * 1.05


def test_mlmc_estimator_costs_one_allreduce_and_one_read(monkeypatch):
    """all levels live in one (levels, 8) tensor: ONE collective per estimator call, whatever the level count"""
    calls = []
    real = M.LevelMoments.all_reduce
    monkeypatch.setattr(M.LevelMoments, "all_reduce", lambda self: (calls.append(tuple(self.buf.shape)), real(self))[1])
    sde = sm.Merton(0.02, 0.2, 1, -0.05, 0.3, torch.tensor([1.]), 1)
    solver = sm.JumpEulerSolver(sde, 3, 1, device=DEV, exact_jumps=True)
    sm.mc_multilevel([4000, 2000, 1000, 500, 250, 120], [1, 2, 4, 8, 16, 32], solver, sm.EuroCall(1.0),
                     sm.ConstantShortRate(0.02))
    assert calls == [(6, 8)]


def test_run_cv_mc_device_sized_main_run():
    """run_cv_mc (mc.py:443-467) with the fused control-variate kernel: the pilot with the nets, ceil_mult(N, nn_bs) on
    the device and the main run are one submission; trials is a multiple of nn_bs and the tolerance is met"""
    torch.manual_seed(0)
    sde = sm.Merton(0.02, 0.2, 1, -0.05, 0.3, torch.tensor([1.]), 1)
    solver = sm.JumpEulerSolver(sde, 3, 60, device=DEV, seed=5)
    prob = _P(solver, sm.EuroCall(1.0), sm.ConstantShortRate(0.02))
    nets = [sm.Mlp(2, [50, 50, 50], 1, batch_norm=False, batch_norm_init=False, device=DEV) for _ in range(2)]
    opt = torch.optim.Adam([w for n in nets for w in n.parameters()], lr=1e-3)
    eps = 3e-3
    stats, train_time, test_time = sm.run_cv_mc(prob, nets, opt, eps, train_size=2000, step_factor=30, sim_bs=1e4,
                                                train_bs=500, nn_bs=1000, epochs=1, print_losses=False, init_trials=20000)
    assert stats.num_trials % 1000 == 0 and stats.num_trials > 20000
    assert 1.96 * stats.sample_std <= eps * 1.05
    exact = sm.merton_call(1, 1, 3, 0.02, 0.2, -0.05, 0.3, 1)
    assert abs(stats.sample_mean - exact) < 1.96 * stats.sample_std + 5e-4
