"""Run-to-tolerance drivers (mc.py:418-467 of the reference) on CPU: the estimators are replaced by stubs, so what is
pinned is the pilot -> trial-count rule, the payoff index the drivers pick and what they pass on.  No kernel runs."""
import math

from common import sm  # noqa: F401  (puts the repo on sys.path)
from sde_mc_b200 import mc as MC


class _Solver:
    def __init__(self, has_jumps):
        self.has_jumps = has_jumps
        self.num_steps = 10


class _Problem:
    def __init__(self, has_jumps):
        self.solver = _Solver(has_jumps)
        self.payoff = "payoff"
        self.discounter = "discounter"


def _stub_mc_simple(calls, std):
    def mc_simple(num_trials, solver, payoff, discounter, bs=None, return_normals=False, payoff_time='terminal'):
        calls.append(dict(n=num_trials, bs=bs, payoff_time=payoff_time))
        return MC.MCStatistics(0.25, std, 0.01, int(num_trials))
    return mc_simple


def test_find_num_trials_rule_and_payoff_index(monkeypatch):
    for has_jumps, want_index in ((True, 'adapted'), (False, 'terminal')):
        calls = []
        monkeypatch.setattr(MC, "mc_simple", _stub_mc_simple(calls, std=2e-3))
        n = MC.find_num_trials(_Problem(has_jumps), eps=1e-3, init_trials=1e5, bs=1e4)
        # mc.py:418-427: ceil((std * 1.96 / eps)^2 * init_trials)
        assert n == math.ceil((2e-3 * 1.96 / 1e-3) ** 2 * 1e5)
        assert calls == [dict(n=1e5, bs=1e4, payoff_time=want_index)]


def test_run_mc_pilot_then_run(monkeypatch):
    calls = []
    monkeypatch.setattr(MC, "mc_simple", _stub_mc_simple(calls, std=1e-3))
    st = MC.run_mc(_Problem(True), eps=5e-4, bs=2e4, init_trials=5e4)
    want = math.ceil((1e-3 * 1.96 / 5e-4) ** 2 * 5e4)
    assert [c["n"] for c in calls] == [5e4, want]
    assert all(c["payoff_time"] == 'adapted' and c["bs"] == 2e4 for c in calls)
    assert st.num_trials == want


def test_terminal_cv_driver(monkeypatch):
    calls = []

    def mc_terminal_cv(num_trials, solver, payoff, discounter, bs=None):
        calls.append((num_trials, bs))
        return MC.MCStatistics(0.25, 4e-4, 0.01, int(num_trials))

    monkeypatch.setattr(MC, "mc_terminal_cv", mc_terminal_cv)
    MC.run_mc_terminal_cv(_Problem(False), eps=1e-4, bs=1e5, init_trials=1e5)
    assert calls == [(1e5, 1e5), (math.ceil((4e-4 * 1.96 / 1e-4) ** 2 * 1e5), 1e5)]


def test_mcstatistics_prints_the_95_percent_interval():
    # the reference's __str__ (mc.py:45-50) on the same numbers, evaluated in the build container
    s = str(MC.MCStatistics(0.2630, 1e-4, 1.5, 10 ** 6))
    assert s == 'Mean: 0.263000  +/- 0.000196    Time taken (s): 1.50    N: 1.00E+06'
