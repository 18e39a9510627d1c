"""Pins oracle/torch_port.py (the CPU 'port' that bench.py times as the reference arm) against seeded runs of the
unmodified reference using torch's own RNG: same seed => same draws in the same order => same trajectories.
CPU only.  Tolerance 1e-6 relative: coefficient folding (e.g. mu - rate*E[J] in double, then fp32) differs from the
reference's op-by-op fp32 evaluation by an ulp; any mismatch in RNG consumption would show up as O(1)."""
import numpy as np
import torch

from common import golden, rel_err, sm
from oracle import torch_port as tp


def test_gbm_seeded_matches_reference():
    g = golden("seeded")
    spec = sm.Gbm(0.02, 0.3, torch.tensor([1.]), 1).kernel_spec()
    torch.manual_seed(5)
    paths, normals = tp.diffusion_solve(spec, 3, 16, 64)
    assert rel_err(paths.numpy(), g["gbm_paths"]) < 1e-6
    assert rel_err(normals.numpy(), g["gbm_normals"], 1e-2) < 1e-6


def test_merton_seeded_matches_reference():
    g = golden("seeded")
    spec = sm.Merton(0.02, 0.2, 1, -0.05, 0.3, torch.tensor([1.]), 1).kernel_spec()
    assert tp.max_jumps(3, 1.0) == 33
    torch.manual_seed(5)
    paths, aux = tp.jump_solve(spec, 3, 20, 64, low_storage=False)
    assert aux[3] == int(g["merton_total_steps"])
    assert rel_err(paths.numpy(), g["merton_paths"]) < 1e-6
    assert rel_err(aux[1][:, :aux[3] + 1, 0].numpy(), g["merton_times"]) < 1e-6
    torch.manual_seed(5)
    paths, _ = tp.jump_solve(spec, 3, 20, 64, low_storage=True)
    assert rel_err(paths[:, -1].numpy(), g["merton_low_last"]) < 1e-6


def test_levy_seeded_matches_reference():
    g = golden("seeded")
    levy = sm.ExpExampleLevy(1, 1, 0.5, 2, 0.02, 0.3, 0.2, 0.01, dim=2)
    spec = sm.LevySde(levy, torch.tensor([1., 1.])).kernel_spec()
    torch.manual_seed(5)
    paths, _ = tp.jump_solve(spec, 3, 8, 32, low_storage=True)
    assert rel_err(paths[:, -1].numpy(), g["levy_low_last"]) < 5e-6


def test_mc_simple_port_runs_and_prices():
    spec = sm.Merton(0.02, 0.2, 1, -0.05, 0.3, torch.tensor([1.]), 1).kernel_spec()
    torch.manual_seed(1)
    mean, sd, secs = tp.mc_simple_batched(spec, 3, 50, 20000, 10000, tp.payoff_call_on("euro_call", 1.0), 0.02, True)
    assert abs(mean - 0.26298) < 4 * sd + 2e-3 and secs > 0
