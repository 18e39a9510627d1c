"""GPU smoke tests of the drop-in API, mirroring what the reference's own suite exercises (tests/test_solvers.py,
tests/test_mc.py, tests/test_varred.py of piers-hinds/sde_mc): shapes, NaN-freeness, and that every estimator /
training pipeline runs end to end on top of the kernels -- with device='cuda' and with the reference's default
device='cpu' (results copied back)."""
import pytest
import torch

from common import sm

pytestmark = pytest.mark.gpu


@pytest.fixture(params=["cuda", "cpu"])
def device(request):
    return request.param


def _solvers(device):
    gbm2 = sm.Gbm(0.02, 0.2, torch.tensor([1., 2.]), dim=2)
    heston = sm.Heston(0.02, 0.5, 0.1, 0.15, -0.5, torch.tensor([1., 2.]))
    merton1 = sm.Merton(0.02, 0.3, 2, -0.05, 0.3, torch.tensor([1.]), dim=1)
    merton2 = sm.Merton(0.02, 0.3, 2, -0.05, 0.3, torch.tensor([1., 1.]), dim=2)
    return (sm.EulerSolver(gbm2, 3, 10, device=device), sm.HestonSolver(heston, 3, 10, device=device),
            sm.JumpEulerSolver(merton1, 3, 10, device=device), sm.JumpEulerSolver(merton2, 3, 10, device=device))


def test_solver_shapes(device):
    gbm, heston, merton1, merton2 = _solvers(device)
    paths, normals = gbm.solve(bs=4, return_normals=True)                    # tests/test_solvers.py:5-9
    assert paths.shape == (4, 11, 2) and normals.shape == (4, 10, 2) and not torch.isnan(paths).any()
    assert paths.device.type == device
    paths, normals = heston.solve(bs=8, return_normals=True)                 # :12-16
    assert paths.shape == (8, 11, 2) and normals.shape == (8, 10, 2) and not torch.isnan(paths).any()
    assert (paths[:, :, 1] >= 0).all()
    for solver in (merton1, merton2):                                        # :19-25
        paths, (normals, time_paths, left_paths, total_steps, jump_paths) = solver.solve(8)
        d = solver.sde.dim
        S = solver.num_steps + solver.max_jumps
        assert paths.shape == (8, total_steps + 1, d) and normals.shape == (8, S, d)
        assert time_paths.shape == (8, S + 1, 1) and left_paths.shape == (8, S + 1, d) == jump_paths.shape
        for a in (paths, normals, time_paths, left_paths, jump_paths):
            assert not torch.isnan(a).any()
        assert torch.all(time_paths[:, total_steps, 0] >= 3.0 - 1e-5)
    paths, aux = merton1.solve(bs=5, low_storage=True)
    assert aux[0] is None and paths.shape[0] == 5


def test_mc_flows(device):
    gbm, heston, merton1, merton2 = _solvers(device)
    call, csr = sm.EuroCall(1), sm.ConstantShortRate(0.02)
    st = sm.mc_simple(100, gbm, call, csr)                                    # tests/test_mc.py:5-11
    assert st.time_elapsed >= 0 and st.sample_mean >= 0 and st.sample_std > 0
    assert not st.payoffs.isnan().any() and st.paths.shape == (100, 11, 2)
    gbm1 = sm.EulerSolver(sm.Gbm(0.02, 0.2, torch.tensor([1.]), dim=1), 3, 10, device=device)
    st = sm.mc_simple(100, gbm1, call, csr, bs=17)                             # :14-18
    assert st.sample_mean >= 0 and st.sample_std > 0 and st.paths is None
    assert sm.mc_terminal_cv(100, gbm1, call, csr, 10).sample_std > 0          # :58-59
    assert sm.mc_terminal_cv(100, gbm1, call, csr).payoffs.shape == (100,)
    dl = sm.simulate_data(16, gbm, call, csr, bs=16)                           # :39-40
    (paths, normals), payoffs = next(iter(dl))
    assert paths.shape == (16, 10, 2) and normals.shape == (16, 10, 2) and payoffs.shape == (16,)
    sm.simulate_data(16, merton1, call, csr, bs=16)                            # :43-44
    dl = sm.simulate_adapted_data(16, merton1, call, csr, bs=8)                # :47-48
    (paths, normals, left, times, jumps), payoffs = next(iter(dl))
    assert paths.shape == normals.shape == left.shape == jumps.shape and times.shape[-1] == 1
    hp = sm.HestonEuroCall.default_params(100, device)
    st = sm.run_mc_terminal_cv(hp, 0.01, bs=1e3, init_trials=1e4)              # :61-62
    assert st.sample_std * 1.96 < 0.02
    st = sm.run_mc(sm.MertonEuroCall.default_params(50, device), 0.005, bs=1e4, init_trials=1e4)
    assert st.sample_std * 1.96 < 0.006


def test_control_variate_pipelines(device):
    gbm1 = sm.EulerSolver(sm.Gbm(0.02, 0.2, torch.tensor([1.]), dim=1), 3, 10, device=device)
    merton1 = sm.JumpEulerSolver(sm.Merton(0.02, 0.3, 2, -0.05, 0.3, torch.tensor([1.]), dim=1), 3, 10, device=device)
    call, csr = sm.EuroCall(1), sm.ConstantShortRate(0.02)
    bcv = sm.Mlp(2, [5, 5], 1, activation=sm.nn.ReLU, device=device)          # BatchNorm nets -> PyTorch application
    jcv = sm.Mlp(2, [5, 5], 1, activation=sm.nn.ReLU, device=device)
    adam = torch.optim.Adam(bcv.parameters())
    st = sm.mc_control_variates(bcv, adam, gbm1, (100, 100), (10, 20), call, csr, sim_bs=(100, 100), bs=(10, 10),
                                print_losses=False)                            # tests/test_mc.py:21-24
    assert st.sample_std > 0
    adam = torch.optim.Adam(list(bcv.parameters()) + list(jcv.parameters()))
    st = sm.mc_adaptive_cv([bcv, jcv], adam, merton1, (100, 100), (10, 20), call, csr, sim_bs=(100, 100),
                           bs=(10, 10), print_losses=False)                    # :51-55
    assert st.sample_std > 0
    # the experiments' BN-free nets take the fused tcgen05 kernel, trained here by the PyTorch pipeline first
    f = sm.Mlp(2, [50, 50, 50], 1, batch_norm=False, batch_norm_init=False, device=device)
    g = sm.Mlp(2, [50, 50, 50], 1, batch_norm=False, batch_norm_init=False, device=device)
    adam = torch.optim.Adam(list(f.parameters()) + list(g.parameters()))
    merton1.num_steps = 20
    dl = sm.simulate_adapted_data(2000, merton1, call, csr, bs=200)
    losses = sm.train_adapted_control_variates([f, g], adam, dl, merton1, csr, 3, False)
    assert len(losses) == 3 and losses[-1] < losses[0] * 1.5
    merton1.num_steps = 100
    assert sm.fused_cv_supported([f, g], merton1)
    plain = sm.mc_simple(200000, merton1, call, csr, bs=10 ** 5, payoff_time='adapted')
    cv = sm.mc_apply_cvs([f, g], merton1, 200000, call, csr, sim_bs=10 ** 5, bs=2000)
    assert abs(cv.sample_mean - plain.sample_mean) < 4 * (cv.sample_std + plain.sample_std)
    assert cv.sample_std < plain.sample_std
