#!/usr/bin/env python
"""Golden fixture for the `tol > 0` trimming of the Brownian control-variate sum (integrate_cv
/root/reference/sde_mc/varred.py:202-209, remove_steps helpers.py:71-74): the UNMODIFIED reference applied, with
tol > 0, to the nets and the injected noise of the committed fixtures cv_gbm_1d.npz / cv_merton_1d.npz.

Run only in the build container (needs /root/reference):   python tests/golden/make_golden_cv_tol.py
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import make_golden as mg  # noqa: E402  (imports the reference, defines the injection adaptors)

ref, torch = mg.ref, mg.torch


def net_from(g, prefix):
    net = ref.Mlp(2, [50, 50, 50], 1, batch_norm=False, batch_norm_init=False)
    lin = [m for m in net.net if isinstance(m, torch.nn.Linear)]
    with torch.no_grad():
        for i, l in enumerate(lin):
            l.weight.copy_(torch.as_tensor(g["%s_w%d" % (prefix, i)]))
            l.bias.copy_(torch.as_tensor(g["%s_b%d" % (prefix, i)]))
    return net.eval()


def main():
    csr, call = ref.ConstantShortRate(0.02), ref.EuroCall(1.0)
    out = {}
    # diffusion: steps = num_steps = 16, tol = 0.5 keeps floor(16 - 0.5 / (3 / 16)) = 13 steps
    g = np.load(os.path.join(mg.HERE, "cv_gbm_1d.npz"))
    fd = net_from(g, "f")
    solver = ref.EulerSolver(ref.Gbm(0.02, 0.3, torch.tensor([1.]), 1), 3, 16)
    mg.inject_diffusion(solver, g["z"])
    bs = g["z"].shape[0]
    dl = ref.simulate_data(bs, solver, call, csr, bs=1, inference=True)
    for tol in (0.5, 1.0):
        gam = [float(ref.apply_diffusion_control_variate(fd, mg._ShapeProxy(dl.dataset, i), solver, csr, tol=tol)[0])
               for i in range(bs)]
        out["gbm_tol%g" % tol] = np.array(gam, np.float32)
        out["gbm_keep%g" % tol] = ref.remove_steps(tol, 16, 3)
    # jump-adapted: steps = the batch's total_steps (mc.py:394-396 trims the stored arrays to it)
    g = np.load(os.path.join(mg.HERE, "cv_merton_1d.npz"))
    f, gn = net_from(g, "f"), net_from(g, "g")
    solver = ref.JumpEulerSolver(ref.Merton(0.02, 0.2, 1, -0.05, 0.3, torch.tensor([1.]), 1), 3, 20)
    mg.JumpInjector(solver, g["z"], None, g["jump_times"], g["marks"], "lognormal")
    bs = g["z"].shape[0]
    dl = ref.simulate_adapted_data(bs, solver, call, csr, bs=1, inference=True)
    assert dl.dataset.total_steps == int(g["total_steps"])
    with torch.inference_mode():
        for tol in (0.5,):
            gam = [float(ref.apply_adapted_control_variates([f, gn], mg._ShapeProxy(dl.dataset, i), solver, csr, tol=tol)[0])
                   for i in range(bs)]
            out["merton_tol%g" % tol] = np.array(gam, np.float32)
            out["merton_keep%g" % tol] = ref.remove_steps(tol, int(g["total_steps"]), 3)
    mg.save("cv_tol", **out)


if __name__ == "__main__":
    main()
