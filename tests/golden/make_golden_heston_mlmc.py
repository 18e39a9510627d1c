#!/usr/bin/env python
"""Golden fixture for the coupled fine/coarse pair of the Heston solver: the UNMODIFIED reference's
HestonSolver.multilevel_solve (DiffusionSolver.multilevel_solve /root/reference/sde_mc/solvers.py:90-119 with
HestonScheme.step schemes.py:16-22) on injected normals.

Run only in the build container (needs /root/reference):   python tests/golden/make_golden_heston_mlmc.py
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import make_golden as mg  # noqa: E402  (imports the reference, defines the injection adaptors)

ref, torch = mg.ref, mg.torch


def main():
    rng = np.random.default_rng(20261018)
    for fine, coarse in ((8, 2), (16, 8)):
        sde = ref.Heston(0.02, 0.25, 0.5, 0.3, -0.3, torch.tensor([1., 0.15]))
        solver = ref.HestonSolver(sde, 3, fine)
        bs = 32
        z = rng.standard_normal((bs, fine, 2, 1)).astype(np.float32)
        mg.inject_diffusion(solver, z)
        (pf, pc), _ = solver.multilevel_solve(bs, (fine, coarse))
        mg.save("mlmc_heston_%d_%d" % (fine, coarse), z=z, paths_fine=pf, paths_coarse=pc, r=0.02, kappa=0.25, theta=0.5,
                xi=0.3, rho=-0.3, x0=[1.0, 0.15], T=3.0)


if __name__ == "__main__":
    main()
