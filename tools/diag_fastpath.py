"""diagnostic (GPU box): worst per-path deviations of jump1d vs the storing kernel, and PACKED vs oracle"""
import math, sys, os
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
import numpy as np, torch
from common import oracle, oracle_sde, sm
from oracle import philox_streams as ps
from sde_mc_b200 import _engine as E, _lib as L, _spec
import test_gpu_fastpath as T

def jump1d_vs_store(steps, n, rate=1.0, qd=0, exact=False):
    def factory():
        s = sm.JumpEulerSolver(T._merton(rate), 3, steps, device="cuda", seed=17, exact_jumps=exact)
        s.jump_strategy, s.queue_depth = L.JUMPS_QUEUE, qd
        return s
    pay, it, term = T._moments_per_path(factory(), sm.EuroCall(1.0), n)
    s = factory()
    paths, aux = s.solve(bs=n, low_storage=False)
    normals, times, left, total, jumps = aux
    P = paths.cpu().numpy()[:, :, 0]; ref_it = s.last_iters.cpu().numpy()
    ref_term = P[np.arange(n), np.minimum(ref_it, P.shape[1] - 1)]
    err = np.abs(term[:, 0] - ref_term) / np.maximum(np.abs(ref_term), 1e-3)
    worst = np.argsort(-err)[:8]
    J = jumps.cpu().numpy()[:, :, 0]; Tm = times.cpu().numpy()[:, :, 0]
    print("steps", steps, "n", n, "paths with err>1e-5:", int((err > 1e-5).sum()), "iters differ:", int((it != ref_it).sum()))
    for w in worst:
        nj = int((J[w] != 0).sum())
        print(" path", w, "err %.3e" % err[w], "it", it[w], ref_it[w], "term", term[w, 0], ref_term[w], "P[-1]", P[w, -1], "njumps", nj,
              "jump iters", np.nonzero(J[w])[0][:12], "t_end", Tm[w, ref_it[w]])

def packed_vs_oracle(steps, exact):
    n, seed, lo = 8192, 37, 4242
    outs = {}
    for short in (L.SHORT_PACKED, L.SHORT_PACKED_GENERIC):
        solver = sm.JumpEulerSolver(T._merton(), 3, steps, device="cuda", seed=seed, exact_jumps=exact)
        solver.short_path, solver._next_path = short, lo
        outs[short] = T._moments_per_path(solver, sm.EuroCall(1.0), n)
    K = int(max(o[1].max() for o in outs.values())) + 2; K += K & 1
    z, gap, raw = T._draws(solver, L.DRAWS_PACKED, lo, n, K, 3)
    jt, marks, iters = ps.candidate_jumps(3.0 / steps, 3.0, 1.0, gap, raw, solver.max_jumps)
    ref = oracle.jump(oracle_sde(solver), z.reshape(n, K, 1), None, jt, marks)
    ref_term = ref["paths"][np.arange(n), ref["iters"], 0]
    for short, (pay, it, term) in outs.items():
        err = np.abs(term[:, 0] - ref_term) / np.maximum(np.abs(ref_term), 1e-3)
        worst = np.argsort(-err)[:5]
        print("packed short", short, "steps", steps, "exact", exact, "iters differ", int((it != ref["iters"]).sum()), "err>1e-5:", int((err > 1e-5).sum()))
        for w in worst:
            k = ref["iters"][w]
            print("  path", w, "err %.3e" % err[w], "it", it[w], k, "term", term[w, 0], ref_term[w], "times", ref["times"][w, :k + 1], "z", z[w, :k], "marks", marks[w, :k])

jump1d_vs_store(100, 1_000_000)
jump1d_vs_store(13, 1_000_000)
jump1d_vs_store(100, 500_000, 3.0, 4)
packed_vs_oracle(1, False)
packed_vs_oracle(4, False)
