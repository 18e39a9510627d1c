import math, sys, os
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
import numpy as np, torch
from common import oracle, oracle_sde, sm
from sde_mc_b200 import _engine as E, _lib as L, _spec
import test_gpu_fastpath as T
np.set_printoptions(linewidth=200, precision=7)
steps, n = 13, 1_000_000
def factory():
    s = sm.JumpEulerSolver(T._merton(1.0), 3, steps, device="cuda", seed=17)
    s.jump_strategy = L.JUMPS_QUEUE
    return s
pay, it, term = T._moments_per_path(factory(), sm.EuroCall(1.0), n)
s = factory()
paths, aux = s.solve(bs=n, low_storage=False)
normals, times, left, total, jumps = aux
P = paths.cpu().numpy()[:, :, 0]; ref_it = s.last_iters.cpu().numpy()
ref_term = P[np.arange(n), np.minimum(ref_it, P.shape[1] - 1)]
err = np.abs(term[:, 0] - ref_term) / np.maximum(np.abs(ref_term), 1e-3)
bad = np.nonzero(err > 1e-5)[0]
print("bad", bad)
J = jumps.cpu().numpy()[:, :, 0]; Tm = times.cpu().numpy()[:, :, 0]; N = normals.cpu().numpy()[:, :, 0]
for w in bad[:4]:
    sol = factory()
    K = 18 + sol.max_jumps; K = -(-K // 6) * 6
    z = T._draws(sol, L.DRAWS_BROWNIAN, int(w), 1, K, 1)[0].reshape(1, K, 1)
    jt, raw = T._draws(sol, L.DRAWS_QUEUE, int(w), 1, 36, 2)
    jt, raw = jt[:, :sol.max_jumps], raw[:, :sol.max_jumps]
    osde = oracle_sde(sol)
    ref = oracle.jump(osde, z, None, jt, T._marks_at_hits(osde, z, None, jt, raw))
    k = ref["iters"][0]
    print("path", w, "jump1d", term[w, 0], it[w], "store", ref_term[w], ref_it[w], "oracle", ref["paths"][0, k, 0], k)
    print(" queue times", jt[0, :8], "raw", raw[0, :8], "J", np.exp(-0.05 + 0.3 * raw[0, :8]) - 1)
    print(" oracle times", ref["times"][0, :k + 1]); print(" store  times", Tm[w, :ref_it[w] + 1])
    print(" oracle jumps", ref["jumps"][0, :k + 1, 0]); print(" store  jumps", J[w, :ref_it[w] + 1])
    print(" oracle paths", ref["paths"][0, :k + 1, 0]); print(" store  paths", P[w, :ref_it[w] + 1])
    print(" z", z[0, :k, 0]); print(" store dW/sqrt(dt)", N[w, :ref_it[w]] / np.sqrt(np.maximum(np.diff(Tm[w, :ref_it[w] + 1]), 1e-30)))
