import sys, math
sys.path.insert(0, '/root/repo')
import torch, numpy as np
import sde_mc_b200 as sm
from sde_mc_b200.mlmc import _pair_terminals, _level_moments
sde = sm.Merton(0.02, 0.2, 1, -0.05, 0.3, torch.tensor([1.]), 1)
call, csr = sm.EuroCall(1.0), sm.ConstantShortRate(0.02)
df = math.exp(-0.06)
levels = [1, 2, 4, 8, 16, 32, 64, 128]
n = 4_000_000
for ex in (False, True):
    solver = sm.JumpEulerSolver(sde, 3, 1, device='cuda', exact_jumps=ex)
    print("exact_jumps", ex)
    prev_f = None
    for i, lv in enumerate(levels):
        m = _level_moments(solver, call, csr, n, lv, 0).read()
        single = m['sum'] / n; se = math.sqrt(max(m['sumsq']/n - single**2, 0)/n)
        line = "L=%3d single %.5f +- %.5f" % (lv, single, se)
        if i:
            term = _pair_terminals(solver, n, (lv, levels[i-1]), None)
            pf = call(term[:, 0]) * df; pc = call(term[:, 1]) * df
            line += "  pair fine %.5f coarse %.5f diff %.5f +- %.5f" % (pf.mean().item(), pc.mean().item(), (pf-pc).mean().item(), (pf-pc).std().item()/math.sqrt(n))
        print(line)
