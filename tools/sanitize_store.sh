#!/bin/bash
# compute-sanitizer over the path-storing kernels (TMA and LSU data paths) at small sizes
out=gpurun_out; mkdir -p $out
cat > /tmp/san_store.py <<'PY'
import sys, torch
sys.path.insert(0, '.')
import sde_mc_b200 as sm
dev = 'cuda'
def run(solver, bs, **kw):
    for align, tma in ((32, True), (32, False), (1, True)):
        solver.row_align, solver.tma_store = align, tma
        out = solver.solve(bs=bs, **kw)
    torch.cuda.synchronize()
m1 = sm.Merton(0.02, 0.2, 1, -0.05, 0.3, torch.tensor([1.0]), 1)
for steps in (1, 24, 100):
    run(sm.JumpEulerSolver(m1, 3.0, steps, device=dev, seed=3), 333)
run(sm.JumpEulerSolver(m1, 3.0, 100, device=dev, seed=3), 97, low_storage=True)
levy = sm.ExpExampleLevy(1, 1, 0.5, 2, 0.02, 0.3, 0.2, 0.05, dim=2)
run(sm.JumpEulerSolver(sm.LevySde(levy, torch.tensor([1., 1.])), 1.0, 32, device=dev, seed=13), 130)
run(sm.JumpEulerSolver(sm.Merton(0.02, 0.2, 2, -0.05, 0.3, torch.ones(3), 3), 1.0, 30, device=dev, seed=12), 65)
for steps in (5, 37, 252):
    run(sm.EulerSolver(sm.Gbm(0.02, 0.3, torch.tensor([1.0]), 1), 3.0, steps, device=dev, seed=7), 257)
run(sm.EulerSolver(sm.Gbm(0.02, 0.3, torch.ones(3), 3, sm.get_corr_matrix([0.3, -0.2, 0.5])), 3.0, 50, device=dev, seed=5), 100)
run(sm.HestonSolver(sm.Heston(0.02, 2.0, 0.04, 0.2, -0.7, torch.tensor([1.0, 0.04])), 3.0, 64, device=dev), 100)
print("done")
PY
for tool in memcheck racecheck; do
  timeout 900 compute-sanitizer --tool $tool --print-limit 20 python /tmp/san_store.py > $out/sanitizer_$tool.log 2>&1
  echo "== $tool: rc=$?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|done|Error|hazard" $out/sanitizer_$tool.log | head -12
done
