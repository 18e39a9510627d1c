#!/bin/bash
# refresh of the path-storing evidence: full GPU suite, bench lines, launch list and ncu captures of both storing kernels
tag=${1:-rXX}
out=gpurun_out; mkdir -p $out
(time timeout 900 python -m pytest tests -m gpu -x -q) 2>&1 | tail -5 | tee $out/pytest_${tag}_store.log
for w in gbm_store merton_store; do
  timeout 600 python bench.py --workload $w --steps 20 > $out/bench_${tag}_$w.json 2> $out/bench_${tag}_$w.err
  python - <<PY
import json
d = json.load(open("$out/bench_${tag}_$w.json")); r = d["roofline"]
print("$w", "ms %.3f" % d["ms_per_step"], "kernel %.3f" % r["kernel_ms_per_launch"], "GB/s %.0f" % r["achieved"], "frac %.3f" % r["frac"], d["clocks"])
PY
done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/launches_${tag}_merton_store.csv python bench.py --workload merton_store --steps 2 --warmup 1 --no-cpu-baseline > $out/launches_${tag}_merton_store.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/launches_${tag}_gbm_store.csv python bench.py --workload gbm_store --steps 2 --warmup 1 --no-cpu-baseline > $out/launches_${tag}_gbm_store.log 2>&1
tools/prof_one.sh $tag gbm_store diffusion_store_tma_kernel 4e6 | grep -E "duration|issue_active|inst_executed.sum|dram__bytes|stall"
tools/prof_one.sh $tag merton_store jump_store_tma_kernel 2e6 | grep -E "duration|issue_active|inst_executed.sum|dram__bytes|stall"
