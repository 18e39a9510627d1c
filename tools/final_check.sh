#!/bin/bash
# end-of-round validation on a GPU box: GPU tests, smoke, the default bench line (+ reference arm), CV and MLMC lines
out=gpurun_out; mkdir -p $out
(time timeout 900 python -m pytest tests -m gpu -x -q) 2>&1 | tail -5
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 600 python bench.py > $out/bench_final_merton.json 2> $out/bench_final_merton.err; cut -c1-300 $out/bench_final_merton.json; echo
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > $out/bench_final_reference.json 2>/dev/null; cut -c1-300 $out/bench_final_reference.json; echo
for w in merton_cv mlmc levy2d gbm; do
  timeout 600 python bench.py --workload $w > $out/bench_final_$w.json 2> $out/bench_final_$w.err
  python - <<PY
import json
try:
    d = json.load(open("$out/bench_final_$w.json"))
    print("$w", "%.4g" % d["value"], "ms %.3f" % d["ms_per_step"], "e2e %.4g" % d["e2e"]["value"], "frac %.3f" % d["roofline"]["frac"], d["clocks"])
except Exception as e:
    print("$w FAILED", e)
PY
done
