#!/bin/bash
# quick GPU validation: GPU tests, smoke, the default bench line (+ reference arm) and the GBM line
# usage (on a GPU box): tools/gpu_check.sh <tag>
tag=${1:-chk}; out=gpurun_out; mkdir -p $out
(time timeout 1500 python -m pytest tests -m gpu -q) > $out/pytest_$tag.log 2>&1; tail -15 $out/pytest_$tag.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 600 python bench.py > $out/bench_${tag}_merton.json 2> $out/bench_${tag}_merton.err; cut -c1-300 $out/bench_${tag}_merton.json
timeout 600 python bench.py --workload gbm --no-cpu-baseline > $out/bench_${tag}_gbm.json 2> $out/bench_${tag}_gbm.err; cut -c1-200 $out/bench_${tag}_gbm.json
