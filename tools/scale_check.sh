#!/bin/bash
# multi-GPU health check of the final library: default bench (weak + strong) and the storing workloads on N GPUs
n=${1:-2}; out=gpurun_out; mkdir -p $out
for args in "--scaling weak" "--scaling strong" "--workload merton_store --steps 10" "--workload gbm_store --steps 10"; do
  tag=$(echo $args | tr -d '-' | tr ' ' '_')
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29700 + n)) \
    bench.py --gpus $n --no-cpu-baseline $args > $out/scalecheck_n${n}_$tag.json 2> $out/scalecheck_n${n}_$tag.err
  python - <<PY
import json
try:
    d = json.load(open("$out/scalecheck_n${n}_$tag.json"))
    print("N=$n $args: %.4g %s, %.3f ms/step, frac %.3f, mean %s" % (d["value"], d["unit"], d["ms_per_step"], d["roofline"]["frac"], d["estimate"].get("mean_repr", d["estimate"].get("mean_terminal_state"))))
except Exception as e:
    print("N=$n $args FAILED", e); print(open("$out/scalecheck_n${n}_$tag.err").read()[-800:])
PY
done
