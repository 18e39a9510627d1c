#!/bin/bash
# Builder scaling lines on one multi-GPU box: tools/scale_run.sh <tag> <Ns...>   (outputs gpurun_out/bench_<tag>_<workload>_<scaling>_n<N>.json)
tag=$1; shift
out=gpurun_out; mkdir -p $out
run() { # N workload scaling extra...
  local n=$1 w=$2 sc=$3; shift 3
  local f=$out/bench_${tag}_${w}_${sc}_n${n}.json
  if [ "$n" = 1 ]; then
    timeout 600 python bench.py --gpus 1 --workload $w --scaling $sc --no-cpu-baseline "$@" > $f 2> $f.err
  else
    timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500 + n)) \
      bench.py --gpus $n --workload $w --scaling $sc --no-cpu-baseline "$@" > $f 2> $f.err
  fi
  python - <<PY
import json
try:
    d = json.load(open("$f"))
    est = d.get("estimate", {})
    print("N=$n $w $sc: %.4g path-steps/s, %.3f ms/step, e2e %.4g, mean %s" % (d["value"], d["ms_per_step"], d["e2e"]["value"], est.get("mean_repr", est.get("mean"))))
except Exception as e:
    print("N=$n $w $sc FAILED", e); print(open("$f.err").read()[-600:])
PY
}
for n in "$@"; do
  run $n merton weak
  run $n merton strong
  run $n gbm strong
  run $n mlmc strong
done
