#!/bin/bash
# MLMC (C5) strong-scaling line on N GPUs of one box: tools/scale_mlmc.sh <tag> <N>
tag=$1; n=$2; out=gpurun_out; mkdir -p $out
f=$out/bench_${tag}_mlmc_strong_n${n}.json
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500 + n)) \
  bench.py --gpus $n --workload mlmc --scaling strong --no-cpu-baseline --steps 20 > $f 2> $f.err
python - <<PY
import json
d = json.load(open("$f"))
print("N=$n mlmc strong: %.3f ms/pass, e2e %.3f ms, mean %s" % (d["ms_per_step"], d["e2e"]["ms_per_step"], d["estimate"].get("mean_repr", d["estimate"]["mean"])))
PY
