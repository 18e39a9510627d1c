#!/bin/bash
# MLMC strong scaling, captured graph vs queued launches: tools/mlmc_scale_ab.sh <N>
n=$1
for g in 1 0; do
  SDEMC_MLMC_GRAPH=$g timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29600 + n)) \
     bench.py --gpus $n --workload mlmc --scaling strong --no-cpu-baseline --steps 20 2>/tmp/mlmc_ab.err | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('N=$n graph=$g', 'ms/pass %.3f' % d['ms_per_step'], 'e2e %.3f' % d['e2e']['ms_per_step'], 'mean', d['estimate']['mean'])" || tail -5 /tmp/mlmc_ab.err
done
