#!/bin/bash
# A/B of library variants on the three headline workloads: tools/ab_three.sh <lib.so|default> ...
for v in "$@"; do
  bash tools/bench_variant.sh $v gbm 1e9
  bash tools/bench_variant.sh $v merton 1e9
  if [ "$v" != default ]; then export SDEMC_B200_LIB=$PWD/$v; else unset SDEMC_B200_LIB; fi
  python bench.py --workload mlmc --no-cpu-baseline 2>/tmp/mlmc.err | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$v', 'mlmc ms/pass', round(d['ms_per_step'],3))" || tail -3 /tmp/mlmc.err
  unset SDEMC_B200_LIB
done
