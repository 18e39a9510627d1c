#!/bin/bash
# A/B of library variants on the path-storing workloads: tools/ab_store.sh <lib.so|default> ...
for v in "$@"; do
  for w in ${WORKLOADS:-gbm_store merton_store}; do
    if [ "$v" != default ]; then export SDEMC_B200_LIB=$PWD/$v; else unset SDEMC_B200_LIB; fi
    python bench.py --workload $w --steps 20 --warmup 3 --no-cpu-baseline 2>/tmp/st.err | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$v', '$w', 'ms %.3f' % d['ms_per_step'], 'GB/s %.0f' % d['roofline']['achieved'], 'frac %.3f' % d['roofline']['frac'])" || tail -3 /tmp/st.err
    unset SDEMC_B200_LIB
  done
done
