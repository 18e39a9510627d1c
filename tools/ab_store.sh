#!/bin/bash
# A/B of library variants on the path-storing workloads: tools/ab_store.sh <lib.so|default> ...
cat > /tmp/ab_fmt.py <<'PY'
import json, sys
d = json.loads(sys.stdin.read())
r = d['roofline']
print(sys.argv[1], sys.argv[2], 'ms %.3f' % d['ms_per_step'], 'kernel %.3f' % (r.get('kernel_ms_per_launch') or 0),
      'GB/s %.0f' % r['achieved'], 'frac %.3f' % r['frac'])
PY
for v in "$@"; do
  for w in ${WORKLOADS:-gbm_store merton_store}; do
    if [ "$v" != default ]; then export SDEMC_B200_LIB=$PWD/$v; else unset SDEMC_B200_LIB; fi
    python bench.py --workload $w --steps 20 --warmup 3 --no-cpu-baseline 2>/tmp/st.err | python /tmp/ab_fmt.py $v $w || tail -3 /tmp/st.err
    unset SDEMC_B200_LIB
  done
done
