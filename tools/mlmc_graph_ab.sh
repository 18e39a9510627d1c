#!/bin/bash
# A/B of the MLMC pass: captured graph vs queued launches vs one stream
for mode in "SDEMC_MLMC_GRAPH=1" "SDEMC_MLMC_GRAPH=0" "SDEMC_MLMC_GRAPH=0 SDEMC_MLMC_STREAMS=0"; do
  env $mode python bench.py --workload mlmc --no-cpu-baseline --steps ${STEPS:-20} 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$mode', 'ms/pass %.3f' % d['ms_per_step'], 'e2e %.3f' % d['e2e']['ms_per_step'], 'mean', d['estimate']['mean'])"
done
