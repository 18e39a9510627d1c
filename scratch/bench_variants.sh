#!/bin/bash
run() { SDEMC_B200_LIB=$1 python bench.py --workload $2 --steps 3 --warmup 2 --paths 2e8 --no-cpu-baseline 2>/dev/null | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$1', '$2', '%.4g' % d['value'])"; }
run $PWD/sde_mc_b200/libsdemc_b200.so gbm
run $PWD/scratch/variants/lib_diff6.so gbm
run $PWD/scratch/variants/lib_diff8.so gbm
run $PWD/sde_mc_b200/libsdemc_b200.so merton
run $PWD/scratch/variants/lib_jump5.so merton
run $PWD/scratch/variants/lib_jump6.so merton
