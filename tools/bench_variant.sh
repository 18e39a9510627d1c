#!/bin/bash
# usage: tools/bench_variant.sh <lib.so|default> <workload> [paths]
lib=$1; w=$2; p=${3:-2e8}
if [ "$lib" != default ]; then export SDEMC_B200_LIB=$PWD/$lib; fi
python bench.py --workload $w --steps 3 --warmup 2 --paths $p --no-cpu-baseline 2>/dev/null | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$lib', '$w', '%.4g path-steps/s' % d['value'], 'frac %.3f' % d['roofline']['frac'], 'mean %.6f +- %.6f' % (d['estimate']['mean'], d['estimate']['stderr']))"
