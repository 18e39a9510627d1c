#!/bin/bash
# A/B of library variants on a GPU box: tools/ab_round.sh <workload> <paths> <lib...>   ("default" = in-tree build)
w=$1; p=$2; shift 2
for lib in "$@"; do
  bash tools/bench_variant.sh $lib $w $p
done
