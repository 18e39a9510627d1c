#!/bin/bash
# one `ncu --set full` capture of a workload's kernel, summarised:  tools/prof_one.sh <tag> <workload> <kernel-regex> <paths>
tag=$1; w=$2; k=$3; p=$4; out=gpurun_out
mkdir -p $out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"$k" -s 1 -c 1 -o $out/prof_${tag}_$w -f python bench.py --workload $w --paths $p --steps 1 --warmup 1 --no-cpu-baseline > $out/ncu_${tag}_$w.log 2>&1
python profiles/ncu_summary.py $out/prof_${tag}_$w.ncu-rep > $out/ncu_${tag}_$w.summary.txt 2>&1
ncu -i $out/prof_${tag}_$w.ncu-rep --page raw --csv 2>/dev/null | gzip > $out/ncu_${tag}_$w.raw.csv.gz
ncu -i $out/prof_${tag}_$w.ncu-rep --page source --csv 2>/dev/null | gzip > $out/ncu_${tag}_$w.source.csv.gz
rm -f $out/prof_${tag}_$w.ncu-rep
cat $out/ncu_${tag}_$w.summary.txt
