#!/bin/bash
# Round profile: bench lines for every workload, the ncu launch list of the default bench and one `ncu --set full`
# capture per hot kernel.  Run on a GPU box:  tools/profile_round.sh <tag>   (outputs under gpurun_out/)
tag=${1:-rXX}
out=gpurun_out
mkdir -p $out
for w in gbm merton levy2d merton_cv mlmc gbm_store merton_store; do
  extra=""
  case $w in gbm_store|merton_store) extra="--steps 20";; esac
  timeout 600 python bench.py --workload $w $extra > $out/bench_${tag}_$w.json 2> $out/bench_${tag}_$w.err
  python - <<PY
import json
try:
    d = json.load(open("$out/bench_${tag}_$w.json"))
    print("$w", "%.4g" % d["value"], "e2e %.4g" % d["e2e"]["value"], "frac %.3f" % d["roofline"]["frac"], d.get("roofline_tensor", {}).get("frac"), d["clocks"])
except Exception as e:
    print("$w FAILED", e)
PY
done
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > $out/bench_${tag}_reference.json 2>/dev/null
timeout 300 python bench.py --gpus 1 --workload mlmc --scaling strong --no-cpu-baseline --steps 20 > $out/bench_${tag}_mlmc_strong_n1.json 2>/dev/null
# launch list of the default bench command = the Merton north-star workload (cold-cache, serialised: shares, not absolutes)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/launches_${tag}_merton.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > $out/launches_${tag}_merton.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/launches_${tag}_gbm.csv python bench.py --workload gbm --steps 2 --warmup 1 --no-cpu-baseline > $out/launches_${tag}_gbm.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/launches_${tag}_merton_store.csv python bench.py --workload merton_store --steps 2 --warmup 1 --no-cpu-baseline > $out/launches_${tag}_merton_store.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/launches_${tag}_gbm_store.csv python bench.py --workload gbm_store --steps 2 --warmup 1 --no-cpu-baseline > $out/launches_${tag}_gbm_store.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/launches_${tag}_mlmc.csv python bench.py --workload mlmc --steps 2 --warmup 1 --no-cpu-baseline > $out/launches_${tag}_mlmc.log 2>&1
prof() { # workload kernel-regex paths
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:"$2" -s 1 -c 1 -o $out/prof_${tag}_$1 -f python bench.py --workload $1 --paths $3 --steps 1 --warmup 1 --no-cpu-baseline > $out/ncu_${tag}_$1.log 2>&1
  # gpurun merges at most 64 MiB back: keep the summary, the raw metric row and the per-instruction page (gzip), not the report
  python profiles/ncu_summary.py $out/prof_${tag}_$1.ncu-rep > $out/ncu_${tag}_$1.summary.txt 2>&1
  ncu -i $out/prof_${tag}_$1.ncu-rep --page raw --csv 2>/dev/null | gzip > $out/ncu_${tag}_$1.raw.csv.gz
  ncu -i $out/prof_${tag}_$1.ncu-rep --page source --csv 2>/dev/null | gzip > $out/ncu_${tag}_$1.source.csv.gz
  rm -f $out/prof_${tag}_$1.ncu-rep
}
prof gbm diffusion_kernel 1e8
prof merton jump1d_kernel 5e7
prof levy2d jump_kernel 5e6
prof mlmc jump_flat1d_kernel 1
prof merton_cv cv_kernel 2e6
prof gbm_store diffusion_store_tma_kernel 4e6
prof merton_store jump_store_tma_kernel 2e6
ls -la $out | tail -30
