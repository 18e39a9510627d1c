#!/bin/bash
# end-of-round validation on a GPU box: GPU tests, smoke, the default bench line (+ reference arm), GBM capture
out=gpurun_out; mkdir -p $out
(time timeout 900 python -m pytest tests -m gpu -x -q) 2>&1 | tail -5
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 600 python bench.py > $out/bench_final_gbm.json 2> $out/bench_final_gbm.err; cut -c1-400 $out/bench_final_gbm.json
timeout 600 python bench.py --workload mlmc > $out/bench_final_mlmc.json 2>/dev/null; cut -c1-200 $out/bench_final_mlmc.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/launches_final_gbm.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > $out/launches_final_gbm.log 2>&1
bash tools/prof_one.sh final gbm diffusion_kernel 1e8 2>&1 | grep -E "time_duration|registers|issue_active|pipe_xu|pipe_alu|pipe_fma.avg|stall"
