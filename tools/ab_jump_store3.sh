#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_store_layouts.py tests/test_gpu_parity.py tests/test_gpu_api_flows.py tests/test_gpu_fastpath.py -m gpu -x -q 2>&1 | tail -15 > gpurun_out/pytest_jtma.log
tail -3 gpurun_out/pytest_jtma.log
for rep in 1 2; do
  WORKLOADS=merton_store tools/ab_store.sh default variants/libsdemc_jnostash.so 2>&1 | tee -a gpurun_out/ab_jtma3.txt
done
tools/prof_one.sh r02f merton_store jump_store_tma 2e6 > /dev/null
head -50 gpurun_out/ncu_r02f_merton_store.summary.txt | grep -E "duration|issue_active|inst_executed.sum|stall|registers"
