#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_store_layouts.py tests/test_gpu_parity.py tests/test_gpu_api_flows.py tests/test_gpu_fastpath.py tests/test_gpu_user_sde.py tests/test_gpu_milstein.py -m gpu -x -q 2>&1 | tail -15 > gpurun_out/pytest_jtma.log
tail -3 gpurun_out/pytest_jtma.log
tools/ab_store.sh default 2>&1 | tee gpurun_out/ab_jtma3.txt
WORKLOADS=merton_store tools/ab_store.sh variants/libsdemc_jinline.so 2>&1 | tee -a gpurun_out/ab_jtma3.txt
WORKLOADS=gbm_store tools/ab_store.sh variants/libsdemc_dinline.so 2>&1 | tee -a gpurun_out/ab_jtma3.txt
tools/ab_store.sh default 2>&1 | tee -a gpurun_out/ab_jtma3.txt
tools/prof_one.sh r02e merton_store jump_store_tma 2e6 > /dev/null
head -50 gpurun_out/ncu_r02e_merton_store.summary.txt | grep -E "duration|issue_active|inst_executed.sum|stall"
