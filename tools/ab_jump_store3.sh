#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_store_layouts.py tests/test_gpu_parity.py tests/test_gpu_api_flows.py tests/test_gpu_fastpath.py -m gpu -x -q 2>&1 | tail -15 > gpurun_out/pytest_jtma.log
tail -3 gpurun_out/pytest_jtma.log
WORKLOADS=merton_store tools/ab_store.sh default "$@" 2>&1 | tee gpurun_out/ab_jtma3.txt
tools/prof_one.sh r02c merton_store jump_store_tma 2e6 > /dev/null
head -50 gpurun_out/ncu_r02c_merton_store.summary.txt
