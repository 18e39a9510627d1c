#!/bin/bash
# GPU check of the TMA jump-store kernel: full GPU suite, then A/B of the buffering variants on merton_store
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/pytest_jtma.log
tail -3 gpurun_out/pytest_jtma.log
WORKLOADS=merton_store tools/ab_store.sh default variants/libsdemc_nb2.so variants/libsdemc_w2.so variants/libsdemc_b128.so variants/libsdemc_b32nb2.so variants/libsdemc_b32.so 2>&1 | tee gpurun_out/ab_jtma.txt
WORKLOADS=gbm_store tools/ab_store.sh default 2>&1 | tee -a gpurun_out/ab_jtma.txt
