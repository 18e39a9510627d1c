#!/bin/bash
mkdir -p gpurun_out
WORKLOADS=merton_store tools/ab_store.sh variants/libsdemc_b32.so variants/libsdemc_b32p4.so variants/libsdemc_b32p14.so variants/libsdemc_b96.so default 2>&1 | tee gpurun_out/ab_jtma2.txt
tools/prof_one.sh b64 merton_store jump_store_tma 2e6 > /dev/null
SDEMC_B200_LIB=$PWD/variants/libsdemc_b32.so tools/prof_one.sh b32 merton_store jump_store_tma 2e6 > /dev/null
head -50 gpurun_out/ncu_b64_merton_store.summary.txt gpurun_out/ncu_b32_merton_store.summary.txt
