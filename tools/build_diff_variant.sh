#!/bin/bash
# A/B variant that only changes the uniform-grid kernels: tools/build_diff_variant.sh <name> [nvcc flags]
# recompiles launch_diffusion.cu with the flags and links it with the objects of the current default build
# -> variants/libsdemc_<name>.so (select with SDEMC_B200_LIB, tools/ab_store.sh)
set -e
name=$1; shift
root=$(cd "$(dirname "$0")/.." && pwd)
obj=/tmp/sdemc_variant_$name; mkdir -p $obj $root/variants
ARCH="-gencode arch=compute_100a,code=sm_100a"
nvcc -O3 -std=c++17 -lineinfo $ARCH -Xcompiler -fPIC -Xptxas -v --expt-relaxed-constexpr "$@" \
     -c $root/sde_mc_b200/csrc/launch_diffusion.cu -o $obj/launch_diffusion.o 2> $obj/launch_diffusion.ptxas.log
c=$root/sde_mc_b200/csrc
nvcc $ARCH -shared -o $root/variants/libsdemc_$name.so $obj/launch_diffusion.o $c/abi.o $c/launch_jump.o $c/launch_pair.o $c/launch_cv.o -cudart static
echo "built variants/libsdemc_$name.so"
