#!/bin/bash
# A/B variant that differs only in launch_jump.cu: tools/build_jump_variant.sh <name> [-D...]
# -> variants/libsdemc_<name>.so (the other objects are those of the current build of sde_mc_b200/csrc)
set -e
name=$1; shift
root=$(cd "$(dirname "$0")/.." && pwd)
obj=/tmp/sdemc_jvariant_$name; mkdir -p $obj $root/variants
ARCH="-gencode arch=compute_100a,code=sm_100a"
nvcc -O3 -std=c++17 -lineinfo $ARCH -Xcompiler -fPIC -Xptxas -v --expt-relaxed-constexpr "$@" \
     -c $root/sde_mc_b200/csrc/launch_jump.cu -o $obj/launch_jump.o 2> $obj/launch_jump.ptxas.log
c=$root/sde_mc_b200/csrc
nvcc $ARCH -shared -o $root/variants/libsdemc_$name.so $c/abi.o $c/launch_diffusion.o $obj/launch_jump.o $c/launch_pair.o $c/launch_cv.o -cudart static
echo "built variants/libsdemc_$name.so"
