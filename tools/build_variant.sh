#!/bin/bash
# build an A/B variant of the engine: tools/build_variant.sh <name> [extra nvcc flags, e.g. -DSDEMC_JUMP1D_MIN_BLOCKS=4]
# -> variants/libsdemc_<name>.so (git-ignored; travels to the GPU box; select with SDEMC_B200_LIB or tools/bench_variant.sh)
set -e
name=$1; shift
root=$(cd "$(dirname "$0")/.." && pwd)
obj=/tmp/sdemc_variant_$name; mkdir -p $obj $root/variants
ARCH="-gencode arch=compute_100a,code=sm_100a"
for f in abi launch_diffusion launch_jump launch_jump_store launch_jump_store_inject launch_jump_store_queue launch_jump_store_inline launch_pair launch_cv; do
  nvcc -O3 -std=c++17 -lineinfo $ARCH -Xcompiler -fPIC -Xptxas -v --expt-relaxed-constexpr "$@" \
       -c $root/sde_mc_b200/csrc/$f.cu -o $obj/$f.o 2> $obj/$f.ptxas.log &
done
wait
nvcc $ARCH -shared -o $root/variants/libsdemc_$name.so $obj/*.o -cudart static
echo "built variants/libsdemc_$name.so"
