#!/bin/bash
# A/B variant that differs in ONE translation unit: tools/build_one_variant.sh <name> <launch_jump|launch_diffusion|...> [-D...]
# -> variants/libsdemc_<name>.so (the other objects are those of the current build of sde_mc_b200/csrc)
set -e
name=$1; tu=$2; shift; shift
root=$(cd "$(dirname "$0")/.." && pwd)
obj=/tmp/sdemc_1variant_$name; mkdir -p $obj $root/variants
ARCH="-gencode arch=compute_100a,code=sm_100a"
nvcc -O3 -std=c++17 -lineinfo $ARCH -Xcompiler -fPIC -Xptxas -v --expt-relaxed-constexpr "$@" \
     -c $root/sde_mc_b200/csrc/$tu.cu -o $obj/$tu.o 2> $obj/$tu.ptxas.log
c=$root/sde_mc_b200/csrc
objs=""
for f in abi launch_diffusion launch_jump launch_jump_store launch_jump_store_inject launch_jump_store_queue launch_jump_store_inline launch_pair launch_cv; do
  if [ $f = $tu ]; then objs="$objs $obj/$f.o"; else objs="$objs $c/$f.o"; fi
done
nvcc $ARCH -shared -o $root/variants/libsdemc_$name.so $objs -cudart static
echo "built variants/libsdemc_$name.so"
