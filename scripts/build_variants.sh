#!/bin/bash
# A/B builds of one translation unit with experiment macros: scripts/build_variants.sh <source.cu> <name>:<-Dflags> ...
# -> larnd-sim-jax_b200/variants/lib_<name>.so (linked with the objects of the regular build; select with LARND_B200_LIB)
set -e
cd "$(dirname "$0")/.."
src=$1; shift
base=$(basename $src .cu)
mkdir -p larnd-sim-jax_b200/variants
for spec in "$@"; do
  name=${spec%%:*}; flags=${spec#*:}
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -Xptxas -v -Iinclude $flags -c larnd-sim-jax_b200/csrc/$base.cu -o /tmp/var_$name.o 2> /tmp/var_$name.log
  objs=$(ls larnd-sim-jax_b200/build/*.o | grep -v "/$base.o")
  nvcc -gencode arch=compute_100a,code=sm_100a -shared -o larnd-sim-jax_b200/variants/lib_$name.so $objs /tmp/var_$name.o
  echo "$name: $(grep -A1 'k_acc_tilesILi1ELi4E\|k_bwd_tilesILi4ELi2ELb1' /tmp/var_$name.log | grep -E 'spill' | head -1) $(grep -A2 'k_acc_tilesILi1ELi4E\|k_bwd_tilesILi4ELi2ELb1' /tmp/var_$name.log | grep -E 'Used' | head -1)"
done
