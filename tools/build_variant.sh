#!/bin/bash
# Builds libp2gpu.so with extra compile flags into plonky2_aes_b200/variants/libp2gpu_<name>.so for A/B
# measurements (select at run time with P2G_LIB_PATH).  Usage: tools/build_variant.sh <name> <flags...>
set -e
name=$1; shift
root=$(cd "$(dirname "$0")/.." && pwd)
src=$root/plonky2_aes_b200/csrc
obj=$(mktemp -d)
mkdir -p $root/plonky2_aes_b200/variants
for f in api ntt merkle prover witgen; do
  /usr/local/cuda/bin/nvcc -O3 -std=c++17 -lineinfo -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC "$@" -c $src/$f.cu -o $obj/$f.o &
done
wait
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -shared -o $root/plonky2_aes_b200/variants/libp2gpu_$name.so $obj/api.o $obj/ntt.o $obj/merkle.o $obj/prover.o $obj/witgen.o -lcudart
rm -rf $obj
echo built plonky2_aes_b200/variants/libp2gpu_$name.so
