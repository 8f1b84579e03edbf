#!/bin/bash
# How much do the integer half (S-boxes, folds, read-outs) and the FP64 half (MDS) of the Poseidon
# permutation overlap?  Builds two diagnostic variants of libp2gpu.so that drop one half each
# (-DP2G_DIAG_NO_SBOX / -DP2G_DIAG_NO_MDS, wrong results by construction) and prints the chained
# permutation rate of each next to the real library.  Run on the GPU box:
#   gpurun -- 'bash tools/poseidon_overlap.sh'
# Round-1 result: FP64 half alone 2.53 G perm/s, integer half alone 1.66 G, both 1.28 G.
set -e
cd "$(dirname "$0")/.."
SRC=plonky2_aes_b200/csrc
OUT=$(mktemp -d)
for v in NO_SBOX NO_MDS; do
  for f in api ntt merkle prover witgen; do
    /usr/local/cuda/bin/nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC -DP2G_DIAG_$v -c $SRC/$f.cu -o $OUT/$f.o &
  done
  wait
  /usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -shared -o $OUT/libp2gpu.so $OUT/api.o $OUT/ntt.o $OUT/merkle.o $OUT/prover.o $OUT/witgen.o -lcudart
  cp plonky2_aes_b200/libp2witness.so $OUT/
  echo -n "$v: "; P2G_LIB_PATH=$OUT/libp2gpu.so python tools/poseidon_peak.py | tail -1
done
echo -n "full: "; python tools/poseidon_peak.py | tail -1
rm -rf "$OUT"
