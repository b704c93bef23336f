#!/bin/bash
# build_variant.sh NAME "-DSW_WPB=5 ..." : a copy of the product library with other kernel parameters -> pdmp3_b200/libp3_NAME.so
# P3_SRC=<dir> builds from another copy of pdmp3_b200/csrc + include (e.g. an export of an older commit) for A/B runs via P3_LIB.
set -e
cd "$(dirname "$0")/.."
N=$1; shift
SRC=${P3_SRC:-.}
mkdir -p build/v_$N
for s in p3_tables.c p3_parse.c p3_api.c; do gcc -O2 -fPIC -ffp-contract=off -I $SRC/include -c $SRC/pdmp3_b200/csrc/$s -o build/v_$N/$s.o; done
for s in p3_kernels.cu p3_fused.cu p3_hop.cu p3_cabi.cu; do nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -diag-suppress 550,177 -Xptxas -v -Xcompiler -fPIC "$@" -I $SRC/include -c $SRC/pdmp3_b200/csrc/$s -o build/v_$N/$s.o 2>&1 | grep -A2 "k_synth_warp\|k_huffman" | grep -E "Used|spill" || true; done
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o pdmp3_b200/libp3_$N.so build/v_$N/*.o -lpthread -lm -ldl
echo built pdmp3_b200/libp3_$N.so
