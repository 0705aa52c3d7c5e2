#!/bin/bash
# Builds tuning variants of libbm25x.so side by side (vectorchord-bm25_b200/variants/, git-ignored) so that one GPU
# session can time them all: tools/time_variants.sh.  Usage: tools/build_variants.sh name:"-DFLAGS" ...
set -e
cd "$(dirname "$0")/../vectorchord-bm25_b200/csrc"
mkdir -p ../variants
for spec in "$@"; do
    name="${spec%%:*}"; flags="${spec#*:}"
    /usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo \
        -Xcompiler -fPIC,-fopenmp,-ffp-contract=off,-O3 $flags -shared -o ../variants/libbm25x_$name.so \
        bm25x_index.cu bm25x_search.cu bm25x_synth.cpp -lgomp &
done
wait
ls -la ../variants/
