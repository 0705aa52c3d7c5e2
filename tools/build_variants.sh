#!/bin/bash
# Builds tuning variants of libbm25x.so side by side (vectorchord-bm25_b200/variants/, git-ignored) so that one GPU
# session can time them all: tools/time_variants.sh.  Usage: tools/build_variants.sh name:"-DFLAGS" ...
#
# Shortlist for the next round (sized with tools/sim_planner.py); the global overrides apply to every
# query class, so read the C3 column for the 3-term class and the mix column for the rest:
#   tools/build_variants.sh 'ship:' 'bal:-DBM25X_BALANCED=1' \
#       'bal_s13:-DBM25X_BALANCED=1 -DBM25X_TWOMAP=0 -DBM25X_LOG_S=13' \
#       'bal_t1:-DBM25X_BALANCED=1 -DBM25X_TWOMAP=1 -DBM25X_LOG_S=12' \
#       'bal_t1s13:-DBM25X_BALANCED=1 -DBM25X_TWOMAP=1 -DBM25X_LOG_S=13' \
#       'bal_cb3_t1s13:-DBM25X_BALANCED=1 -DBM25X_CBMUL=3 -DBM25X_TWOMAP=1 -DBM25X_LOG_S=13'
#   gpurun -- 'python tools/time_variants.py ship bal bal_s13 bal_t1 bal_t1s13 bal_cb3_t1s13'
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
