#!/bin/bash
# Round-2 GPU pass e: single-buffered ring variants.
mkdir -p gpurun_out
O=gpurun_out/r2e
VAR_TAG=r2e_variants VAR_WORKLOADS=c3,c3k100,c5mix,c2 timeout 600 python tools/time_variants.py main@ring sb8@ring sb8s12@ring sb9@ring sb7s12w24@ring 2>&1 | tail -6 | tee ${O}_variants.log
timeout 200 ncu --set full --clock-control none --import-source on -k regex:k_search_ring -s 2 -c 1 -f -o gpurun_out/prof_r2e \
    env BM25X_LIBRARY=$PWD/vectorchord-bm25_b200/variants/libbm25x_sb8.so python bench.py --steps 1 --warmup 2 --no-cpu-baseline --no-strong --queries 20000 > ${O}_ncu_full.log 2>&1; echo "ncu full rc=$?"
timeout 100 python -m pytest tests/test_gpu_blocks.py -q -m gpu -k "wand" 2>&1 | tail -3
