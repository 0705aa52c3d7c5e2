#!/bin/bash
# Round-2 GPU pass d: full GPU test suite, ring variants, ncu capture, C4 (Zipf) pruning variants, sanitizer.
mkdir -p gpurun_out
O=gpurun_out/r2d
timeout 150 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "matches_oracle or edge or prefilter or pruning or lookup or large_limit" 2>&1 | tail -8 > ${O}_pytest_gate.log
cat ${O}_pytest_gate.log
VAR_TAG=r2d_variants VAR_WORKLOADS=c3,c3k100,c5mix,c2 timeout 400 python tools/time_variants.py main@ring r9s12@ring r8s13@ring r8s12w20@ring 2>&1 | tail -5 | tee ${O}_variants.log
timeout 200 ncu --set full --clock-control none --import-source on -k regex:k_search_ring -s 2 -c 1 -f -o gpurun_out/prof_r2d \
    python bench.py --steps 1 --warmup 2 --no-cpu-baseline --no-strong --queries 20000 > ${O}_ncu_full.log 2>&1; echo "ncu full rc=$?"
VAR_TAG=r2d_c4 VAR_CORPUS=zipf VAR_WORKLOADS=c4,c4mix,c4np VAR_TIMEOUT=240 timeout 1300 python tools/time_variants.py main@wq main@ring a05@ring a25@ring a75@ring 2>&1 | tail -6 | tee ${O}_c4_variants.log
timeout 500 python -m pytest tests -q -m gpu 2>&1 | tail -15 > ${O}_pytest.log; cat ${O}_pytest.log
bash tools/gpu_sanitize.sh r2d 2>&1 | tail -16
