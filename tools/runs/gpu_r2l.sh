#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out/r2l
timeout 150 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "matches_oracle or edge or prefilter or pruning or lookup or large_limit" 2>&1 | tail -4 > ${O}_pytest_gate.log
cat ${O}_pytest_gate.log
VAR_TAG=r2l_variants VAR_WORKLOADS=c3,c3k100,c5mix,c2 timeout 500 python tools/time_variants.py main bm9 bm8 bm9s12 bm8s11 2>&1 | tail -6 | tee ${O}_variants.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_search_ring -s 4 -c 1 -f -o gpurun_out/prof_r2l_m8 \
    python bench.py --workload c5 --docs 10000000 --queries 40000 --steps 1 --warmup 1 --no-cpu-baseline --no-strong > ${O}_ncu_m8.log 2>&1; echo "ncu m8 rc=$?"
