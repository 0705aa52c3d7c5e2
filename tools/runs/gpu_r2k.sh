#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out/r2k
timeout 150 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "matches_oracle or edge or prefilter or pruning or lookup or large_limit" 2>&1 | tail -4 > ${O}_pytest_gate.log
cat ${O}_pytest_gate.log
VAR_TAG=r2k_variants VAR_WORKLOADS=c3,c3k100,c5mix,c2 timeout 300 python tools/time_variants.py main mp6144 2>&1 | tail -3 | tee ${O}_variants.log
VAR_TAG=r2k_c4 VAR_CORPUS=zipf VAR_WORKLOADS=c4,c4mix,c4np VAR_TIMEOUT=240 timeout 300 python tools/time_variants.py main 2>&1 | tail -2 | tee ${O}_c4_variants.log
timeout 200 ncu --set full --clock-control none --import-source on -k regex:k_search_ring -s 2 -c 1 -f -o gpurun_out/prof_r2k \
    python bench.py --steps 1 --warmup 2 --no-cpu-baseline --no-strong --queries 20000 > ${O}_ncu_full.log 2>&1; echo "ncu full rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on -k "regex:int.8," -s 1 -c 1 -f -o gpurun_out/prof_r2k_m8 \
    python bench.py --workload c5 --docs 10000000 --queries 40000 --steps 1 --warmup 1 --no-cpu-baseline --no-strong > ${O}_ncu_m8.log 2>&1; echo "ncu m8 rc=$?"
