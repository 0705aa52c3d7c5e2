#!/bin/bash
# Round-2 second GPU pass: quick parity gate, variants, ncu capture of the ring kernel, both bench arms.
mkdir -p gpurun_out
O=gpurun_out/r2b
timeout 120 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "matches_oracle or edge or prefilter or pruning or lookup" 2>&1 | tail -5 > ${O}_pytest_gate.log
cat ${O}_pytest_gate.log
VAR_TAG=r2b_variants VAR_WORKLOADS=c3,c3k100,c5mix,c2 timeout 500 python tools/time_variants.py main@ring r8s12w20@ring r7s12w24@ring r8s11w24@ring r8s13@ring 2>&1 | tail -6 | tee ${O}_variants.log
timeout 200 ncu --set full --clock-control none --import-source on -k regex:k_search_ring -s 2 -c 1 -f -o gpurun_out/prof_r2b \
    python bench.py --steps 1 --warmup 2 --no-cpu-baseline --no-strong --queries 20000 > ${O}_ncu_full.log 2>&1; echo "ncu full rc=$?"
timeout 200 python bench.py --impl reference --steps 3 --warmup 1 > ${O}_bench_ref.json 2> ${O}_bench_ref.err; echo "ref rc=$?"; cut -c1-400 ${O}_bench_ref.json
timeout 200 python bench.py --steps 10 --warmup 3 > ${O}_bench_c3.json 2> ${O}_bench_c3.err; echo "bench rc=$?"; tail -c 1500 ${O}_bench_c3.json; tail -3 ${O}_bench_c3.err
