#!/bin/bash
# Round-2 GPU pass f: adaptive rings + single-buffered default + ballot window search + heavy-first order.
mkdir -p gpurun_out
O=gpurun_out/r2f
timeout 150 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "matches_oracle or edge or prefilter or pruning or lookup or large_limit" 2>&1 | tail -8 > ${O}_pytest_gate.log
cat ${O}_pytest_gate.log
VAR_TAG=r2f_variants VAR_WORKLOADS=c3,c3k100,c5mix,c2 timeout 400 python tools/time_variants.py main@ring db9@ring s12w20@ring 2>&1 | tail -4 | tee ${O}_variants.log
VAR_TAG=r2f_c4 VAR_CORPUS=zipf VAR_WORKLOADS=c4,c4mix,c4np VAR_TIMEOUT=240 timeout 600 python tools/time_variants.py main@ring db9@ring 2>&1 | tail -3 | tee ${O}_c4_variants.log
timeout 200 ncu --set full --clock-control none --import-source on -k regex:k_search_ring -s 2 -c 1 -f -o gpurun_out/prof_r2f \
    python bench.py --steps 1 --warmup 2 --no-cpu-baseline --no-strong --queries 20000 > ${O}_ncu_full.log 2>&1; echo "ncu full rc=$?"
timeout 500 python -m pytest tests -q -m gpu 2>&1 | tail -15 > ${O}_pytest.log; cat ${O}_pytest.log
timeout 200 python bench.py --steps 10 --warmup 3 > ${O}_bench_c3.json 2> ${O}_bench_c3.err; echo "bench rc=$?"; python - <<PY
import json
l=json.loads(open("${O}_bench_c3.json").read().strip().splitlines()[-1])
print({k:l[k] for k in ("value","ms_per_step")}, l["roofline"]["frac"], l["e2e"]["value"], l["e2e"]["ms_per_step"], l["cpu_baseline"]["value"], l["cpu_baseline"]["cores"], l["strong_scaling"]["value"], l["top100"])
PY
