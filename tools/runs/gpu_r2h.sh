#!/bin/bash
# Round-2 GPU pass h: timing of the trimmed kernel (+ A/B of the per-query ring sizes), then the secondary workloads at size.
mkdir -p gpurun_out
O=gpurun_out/r2h
timeout 150 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "matches_oracle or edge or prefilter or pruning or lookup or large_limit" 2>&1 | tail -4 > ${O}_pytest_gate.log
cat ${O}_pytest_gate.log
VAR_TAG=r2h_variants VAR_WORKLOADS=c3,c3k100,c5mix,c2 timeout 300 python tools/time_variants.py main@ring noadapt@ring 2>&1 | tail -3 | tee ${O}_variants.log
VAR_TAG=r2h_c4 VAR_CORPUS=zipf VAR_WORKLOADS=c4,c4mix,c4np VAR_TIMEOUT=240 timeout 400 python tools/time_variants.py main@ring noadapt@ring 2>&1 | tail -3 | tee ${O}_c4_variants.log
bash tools/gpu_r2g.sh
