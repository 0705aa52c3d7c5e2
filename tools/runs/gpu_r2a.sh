#!/bin/bash
# Round-2 first GPU pass: parity of the ring kernel (watchdog build), then kernel generations / ring variants side by side.
mkdir -p gpurun_out
O=gpurun_out/r2a
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv,noheader
timeout 120 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "matches_oracle or golden or edge or prefilter or pruning" 2>&1 | tail -25 > ${O}_pytest_gate.log
cat ${O}_pytest_gate.log
if grep -q "passed" ${O}_pytest_gate.log && ! grep -q "failed\|error" ${O}_pytest_gate.log; then
  timeout 400 python -m pytest tests -q -m gpu 2>&1 | tail -25 > ${O}_pytest.log; cat ${O}_pytest.log
fi
VAR_TAG=r2a_variants VAR_WORKLOADS=c3,c3k100,c5mix,c2 timeout 500 python tools/time_variants.py main@wq main@ring r8s12w20@ring r9s12@ring r8s13@ring u1@ring 2>&1 | tee ${O}_variants.log
