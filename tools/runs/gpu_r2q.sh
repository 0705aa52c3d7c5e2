#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out/r2q
timeout 150 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "matches_oracle or edge or prefilter or pruning or lookup or large_limit" 2>&1 | tail -4 > ${O}_pytest_gate.log
cat ${O}_pytest_gate.log
VAR_TAG=r2q_variants VAR_WORKLOADS=c3,c3k100,c5mix,c2 timeout 300 python tools/time_variants.py main u1 2>&1 | tail -3 | tee ${O}_variants.log
