#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out/r2m
timeout 150 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "matches_oracle or edge or prefilter or pruning or lookup or large_limit" 2>&1 | tail -4 > ${O}_pytest_gate.log
cat ${O}_pytest_gate.log
VAR_TAG=r2m_variants VAR_WORKLOADS=c3,c3k100,c5mix,c2 timeout 700 python tools/time_variants.py main w20 w24 s12 s12w20 r9s11 r9s12 s10w20 r7 2>&1 | tail -10 | tee ${O}_variants.log
VAR_TAG=r2m_c4 VAR_CORPUS=zipf VAR_WORKLOADS=c4,c4mix,c4np VAR_TIMEOUT=240 timeout 300 python tools/time_variants.py main 2>&1 | tail -2 | tee ${O}_c4_variants.log
