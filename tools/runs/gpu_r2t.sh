#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out/r2t
timeout 300 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "more_than_32 or matches_oracle or pruning or edge" 2>&1 | tail -25 > ${O}_pytest_gate.log
cat ${O}_pytest_gate.log
VAR_TAG=r2t_variants VAR_WORKLOADS=c3,c3k100,c5mix,c2 timeout 300 python tools/time_variants.py main 2>&1 | tail -2 | tee ${O}_variants.log
timeout 300 python -m pytest tests/test_gpu_zx_stress.py -q -m gpu 2>&1 | tail -4
