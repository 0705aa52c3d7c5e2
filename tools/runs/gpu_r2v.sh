#!/bin/bash
# round 2, second session: doc-id-only rings (DOCRING) + two-bit cells (K2) against the previous kernel, side by side
mkdir -p gpurun_out
O=gpurun_out/r2v
timeout 400 python -m pytest tests/test_gpu_parity.py tests/test_gpu_blocks.py -q -m gpu -x 2>&1 | tail -15 > ${O}_pytest_gate.log
cat ${O}_pytest_gate.log
VAR_TAG=r2v_variants VAR_WORKLOADS=c3,c3k100,c5mix VAR_TIMEOUT=150 timeout 1500 python tools/time_variants.py base main k2 dk1 dg2 dm4 dr10 dt4 dw22 2>&1 | tail -10 | tee ${O}_variants.log
