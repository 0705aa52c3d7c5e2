#!/bin/bash
# seeded launches (champion lists + doc-id-only stream) against the unseeded kernel and the two-phase launches
mkdir -p gpurun_out
O=gpurun_out/r2x
timeout 500 python -m pytest tests/test_gpu_parity.py tests/test_gpu_blocks.py tests/test_gpu_zx_stress.py tests/test_gpu_zz_growing.py -q -m gpu -x 2>&1 | tail -15 > ${O}_pytest_seed.log; tail -3 ${O}_pytest_seed.log
BM25X_SEED=0 BM25X_TWOPHASE=1 timeout 500 python -m pytest tests/test_gpu_parity.py tests/test_gpu_zx_stress.py -q -m gpu -x 2>&1 | tail -15 > ${O}_pytest_twophase.log; tail -3 ${O}_pytest_twophase.log
VAR_TAG=r2x_variants VAR_WORKLOADS=c3,c3k100,c5mix VAR_TIMEOUT=150 timeout 1500 python tools/time_variants.py main@seed=0 main main@seed=0,twophase=1 st4 sm4 st4m4 sw22 sw24 2>&1 | tail -9 | tee ${O}_variants.log
