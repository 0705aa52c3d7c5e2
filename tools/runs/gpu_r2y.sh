#!/bin/bash
# seeded launches, second flavour: seeds join the verification of their doc window (no probes)
mkdir -p gpurun_out
O=gpurun_out/r2y
timeout 500 python -m pytest tests/test_gpu_parity.py tests/test_gpu_blocks.py tests/test_gpu_zx_stress.py tests/test_gpu_zz_growing.py -q -m gpu -x 2>&1 | tail -15 > ${O}_pytest_seed.log; tail -3 ${O}_pytest_seed.log
VAR_TAG=r2y_variants VAR_WORKLOADS=c3,c3k100,c5mix VAR_TIMEOUT=150 timeout 1500 python tools/time_variants.py main@seed=0 main st4 sm4 st4m4 si32 st4m4w18 st4m4k1 2>&1 | tail -9 | tee ${O}_variants.log
