#!/bin/bash
# seeded launches without pruning / single-posting test (skewed queries handed back to the plain kernel)
mkdir -p gpurun_out
O=gpurun_out/r3a
timeout 500 python -m pytest tests/test_gpu_parity.py tests/test_gpu_zx_stress.py tests/test_gpu_zz_growing.py tests/test_gpu_blocks.py -q -m gpu -x 2>&1 | tail -15 > ${O}_pytest.log; tail -3 ${O}_pytest.log
VAR_TAG=r3a_variants VAR_WORKLOADS=c3,c3k100,c5mix VAR_TIMEOUT=150 timeout 1500 python tools/time_variants.py main@seed=0 main st4 sm4 st4m4 w21 st4w21 2>&1 | tail -8 | tee ${O}_variants.log
