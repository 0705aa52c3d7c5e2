#!/bin/bash
# seeded launches, third flavour: the seeds join the candidate list inside the producer loop (one verification site)
mkdir -p gpurun_out
O=gpurun_out/r2z
timeout 500 python -m pytest tests/test_gpu_parity.py tests/test_gpu_zx_stress.py -q -m gpu -x 2>&1 | tail -15 > ${O}_pytest_seed.log; tail -3 ${O}_pytest_seed.log
VAR_TAG=r2z_variants VAR_WORKLOADS=c3,c3k100,c5mix VAR_TIMEOUT=150 timeout 1500 python tools/time_variants.py main@seed=0 main noseeds st4 sm4 st4m4 2>&1 | tail -7 | tee ${O}_variants.log
