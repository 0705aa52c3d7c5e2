#!/bin/bash
# r3c: density hand-back, scan / mask micro-optimisations; forced-seeded tests; Zipf workloads
mkdir -p gpurun_out
O=gpurun_out/r3c
BM25X_SEED_FORCE=1 timeout 400 python -m pytest tests/test_gpu_parity.py tests/test_gpu_zx_stress.py tests/test_gpu_zz_growing.py -q -m gpu -x 2>&1 | tail -6 > ${O}_pytest_forced.log; tail -2 ${O}_pytest_forced.log
timeout 400 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "kernel_paths or pruning" 2>&1 | tail -3
VAR_TAG=r3c_variants VAR_WORKLOADS=c3,c3k100,c5mix VAR_TIMEOUT=150 timeout 900 python tools/time_variants.py main@seed=0 main k2s7 g2t2 m4 2>&1 | tail -6 | tee ${O}_variants.log
VAR_CORPUS=zipf VAR_TAG=r3c_variants_zipf VAR_WORKLOADS=c4,c4np,c4mix VAR_TIMEOUT=200 timeout 600 python tools/time_variants.py main@seed=0 main 2>&1 | tail -3 | tee ${O}_variants_zipf.log
