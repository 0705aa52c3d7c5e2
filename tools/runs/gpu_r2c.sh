#!/bin/bash
# Round-2 third GPU pass: sanitizer on the shipped kernels + C4 (Zipf) pruning variants.
mkdir -p gpurun_out
O=gpurun_out/r2c
bash tools/gpu_sanitize.sh r2c 2>&1 | tail -20
VAR_TAG=r2c_c4 VAR_CORPUS=zipf VAR_WORKLOADS=c4,c4mix,c4np VAR_TIMEOUT=200 timeout 900 python tools/time_variants.py main@wq main@ring a05@ring a25@ring a75@ring 2>&1 | tail -8 | tee ${O}_c4_variants.log
