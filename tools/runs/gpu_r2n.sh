#!/bin/bash
mkdir -p gpurun_out
VAR_TAG=r2n_variants VAR_WORKLOADS=c3,c3k100,c5mix,c2 timeout 300 python tools/time_variants.py main w18 w22 2>&1 | tail -4 | tee gpurun_out/r2n_variants.log
bash tools/final_validate.sh r2n
