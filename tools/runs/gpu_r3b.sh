#!/bin/bash
# validation of the seeded kernel: all GPU tests, smoke, both bench arms, ncu captures; Zipf workloads seeded vs unseeded
bash tools/final_validate.sh r3
VAR_CORPUS=zipf VAR_TAG=r3_variants_zipf VAR_WORKLOADS=c4,c4np,c4mix VAR_TIMEOUT=200 timeout 600 python tools/time_variants.py main@seed=0 main main@seed_max_terms=4 2>&1 | tail -4 | tee gpurun_out/r3_variants_zipf.log
