#!/bin/bash
# diagnostics: steady-state speed of the doc-id-only rings (single-term test compiled out: results are wrong on purpose)
# and the fraction of postings consumed while the single-term test is still needed (k2_sf: fetched_ratio)
mkdir -p gpurun_out
VAR_TAG=r2w_variants VAR_WORKLOADS=c3,c3k100,c5mix VAR_TIMEOUT=150 timeout 1500 python tools/time_variants.py k2 base_ns main_ns dm4_ns dt4_ns dg2_ns dr10_ns k2_sf 2>&1 | tail -9 | tee gpurun_out/r2w_variants.log
