#!/bin/bash
# compute-sanitizer, second pass on the seeded / handed-back / two-phase paths: racecheck on the forced-seeded parity tests,
# memcheck on the kernel-path identity test
mkdir -p gpurun_out
BM25X_SEED_FORCE=1 timeout 70 compute-sanitizer --tool racecheck --error-exitcode 9 --print-limit 20 \
    python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "matches_oracle and (C1 or ties)" > gpurun_out/r3g_sanitizer_racecheck.log 2>&1
echo "racecheck rc=$?"; tail -4 gpurun_out/r3g_sanitizer_racecheck.log
timeout 70 compute-sanitizer --tool memcheck --error-exitcode 9 --print-limit 20 \
    python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "kernel_paths and uniform or replica or sliced" > gpurun_out/r3g_sanitizer_memcheck_paths.log 2>&1
echo "memcheck rc=$?"; tail -4 gpurun_out/r3g_sanitizer_memcheck_paths.log
