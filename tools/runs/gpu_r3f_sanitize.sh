#!/bin/bash
# compute-sanitizer (memcheck, synccheck) on the seeded kernel: every eligible query forced through it
mkdir -p gpurun_out
SEL='matches_oracle and (C1 or ties)'
for tool in memcheck synccheck; do
  BM25X_SEED_FORCE=1 timeout 140 compute-sanitizer --tool $tool --error-exitcode 9 --print-limit 30 \
      python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "$SEL" > gpurun_out/r3f_sanitizer_${tool}.log 2>&1
  echo "$tool rc=$?"; tail -4 gpurun_out/r3f_sanitizer_${tool}.log
done
