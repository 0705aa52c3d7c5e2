#!/bin/bash
mkdir -p gpurun_out
timeout 40 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "broker" 2>&1 | tail -6 | tee gpurun_out/r3h_pytest_broker.log
