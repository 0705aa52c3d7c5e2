#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out/r2i
timeout 150 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "matches_oracle or edge or prefilter or pruning or lookup or large_limit" 2>&1 | tail -4 > ${O}_pytest_gate.log
cat ${O}_pytest_gate.log
VAR_TAG=r2i_variants VAR_WORKLOADS=c3,c3k100,c5mix,c2 timeout 300 python tools/time_variants.py main@ring noadapt@ring nosub@ring 2>&1 | tail -4 | tee ${O}_variants.log
VAR_TAG=r2i_c4 VAR_CORPUS=zipf VAR_WORKLOADS=c4,c4mix,c4np VAR_TIMEOUT=240 timeout 500 python tools/time_variants.py main@ring noadapt@ring nosub@ring 2>&1 | tail -4 | tee ${O}_c4_variants.log
timeout 500 python -m pytest tests -q -m gpu 2>&1 | tail -6 > ${O}_pytest.log; cat ${O}_pytest.log
