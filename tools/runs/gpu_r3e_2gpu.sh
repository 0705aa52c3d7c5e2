#!/bin/bash
# 2 GPUs, final kernel of round 2: replica test, then the default bench line (weak scaling + strong-scaling leg with the NCCL gather)
mkdir -p gpurun_out
O=gpurun_out/r3e
timeout 200 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "replica or sliced or kernel_paths" 2>&1 | tail -4 | tee ${O}_pytest_replica.log
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > ${O}_bench_c3_n2.json 2> ${O}_bench_c3_n2.err; echo "c3 n2 rc=$?"
tail -c 2500 ${O}_bench_c3_n2.json; tail -5 ${O}_bench_c3_n2.err
