#!/bin/bash
# 2 GPUs: the default bench line (weak scaling + the strong-scaling leg with the NCCL gather) and the 50M-doc C5 workload.
mkdir -p gpurun_out
O=gpurun_out/r2j
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > ${O}_bench_c3_n2.json 2> ${O}_bench_c3_n2.err; echo "c3 n2 rc=$?"
tail -c 1200 ${O}_bench_c3_n2.json; tail -5 ${O}_bench_c3_n2.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --workload c5 --steps 3 --warmup 1 --no-cpu-baseline > ${O}_bench_c5_n2.json 2> ${O}_bench_c5_n2.err; echo "c5 n2 rc=$?"
tail -c 1500 ${O}_bench_c5_n2.json; tail -5 ${O}_bench_c5_n2.err
