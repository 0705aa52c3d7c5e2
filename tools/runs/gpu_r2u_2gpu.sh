#!/bin/bash
# 2 GPUs, final kernel: default bench line (weak scaling + strong-scaling leg with the NCCL gather) and the reference arm under torchrun.
mkdir -p gpurun_out
O=gpurun_out/r2u
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > ${O}_bench_c3_n2.json 2> ${O}_bench_c3_n2.err; echo "c3 n2 rc=$?"
tail -c 1500 ${O}_bench_c3_n2.json; tail -5 ${O}_bench_c3_n2.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > ${O}_bench_ref_n2.json 2> ${O}_bench_ref_n2.err; echo "ref n2 rc=$?"
tail -c 800 ${O}_bench_ref_n2.json; tail -3 ${O}_bench_ref_n2.err
