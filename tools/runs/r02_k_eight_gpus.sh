#!/bin/bash
# 8 GPUs: our arm at N=8 (+ DDP bucket timeline), N=4
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --steps 15 --warmup 4 > gpurun_out/k_bench_n8.log 2>&1; tail -1 gpurun_out/k_bench_n8.log | cut -c1-260
UD_DDP_DEBUG=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus 8 --steps 5 --warmup 3 > gpurun_out/k_bench_n8_timeline.log 2>&1; grep "ddp timeline" gpurun_out/k_bench_n8_timeline.log | tail -30
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29523 bench.py --gpus 4 --steps 15 --warmup 4 > gpurun_out/k_bench_n4.log 2>&1; tail -1 gpurun_out/k_bench_n4.log | cut -c1-260
timeout 300 python bench.py --steps 15 --warmup 4 --no-cpu-baseline > gpurun_out/k_bench_n1.log 2>&1; tail -1 gpurun_out/k_bench_n1.log | cut -c1-260
