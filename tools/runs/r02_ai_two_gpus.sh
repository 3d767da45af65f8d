#!/bin/bash
# 2 GPUs: the 2-rank NCCL parity test (ThinDDP + FusedAdamW with the masked / split head), bench at N=1 and N=2 on the same box
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_ddp_nccl_gpu.py -x -q > gpurun_out/ai_tests_ddp.log 2>&1; echo "tests rc=$?" >> gpurun_out/ai_tests_ddp.log; tail -3 gpurun_out/ai_tests_ddp.log
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/ai_bench_n1.log 2>&1; tail -1 gpurun_out/ai_bench_n1.log | cut -c1-200
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/ai_bench_n2.log 2>&1; tail -1 gpurun_out/ai_bench_n2.log | cut -c1-200
