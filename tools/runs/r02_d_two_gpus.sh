#!/bin/bash
# GPU run D (2 GPUs): full parity suite incl. the 2-rank NCCL test, GEMM kbench after the epilogue rework, bench at N=1 and N=2
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/d_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/d_tests.log; tail -4 gpurun_out/d_tests.log
python tools/kbench.py gemm > gpurun_out/d_kbench_gemm.log 2>&1; cat gpurun_out/d_kbench_gemm.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/d_bench_n1.log 2>&1; tail -1 gpurun_out/d_bench_n1.log | cut -c1-300
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/d_bench_n2.log 2>&1; tail -1 gpurun_out/d_bench_n2.log | cut -c1-300
UD_DDP_NO_WGRAD_STAGE=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/d_bench_n2_nostage.log 2>&1; tail -1 gpurun_out/d_bench_n2_nostage.log | cut -c1-300
