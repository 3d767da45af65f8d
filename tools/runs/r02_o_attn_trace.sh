#!/bin/bash
# in-kernel timeline of the attention-backward kernels + in-situ per-launch times of one training step
mkdir -p gpurun_out
timeout 300 python tools/attn_trace.py > gpurun_out/o_attn_trace.log 2>&1
UD_KERNEL_TIMES=1 UD_PHASE_TIMING=1 timeout 600 python bench.py --steps 6 --warmup 3 > gpurun_out/o_bench_ktimes.log 2>&1
tail -5 gpurun_out/o_attn_trace.log
