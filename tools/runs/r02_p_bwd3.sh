#!/bin/bash
# attention backward v3 (128-row streamed tiles): parity tests, timing against v2 / cuDNN, in-kernel timeline
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_attention_gpu.py -x -q > gpurun_out/p_tests_attn.log 2>&1; tail -3 gpurun_out/p_tests_attn.log
timeout 300 python tools/kbench.py attn --vs-cudnn > gpurun_out/p_kbench_attn_v3.log 2>&1; tail -8 gpurun_out/p_kbench_attn_v3.log
UD_ATTN_BWD=2 timeout 300 python tools/kbench.py attn > gpurun_out/p_kbench_attn_v2.log 2>&1; tail -4 gpurun_out/p_kbench_attn_v2.log
timeout 300 python tools/attn_trace.py > gpurun_out/p_attn_trace_v3.log 2>&1
timeout 600 python tools/attn_stress.py > gpurun_out/p_attn_stress.log 2>&1; tail -3 gpurun_out/p_attn_stress.log
