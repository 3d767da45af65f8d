#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_attention_gpu.py -x -q > gpurun_out/t_tests.log 2>&1; tail -2 gpurun_out/t_tests.log
timeout 300 python tools/kbench.py attn > gpurun_out/t_kbench_attn_hybrid.log 2>&1; tail -2 gpurun_out/t_kbench_attn_hybrid.log
UD_ATTN_BWD=2 timeout 300 python tools/kbench.py attn > gpurun_out/t_kbench_attn_v2.log 2>&1; tail -1 gpurun_out/t_kbench_attn_v2.log
UD_ATTN_BWD=3 timeout 300 python tools/kbench.py attn > gpurun_out/t_kbench_attn_v3.log 2>&1; tail -1 gpurun_out/t_kbench_attn_v3.log
UD_ATTN_BWD=3 timeout 300 python tools/attn_trace.py > gpurun_out/t_attn_trace_v3.log 2>&1
UD_ATTN_BWD=2 timeout 300 python tools/attn_trace.py > gpurun_out/t_attn_trace_v2.log 2>&1
