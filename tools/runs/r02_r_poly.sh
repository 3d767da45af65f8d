#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_attention_gpu.py tests/test_model_gpu.py -x -q > gpurun_out/r_tests.log 2>&1; tail -3 gpurun_out/r_tests.log
timeout 300 python tools/kbench.py attn > gpurun_out/r_kbench_attn_v3.log 2>&1; tail -3 gpurun_out/r_kbench_attn_v3.log
UD_ATTN_BWD=2 timeout 300 python tools/kbench.py attn > gpurun_out/r_kbench_attn_v2.log 2>&1; tail -2 gpurun_out/r_kbench_attn_v2.log
timeout 300 python tools/attn_trace.py > gpurun_out/r_attn_trace_v3.log 2>&1
