#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_attention_gpu.py -x -q > gpurun_out/q_tests_attn.log 2>&1; tail -3 gpurun_out/q_tests_attn.log
timeout 300 python tools/kbench.py attn > gpurun_out/q_kbench_attn_v3.log 2>&1; tail -3 gpurun_out/q_kbench_attn_v3.log
timeout 300 python tools/attn_trace.py > gpurun_out/q_attn_trace_v3.log 2>&1
