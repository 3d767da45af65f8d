#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_attention_gpu.py -x -q > gpurun_out/u_tests.log 2>&1; tail -2 gpurun_out/u_tests.log
timeout 300 python tools/kbench.py attn > gpurun_out/u_kbench_attn.log 2>&1; tail -2 gpurun_out/u_kbench_attn.log
timeout 300 python tools/attn_trace.py > gpurun_out/u_attn_trace.log 2>&1
