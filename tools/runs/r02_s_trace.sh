#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/kbench.py attn > gpurun_out/s_kbench_attn_v3.log 2>&1; tail -2 gpurun_out/s_kbench_attn_v3.log
timeout 300 python tools/attn_trace.py > gpurun_out/s_attn_trace_v3.log 2>&1
UD_ATTN_BWD=2 timeout 300 python tools/attn_trace.py > gpurun_out/s_attn_trace_v2.log 2>&1
