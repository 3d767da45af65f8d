#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_attention_gpu.py -x -q > gpurun_out/w_tests.log 2>&1; tail -2 gpurun_out/w_tests.log
UD_ATTN_BWD=3 timeout 900 python -m pytest tests/test_attention_gpu.py -x -q > gpurun_out/w_tests_v3.log 2>&1; tail -2 gpurun_out/w_tests_v3.log
timeout 300 python tools/kbench.py attn > gpurun_out/w_kbench_attn.log 2>&1; tail -1 gpurun_out/w_kbench_attn.log
UD_ATTN_BWD=2 timeout 300 python tools/kbench.py attn > gpurun_out/w_kbench_attn_v2.log 2>&1; tail -1 gpurun_out/w_kbench_attn_v2.log
timeout 300 python tools/attn_trace.py > gpurun_out/w_attn_trace.log 2>&1
timeout 600 python bench.py --workload unidisc-1.4B-interleaved --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/w_interleaved.log 2>&1; tail -1 gpurun_out/w_interleaved.log | cut -c1-200
UD_ATTN_BWD=3 timeout 600 python bench.py --workload unidisc-1.4B-interleaved --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/w_interleaved_v3.log 2>&1; tail -1 gpurun_out/w_interleaved_v3.log | cut -c1-200
