#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/j_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/j_tests.log; tail -3 gpurun_out/j_tests.log
timeout 300 python tools/attn_stress.py 40 > gpurun_out/j_stress.log 2>&1; tail -5 gpurun_out/j_stress.log
python tools/kbench.py attn rows > gpurun_out/j_kbench.log 2>&1; cat gpurun_out/j_kbench.log
UD_ATTN_FWD=3 python tools/kbench.py attn > gpurun_out/j_kbench_fwd3.log 2>&1; cat gpurun_out/j_kbench_fwd3.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/j_bench.log 2>&1; tail -1 gpurun_out/j_bench.log | cut -c1-200
