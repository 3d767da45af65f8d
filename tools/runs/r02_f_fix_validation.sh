#!/bin/bash
mkdir -p gpurun_out
python tools/attn_stress.py 200 > gpurun_out/f_stress.log 2>&1; cat gpurun_out/f_stress.log | tail -6
UD_ATTN_FWD=3 python tools/attn_stress.py 40 > gpurun_out/f_stress_fwd3.log 2>&1; cat gpurun_out/f_stress_fwd3.log | tail -6
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/f_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/f_tests.log; tail -6 gpurun_out/f_tests.log
python tools/kbench.py attn > gpurun_out/f_kbench_attn_fwd4.log 2>&1; cat gpurun_out/f_kbench_attn_fwd4.log
UD_ATTN_FWD=3 python tools/kbench.py attn > gpurun_out/f_kbench_attn_fwd3.log 2>&1; cat gpurun_out/f_kbench_attn_fwd3.log
