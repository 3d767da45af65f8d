#!/bin/bash
mkdir -p gpurun_out
python tools/attn_stress.py 200 > gpurun_out/e_stress_fast.log 2>&1; cat gpurun_out/e_stress_fast.log | tail -6
UD_ATTN_BWD_SAFE=1 python tools/attn_stress.py 200 > gpurun_out/e_stress_safe.log 2>&1; cat gpurun_out/e_stress_safe.log | tail -6
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/e_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/e_tests.log; tail -6 gpurun_out/e_tests.log
