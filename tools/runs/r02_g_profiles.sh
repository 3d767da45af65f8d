#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/g_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/g_tests.log; tail -5 gpurun_out/g_tests.log
python tools/attn_stress.py 100 > gpurun_out/g_stress.log 2>&1; tail -5 gpurun_out/g_stress.log
python tools/kbench.py attn rows sampler roof > gpurun_out/g_kbench.log 2>&1; cat gpurun_out/g_kbench.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/g_bench.log 2>&1; tail -1 gpurun_out/g_bench.log | cut -c1-250
bash tools/profile_round.sh r02 > gpurun_out/g_profile.log 2>&1; tail -8 gpurun_out/g_profile.log
