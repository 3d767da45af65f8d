#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/ab_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/ab_tests.log; tail -3 gpurun_out/ab_tests.log
python tools/attn_stress.py 50 > gpurun_out/ab_stress.log 2>&1; tail -5 gpurun_out/ab_stress.log | cut -c1-160
python tools/kbench.py attn rows roof > gpurun_out/ab_kbench.log 2>&1; cat gpurun_out/ab_kbench.log
bash tools/profile_round.sh r02b nosampler > gpurun_out/ab_profile.log 2>&1; tail -10 gpurun_out/ab_profile.log
