#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x > gpurun_out/ac_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/ac_tests.log; tail -3 gpurun_out/ac_tests.log
timeout 300 python tools/kbench.py rows > gpurun_out/ac_kbench_rows.log 2>&1; grep -i "qk_ln\|norm_res\|colsum" gpurun_out/ac_kbench_rows.log
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/ac_bench.log 2>&1; tail -1 gpurun_out/ac_bench.log | cut -c1-220
