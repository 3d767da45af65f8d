#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/ak_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/ak_tests.log; tail -3 gpurun_out/ak_tests.log
grep -E "FAILED|ERROR" gpurun_out/ak_tests.log | head -10
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/ak_bench.log 2>&1; tail -1 gpurun_out/ak_bench.log | cut -c1-200
