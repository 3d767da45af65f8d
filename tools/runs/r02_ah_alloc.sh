#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_model_gpu.py -q -x > gpurun_out/ah_tests.log 2>&1; tail -2 gpurun_out/ah_tests.log
run() { timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline $2 > gpurun_out/ah_$1.log 2>&1; tail -1 gpurun_out/ah_$1.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$1', d['ms_per_step'], d['value'], d['e2e']['ms_per_step'], d['e2e']['value'], d['clocks']['sm_mhz'])"; }
run split_a
run full_a --full-head
run split_b
