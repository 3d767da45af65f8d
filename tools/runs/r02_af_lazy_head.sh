#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x > gpurun_out/af_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/af_tests.log; tail -4 gpurun_out/af_tests.log
run() { timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline $2 > gpurun_out/af_$1.log 2>&1; tail -1 gpurun_out/af_$1.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$1', d['ms_per_step'], d['value'], d['e2e']['ms_per_step'], d['e2e']['value'], d['head_rows_fraction'], d.get('head_cols_per_row'), d['clocks']['sm_mhz'])"; }
run full_a --full-head
run split_a
run nosplit_a --no-split-head
run full_b --full-head
run split_b
