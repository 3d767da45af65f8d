#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x > gpurun_out/ad_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/ad_tests.log; tail -3 gpurun_out/ad_tests.log
run() { timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline $2 > gpurun_out/ad_$1.log 2>&1; tail -1 gpurun_out/ad_$1.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$1', d['ms_per_step'], d['value'], d['e2e']['value'], d['head_rows_fraction'], d['e2e']['loss'], d['clocks']['sm_mhz'])"; }
run full_a --full-head
run masked_a
run full_b --full-head
run masked_b
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/ad_smoke.log 2>&1; tail -2 gpurun_out/ad_smoke.log
