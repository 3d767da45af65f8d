#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/aa_tests.log 2>&1; tail -3 gpurun_out/aa_tests.log
timeout 300 python tools/kbench.py rows > gpurun_out/aa_kbench_rows.log 2>&1; grep -i "qk_ln\|norm_res\|colsum" gpurun_out/aa_kbench_rows.log
cp unidisc_b200/libunidisc_b200.so /tmp/new.so
run() { timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline $2 > gpurun_out/aa_$1.log 2>&1; tail -1 gpurun_out/aa_$1.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$1', d['ms_per_step'], d['value'], d['clocks'])"; }
cp tools/_build/libunidisc_b200_old.so unidisc_b200/libunidisc_b200.so; run old_a
cp /tmp/new.so unidisc_b200/libunidisc_b200.so; run new_a
cp tools/_build/libunidisc_b200_old.so unidisc_b200/libunidisc_b200.so; run old_b
cp /tmp/new.so unidisc_b200/libunidisc_b200.so; run new_b
