#!/bin/bash
# same-box A/B of the whole training step: library before the attention work of this session vs current
mkdir -p gpurun_out
cp unidisc_b200/libunidisc_b200.so /tmp/new.so
run() { timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline $2 > gpurun_out/x_$1.log 2>&1; tail -1 gpurun_out/x_$1.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$1', d['ms_per_step'], d['value'], d['clocks'])"; }
cp tools/_build/libunidisc_b200_old.so unidisc_b200/libunidisc_b200.so; run old_a
cp /tmp/new.so unidisc_b200/libunidisc_b200.so; run new_a
cp tools/_build/libunidisc_b200_old.so unidisc_b200/libunidisc_b200.so; run old_b; run old_interleaved "--workload unidisc-1.4B-interleaved"
cp /tmp/new.so unidisc_b200/libunidisc_b200.so; run new_b; run new_interleaved "--workload unidisc-1.4B-interleaved"
timeout 300 python tools/kbench.py attn > gpurun_out/x_kbench_attn.log 2>&1; tail -2 gpurun_out/x_kbench_attn.log
