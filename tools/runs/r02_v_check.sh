#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x > gpurun_out/v_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/v_tests.log; tail -3 gpurun_out/v_tests.log
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/v_bench.log 2>&1; tail -1 gpurun_out/v_bench.log | cut -c1-200
UD_ATTN_BWD=2 timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/v_bench_bwd2.log 2>&1; tail -1 gpurun_out/v_bench_bwd2.log | cut -c1-200
timeout 600 python bench.py --workload unidisc-1.4B-interleaved --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/v_interleaved.log 2>&1; tail -1 gpurun_out/v_interleaved.log | cut -c1-200
timeout 600 python bench.py --workload dit-b --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/v_ditb.log 2>&1; tail -1 gpurun_out/v_ditb.log | cut -c1-200
