#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/ag_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/ag_tests.log; tail -3 gpurun_out/ag_tests.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/ag_smoke.log 2>&1; tail -2 gpurun_out/ag_smoke.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/ag_bench.log 2>&1; tail -1 gpurun_out/ag_bench.log | cut -c1-200
timeout 600 python bench.py --steps 20 --warmup 5 --full-head --no-cpu-baseline > gpurun_out/ag_bench_full_head.log 2>&1; tail -1 gpurun_out/ag_bench_full_head.log | cut -c1-200
timeout 600 python bench.py --workload unidisc-1.4B-interleaved --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/ag_interleaved.log 2>&1; tail -1 gpurun_out/ag_interleaved.log | cut -c1-200
timeout 600 python bench.py --workload dit-b --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/ag_ditb.log 2>&1; tail -1 gpurun_out/ag_ditb.log | cut -c1-200
timeout 600 python bench.py --workload unidisc-1.4B-sample --predictor ddpm_cache --steps 1 --warmup 1 > gpurun_out/ag_sample_ddpm_cache.log 2>&1; tail -1 gpurun_out/ag_sample_ddpm_cache.log | cut -c1-200
timeout 600 python bench.py --workload unidisc-1.4B-sample --predictor maskgit --steps 1 --warmup 1 > gpurun_out/ag_sample_maskgit.log 2>&1; tail -1 gpurun_out/ag_sample_maskgit.log | cut -c1-200
