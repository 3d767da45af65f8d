#!/bin/bash
# GPU run A (round 2): parity tests, cfg4 sampling (ddpm_cache, maskgit), cfg5 interleaved, default bench
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/a_build.log 2>&1
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/a_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/a_tests.log
tail -5 gpurun_out/a_tests.log
timeout 600 python bench.py --workload unidisc-1.4B-sample --predictor ddpm_cache --steps 1 --warmup 1 > gpurun_out/a_sample_ddpm_cache.log 2>&1; tail -1 gpurun_out/a_sample_ddpm_cache.log | cut -c1-600
timeout 600 python bench.py --workload unidisc-1.4B-sample --predictor maskgit --steps 1 --warmup 1 > gpurun_out/a_sample_maskgit.log 2>&1; tail -1 gpurun_out/a_sample_maskgit.log | cut -c1-600
timeout 600 python bench.py --workload unidisc-1.4B-interleaved --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/a_interleaved.log 2>&1; tail -1 gpurun_out/a_interleaved.log | cut -c1-600
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/a_bench.log 2>&1; tail -1 gpurun_out/a_bench.log | cut -c1-400
