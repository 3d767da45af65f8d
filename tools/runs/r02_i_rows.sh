#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py tests/test_model_gpu.py -m gpu -q > gpurun_out/i_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/i_tests.log; tail -3 gpurun_out/i_tests.log
python tools/kbench.py rows > gpurun_out/i_kbench_rows.log 2>&1; cat gpurun_out/i_kbench_rows.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/i_bench.log 2>&1; tail -1 gpurun_out/i_bench.log | cut -c1-200
