#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_kernels_gpu.py tests/test_model_gpu.py tests/test_timecond_gpu.py -x -q > gpurun_out/z_tests.log 2>&1; tail -3 gpurun_out/z_tests.log
timeout 300 python tools/kbench.py rows > gpurun_out/z_kbench_rows_tma.log 2>&1; grep -i "qk_ln\|norm_res" gpurun_out/z_kbench_rows_tma.log
UD_NORM_BWD=1 timeout 300 python tools/kbench.py rows > gpurun_out/z_kbench_rows_old.log 2>&1; grep -i "norm_residual_bwd" gpurun_out/z_kbench_rows_old.log
