#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py -x -q -k "qk_ln" > gpurun_out/y_tests.log 2>&1; tail -3 gpurun_out/y_tests.log
timeout 300 python tools/kbench.py rows > gpurun_out/y_kbench_rows_tma.log 2>&1; grep -i "qk_ln\|norm_res" gpurun_out/y_kbench_rows_tma.log
UD_QKLN_BWD=1 timeout 300 python tools/kbench.py rows > gpurun_out/y_kbench_rows_old.log 2>&1; grep -i "qk_ln" gpurun_out/y_kbench_rows_old.log
