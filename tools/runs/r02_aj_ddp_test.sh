#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_ddp_nccl_gpu.py -x -q > gpurun_out/aj_tests_ddp.log 2>&1; echo "tests rc=$?" >> gpurun_out/aj_tests_ddp.log; tail -3 gpurun_out/aj_tests_ddp.log
grep -h "differ" gpurun_out/aj_tests_ddp.log | head -3
