#!/bin/bash
# GPU run C: new parity tests (attention caching, ...) + the reference backbone under torch.compile(max-autotune-no-cudagraphs)
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "caching or throughput or maskgit or sample" > gpurun_out/c_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/c_tests.log; tail -4 gpurun_out/c_tests.log
timeout 1500 python bench.py --impl reference --ref-device cuda --ref-compile --steps 30 --warmup 5 > gpurun_out/c_ref_compile_maxautotune.log 2>&1; rc=$?; echo "rc=$rc"; grep -v "Rank:0" gpurun_out/c_ref_compile_maxautotune.log | tail -2 | cut -c1-900
if [ $rc -ne 0 ]; then
  timeout 900 python bench.py --impl reference --ref-device cuda --ref-compile --ref-sd3-config 0 --steps 30 --warmup 5 > gpurun_out/c_ref_compile_maxautotune_nosd3.log 2>&1; echo "rc=$?"; grep -v "Rank:0" gpurun_out/c_ref_compile_maxautotune_nosd3.log | tail -2 | cut -c1-900
fi
