#!/bin/bash
# GPU run B (round 2): the unmodified reference backbone on the GPU (eager, then torch.compile like the reference config),
# attention kernels vs torch SDPA on the cuDNN / flash backends
mkdir -p gpurun_out
python tools/kbench.py attn --vs-cudnn > gpurun_out/b_kbench_attn.log 2>&1; cat gpurun_out/b_kbench_attn.log | tail -6
timeout 600 python bench.py --impl reference --ref-device cuda --steps 30 --warmup 5 > gpurun_out/b_ref_eager.log 2>&1; echo "rc=$?"; tail -2 gpurun_out/b_ref_eager.log | cut -c1-700
timeout 1200 python bench.py --impl reference --ref-device cuda --ref-compile --steps 30 --warmup 5 > gpurun_out/b_ref_compile.log 2>&1; rc=$?; echo "rc=$rc"; tail -2 gpurun_out/b_ref_compile.log | cut -c1-700
if [ $rc -ne 0 ]; then
  timeout 600 python bench.py --impl reference --ref-device cuda --ref-compile --ref-sd3-config 0 --ref-compile-mode default --steps 30 --warmup 5 > gpurun_out/b_ref_compile_default.log 2>&1; echo "rc=$?"; tail -2 gpurun_out/b_ref_compile_default.log | cut -c1-700
fi
