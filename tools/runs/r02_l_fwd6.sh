#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_attention_gpu.py tests/test_model_gpu.py tests/test_interleaved_gpu.py -m gpu -q > gpurun_out/l_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/l_tests.log; tail -4 gpurun_out/l_tests.log
timeout 300 python tools/attn_stress.py 60 > gpurun_out/l_stress.log 2>&1; tail -5 gpurun_out/l_stress.log
for v in 6 5 3; do UD_ATTN_FWD=$v python tools/kbench.py attn > gpurun_out/l_kbench_fwd$v.log 2>&1; echo "fwd variant $v"; head -1 gpurun_out/l_kbench_fwd$v.log; done
