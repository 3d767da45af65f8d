#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_attention_gpu.py tests/test_model_gpu.py tests/test_interleaved_gpu.py -m gpu -q > gpurun_out/h_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/h_tests.log; tail -5 gpurun_out/h_tests.log
timeout 300 python tools/attn_stress.py 60 > gpurun_out/h_stress.log 2>&1; tail -5 gpurun_out/h_stress.log
python tools/kbench.py attn sampler > gpurun_out/h_kbench_fwd5.log 2>&1; cat gpurun_out/h_kbench_fwd5.log
UD_ATTN_FWD=3 python tools/kbench.py attn > gpurun_out/h_kbench_fwd3.log 2>&1; cat gpurun_out/h_kbench_fwd3.log
