#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_attention_gpu.py tests/test_model_gpu.py tests/test_interleaved_gpu.py -m gpu -q > gpurun_out/m_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/m_tests.log; tail -4 gpurun_out/m_tests.log
timeout 300 python tools/attn_stress.py 100 > gpurun_out/m_stress.log 2>&1; tail -5 gpurun_out/m_stress.log
python tools/kbench.py attn > gpurun_out/m_kbench_dq3.log 2>&1; cat gpurun_out/m_kbench_dq3.log
UD_ATTN_BWD_DQ=2 python tools/kbench.py attn > gpurun_out/m_kbench_dq2.log 2>&1; cat gpurun_out/m_kbench_dq2.log
