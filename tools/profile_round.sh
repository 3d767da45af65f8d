#!/bin/bash
# Run on the GPU box (under gpurun): launch list + ncu --set full captures of the hot kernels. Outputs in gpurun_out/.
set -x
B="python bench.py --steps 1 --warmup 3 --no-cpu-baseline"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 8000 --csv --log-file gpurun_out/launches.csv $B > gpurun_out/ncu_bench.log 2>&1
for spec in "gemm2:gemm2_kernel:300:3" "attnbwd:attn_bwd_kernel:40:2" "attnfwd:attn_fwd_kernel:30:1" "sampler:ddpm_update_logits_fast_kernel:1:1" "normbwd:norm_residual_bwd_kernel:60:1" "qkln:qk_ln_rope_fwd_kernel:30:1" "nll:subs_nll:2:2"; do
  IFS=: read name pat skip cnt <<< "$spec"
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$pat -s $skip -c $cnt -f -o gpurun_out/prof_$name $B > gpurun_out/ncu_$name.log 2>&1
done
ls -la gpurun_out/*.ncu-rep
