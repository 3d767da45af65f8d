#!/bin/bash
# Run on the GPU box (under gpurun): launch list of one bench step + ncu --set full captures of bench.py's roofline kernel
# (the mlp.0 GEMM with bias + GELU epilogue), the attention kernels and the vocabulary-pass sampler.
# Outputs (CSV only, the .ncu-rep files are dropped) in gpurun_out/.   usage: bash tools/profile_round.sh [tag]
TAG=${1:-r02}
B="python bench.py --steps 1 --warmup 3 --no-cpu-baseline"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 8000 --csv --log-file gpurun_out/${TAG}_launches.csv $B > gpurun_out/${TAG}_ncu_bench.log 2>&1
cap() {  # name regex skip kbench-section
  timeout 300 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:$2" -s $3 -c 1 -f -o gpurun_out/prof_$1 python tools/kbench.py $4 > gpurun_out/${TAG}_ncu_$1.log 2>&1
  ncu -i gpurun_out/prof_$1.ncu-rep --page raw --csv > gpurun_out/${TAG}_ncu_full_$1.csv 2>/dev/null
  ncu -i gpurun_out/prof_$1.ncu-rep --page source --csv > gpurun_out/${TAG}_ncu_source_$1.csv 2>/dev/null
  rm -f gpurun_out/prof_$1.ncu-rep
}
cap gemm_roof 'gemm2_kernel<\(bool\)0, \(bool\)0, \(int\)256, \(int\)1>' 3 roof
cap attn_fwd 'attn_fwd6_kernel' 3 attn
cap attn_bwd_dq 'attn_bwd3_kernel<\(int\)128, \(int\)1>' 3 attn
cap attn_bwd_dkv 'attn_bwd2_kernel<\(int\)128, \(int\)0>' 3 attn
cap qkln_bwd 'qk_ln_rope_bwd_tma_kernel' 2 rows
cap norm_bwd 'norm_residual_bwd_tma_kernel<\(int\)2, \(bool\)1' 2 rows
[ "$2" = "nosampler" ] || cap sampler 'ddpm_update_logits_fast_kernel' 2 sampler
wc -l gpurun_out/${TAG}_launches.csv
ls -la gpurun_out/${TAG}_ncu_full_*.csv
