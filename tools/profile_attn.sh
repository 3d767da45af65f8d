#!/bin/bash
# GPU box: ncu --set full captures (with source) of single launches of the attention kernels, driven by tools/kbench.py
# usage: bash tools/profile_attn.sh [dkv] [dq] [fwd]
cap() {  # name regex skip
  timeout 300 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:$2" -s $3 -c 1 -f -o gpurun_out/prof_$1 python tools/kbench.py attn > gpurun_out/ncu_$1.log 2>&1
  ncu -i gpurun_out/prof_$1.ncu-rep --page raw --csv > gpurun_out/prof_$1.raw.csv 2>/dev/null
  ncu -i gpurun_out/prof_$1.ncu-rep --page source --csv > gpurun_out/prof_$1.source.csv 2>/dev/null
  rm -f gpurun_out/prof_$1.ncu-rep
}
for w in ${@:-dkv dq fwd}; do
  case $w in
    dkv) cap attn_bwd_dkv 'attn_bwd2_kernel<\(int\)128, \(int\)0>' 3;;
    dq) cap attn_bwd_dq 'attn_bwd2_kernel<\(int\)128, \(int\)1>' 3;;
    fwd) cap attn_fwd 'attn_fwd3_kernel' 3;;
  esac
done
