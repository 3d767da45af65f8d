#!/bin/bash
# GPU box: ncu --set full captures (with source) of single launches of the hot kernels, driven by tools/kbench.py
cap() {  # name regex skip section
  timeout 300 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:$2" -s $3 -c 1 -f -o gpurun_out/prof_$1 python tools/kbench.py $4 > gpurun_out/ncu_$1.log 2>&1
  ncu -i gpurun_out/prof_$1.ncu-rep --page raw --csv > gpurun_out/prof_$1.raw.csv 2>/dev/null
  ncu -i gpurun_out/prof_$1.ncu-rep --page source --csv > gpurun_out/prof_$1.source.csv 2>/dev/null
}
cap gemm_gelu 'gemm2_kernel<\(bool\)0, \(bool\)0, \(int\)256, \(int\)1>' 5 gemm
cap gemm_dgelu 'gemm2_kernel<\(bool\)0, \(bool\)1, \(int\)256, \(int\)2>' 5 gemm
ls -la gpurun_out/*.ncu-rep
