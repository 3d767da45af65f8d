#!/bin/bash
# Run on the GPU box (under gpurun): launch list of one bench step + ncu --set full capture of bench.py's roofline kernel.
# Outputs (CSV only, the .ncu-rep files are dropped) in gpurun_out/.
B="python bench.py --steps 1 --warmup 3 --no-cpu-baseline"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 8000 --csv --log-file gpurun_out/launches.csv $B > gpurun_out/ncu_bench.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:gemm2_kernel" -s 3 -c 1 -f -o gpurun_out/prof_gemm_roof python tools/kbench.py roof > gpurun_out/ncu_gemm_roof.log 2>&1
ncu -i gpurun_out/prof_gemm_roof.ncu-rep --page raw --csv > gpurun_out/prof_gemm_roof.raw.csv 2>/dev/null
rm -f gpurun_out/prof_gemm_roof.ncu-rep
tail -2 gpurun_out/ncu_gemm_roof.log
wc -l gpurun_out/launches.csv
