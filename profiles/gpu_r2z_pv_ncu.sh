#!/bin/bash
# ncu --set full of the PV projection on the final code (direct plane stores on by default)
mkdir -p gpurun_out
timeout 300 ncu --clock-control none --set full --import-source on -k regex:gemm_tc -s 3 -c 1 -o gpurun_out/r2z_pv_full -f python profiles/prof_pv.py > gpurun_out/r2z_pv_ncu.log 2>&1; echo "pv rc=$?"
ncu -i gpurun_out/r2z_pv_full.ncu-rep --page raw --csv > gpurun_out/r2z_pv_raw.csv 2>/dev/null; wc -l gpurun_out/r2z_pv_raw.csv
