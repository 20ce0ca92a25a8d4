#!/bin/bash
# round-2 visit h: what bounds hv_kernel?  TMEM read-rate micro-benchmark + full ncu capture of the new kernels
mkdir -p gpurun_out
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/tmem_ld_rate profiles/micro/tmem_ld_rate.cu && timeout 60 /tmp/tmem_ld_rate | tee gpurun_out/r2h_tmem_ld_rate.txt
NCU="ncu --clock-control none"
timeout 600 $NCU --set full --import-source on -k regex:hv_kernel -s 2 -c 2 -o gpurun_out/r2h_hv_full -f python profiles/prof_coattn.py 2 > gpurun_out/r2h_hv_ncu.log 2>&1; echo "hv rc=$?"
ncu -i gpurun_out/r2h_hv_full.ncu-rep --page raw --csv > gpurun_out/r2h_hv_raw.csv 2>/dev/null
tail -3 gpurun_out/r2h_hv_ncu.log
