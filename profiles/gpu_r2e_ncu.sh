#!/bin/bash
# GPU visit r2e: ncu evidence -- launch list of one step with DRAM bytes, and --set full captures of the heavy kernels.
mkdir -p gpurun_out
NCU="ncu --clock-control none"
HCA_PDL=0 timeout 900 $NCU --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum -c 1500 --csv --log-file gpurun_out/r2_step_metrics.csv \
  python bench.py --steps 2 --warmup 3 --no-graph --skip-cpu-baseline --skip-gpu-baseline --skip-legs > gpurun_out/r2_step_ncu.log 2>&1; echo "step list rc=$?"
cap() {  # name, kernel regex, skip, count, script
  timeout 600 $NCU --set full --import-source on -k regex:$2 -s $3 -c $4 -o gpurun_out/r2_$1_full -f python $5 > gpurun_out/r2_$1_ncu.log 2>&1; echo "$1 rc=$?"
  ncu -i gpurun_out/r2_$1_full.ncu-rep --page raw --csv > gpurun_out/r2_$1_raw.csv 2>/dev/null
}
cap lstm lstm_rec_kernel 2 2 profiles/prof_lstm.py
cap wgrad gemm_tc 3 1 profiles/prof_wgrad.py
cap pv gemm_tc 3 1 profiles/prof_pv.py
cap hv hv_kernel 2 2 "profiles/prof_coattn.py 2"
ls -la gpurun_out/*.ncu-rep | awk '{print $5, $9}'
