#!/bin/bash
# ncu launch list of one step with DRAM bytes (final code of round 2)
mkdir -p gpurun_out
NCU="ncu --clock-control none"
HCA_PDL=0 timeout 900 $NCU --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum -c 1500 --csv --log-file gpurun_out/r2n_step_metrics.csv \
  python bench.py --steps 2 --warmup 3 --no-graph --skip-cpu-baseline --skip-gpu-baseline --skip-legs > gpurun_out/r2n_step_ncu.log 2>&1; echo "step list rc=$?"
