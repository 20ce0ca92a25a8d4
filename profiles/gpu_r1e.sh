#!/bin/bash
# GPU visit r1e: parity tests, bench line (with CPU baseline), reference arm, ncu launch list of one step, ncu --set full of the PV GEMM.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -3 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 200 --warmup 5 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"
cut -c1-330 gpurun_out/bench.json; tail -3 gpurun_out/bench.err
timeout 300 python bench.py --impl reference --steps 5 --warmup 2 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "ref rc=$?"
cut -c1-200 gpurun_out/bench_ref.json
HCA_PDL=0 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/launches.csv \
  python bench.py --steps 2 --warmup 3 --no-graph --skip-cpu-baseline > gpurun_out/bench_ncu.log 2>&1; echo "ncu rc=$?"
python profiles/summarize_launches.py gpurun_out/launches.csv | head -14
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_tc -s 4 -c 1 -o gpurun_out/pv_gemm_full -f \
  python profiles/prof_pv.py > gpurun_out/pv_ncu.log 2>&1; echo "ncu full rc=$?"
ncu -i gpurun_out/pv_gemm_full.ncu-rep --page raw --csv > gpurun_out/pv_gemm_full_raw.csv 2>/dev/null
