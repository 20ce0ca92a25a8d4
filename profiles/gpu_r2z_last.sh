#!/bin/bash
# Last visit: smoke() of __graft_entry__, default-argument bench run, ncu launch list of the final code.
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 600 python bench.py > gpurun_out/r2z_bench_default_args.json 2> gpurun_out/r2z_bench_default.err; echo "default bench rc=$?"; cut -c1-230 gpurun_out/r2z_bench_default_args.json
bash profiles/gpu_r2n_ncu.sh
