#!/bin/bash
# Final 8-GPU visit of round 2: one data-parallel bench line (fused NVLink reduce + Adam kernel, early slice overlapped).
mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port $((29600 + RANDOM % 200)) \
  bench.py --gpus 8 --steps 100 --warmup 5 --skip-legs > gpurun_out/r2z_bench_8gpu.json 2> gpurun_out/r2z_bench_8gpu.err
echo "rc=$?"; cut -c1-300 gpurun_out/r2z_bench_8gpu.json; tail -2 gpurun_out/r2z_bench_8gpu.err
