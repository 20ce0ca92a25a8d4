#!/bin/bash
# Final 2-GPU visit of round 2: the 2-rank data-parallel parity tests (both transports, early slice) and one bench line.
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_training.py -q -m gpu -x -s -k two_rank 2>&1 | grep -E "2-rank|passed|failed|Error" | cut -c1-260 | tee gpurun_out/r2z_tests_two_rank_dp.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $((29600 + RANDOM % 200)) \
  bench.py --gpus 2 --steps 100 --warmup 5 --skip-legs > gpurun_out/r2z_bench_2gpu.json 2> gpurun_out/r2z_bench_2gpu.err
echo "rc=$?"; cut -c1-300 gpurun_out/r2z_bench_2gpu.json; tail -2 gpurun_out/r2z_bench_2gpu.err
