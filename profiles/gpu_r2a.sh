#!/bin/bash
# GPU visit r2a (2 GPUs): the new parity / training / DP tests, the old suite, symmetric-memory probe, bench at N=1 and N=2 (fused and NCCL).
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r2a_gpus.txt 2>&1
timeout 1500 python -m pytest tests/test_gpu_headline.py -q -m gpu -s > gpurun_out/r2a_tests_headline.log 2>&1; echo "headline rc=$?"
tail -5 gpurun_out/r2a_tests_headline.log
timeout 900 python -m pytest tests/test_gpu_training.py -q -m gpu -s -k "not two_rank" > gpurun_out/r2a_tests_training.log 2>&1; echo "training rc=$?"
tail -5 gpurun_out/r2a_tests_training.log
timeout 120 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 profiles/probe_symm.py > gpurun_out/r2a_probe.log 2>&1; echo "probe rc=$?"
grep -a "multicast" gpurun_out/r2a_probe.log | cut -c1-400
timeout 600 python -m pytest tests/test_gpu_training.py -q -m gpu -s -k "two_rank" > gpurun_out/r2a_tests_dp.log 2>&1; echo "dp rc=$?"
tail -8 gpurun_out/r2a_tests_dp.log | cut -c1-600
timeout 900 python -m pytest tests -q -m gpu --deselect tests/test_gpu_headline.py --deselect tests/test_gpu_training.py > gpurun_out/r2a_tests_old.log 2>&1; echo "old rc=$?"
tail -4 gpurun_out/r2a_tests_old.log
timeout 900 python bench.py --steps 100 --warmup 5 > gpurun_out/r2a_bench.json 2> gpurun_out/r2a_bench.err; echo "bench rc=$?"
cut -c1-400 gpurun_out/r2a_bench.json; tail -3 gpurun_out/r2a_bench.err
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 100 --warmup 5 --skip-legs \
  > gpurun_out/r2a_bench_2gpu.json 2> gpurun_out/r2a_bench_2gpu.err; echo "bench2 rc=$?"
cut -c1-300 gpurun_out/r2a_bench_2gpu.json; tail -3 gpurun_out/r2a_bench_2gpu.err
HCA_DP_FUSED=0 timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 100 --warmup 5 --skip-legs \
  > gpurun_out/r2a_bench_2gpu_nccl.json 2> gpurun_out/r2a_bench_2gpu_nccl.err; echo "bench2 nccl rc=$?"
cut -c1-300 gpurun_out/r2a_bench_2gpu_nccl.json
