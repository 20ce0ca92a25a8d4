#!/bin/bash
# GPU visit r2f: the round's bench lines (our arm with every leg, the reference arm), the config-2 context number, the full GPU test suite.
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu > gpurun_out/r2f_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/r2f_tests.log
timeout 900 python bench.py --steps 200 --warmup 5 > gpurun_out/r2f_bench.json 2> gpurun_out/r2f_bench.err; echo "bench rc=$?"; cut -c1-260 gpurun_out/r2f_bench.json
timeout 600 python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/r2f_bench_reference_arm.json 2> gpurun_out/r2f_ref.err; echo "ref rc=$?"; cut -c1-200 gpurun_out/r2f_bench_reference_arm.json
timeout 600 python bench.py --scaling strong --steps 30 --warmup 3 --skip-legs --skip-cpu-baseline --skip-gpu-baseline > gpurun_out/r2f_bench_strong_1gpu.json 2> gpurun_out/r2f_strong.err; echo "strong rc=$?"; cut -c1-200 gpurun_out/r2f_bench_strong_1gpu.json
timeout 600 python profiles/baseline_net_context.py 2>/dev/null | tail -1 > gpurun_out/r2f_baseline_net_context.json; cat gpurun_out/r2f_baseline_net_context.json | cut -c1-400
