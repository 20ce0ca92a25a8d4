#!/bin/bash
mkdir -p gpurun_out
timeout 400 python profiles/infer_sweep.py > gpurun_out/infer_sweep.md 2> gpurun_out/infer_sweep.err; echo "sweep rc=$?"
cat gpurun_out/infer_sweep.md; tail -3 gpurun_out/infer_sweep.err
timeout 500 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -x -q -k "small_golden or fused_cross_entropy or flat_gradient_sink" > gpurun_out/sanitizer_memcheck.log 2>&1; echo "memcheck rc=$?"
grep -E "ERROR SUMMARY|passed|failed|Invalid|out of bounds" gpurun_out/sanitizer_memcheck.log | head -10
