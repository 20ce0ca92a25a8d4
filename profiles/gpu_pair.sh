#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_gemm.py -m gpu -x -q > gpurun_out/pytest_pair.log 2>&1; echo "pytest rc=$?"
tail -5 gpurun_out/pytest_pair.log | cut -c1-400
timeout 120 python profiles/time_pv.py 2>&1 | tail -8
timeout 120 python profiles/timeline_pv.py 2>&1 | grep -A14 "===" | cut -c1-200 | head -60
