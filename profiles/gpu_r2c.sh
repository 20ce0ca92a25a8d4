#!/bin/bash
# GPU visit r2c: full GPU suite + bench after the LSTM rebuild.
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu -x --deselect tests/test_gpu_headline.py::test_argmax_agreement_on_10240_samples > gpurun_out/r2c_tests.log 2>&1; echo "tests rc=$?"
tail -4 gpurun_out/r2c_tests.log
timeout 900 python bench.py --steps 100 --warmup 5 > gpurun_out/r2c_bench.json 2> gpurun_out/r2c_bench.err; echo "bench rc=$?"
cut -c1-330 gpurun_out/r2c_bench.json; tail -3 gpurun_out/r2c_bench.err
