#!/bin/bash
# GPU visit r2b: the rebuilt LSTM recurrence -- tests, timing, full-step parity.
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_lstm.py -q -m gpu -x 2>&1 | tail -15
timeout 300 python profiles/time_lstm.py 2>&1 | grep -v Warn | tail -8
timeout 600 python -m pytest tests/test_gpu_parity.py -q -m gpu -x 2>&1 | tail -6
