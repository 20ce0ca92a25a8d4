"""Times the two LSTM recurrence kernels (CUDA events around the kernel on its stream, via hca_debug_lstm_events) at the bench shape."""
import importlib, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
sys.argv = [sys.argv[0]]
import bench
pkg = importlib.import_module("visual-question-answering_b200")
dev = torch.device("cuda:0")
pk = bench.peaks()
for B in (160, 320, 1280):
    legs = bench.time_lstm_legs(pkg, dev, 20, pk, B)
    for l in legs:
        print(f"B={B} {l['name']}: {l['ms_per_launch'] * 1e3:.1f} us, {l['achieved']:.1f} TFLOP/s algorithmic ({l['frac']:.3f} of bf16 peak)", flush=True)
