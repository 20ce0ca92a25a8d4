#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_lstm.py -m gpu -x -q > gpurun_out/pytest_lstm.log 2>&1; echo "pytest rc=$?"
tail -6 gpurun_out/pytest_lstm.log | cut -c1-300
timeout 120 python - <<'PY'
import importlib, os, sys, torch
sys.path.insert(0, os.getcwd())
pkg = importlib.import_module("visual-question-answering_b200")
B, T, E, H = 160, 26, 512, 512
g = torch.Generator().manual_seed(0)
syn = pkg.synthetic
lens = torch.from_numpy(syn.make_inputs(B, 4, T, 8, 100, 10, seed=1)["lens"]).cuda()
x = torch.randn(B, T, E, generator=g).cuda().requires_grad_(True)
k = 1 / H ** 0.5
w = [((torch.rand(s, generator=g) * 2 - 1) * k).cuda().requires_grad_(True) for s in [(4 * H, E), (4 * H, H), (4 * H,), (4 * H,)]]
dy = torch.randn(B, T, H, generator=g).cuda()
res = {}
for mode in ("1", "0"):
    os.environ["HCA_LSTM_STK"] = mode
    for it in range(3):
        out, _ = pkg.ops.lstm(x, lens, *w)
        out.backward(dy)
    torch.cuda.synchronize()
    with torch.profiler.profile(activities=[torch.profiler.ProfilerActivity.CUDA]) as prof:
        for it in range(5):
            for t in [x] + w: t.grad = None
            out, _ = pkg.ops.lstm(x, lens, *w)
            out.backward(dy)
        torch.cuda.synchronize()
    for ev in prof.key_averages():
        if "lstm_rec" in ev.key: print("STK=" + mode, ev.key[:60], round(ev.device_time_total / ev.count, 1), "us")
    res[mode] = [t.grad.clone() for t in [x] + w]
print("max |grad diff| stacked vs plain:", max(float((a - b).abs().max()) for a, b in zip(res["1"], res["0"])))
PY
