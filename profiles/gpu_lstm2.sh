#!/bin/bash
for cfg in "HCA_LSTM_STK=0" "HCA_LSTM_STK=0 HCA_LSTM_NKB_BWD=2" "HCA_LSTM_STK=0 HCA_LSTM_NKB_BWD=1" "HCA_LSTM_STK=0 HCA_LSTM_NKB_FWD=1" "HCA_LSTM_STK=1"; do
  echo "== $cfg"
  env $cfg timeout 120 python - <<'PY' 2>&1 | grep lstm_rec
import importlib, os, sys, torch
sys.path.insert(0, os.getcwd())
pkg = importlib.import_module("visual-question-answering_b200")
B, T, E, H = 160, 26, 512, 512
g = torch.Generator().manual_seed(0)
lens = torch.from_numpy(pkg.synthetic.make_inputs(B, 4, T, 8, 100, 10, seed=1)["lens"]).cuda()
x = torch.randn(B, T, E, generator=g).cuda().requires_grad_(True)
k = 1 / H ** 0.5
w = [((torch.rand(s, generator=g) * 2 - 1) * k).cuda().requires_grad_(True) for s in [(4 * H, E), (4 * H, H), (4 * H,), (4 * H,)]]
dy = torch.randn(B, T, H, generator=g).cuda()
for it in range(3):
    out, _ = pkg.ops.lstm(x, lens, *w); out.backward(dy)
torch.cuda.synchronize()
with torch.profiler.profile(activities=[torch.profiler.ProfilerActivity.CUDA]) as prof:
    for it in range(5):
        out, _ = pkg.ops.lstm(x, lens, *w); out.backward(dy)
    torch.cuda.synchronize()
for ev in prof.key_averages():
    if "lstm_rec" in ev.key: print(ev.key.replace("void hca::(anonymous namespace)::","")[:40], round(ev.device_time_total / ev.count, 1), "us")
PY
done
