"""Per-step timeline (clock64 of CTA 0) of the persistent LSTM recurrence kernels at the headline shape."""
import importlib, sys, os, torch, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
pkg = importlib.import_module("visual-question-answering_b200")
L = pkg._lib.lib()
B, T, E, H = 160, 26, 512, 512
g = torch.Generator().manual_seed(0)
syn = pkg.synthetic
lens = torch.from_numpy(syn.make_inputs(B, 4, T, 8, 100, 10, seed=1)["lens"]).cuda()
x = torch.randn(B, T, E, generator=g).cuda().requires_grad_(True)
k = 1 / H ** 0.5
w = [((torch.rand(s, generator=g) * 2 - 1) * k).cuda().requires_grad_(True) for s in [(4 * H, E), (4 * H, H), (4 * H,), (4 * H,)]]
dy = torch.randn(B, T, H, generator=g).cuda()
names = ["cnt_seen", "tma_issued", "first_landed", "mma_issued", "acc_seen", "cell_done", "barrier", "published"]
for which in ("fwd", "bwd"):
    for it in range(3):
        buf = torch.zeros(7 * 8, dtype=torch.int64, device="cuda")
        out, _ = pkg.ops.lstm(x, lens, *w)
        if which == "fwd" and it == 2:
            L.hca_debug_lstm_timeline(buf.data_ptr())
            out, _ = pkg.ops.lstm(x, lens, *w)
            L.hca_debug_lstm_timeline(None)
        if which == "bwd" and it == 2:
            L.hca_debug_lstm_timeline(buf.data_ptr())
        out.backward(dy)
        L.hca_debug_lstm_timeline(None)
        torch.cuda.synchronize()
    tt = buf.cpu().numpy().reshape(7, 8)
    print(f"=== {which}: stamps relative to round 0 'cnt_seen' (cycles)")
    print("round " + " ".join(f"{n:>12s}" for n in names))
    for n in range(7):
        print(f"{n:5d} " + " ".join(f"{int(v - tt[0, 0]):12d}" for v in tt[n]))
