"""ncu target: a few forward + backward passes of the sentence LSTM at the headline shape."""
import importlib, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
pkg = importlib.import_module("visual-question-answering_b200")
B, T, E, H = 160, 26, 512, 512
g = torch.Generator().manual_seed(0)
lens = torch.from_numpy(pkg.synthetic.make_inputs(B, 4, T, 8, 100, 10, seed=1)["lens"]).cuda()
x = torch.randn(B, T, E, generator=g).cuda().requires_grad_(True)
k = 1 / H ** 0.5
w = [((torch.rand(s, generator=g) * 2 - 1) * k).cuda().requires_grad_(True) for s in [(4 * H, E), (4 * H, H), (4 * H,), (4 * H,)]]
dy = torch.randn(B, T, H, generator=g).cuda()
for it in range(3):
    out, _ = pkg.ops.lstm(x, lens, *w)
    out.backward(dy)
torch.cuda.synchronize()
