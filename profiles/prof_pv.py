"""ncu --set full target: the roofline kernel of bench.py (PV = V.Wv^T + bv on bf16 hi/lo planes, planes out), a few launches."""
import importlib, sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
pkg = importlib.import_module("visual-question-answering_b200")
g = torch.Generator().manual_seed(0)
M, N, K = 160 * 196, 512, 512
Ap = [pkg.ops.split_planes(torch.randn(M, K, generator=g).cuda()) for _ in range(3)]
Wp = pkg.ops.split_planes((torch.randn(N, K, generator=g) * 0.04).cuda())
b = torch.randn(N, generator=g).cuda()
outs = [torch.empty(2, M, N, dtype=torch.bfloat16, device="cuda") for _ in range(3)]
for i in range(6):
    pkg.ops.proj_planes(Ap[i % 3], Wp, b, outs[i % 3])
torch.cuda.synchronize()
