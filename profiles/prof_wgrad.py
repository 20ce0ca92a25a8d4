"""ncu --set full target: the largest weight-gradient product (dW_v = dPV^T . V, M = N = 512, K = 160 * 196) on operand planes, a few launches."""
import importlib, sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
pkg = importlib.import_module("visual-question-answering_b200")
L = pkg._lib.lib()
g = torch.Generator().manual_seed(0)
K, M, N = 160 * 196, 512, 512
Ap = [pkg.ops.split_planes(torch.randn(K, M, generator=g).cuda()) for _ in range(3)]
Bp = [pkg.ops.split_planes(torch.randn(K, N, generator=g).cuda()) for _ in range(3)]
D = torch.zeros(M, N, device="cuda")
st = torch.cuda.current_stream().cuda_stream
for i in range(6):
    pkg._lib.check(L.hca_wgrad_planes(Ap[i % 3].data_ptr(), Bp[i % 3].data_ptr(), M, N, K, D.data_ptr(), st), "wgrad")
torch.cuda.synchronize()
