"""Per-k-block timeline (clock64 of the producer and MMA threads) of the split-K weight-gradient kernel dW_v = dPV^T . V.
Needs the debug build (HCA_BUILD_TIMELINE=1) and HCA_TC_DBG=4."""
import importlib, sys, os, torch, numpy as np
os.environ["HCA_TC_DBG"] = "4"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
pkg = importlib.import_module("visual-question-answering_b200")
L = pkg._lib.lib()
g = torch.Generator().manual_seed(0)
K, M, N = 160 * 196, 512, 512
Ap = pkg.ops.split_planes(torch.randn(K, M, generator=g).cuda())
Bp = pkg.ops.split_planes(torch.randn(K, N, generator=g).cuda())
D = torch.zeros(M, N, device="cuda")
st = torch.cuda.current_stream().cuda_stream
run = lambda: pkg._lib.check(L.hca_wgrad_planes(Ap.data_ptr(), Bp.data_ptr(), M, N, K, D.data_ptr(), st), "wgrad")
for _ in range(3):
    run()
torch.cuda.synchronize()
ncta = 148
buf = torch.zeros(ncta, 64, dtype=torch.int64, device="cuda")
L.hca_debug_gemm_timeline_select(buf.data_ptr(), ncta, 0)
run()
torch.cuda.synchronize()
L.hca_debug_gemm_timeline_select(None, 0, -1)
tt = buf.cpu().numpy().astype(np.int64)
life = tt[:, 6] - tt[:, 0]
print(f"CTA lifetime median {np.median(life):.0f} max {life.max()} cycles; setup {np.median(tt[:,1]-tt[:,0]):.0f}; first landed {np.median(tt[:,2]-tt[:,0]):.0f}; "
      f"mma issued (tile 0) {np.median(tt[:,3]-tt[:,0]):.0f}; epi start {np.median(tt[:,4]-tt[:,0]):.0f}; epi end {np.median(tt[:,5]-tt[:,0]):.0f}")
for c in (0, 2, 76):
    r = tt[c]
    print(f"CTA {c} (SM {r[7]}): k-block: producer-issue, landed(MMA passed full), mma-issued   [cycles since CTA start]")
    for it in range(16):
        if r[8 + it] or r[24 + it]:
            print(f"   kb {it:2d}: {int(r[8+it]-r[0]) if r[8+it] else -1:7d} {int(r[24+it]-r[0]) if r[24+it] else -1:7d} {int(r[40+it]-r[0]) if r[40+it] else -1:7d}")
