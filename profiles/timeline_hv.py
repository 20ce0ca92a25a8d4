"""Per-tile timeline of hv_kernel (debug build: HCA_BUILD_TIMELINE=1): clock64 stamps of CTA 0 for its first tiles.
    HCA_BUILD_TIMELINE=1 python visual-question-answering_b200/build.py --force && python profiles/timeline_hv.py
Stamps per tile: 0/1 producer (first / last k-block about to be issued), 2 MMA warp saw tmem_empty, 3 / 4 first / last k-block landed,
5 tile committed, 6 epilogue warp 0 saw tmem_full, 7 its PV slice had landed, 8 its chunk(s) done."""
import importlib, sys, os, torch, numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pkg = importlib.import_module("visual-question-answering_b200")
L = pkg._lib.lib()
B, N, T, d = 160, 196, 26, 512
g = torch.Generator().manual_seed(0)
ca = pkg.modules.ParallelCoAttention(d).cuda()
V = torch.randn(B, N, d, generator=g).cuda()
qs = [torch.randn(B, T, d, generator=g).cuda().requires_grad_(True) for _ in range(3)]
for phase in ("fwd", "bwd"):
    for it in range(2):
        buf = torch.zeros(148, 64, dtype=torch.int64, device="cuda")
        if phase == "fwd" and it == 1:
            L.hca_debug_gemm_timeline_select(buf.data_ptr(), 148, 9999)      # (no gemm_tc launch matches: only hv_kernel writes)
        vhat, qhat = ca.forward_stacked(V, qs)
        torch.cuda.synchronize()
        if phase == "fwd":
            L.hca_debug_gemm_timeline_select(None, 0, -1)
        if phase == "bwd" and it == 1:
            L.hca_debug_gemm_timeline_select(buf.data_ptr(), 148, 9999)      # (no gemm_tc launch matches: only hv_kernel writes)
        (vhat.sum() + qhat.sum()).backward()
        torch.cuda.synchronize()
        L.hca_debug_gemm_timeline_select(None, 0, -1)
    tt = buf.cpu().numpy().astype(np.int64).reshape(-1)[:12 * 16].reshape(12, 16)
    t0 = tt[0, 0]
    print(f"=== hv {phase}: CTA 0, cycles since its first load (producer first/last issue | mma: empty seen, first landed, last landed, committed | epi: full seen, aux seen, done)")
    for i in range(9):
        r = tt[i]
        if r[2] == 0:
            break
        print(f"  tile {i}: prod {r[0]-t0:7d} {r[1]-t0:7d} | mma {r[2]-t0:7d} {r[3]-t0:7d} {r[4]-t0:7d} {r[5]-t0:7d} | epi {r[6]-t0:7d} {r[7]-t0:7d} {r[8]-t0:7d}   (mma phase {r[6]-r[2]}, epilogue {r[8]-r[6]})")
