"""Per-chunk epilogue timeline of one co-attention gemm_tc launch (index = argv[1] within fwd+bwd: 0..4 fwd, 5..14 bwd)."""
import importlib, sys, os, torch, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
pkg = importlib.import_module("visual-question-answering_b200")
L = pkg._lib.lib()
B, N, T, d = 160, 196, 26, 512
g = torch.Generator().manual_seed(0)
ca = pkg.modules.ParallelCoAttention(d).cuda()
V = torch.randn(B, N, d, generator=g).cuda()
qs = [torch.randn(B, T, d, generator=g).cuda().requires_grad_(True) for _ in range(3)]
def run():
    vhat, qhat = ca.forward_stacked(V, qs)
    (vhat.sum() + qhat.sum()).backward()
    torch.cuda.synchronize()
run(); run()
names = ["buf_free", "aux_landed", "in_regs", "staged", "barrier"]
for idx in [int(a) for a in sys.argv[1:]] or [6]:
    ncta = 148
    buf = torch.zeros(ncta, 64, dtype=torch.int64, device="cuda")
    L.hca_debug_gemm_timeline_select(buf.data_ptr(), ncta, idx)
    run()
    L.hca_debug_gemm_timeline_select(None, 0, -1)
    tt = buf.cpu().numpy().astype(np.int64)
    print(f"=== launch {idx}: CTA lifetime median {np.median(tt[:,6]-tt[:,0]):.0f} cycles; setup {np.median(tt[:,1]-tt[:,0]):.0f}; first_full {np.median(tt[:,2]-tt[:,0]):.0f}; "
          f"mma_issued {np.median(tt[:,3]-tt[:,0]):.0f}; epi_start {np.median(tt[:,4]-tt[:,0]):.0f}; epi_end(first tile) {np.median(tt[:,5]-tt[:,0]):.0f}")
    for c in (0, 77):
        r = tt[c]
        st = [int(x - r[0]) for x in r[8:64] if x]
        print(f" CTA {c}: stamps per chunk (cycles since CTA start; deltas):")
        for k in range(0, len(st) - 4, 5):
            ch = st[k:k + 5]
            print("   ", ch, " d:", [ch[i + 1] - ch[i] for i in range(4)], " next:", (st[k + 5] - ch[4]) if k + 5 < len(st) else None)
