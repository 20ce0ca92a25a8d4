"""Timeline of the classifier's transposed-tile products (debug build + HCA_TC_DBG=4): where do ~12 us per launch go?
gemm_tc launches inside hca_mlp_fwd, in order: 0 W_w, 1 W_p, 2 W_s, 3 W_h.  Also CUDA-event timing of the whole forward."""
import importlib, sys, os, torch, numpy as np
os.environ["HCA_TC_DBG"] = "4"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pkg = importlib.import_module("visual-question-answering_b200")
L = pkg._lib.lib()
B, d, mlp, K = 160, 512, 1024, 1001
g = torch.Generator().manual_seed(0)
net = pkg.modules.MLPClassifier(d, mlp, K).cuda()
vhat = torch.randn(3, B, d, generator=g).cuda().requires_grad_(True)
qhat = torch.randn(3, B, d, generator=g).cuda().requires_grad_(True)
for which in range(4):
    for it in range(2):
        ncta = 148
        buf = torch.zeros(ncta, 64, dtype=torch.int64, device="cuda")
        torch.cuda.synchronize()
        if it == 1:
            L.hca_debug_gemm_timeline_select(buf.data_ptr(), ncta, which)
        out = net.forward_stacked(vhat, qhat)
        torch.cuda.synchronize()
        L.hca_debug_gemm_timeline_select(None, 0, -1)
    tt = buf.cpu().numpy().astype(np.int64)
    tt = tt[tt[:, 0] != 0]
    life = tt[:, 6] - tt[:, 0]
    print(f"=== fwd gemm {which}: {len(tt)} CTAs; lifetime median {np.median(life):.0f} max {life.max()} cycles; setup {np.median(tt[:,1]-tt[:,0]):.0f}; "
          f"first landed {np.median(tt[:,2]-tt[:,0]):.0f}, mma issued {np.median(tt[:,3]-tt[:,0]):.0f}, epi start {np.median(tt[:,4]-tt[:,0]):.0f}, epi end {np.median(tt[:,5]-tt[:,0]):.0f}")
    r = tt[0]
    print("   CTA 0 k-blocks [producer issue, landed, mma issued]:", [(int(r[8+i]-r[0]), int(r[24+i]-r[0]), int(r[40+i]-r[0])) for i in range(16) if r[8+i]])
ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
for it in range(3):
    ev[0].record()
    for _ in range(20):
        out = net.forward_stacked(vhat, qhat)
    ev[1].record()
    torch.cuda.synchronize()
    print(f"mlp forward (eager, PDL on): {ev[0].elapsed_time(ev[1]) / 20 * 1e3:.1f} us")
