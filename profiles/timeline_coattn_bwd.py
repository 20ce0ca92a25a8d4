"""Per-k-block / per-tile timeline of the transposed-tile products of the co-attention backward (debug build + HCA_TC_DBG=4).
gemm_tc launches inside hca_coattn_bwd, in order: 0 dZq, 1 dPQ^T, 2 dS^T, 3 dQ, 4 dWq, 5 dWv."""
import importlib, sys, os, torch, numpy as np
os.environ["HCA_TC_DBG"] = "4"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pkg = importlib.import_module("visual-question-answering_b200")
L = pkg._lib.lib()
syn = pkg.synthetic
d, N, T, vocab, K, mlp, B = 512, 196, 26, 10000, 1001, 1024, 160
p = syn.make_params(d, vocab, K, mlp, seed=0)
x = syn.make_inputs(B, N, T, d, vocab, K, seed=1)
net = pkg.HieCoAttnHotPath(vocab, d, K, mlp)
net.load_state_dict({k: torch.from_numpy(v) for k, v in p.items()}, strict=False)
net.cuda()
feats, tokens, labels = (torch.from_numpy(x[k]).cuda() for k in ("feats", "tokens", "labels"))
lens = pkg.QuestionLens(torch.from_numpy(x["lens"]), "cuda")
names = {0: "dZq", 1: "dPQ^T", 2: "dS^T", 3: "dQ"}
for which in (1, 2, 0, 3):
    for it in range(2):
        net.zero_grad(set_to_none=True)
        hier = net.question_encoder(tokens, lens)
        vhat, qhat = net.co_attention.forward_stacked(feats, hier)
        loss = (vhat.sum() + qhat.sum())
        torch.cuda.synchronize()
        ncta = 148
        buf = torch.zeros(ncta, 64, dtype=torch.int64, device="cuda")
        if it == 1:
            L.hca_debug_gemm_timeline_select(buf.data_ptr(), ncta, which)
        loss.backward()
        torch.cuda.synchronize()
        L.hca_debug_gemm_timeline_select(None, 0, -1)
    tt = buf.cpu().numpy().astype(np.int64)
    life = tt[:, 6] - tt[:, 0]
    print(f"=== {names[which]}: CTA lifetime median {np.median(life):.0f} min {life.min()} max {life.max()} cycles; setup {np.median(tt[:,1]-tt[:,0]):.0f}; "
          f"tile0: first landed {np.median(tt[:,2]-tt[:,0]):.0f}, mma issued {np.median(tt[:,3]-tt[:,0]):.0f}, epi start {np.median(tt[:,4]-tt[:,0]):.0f}, epi end {np.median(tt[:,5]-tt[:,0]):.0f}")
    r = tt[3]
    print("   CTA 3 k-blocks of tile 0 [producer issue, landed, mma issued]:", [(int(r[8+i]-r[0]), int(r[24+i]-r[0]), int(r[40+i]-r[0])) for i in range(8) if r[8+i]])
