"""Per-chunk epilogue timeline of the PV projection kernel (clock64 stamps recorded by gemm_tc_kernel) under the launcher switches."""
import importlib, sys, os, torch, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
pkg = importlib.import_module("visual-question-answering_b200")
L = pkg._lib.lib()
g = torch.Generator().manual_seed(0)
M, N, K = 160 * 196, 512, 512
Ap = pkg.ops.split_planes(torch.randn(M, K, generator=g).cuda())
Wp = pkg.ops.split_planes((torch.randn(N, K, generator=g) * 0.04).cuda())
b = torch.randn(N, generator=g).cuda()
out = torch.empty(2, M, N, dtype=torch.bfloat16, device="cuda")


def run(label, env):
    for k in ("HCA_TC_EG", "HCA_TC_STAGES", "HCA_TC_PAIR", "HCA_TC_PLDIRECT"):
        os.environ.pop(k, None)
    os.environ.update(env)
    for _ in range(3):
        pkg.ops.proj_planes(Ap, Wp, b, out)
    torch.cuda.synchronize()
    ncta = 148
    buf = torch.zeros(ncta, 64, dtype=torch.int64, device="cuda")
    L.hca_debug_gemm_timeline_select(buf.data_ptr(), ncta, 0)
    pkg.ops.proj_planes(Ap, Wp, b, out)
    torch.cuda.synchronize()
    L.hca_debug_gemm_timeline_select(None, 0, -1)
    tt = buf.cpu().numpy().astype(np.int64)
    med = lambda a: float(np.median(a))
    print(f"=== {label}: CTA lifetime median {med(tt[:,6]-tt[:,0]):.0f} max {float((tt[:,6]-tt[:,0]).max()):.0f} cycles; setup {med(tt[:,1]-tt[:,0]):.0f}; "
          f"first_full {med(tt[:,2]-tt[:,0]):.0f}; tile0 mma_issued {med(tt[:,3]-tt[:,0]):.0f}; epi_start {med(tt[:,4]-tt[:,0]):.0f}; "
          f"epi_end(tile 0) {med(tt[:,5]-tt[:,0]):.0f}")
    for c in (0, 77):
        r = tt[c]
        st = [int(x - r[0]) for x in r[8:64] if x]
        print(f" CTA {c}: [buf_free, aux, in_regs, staged, barrier] cycles since CTA start; deltas; gap to next chunk")
        for k in range(0, len(st) - 4, 5):
            ch = st[k:k + 5]
            print("   ", ch, " d:", [ch[i + 1] - ch[i] for i in range(4)], " next:", (st[k + 5] - ch[4]) if k + 5 < len(st) else None)


run("pair 256x256, 1 group", {})
run("pair, 2 groups", {"HCA_TC_EG": "2"})
run("single 128x128, 1 group", {"HCA_TC_PAIR": "0"})
