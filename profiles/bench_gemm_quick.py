import importlib, sys, os, torch, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
pkg = importlib.import_module("visual-question-answering_b200")
L = pkg._lib.lib()
g = torch.Generator().manual_seed(0)
M, N, K = 31360, 512, 512
A = torch.randn(M, K, generator=g).cuda(); W = torch.randn(N, K, generator=g).cuda(); b = torch.randn(N, generator=g).cuda()
ncta = 600
buf = torch.zeros(ncta, 64, dtype=torch.int64, device="cuda")
bb = b if os.environ.get("NOBIAS") is None else None
pkg.ops.gemm(A, W, bb, "nt", 1)
L.hca_debug_gemm_timeline(buf.data_ptr(), ncta)
pkg.ops.gemm(A, W, bb, "nt", 1)
torch.cuda.synchronize()
L.hca_debug_gemm_timeline(None, 0)
tt = buf.cpu().numpy().astype(np.int64)
print("DBG", os.environ.get("HCA_TC_DBG"), "mainloop median", np.median(tt[:, 3] - tt[:, 2]), "epilogue median", np.median(tt[:, 5] - tt[:, 4]))
r = tt[300]
print("  landed", [int(x - r[0]) for x in r[24:32]]); print("  issued", [int(x - r[0]) for x in r[40:48]])
print("  epilogue stamps (chunk start, after tmem ld, after smem+barrier) rel epi_start:", [int(x - r[4]) for x in r[52:64]], "epi_end", int(r[5]-r[4]))
