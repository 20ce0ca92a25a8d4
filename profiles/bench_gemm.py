"""Microbenchmark + per-CTA timeline of the tensor-core GEMM (run under gpurun)."""
import importlib, sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
pkg = importlib.import_module("visual-question-answering_b200")
L = pkg._lib.lib()
g = torch.Generator().manual_seed(0)


def timeit(fn, n=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3


for (M, N, K) in [(31360, 512, 512), (12480, 512, 512), (4096, 4096, 4096)]:
    A = torch.randn(M, K, generator=g).cuda(); W = torch.randn(N, K, generator=g).cuda(); b = torch.randn(N, generator=g).cuda()
    for path in (1, 2):
        for bias in (None, b):
            us = timeit(lambda: pkg.ops.gemm(A, W, bias, "nt", path))
            print(f"nt M={M} N={N} K={K} path={path} bias={bias is not None}: {us:8.1f} us  {2*M*N*K/us/1e6:7.1f} TFLOP/s (algorithmic, incl. operand split)")
    us = timeit(lambda: torch.matmul(A, W.T))
    print(f"   torch fp32 matmul: {us:8.1f} us")
    Ab, Wb = A.bfloat16(), W.bfloat16()
    us = timeit(lambda: torch.matmul(Ab, Wb.T))
    print(f"   torch bf16 matmul: {us:8.1f} us {2*M*N*K/us/1e6:7.1f} TFLOP/s")

# per-CTA timeline for the proj_v shape
import numpy as np
for (M, N, K, path) in [(31360, 512, 512, 1), (31360, 512, 512, 3), (4096, 4096, 4096, 1)]:
    A = torch.randn(M, K, generator=g).cuda(); W = torch.randn(N, K, generator=g).cuda(); b = torch.randn(N, generator=g).cuda()
    ncta = 600
    buf = torch.zeros(ncta, 64, dtype=torch.int64, device="cuda")
    P = {1: 1, 3: 0}.get(path, 1)
    run = (lambda: pkg.ops.gemm(A, W, b, "nt", path)) if path != 3 else (lambda: pkg.ops.gemm(A.bfloat16().float(), W.bfloat16().float(), b, "nt", 1))
    run()
    L.hca_debug_gemm_timeline(buf.data_ptr(), ncta)
    run()
    torch.cuda.synchronize()
    L.hca_debug_gemm_timeline(None, 0)
    tt = buf.cpu().numpy().astype(np.int64)
    names = ["start", "setup", "first_full", "mma_issued", "epi_start", "epi_end", "cta_end"]
    print(f"--- timeline M={M} N={N} K={K} path={path}")
    for i, n in enumerate(names[1:], 1):
        d = tt[:, i] - tt[:, 0]
        print(f"{n:12s} median {np.median(d):9.0f}  p10 {np.percentile(d,10):9.0f}  p90 {np.percentile(d,90):9.0f} cycles after CTA start")
    for c in (0, 300, 599):
        r = tt[c]
        print(f" CTA {c} (sm {r[7]}): producer issue  ", [int(x - r[0]) for x in r[8:20] if x])
        print(f"                  landed          ", [int(x - r[0]) for x in r[24:36] if x])
        print(f"                  mma issued      ", [int(x - r[0]) for x in r[40:52] if x])
