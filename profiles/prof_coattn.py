"""ncu target: co-attention forward + backward alone at the headline shapes (B=160, N=196, T=26, d=512), 3 iterations.
One iteration launches 5 (fwd) + 10 (bwd) gemm_tc kernels; run under gpurun, e.g.
    ncu --set full --clock-control none --import-source on -k regex:gemm_tc -s 30 -c 15 -o gpurun_out/coattn python profiles/prof_coattn.py
"""
import importlib, sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
pkg = importlib.import_module("visual-question-answering_b200")
B, N, T, d = int(os.environ.get("B", 160)), 196, 26, 512
g = torch.Generator().manual_seed(0)
ca = pkg.modules.ParallelCoAttention(d).cuda()
V = torch.randn(B, N, d, generator=g).cuda()
qs = [torch.randn(B, T, d, generator=g).cuda().requires_grad_(True) for _ in range(3)]
iters = int(sys.argv[1]) if len(sys.argv) > 1 else 3
ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
for it in range(iters):
    ev[0].record()
    vhat, qhat = ca.forward_stacked(V, qs)
    ev[1].record()
    (vhat.sum() + qhat.sum()).backward()
    ev[2].record()
    torch.cuda.synchronize()
    print(f"iter {it}: fwd {ev[0].elapsed_time(ev[1])*1e3:.0f} us  bwd {ev[1].elapsed_time(ev[2])*1e3:.0f} us")
