"""CUDA-event timing of the PV projection kernel under the launcher's debug switches (HCA_TC_EG / HCA_TC_STAGES)."""
import importlib, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
pkg = importlib.import_module("visual-question-answering_b200")
g = torch.Generator().manual_seed(0)
M, N, K = 160 * 196, 512, 512
Ap = [pkg.ops.split_planes(torch.randn(M, K, generator=g).cuda()) for _ in range(3)]
Wp = pkg.ops.split_planes((torch.randn(N, K, generator=g) * 0.04).cuda())
b = torch.randn(N, generator=g).cuda()
outs = [torch.empty(2, M, N, dtype=torch.bfloat16, device="cuda") for _ in range(3)]


def run(label, env):
    for k in ("HCA_TC_EG", "HCA_TC_STAGES", "HCA_TC_PAIR", "HCA_TC_PLDIRECT"):
        os.environ.pop(k, None)
    os.environ.update(env)
    for i in range(5):
        pkg.ops.proj_planes(Ap[i % 3], Wp, b, outs[i % 3])
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(30):
        pkg.ops.proj_planes(Ap[i % 3], Wp, b, outs[i % 3])
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / 30 * 1e3
    print(f"{label:28s} {us:7.1f} us  {2.0 * M * N * K * 3 / us / 1e6:7.1f} issued TFLOP/s")


run("default (pair 256x256, direct plane stores)", {})
run("pair, 2 epilogue groups", {"HCA_TC_EG": "2"})
run("pair, TMA plane stores", {"HCA_TC_PLDIRECT": "0"})
run("single CTA 128x128", {"HCA_TC_PAIR": "0"})
run("single CTA, 2 groups", {"HCA_TC_PAIR": "0", "HCA_TC_EG": "2"})
run("single CTA, TMA plane stores", {"HCA_TC_PAIR": "0", "HCA_TC_PLDIRECT": "0"})
