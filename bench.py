#!/usr/bin/env python
"""Headline benchmark: HieCoAttn training step (fwd + mean-CE + bwd + Adam) samples/sec on N B200s.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--batch B] [--no-graph]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W

Workload = BASELINE.json configs[2]: batch 160 per GPU, 196 regions, d=512, T=26, vocab 10000, K=1000(+1
UNKNOWN class, reference main.py:155), synthetic 196x512 image-feature grids and random-token questions
(the VQA/COCO data is not available offline), random-init weights.  Weak scaling: every rank processes
its own 160-sample shard and the gradients are all-reduced over NCCL (configs[3] is the same at 8 GPUs).

Prints ONE JSON line (rank 0).  `value` = samples/s with inputs resident in HBM; `e2e` = the same metric
through the public module API with host buffers (pinned H2D of every step's inputs and a D2H read of the
loss inside the timed region); `roofline` = the dominant kernel (the W_v.V projection GEMM) timed alone
with CUDA events; `cpu_baseline` = the reference's algorithm on this box's host cores (oracle/torch_port.py,
a bounded sample).  `--impl reference` times that CPU port as the whole job instead.
"""
from __future__ import annotations

import argparse
import importlib
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

CFG = dict(N=196, d=512, T=26, vocab=10000, K=1001, mlp=1024)
_REAL_STDOUT = sys.stdout
METRIC = "hiecoattn_train_samples_per_sec"
UNIT = "samples/s"
FLOPS_PER_SAMPLE = 720e6          # algorithmic fwd+bwd, SURVEY.md section 8(a) ledger (no LSTM, no VGG, no Adam)


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        p = json.load(open(path))
        return dict(hbm_gbs=p["hbm_gbs"], bf16_tflops=p["bf16_tflops"], bf16_tflops_sustained=p.get("bf16_tflops_sustained"),
                    source="measured")
    return dict(hbm_gbs=6650.0, bf16_tflops=1590.0, bf16_tflops_sustained=1400.0, source="fallback")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.rows, self.proc, self.idx = [], None, gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "20",
                                          "-i", str(self.idx)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2]))
            except Exception:
                continue
            for name, col in (("hw_slowdown", 5), ("hw_thermal_slowdown", 6), ("sw_thermal_slowdown", 7), ("sw_power_cap", 8)):
                if len(r) > col and r[col].lower().startswith("active"):
                    reasons.add(name)
        busy = [s for s in sm if s > 0]
        return {"sm_mhz": float(np.median(busy)) if busy else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def dist_env():
    rank = int(os.environ.get("RANK", 0)); world = int(os.environ.get("WORLD_SIZE", 1)); local = int(os.environ.get("LOCAL_RANK", 0))
    return rank, world, local


# ------------------------------------------------------------------------------------------------ CPU baseline
def cpu_reference(batch, steps, warmup, with_adam=True, threads=None):
    """The reference's algorithm on host cores (oracle/torch_port.py): returns (samples/s, ms/step, threads)."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import torch_port as TP
    syn = importlib.import_module("visual-question-answering_b200.synthetic")
    if threads:
        torch.set_num_threads(threads)
    p = TP.make_params(syn.make_params(CFG["d"], CFG["vocab"], CFG["K"], CFG["mlp"], seed=0))
    x = syn.make_inputs(batch, CFG["N"], CFG["T"], CFG["d"], CFG["vocab"], CFG["K"], seed=1)
    feats, tokens = torch.from_numpy(x["feats"]), torch.from_numpy(x["tokens"])
    lens, labels = torch.from_numpy(x["lens"]), torch.from_numpy(x["labels"])
    opt = torch.optim.Adam([v for v in p.values() if v.requires_grad], lr=1e-4) if with_adam else None
    for _ in range(warmup):
        TP.train_step(p, feats, tokens, lens, labels, opt)
    times = []
    for _ in range(steps):
        t0 = time.perf_counter()
        TP.train_step(p, feats, tokens, lens, labels, opt)
        times.append(time.perf_counter() - t0)
    ms = 1e3 * float(np.mean(times))
    return batch / (ms / 1e3), ms, torch.get_num_threads()


def run_reference_arm(args):
    rank, world, _ = dist_env()
    if rank != 0:
        return
    B = 32                                        # bounded sample of the workload: 32 of the 160 samples per step
    # torchrun exports OMP_NUM_THREADS=1 to its workers: the CPU arm is meant to use every host core it can
    val, ms, threads = cpu_reference(B, args.steps, max(args.warmup, 1), threads=os.cpu_count())
    line = {"metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "impl": "reference",
            "config": workload_config(args.batch, 1, extra={"sample_batch": B}),
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": threads, "kind": "port",
                             "sample": f"{args.steps} steps of batch {B} (of the 160-sample batch), torch CPU fp32, oracle/torch_port.py"},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), file=_REAL_STDOUT, flush=True)


def workload_config(batch, world, extra=None):
    c = {"workload": "BASELINE.json configs[2]: HieCoAttn parallel co-attention training, question encoder + co-attention x3 + MLP "
                     "+ CE, fwd+bwd+Adam", "batch_per_gpu": batch, "global_batch": batch * world, "regions": CFG["N"], "d": CFG["d"],
         "T": CFG["T"], "vocab": CFG["vocab"], "K": CFG["K"], "parallelism": f"dp{world}",
         "l2": "3 rotating input batches (193 MB) + >300 MB of intermediates per step exceed the 126 MB L2"}
    if extra:
        c.update(extra)
    return c


# ---------------------------------------------------------------------------------------------------- our arm
class Stepper:
    """One training step of the public modules, optionally captured in a CUDA graph per input slot."""

    def __init__(self, pkg, device, batch, world, group, use_graph, slots=3, seed0=1):
        self.pkg, self.device, self.batch, self.world = pkg, device, batch, world
        syn = pkg.synthetic
        self.net = pkg.HieCoAttnHotPath(CFG["vocab"], CFG["d"], CFG["K"], CFG["mlp"])
        p = syn.make_params(CFG["d"], CFG["vocab"], CFG["K"], CFG["mlp"], seed=0)
        self.net.load_state_dict({k: torch.from_numpy(v) for k, v in p.items()}, strict=False)
        self.net.to(device)
        # graph mode: the bucketed all-reduce follows backward inside the captured step (hook-launched NCCL did not replay
        # reliably from a captured graph: observed a hang); eager mode launches each bucket from the gradient hooks instead
        # (graph mode reduces after backward, so ONE collective over the whole 48.7 MB buffer beats four 16 MB buckets: fewer launches,
        # better NVLink / NVSwitch efficiency per message)
        self.dp = pkg.dp.FlatGradAllReduce(self.net.named_parameters(), group, overlap=not use_graph, flat_params=True,
                                           bucket_bytes=(1 << 30) if use_graph else (16 << 20))
        self.opt = pkg.optim.FlatAdam(self.dp, lr=1e-4)          # Adam(lr=1e-4), README.md:95-100 / main.py:180
        self.criterion = pkg.CrossEntropyLoss(scale=self.dp.loss_scale)   # nn.CrossEntropyLoss(), main.py:179
        self.slots = []
        rank = torch.distributed.get_rank() if world > 1 else 0
        for s in range(slots):
            x = syn.make_inputs(batch, CFG["N"], CFG["T"], CFG["d"], CFG["vocab"], CFG["K"], seed=seed0 + 17 * s + 1000 * rank)
            host = {k: torch.from_numpy(v).pin_memory() for k, v in x.items()}
            dev = {k: v.to(device) for k, v in host.items()}
            self.slots.append(dict(host=host, dev=dev, lens=pkg.QuestionLens(host["lens"], device, dev["lens"]), graph=None,
                                   loss=torch.zeros((), device=device)))
        self.use_graph = use_graph
        self.h2d_bytes = sum(v.numel() * v.element_size() for v in self.slots[0]["host"].values())

    def _step_body(self, slot):
        d = slot["dev"]
        self.dp.zero_grad()
        logits = self.net(d["feats"], d["tokens"], slot["lens"])
        loss = self.criterion(logits, d["labels"])           # mean CE x 1/world, loss + gradient in one launch
        loss.backward()
        self.dp.finish()
        self.opt.step()
        slot["loss"].copy_(loss.detach())

    def warm(self, n):
        for i in range(n):
            self._step_body(self.slots[i % len(self.slots)])
        torch.cuda.synchronize()

    def count_launches(self):
        before = self.pkg._lib.launch_count()
        self._step_body(self.slots[0])
        torch.cuda.synchronize()
        return self.pkg._lib.launch_count() - before

    def capture(self):
        for slot in self.slots:
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                self._step_body(slot)
            slot["graph"] = g
        torch.cuda.synchronize()

    def step(self, i):
        slot = self.slots[i % len(self.slots)]
        if slot["graph"] is not None:
            slot["graph"].replay()
        else:
            self._step_body(slot)
        return slot


# ncu --set full capture of this kernel (profiles/): DRAM bytes per launch, read + write.  None until a capture is committed.
ROOFLINE_TRAFFIC_BYTES = 81.0e6
ROOFLINE_TRAFFIC_SOURCE = "profiles/r1e_pv_gemm_ncu_full.md: dram__bytes_read.sum 65.4 MB + dram__bytes_write.sum 15.6 MB, one launch"


def time_roofline_kernel(pkg, device, steps, pk):
    """The largest dense contraction of the path timed alone with CUDA events: PV = V . W_v^T + b_v, M = 160*196, N = K = 512
    (SURVEY section 8a row a6), ONE launch of gemm_tc_kernel on operands already in bf16 hi/lo planes, planes out."""
    M, N, K = 160 * CFG["N"], CFG["d"], CFG["d"]
    g = torch.Generator(device="cpu").manual_seed(0)
    Ap = [pkg.ops.split_planes(torch.randn(M, K, generator=g).to(device)) for _ in range(3)]     # 3 x 64 MB in + 3 x 64 MB out > L2
    Wp = pkg.ops.split_planes((torch.randn(N, K, generator=g) * 0.04).to(device))
    b = torch.randn(N, generator=g).to(device)
    outs = [torch.empty(2, M, N, dtype=torch.bfloat16, device=device) for _ in range(3)]
    for i in range(3):
        pkg.ops.proj_planes(Ap[i % 3], Wp, b, outs[i % 3])
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    ev[0].record()
    for i in range(steps):
        pkg.ops.proj_planes(Ap[i % 3], Wp, b, outs[i % 3])
    ev[1].record()
    torch.cuda.synchronize()
    ms = ev[0].elapsed_time(ev[1]) / steps
    flops = 2.0 * M * N * K
    achieved = flops / (ms * 1e-3) / 1e12
    # the split-precision path issues 3 bf16 MMAs per algorithmic product; `achieved` is ALGORITHMIC TFLOP/s against the bf16 peak
    return {"bound": "tensor", "achieved": achieved, "peak": pk["bf16_tflops"], "unit": "TFLOP/s", "frac": achieved / pk["bf16_tflops"],
            "traffic": ROOFLINE_TRAFFIC_BYTES, "traffic_source": ROOFLINE_TRAFFIC_SOURCE,
            "kernel": "gemm_tc_kernel<256,2,K-major,K-major,64,CG=2,EPI=0> (CTA pair, cta_group::2): PV = V.Wv^T + bv, M=31360 N=512 K=512, "
                      "bf16 hi/lo planes in and out",
            "ms_per_launch": ms, "algorithmic_flops_per_launch": flops, "issued_mma_flops_per_launch": 3.0 * flops,
            "frac_issued": 3.0 * achieved / pk["bf16_tflops"],
            "algorithmic_bytes_per_launch": 2.0 * (2 * M * K * 2) + 2 * N * K * 2,
            "hbm_gbs_if_algorithmic": (2.0 * (2 * M * K * 2) + 2 * N * K * 2) / (ms * 1e-3) / 1e9,
            "peak_source": pk["source"] + " bf16 burst (kernel timed alone)"}


def kernel_shares(st, steps=4, top=48):
    """Per-kernel device time of the step under CUPTI activity tracing (torch.profiler): which kernels the step is made of.
    Runs the step EAGERLY with programmatic dependent launch off (HCA_PDL=0): in the timed graph every kernel starts while its
    predecessor drains and waits in griddepcontrol.wait, so its CUPTI duration would include that wait."""
    try:
        import collections
        os.environ["HCA_PDL"] = "0"
        st._step_body(st.slots[0])
        torch.cuda.synchronize()
        with torch.profiler.profile(activities=[torch.profiler.ProfilerActivity.CUDA]) as prof:
            for i in range(steps):
                st._step_body(st.slots[i % len(st.slots)])
            torch.cuda.synchronize()
        os.environ.pop("HCA_PDL", None)
        agg = collections.OrderedDict()
        for ev in prof.events():
            if ev.device_type == torch.autograd.DeviceType.CUDA:
                a = agg.setdefault(ev.name, [0, 0.0])
                a[0] += 1
                a[1] += ev.device_time
        tot = sum(v[1] for v in agg.values())
        rows = sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]
        short = lambda n: n.replace("hca::(anonymous namespace)::", "").replace("void ", "")[:60]
        return {"kernel_us_per_step": tot / steps,
                "top": [{"kernel": short(k), "us_per_step": us / steps, "launches_per_step": n / steps, "share": us / tot} for k, (n, us) in rows]}
    except Exception as e:                      # evidence only: never fail the bench over it
        os.environ.pop("HCA_PDL", None)
        return {"error": f"{type(e).__name__}: {str(e)[:100]}"}


def run_ours(args):
    rank, world, local = dist_env()
    if args.gpus != world and world > 1:
        raise SystemExit(f"--gpus {args.gpus} != WORLD_SIZE {world}")
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (the product path has no CPU fallback); use --impl reference for the CPU arm")
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    if world > 1:
        # a rank that dies or a collective that never completes must not hold the GPUs until the launcher's own timeout
        def _watchdog():
            sys.stderr.write(f"bench.py rank {rank}: no result after 300 s, giving up\n")
            sys.stderr.flush()
            os._exit(3)
        wd = threading.Timer(300.0, _watchdog)
        wd.daemon = True
        wd.start()
    group = None
    if world > 1:
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")      # stdout carries the one JSON line only
        torch.distributed.init_process_group("nccl", device_id=device)
    pkg = importlib.import_module("visual-question-answering_b200")
    importlib.import_module("visual-question-answering_b200.dp")
    pk = peaks()
    use_graph = not args.no_graph
    st = Stepper(pkg, device, args.batch, world, group, use_graph)
    st.warm(2)
    launches_per_step = st.count_launches()
    graph_note = "cuda-graph replay"
    if use_graph:
        try:
            s = torch.cuda.Stream()
            s.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(s):
                st.warm(3)                      # warm up on the side stream torch.cuda.graph captures from
            torch.cuda.current_stream().wait_stream(s)
            st.capture()
        except Exception as e:                  # capture is an optimisation, never a requirement
            graph_note = f"eager (graph capture failed: {type(e).__name__}: {str(e)[:120]})"
            for slot in st.slots:
                slot["graph"] = None
            torch.cuda.synchronize()
    else:
        graph_note = "eager"

    def barrier():
        if world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize()

    # ---- device-resident timing ---------------------------------------------------------------------------
    for i in range(max(args.warmup, 3)):
        st.step(i)
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for i in range(args.steps):
        st.step(i)
    e1.record()
    barrier()
    ms_total = e0.elapsed_time(e1)
    clocks = sampler.stop() if rank == 0 else None
    t = torch.tensor([ms_total], device=device)
    if world > 1:
        torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
    ms_step = float(t.item()) / args.steps
    value = args.batch * world / (ms_step * 1e-3)

    # ---- end to end: pinned H2D of every step's inputs + D2H of the loss, copies overlapped on a side stream ---
    copy_stream = torch.cuda.Stream()
    main = torch.cuda.current_stream()
    nslots = len(st.slots)
    ready = [torch.cuda.Event() for _ in range(nslots)]
    consumed = [torch.cuda.Event() for _ in range(nslots)]

    def h2d(i):
        slot = st.slots[i % nslots]
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(consumed[i % nslots])
            for k, v in slot["host"].items():
                slot["dev"][k].copy_(v, non_blocking=True)
            ready[i % nslots].record(copy_stream)

    for ev in consumed:
        ev.record(main)
    loss_host = torch.zeros((), pin_memory=True)
    barrier()
    e2, e3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e2.record()
    h2d(0)
    for i in range(args.steps):
        if i + 1 < args.steps:
            h2d(i + 1)
        main.wait_event(ready[i % nslots])
        slot = st.step(i)
        consumed[i % nslots].record(main)
        loss_host.copy_(slot["loss"], non_blocking=True)
    e3.record()
    barrier()
    t2 = torch.tensor([e2.elapsed_time(e3)], device=device)
    if world > 1:
        torch.distributed.all_reduce(t2, op=torch.distributed.ReduceOp.MAX)
    e2e_ms = float(t2.item()) / args.steps
    e2e_value = args.batch * world / (e2e_ms * 1e-3)
    final_loss = float(loss_host) / st.dp.loss_scale

    if rank == 0:
        roof = time_roofline_kernel(pkg, device, max(args.steps, 10), pk)
        shares = kernel_shares(st) if world == 1 else None      # (the step holds collectives when world > 1: not a rank-0-only job)
        step_tflops = FLOPS_PER_SAMPLE * args.batch / (ms_step * 1e-3) / 1e12
        cpu = None
        if world == 1 and not args.skip_cpu_baseline:
            v, ms, threads = cpu_reference(32, 5, 2, threads=os.cpu_count())
            cpu = {"value": v, "unit": UNIT, "cores": threads, "kind": "port",
                   "sample": "5 steps of batch 32 (of the 160-sample batch), torch CPU fp32, oracle/torch_port.py", "ms_per_step": ms}
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
                "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f32 (fp32 CUDA cores)" if pkg._lib.get_option("gemm") == "ffma" else "bf16x2 (fp32 operands as bf16 hi+lo planes, 3 tcgen05 MMAs per product, fp32 accumulate)",
                "data": "synthetic", "config": workload_config(args.batch, world, {"mode": graph_note}),
                "clocks": clocks,
                "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": st.h2d_bytes, "d2h_bytes_per_step": 4,
                        "ms_per_step": e2e_ms},
                "gpu_launches": int(launches_per_step * args.steps), "gpu_launches_per_step": int(launches_per_step),
                "roofline": roof, "cpu_baseline": cpu, "kernel_shares": shares,
                "step_algorithmic_tflops": step_tflops, "step_frac_of_bf16_peak": step_tflops / pk["bf16_tflops_sustained"],
                "final_loss": final_loss, "grad_allreduce_bytes": st.dp.grad_bytes() if world > 1 else 0}
        print(json.dumps(line), file=_REAL_STDOUT, flush=True)
    if world > 1:
        # NCCL communicators that were captured into CUDA graphs do not always tear down cleanly (observed: the job
        # printed its line and then sat in destroy_process_group until the launcher's timeout).  Everything this process
        # owes the caller has been written: drop the graphs, drain the device and leave without the NCCL destructor.
        for slot in st.slots:
            slot["graph"] = None
        torch.cuda.synchronize()
        sys.stdout.flush()
        sys.stderr.flush()
        os._exit(0)


def _claim_stdout():
    """Route this process's fd 1 to stderr and return a file object on the ORIGINAL stdout: libraries that print banners to
    stdout (NCCL's version line) must not end up next to the one JSON line the caller parses."""
    sys.stdout.flush()
    real = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    return real


def main():
    global _REAL_STDOUT
    _REAL_STDOUT = _claim_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", choices=["ours", "reference"], default="ours")
    ap.add_argument("--batch", type=int, default=160, help="samples per GPU")
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--skip-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
