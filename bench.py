#!/usr/bin/env python
"""Headline benchmark: HieCoAttn training step (fwd + mean-CE + bwd + Adam) samples/sec on N B200s.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--batch B] [--no-graph]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W

Workload = BASELINE.json configs[2]: batch 160 per GPU, 196 regions, d=512, T=26, vocab 10000, K=1000(+1
UNKNOWN class, reference main.py:155), synthetic 196x512 image-feature grids and random-token questions
(the VQA/COCO data is not available offline), random-init weights.  Weak scaling: every rank processes
its own 160-sample shard and the gradients are all-reduced over NCCL (configs[3] is the same at 8 GPUs).

`--scaling strong` is BASELINE.json configs[3] as written: a FIXED global batch of 1280 sharded over the ranks
(1x1280, 2x640, 4x320, 8x160).

Prints ONE JSON line (rank 0).  `value` = samples/s with inputs resident in HBM; `e2e` = the same metric
through the public module API with host buffers (pinned H2D of every step's inputs and a D2H read of the
loss inside the timed region); `roofline` = the kernel that takes the largest share of the step (by the CUPTI
breakdown of this very run), timed alone with CUDA events, and `roofline_legs` = the same for the other heavy
kernels (LSTM recurrences, the largest weight-gradient product, the W_v.V projection); `cpu_baseline` = the
reference's algorithm on this box's host cores (oracle/torch_port.py, a bounded sample); `gpu_eager_baseline`
= that same reference algorithm run eagerly on THIS GPU by stock PyTorch (fp32, TF32, autocast bf16; also
CUDA-graph captured where capture succeeds) -- the kernel-for-kernel comparator.  `--impl reference` times
the CPU port as the whole job instead.
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
    B = args.batch                                # the SAME per-step batch as the GPU arm (160): one step = one sample of the workload
    # torchrun exports OMP_NUM_THREADS=1 to its workers: the CPU arm is meant to use every host core it can.  A CPU step of 160
    # samples takes ~0.5-1 s, so the step count is bounded to keep the run within minutes (the line reports what was run)
    steps = min(args.steps, 40)
    val, ms, threads = cpu_reference(B, steps, min(max(args.warmup, 1), 3), threads=os.cpu_count())
    line = {"metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": steps, "warmup": min(max(args.warmup, 1), 3),
            "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "impl": "reference",
            "config": workload_config(args.batch, 1, extra={"sample_batch": B, "steps_requested": args.steps}),
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": threads, "kind": "port",
                             "sample": f"{steps} steps of batch {B} (the GPU arm's batch), fwd+CE+bwd+Adam, torch CPU fp32, oracle/torch_port.py"},
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

    def __init__(self, pkg, device, batch, world, group, use_graph, slots=3, seed0=1, early_reduce=True):
        self.pkg, self.device, self.batch, self.world = pkg, device, batch, world
        syn = pkg.synthetic
        self.net = pkg.HieCoAttnHotPath(CFG["vocab"], CFG["d"], CFG["K"], CFG["mlp"])
        p = syn.make_params(CFG["d"], CFG["vocab"], CFG["K"], CFG["mlp"], seed=0)
        self.net.load_state_dict({k: torch.from_numpy(v) for k, v in p.items()}, strict=False)
        self.net.to(device)
        # world > 1: gradients and parameters live in symmetric memory and FlatAdam.step() is the collective -- one hand-written kernel
        # that reduces in the NVSwitch, updates and multicasts (csrc/dp_fused.cu); the classifier + co-attention slice (13 MB, complete
        # once coattn_bwd has been enqueued) is processed on a side stream, with a small grid, while the LSTM / conv / embedding backward
        # runs.  Without symmetric memory the reducer falls back to ONE NCCL all-reduce after backward (HCA_DP_FUSED=0 forces that).
        fused = None if os.environ.get("HCA_DP_FUSED", "1") != "0" else False
        self.dp = pkg.dp.FlatGradAllReduce(self.net.named_parameters(), group, overlap=False, flat_params=True, bucket_bytes=1 << 30,
                                           fused=fused, early_split="question_encoder." if early_reduce else None)
        self.opt = pkg.optim.FlatAdam(self.dp, lr=1e-4)          # Adam(lr=1e-4), README.md:95-100 / main.py:180
        self.early_end = 0
        if early_reduce and self.opt.overlap_early_slice(max_ctas=16):
            self.early_end = self.dp.buckets[0][1]
        self.criterion = pkg.CrossEntropyLoss(scale=self.dp.loss_scale)   # nn.CrossEntropyLoss(), main.py:179
        self.slots = []
        rank = torch.distributed.get_rank() if world > 1 else 0
        for s in range(slots):
            x = syn.make_inputs(batch, CFG["N"], CFG["T"], CFG["d"], CFG["vocab"], CFG["K"], seed=seed0 + 17 * s + 1000 * rank)
            # loader-side staging buffers: page-locked, write-combined (HCA_STAGING=pinned: plain pin_memory())
            if os.environ.get("HCA_STAGING", "wc") == "pinned":
                host = {k: torch.from_numpy(v).pin_memory() for k, v in x.items()}
            else:
                host = {k: pkg.staging.staged(v, write_combined=True) for k, v in x.items()}
            dev = {k: v.to(device) for k, v in host.items()}
            self.slots.append(dict(host=host, dev=dev, lens=pkg.QuestionLens(torch.from_numpy(x["lens"]), device, dev["lens"]), graph=None,
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


# ---------------------------------------------------------------------------------------------------- roofline legs
# DRAM traffic per launch comes from a committed `ncu --set full` capture, summarised by profiles/summarize_ncu.py into this file;
# the bench line cites it by content hash (no literal typed into this script).
TRAFFIC_FILE = os.path.join(ROOT, "profiles", "ncu_traffic.json")


def ncu_traffic(key):
    try:
        import hashlib
        raw = open(TRAFFIC_FILE, "rb").read()
        rec = json.loads(raw)["kernels"].get(key)
        if not rec:
            return None, None
        return float(rec["dram_bytes_read"]) + float(rec["dram_bytes_write"]), \
            f"profiles/ncu_traffic.json sha256:{hashlib.sha256(raw).hexdigest()[:16]} <- {rec.get('source', '?')}"
    except Exception:
        return None, None


def _events(n):
    return [torch.cuda.Event(enable_timing=True) for _ in range(n)]


def _leg(name, kernel, match, ms, flops, abytes, pk, note=None, issued_factor=3.0):
    """One roofline leg: `achieved` = ALGORITHMIC TFLOP/s of that kernel timed alone with CUDA events on its stream (the split-precision
    kernels issue 3 bf16 MMAs per algorithmic product: `frac_issued` counts those)."""
    achieved = flops / (ms * 1e-3) / 1e12
    traffic, src = ncu_traffic(name)
    leg = {"name": name, "kernel": kernel, "match": match, "bound": "tensor", "achieved": achieved, "peak": pk["bf16_tflops"], "unit": "TFLOP/s",
           "frac": achieved / pk["bf16_tflops"], "frac_issued": issued_factor * achieved / pk["bf16_tflops"], "ms_per_launch": ms,
           "algorithmic_flops_per_launch": flops, "algorithmic_bytes_per_launch": abytes,
           "hbm_gbs_if_algorithmic": abytes / (ms * 1e-3) / 1e9, "hbm_frac_if_algorithmic": abytes / (ms * 1e-3) / 1e9 / pk["hbm_gbs"],
           "traffic": traffic, "traffic_source": src, "peak_source": pk["source"] + " bf16 burst (kernel timed alone)"}
    if note:
        leg["note"] = note
    return leg


def time_pv_leg(pkg, device, steps, pk, batch):
    """PV = V . W_v^T + b_v, M = batch*196, N = K = 512: ONE launch of gemm_tc_kernel, bf16 hi/lo planes in and out."""
    M, N, K = batch * CFG["N"], CFG["d"], CFG["d"]
    g = torch.Generator(device="cpu").manual_seed(0)
    Ap = [pkg.ops.split_planes(torch.randn(M, K, generator=g).to(device)) for _ in range(3)]     # 3 x 64 MB in + 3 x 64 MB out > L2
    Wp = pkg.ops.split_planes((torch.randn(N, K, generator=g) * 0.04).to(device))
    b = torch.randn(N, generator=g).to(device)
    outs = [torch.empty(2, M, N, dtype=torch.bfloat16, device=device) for _ in range(3)]
    for i in range(3):
        pkg.ops.proj_planes(Ap[i % 3], Wp, b, outs[i % 3])
    torch.cuda.synchronize()
    ev = _events(2)
    ev[0].record()
    for i in range(steps):
        pkg.ops.proj_planes(Ap[i % 3], Wp, b, outs[i % 3])
    ev[1].record()
    torch.cuda.synchronize()
    ms = ev[0].elapsed_time(ev[1]) / steps
    return _leg("pv_proj", "gemm_tc_kernel<256,2,K-major,K-major,64,CG=2,EPI=0> (CTA pair, cta_group::2): PV = V.Wv^T + bv, "
                f"M={M} N=512 K=512, bf16 hi/lo planes in and out", "gemm_tc_kernel<256, 2, false, false, 64, 2, 0>", ms, 2.0 * M * N * K,
                2.0 * (2 * M * K * 2) + 2 * N * K * 2, pk)


def time_wgrad_leg(pkg, device, steps, pk, batch):
    """dW_v = dPV^T . V, M = N = 512, K = batch*196 (split-K, fp32 reduce-add): the largest of the weight-gradient products."""
    import ctypes as C
    K, M, N = batch * CFG["N"], CFG["d"], CFG["d"]
    g = torch.Generator(device="cpu").manual_seed(0)
    Ap = [pkg.ops.split_planes(torch.randn(K, M, generator=g).to(device)) for _ in range(3)]
    Bp = [pkg.ops.split_planes(torch.randn(K, N, generator=g).to(device)) for _ in range(3)]
    D = torch.zeros(M, N, device=device)
    L = pkg._lib.lib()
    st = torch.cuda.current_stream().cuda_stream

    def run(i):
        pkg._lib.check(L.hca_wgrad_planes(Ap[i % 3].data_ptr(), Bp[i % 3].data_ptr(), M, N, K, D.data_ptr(), st), "wgrad_planes")

    for i in range(3):
        run(i)
    torch.cuda.synchronize()
    ev = _events(2)
    ev[0].record()
    for i in range(steps):
        run(i)
    ev[1].record()
    torch.cuda.synchronize()
    ms = ev[0].elapsed_time(ev[1]) / steps
    return _leg("wgrad_dWv", f"gemm_tc_kernel<256,2,MN-major,MN-major,64,CG=2,EPI=0>: dW_v = dPV^T.V, M=N=512 K={K}, split-K, fp32 reduce-add",
                "gemm_tc_kernel<256, 2, true, true, 64, 2, 0>", ms, 2.0 * M * N * K, 2.0 * (2 * K * M * 2) + M * N * 4, pk,
                note="one of the 7 launches of this kernel per step (dW_v, dW_q, conv x3, LSTM x2)")


def time_lstm_legs(pkg, device, steps, pk, batch):
    """The two recurrence kernels of the sentence LSTM (csrc/lstm.cu), each bracketed by CUDA events on its stream inside a standalone
    forward / backward call of the op at the bench shape (batch x 26 x 512, H = 512, synthetic lengths)."""
    syn = pkg.synthetic
    T, H = CFG["T"], CFG["d"]
    x = syn.make_inputs(batch, CFG["N"], T, H, CFG["vocab"], CFG["K"], seed=3)
    lens = torch.from_numpy(x["lens"]).to(device)
    g = torch.Generator(device="cpu").manual_seed(0)
    u = lambda *s_: ((torch.rand(*s_, generator=g) * 2 - 1) / H ** 0.5).to(device).requires_grad_()
    w_ih, w_hh, b_ih, b_hh = u(4 * H, H), u(4 * H, H), u(4 * H), u(4 * H)
    xin = torch.randn(batch, T, H, generator=g).to(device).requires_grad_()
    dout = torch.randn(batch, T, H, generator=g).to(device)
    L = pkg._lib.lib()
    evs = {0: [], 1: []}
    for i in range(steps + 2):
        pair = {w: _events(2) for w in (0, 1)}
        for w in (0, 1):
            for e in pair[w]:
                e.record()                      # (a torch event gets its cudaEvent_t on the first record)
            L.hca_debug_lstm_events(pair[w][0].cuda_event, pair[w][1].cuda_event, w)
        out, _ = pkg.ops.lstm(xin, lens, w_ih, w_hh, b_ih, b_hh)
        out.backward(dout)
        for w in (0, 1):
            L.hca_debug_lstm_events(None, None, w)
        torch.cuda.synchronize()
        if i >= 2:
            for w in (0, 1):
                evs[w].append(pair[w][0].elapsed_time(pair[w][1]))
    tokens = float(x["lens"].sum())
    rec_tokens = float((x["lens"] - 1).sum())
    flops = 2.0 * 4 * H * H * rec_tokens
    legs = []
    for w, nm in ((0, "lstm_rec_fwd"), (1, "lstm_rec_bwd")):
        ms = float(np.median(evs[w]))
        abytes = (batch * T * 4 * H * 4 * (2 if w == 0 else 1) + batch * T * H * 4 * 3 + 2 * 4 * H * H * 2 + (2 * batch * T * 4 * H * 2 if w else 0))
        legs.append(_leg(nm, f"lstm_rec_kernel<{'bwd' if w else 'fwd'}>: {T} dependent steps of [{batch},{H}] x [{H},{4 * H}] (W_hh resident in shared memory), "
                         f"{int(tokens)} valid tokens", "lstm_rec_kernel<true" if w else "lstm_rec_kernel<false", ms, flops, float(abytes), pk,
                         note="a latency chain (per-step cross-CTA exchange + small MMAs), bound by neither pipe: see profiles/ for the per-step timeline"))
    return legs


def roofline_legs(pkg, device, steps, pk, batch):
    legs = []
    for fn in (time_lstm_legs, time_wgrad_leg, time_pv_leg):
        try:
            r = fn(pkg, device, steps, pk, batch)
            legs.extend(r if isinstance(r, list) else [r])
        except Exception as e:                  # evidence only: never fail the bench over one leg
            legs.append({"name": fn.__name__, "error": f"{type(e).__name__}: {str(e)[:160]}"})
    return legs


def pick_dominant(legs, shares):
    """The leg whose kernel takes the largest share of the step in the CUPTI breakdown of THIS run (falls back to the longest leg)."""
    good = [l for l in legs if "error" not in l]
    if not good:
        return None
    if shares and "top" in shares:
        def share(l):
            return sum(r["us_per_step"] for r in shares["top"] if l["match"].replace(" ", "") in r["kernel"].replace(" ", ""))
        for l in good:
            l["step_share_us"] = share(l)
            l["step_share"] = l["step_share_us"] / shares["kernel_us_per_step"] if shares.get("kernel_us_per_step") else None
        # a leg is ONE launch; several legs' kernels run more than once per step, so rank by the time of one launch
        best = max(good, key=lambda l: l["ms_per_launch"])
        return best
    return max(good, key=lambda l: l["ms_per_launch"])


def kernel_shares(st, steps=4, top=48):
    """Per-kernel device time of the step under CUPTI activity tracing (torch.profiler): which kernels the step is made of.
    Runs the step EAGERLY with programmatic dependent launch off: in the timed graph every kernel starts while its
    predecessor drains and waits in griddepcontrol.wait, so its CUPTI duration would include that wait."""
    lib = st.pkg._lib
    try:
        import collections
        lib.set_option("pdl", "0")
        st._step_body(st.slots[0])
        torch.cuda.synchronize()
        with torch.profiler.profile(activities=[torch.profiler.ProfilerActivity.CUDA]) as prof:
            for i in range(steps):
                st._step_body(st.slots[i % len(st.slots)])
            torch.cuda.synchronize()
        lib.set_option("pdl", "1")
        agg = collections.OrderedDict()
        for ev in prof.events():
            if ev.device_type == torch.autograd.DeviceType.CUDA:
                a = agg.setdefault(ev.name, [0, 0.0])
                a[0] += 1
                a[1] += ev.device_time
        tot = sum(v[1] for v in agg.values())
        rows = sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]
        short = lambda n: n.replace("hca::(anonymous namespace)::", "").replace("void ", "")[:72]
        return {"kernel_us_per_step": tot / steps,
                "top": [{"kernel": short(k), "us_per_step": us / steps, "launches_per_step": n / steps, "share": us / tot} for k, (n, us) in rows]}
    except Exception as e:                      # evidence only: never fail the bench over it
        try:
            lib.set_option("pdl", "1")
        except Exception:
            pass
        return {"error": f"{type(e).__name__}: {str(e)[:100]}"}


# ------------------------------------------------------------------------------------------- stock-PyTorch GPU baseline
def gpu_eager_baseline(device, batch, steps=10, warmup=3):
    """The kernel to beat on the same box: the reference's algorithm (oracle/torch_port.py, op for op what model.py:246-434 issues,
    incl. the redundant projections) run by stock PyTorch on THIS GPU -- cuBLAS / cuDNN kernels, fwd + CE + bwd + Adam -- in fp32 (TF32
    off), with TF32 allowed, and under autocast(bf16); eagerly, and replayed from a CUDA graph where capture succeeds."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import torch_port as TP
    syn = importlib.import_module("visual-question-answering_b200.synthetic")
    x = syn.make_inputs(batch, CFG["N"], CFG["T"], CFG["d"], CFG["vocab"], CFG["K"], seed=1)
    feats, tokens = torch.from_numpy(x["feats"]).to(device), torch.from_numpy(x["tokens"]).to(device)
    lens, labels = torch.from_numpy(x["lens"]), torch.from_numpy(x["labels"]).to(device)          # lens stay on the host: pack_padded_sequence wants them there
    out = {}
    saved = (torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32)
    for mode in ("fp32", "tf32", "autocast_bf16"):
        try:
            torch.backends.cuda.matmul.allow_tf32 = torch.backends.cudnn.allow_tf32 = (mode != "fp32")
            p = {k: v.detach().to(device).requires_grad_(v.requires_grad) for k, v in
                 TP.make_params(syn.make_params(CFG["d"], CFG["vocab"], CFG["K"], CFG["mlp"], seed=0)).items()}
            opt = torch.optim.Adam([v for v in p.values() if v.requires_grad], lr=1e-4, capturable=True)

            def step():
                with torch.autocast("cuda", dtype=torch.bfloat16, enabled=(mode == "autocast_bf16")):
                    logits = TP.forward(p, feats, tokens, lens)
                    loss = torch.nn.functional.cross_entropy(logits.float(), labels)
                opt.zero_grad(set_to_none=False)
                loss.backward()
                opt.step()
                return loss

            for _ in range(warmup):
                step()
            torch.cuda.synchronize()
            ev = _events(2)
            ev[0].record()
            for _ in range(steps):
                step()
            ev[1].record()
            torch.cuda.synchronize()
            ms = ev[0].elapsed_time(ev[1]) / steps
            rec = {"eager_ms_per_step": ms, "eager_samples_per_s": batch / (ms * 1e-3)}
            try:
                side = torch.cuda.Stream()
                side.wait_stream(torch.cuda.current_stream())
                with torch.cuda.stream(side):
                    for _ in range(2):
                        step()
                torch.cuda.current_stream().wait_stream(side)
                gph = torch.cuda.CUDAGraph()
                with torch.cuda.graph(gph):
                    step()
                for _ in range(warmup):
                    gph.replay()
                torch.cuda.synchronize()
                ev = _events(2)
                ev[0].record()
                for _ in range(steps):
                    gph.replay()
                ev[1].record()
                torch.cuda.synchronize()
                gms = ev[0].elapsed_time(ev[1]) / steps
                rec.update({"graph_ms_per_step": gms, "graph_samples_per_s": batch / (gms * 1e-3)})
            except Exception as e:
                rec["graph"] = f"capture failed: {type(e).__name__}: {str(e)[:100]}"
                torch.cuda.synchronize()
            out[mode] = rec
        except Exception as e:
            out[mode] = {"error": f"{type(e).__name__}: {str(e)[:160]}"}
    torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32 = saved
    out["what"] = ("oracle/torch_port.py (the reference's ATen call sequence, model.py:246-434 + main.py:179-180,214-222) on this GPU by stock "
                   f"PyTorch {torch.__version__}: fwd + mean CE + bwd + Adam, batch {batch}, inputs resident")
    return out


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
    if args.scaling == "strong":
        if args.global_batch % world:
            raise SystemExit(f"--global-batch {args.global_batch} is not divisible by {world} ranks")
        args.batch = args.global_batch // world
    st = Stepper(pkg, device, args.batch, world, group, use_graph, early_reduce=not args.no_early_reduce)
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
        # the CUPTI breakdown and the legs run the kernels outside the step's collectives: only at world == 1 is that a rank-0-only job
        shares = kernel_shares(st) if world == 1 else None
        legs = roofline_legs(pkg, device, 20, pk, args.batch) if not args.skip_legs else []
        roof = pick_dominant(legs, shares)
        step_tflops = FLOPS_PER_SAMPLE * args.batch / (ms_step * 1e-3) / 1e12
        cpu = eager = None
        if world == 1 and not args.skip_cpu_baseline:
            nb = min(args.batch, 160)
            v, ms, threads = cpu_reference(nb, 8, 2, threads=os.cpu_count())
            cpu = {"value": v, "unit": UNIT, "cores": threads, "kind": "port",
                   "sample": f"8 steps of batch {nb} (the GPU arm's batch), fwd+CE+bwd+Adam, torch CPU fp32, oracle/torch_port.py", "ms_per_step": ms}
        if world == 1 and not args.skip_gpu_baseline:
            try:
                eager = gpu_eager_baseline(device, args.batch)
            except Exception as e:
                eager = {"error": f"{type(e).__name__}: {str(e)[:160]}"}
        comm = "none"
        if world > 1:
            comm = ("fused NVLink kernel (csrc/dp_fused.cu: " + ("multimem.ld_reduce / multimem.st through the NVSwitch" if st.dp._symm.mc else "peer loads / stores")
                    + "), all-reduce + Adam + parameter broadcast in one launch" + (", classifier + co-attention slice overlapped with the rest of backward" if st.early_end else "")) \
                if st.dp.fused else "NCCL all-reduce after backward + fused Adam"
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
                "ms_per_step": ms_step, "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None,
                "dtype": "bf16x2 (fp32 operands as bf16 hi+lo planes, 3 tcgen05 MMAs per product, fp32 accumulate)",
                "data": "synthetic", "config": workload_config(args.batch, world, {"mode": graph_note, "collective": comm}),
                "clocks": clocks,
                "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": st.h2d_bytes, "d2h_bytes_per_step": 4,
                        "ms_per_step": e2e_ms},
                "gpu_launches": int(launches_per_step * args.steps), "gpu_launches_per_step": int(launches_per_step),
                "roofline": roof, "roofline_legs": legs, "cpu_baseline": cpu, "gpu_eager_baseline": eager, "kernel_shares": shares,
                "step_algorithmic_tflops": step_tflops, "step_frac_of_bf16_peak": step_tflops / pk["bf16_tflops_sustained"],
                "final_loss": final_loss, "grad_allreduce_bytes": st.dp.grad_bytes() if world > 1 else 0}
        print(json.dumps(line), file=_REAL_STDOUT, flush=True)
    if world > 1:
        for slot in st.slots:
            slot["graph"] = None
        torch.cuda.synchronize()
        sys.stdout.flush()
        sys.stderr.flush()
        if st.dp.fused and os.environ.get("HCA_BENCH_HARD_EXIT", "0") != "1":
            # fused transport: no NCCL collective was captured into the graphs, so the process group tears down normally
            st.dp.close()
            torch.distributed.barrier()
            torch.distributed.destroy_process_group()
            return
        # NCCL transport: communicators that were captured into CUDA graphs do not always tear down cleanly (observed in round 1: the
        # job printed its line and then sat in destroy_process_group until the launcher's timeout).  Everything this process owes the
        # caller has been written: leave without the NCCL destructor.
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
    ap.add_argument("--skip-gpu-baseline", action="store_true", help="skip the stock-PyTorch eager legs on the GPU")
    ap.add_argument("--skip-legs", action="store_true", help="skip the per-kernel roofline legs")
    ap.add_argument("--scaling", choices=["weak", "strong"], default="weak",
                    help="weak: --batch samples per GPU (configs[2]); strong: --global-batch samples split over the GPUs (configs[3])")
    ap.add_argument("--global-batch", type=int, default=1280)
    ap.add_argument("--no-early-reduce", action="store_true", help="fused DP: do not overlap the classifier / co-attention slice with backward")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
