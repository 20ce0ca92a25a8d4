"""BASELINE.json configs[4]: inference sweep of the hot path (no_grad forward, no collective): batch 1..4096, 196 / 576 regions,
T = 26 / 64, K = 3001.  Prints a markdown table: latency per batch and samples/s, CUDA events, eager launches (no graph)."""
import importlib, os, sys, time, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
pkg = importlib.import_module("visual-question-answering_b200")
d, vocab, K, mlp = 512, 10000, 3001, 1024
net = pkg.HieCoAttnHotPath(vocab, d, K, mlp).cuda().eval()
g = torch.Generator().manual_seed(0)
print("| regions | T | batch | ms / batch | samples/s | top-1 finite |")
print("|---:|---:|---:|---:|---:|---|")
for N, T in ((196, 26), (576, 64)):
    for B in (1, 8, 64, 512, 4096):
        lens = torch.randint(3, T + 1, (B,), generator=g).sort(descending=True).values
        tok = torch.zeros(B, T, dtype=torch.long)
        for b in range(B):
            tok[b, :lens[b]] = torch.randint(1, vocab, (int(lens[b]),), generator=g)
        feats = torch.randn(B, N, d, generator=g).clamp_min(0).cuda()
        ql = pkg.QuestionLens(lens, "cuda")
        tok = tok.cuda()
        with torch.no_grad():
            for _ in range(2):
                prob, idx = net.predict(feats, tok, ql, topk=5)
            torch.cuda.synchronize()
            it = 10 if B <= 512 else 3
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(it):
                prob, idx = net.predict(feats, tok, ql, topk=5)
            e1.record()
            torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / it
        print(f"| {N} | {T} | {B} | {ms:.3f} | {B / ms * 1e3:,.0f} | {bool(torch.isfinite(prob).all())} |", flush=True)
        del feats, prob, idx
        torch.cuda.empty_cache()
