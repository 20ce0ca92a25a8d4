"""BASELINE.json configs[4]: inference sweep of the hot path (no_grad forward, no collective): batch 1..4096 x {196, 576} regions x
{26, 64} tokens, K = 3001.  Prints a markdown table: latency per batch and samples/s with CUDA events, eager launches and CUDA-graph
replay (InferenceSession), plus a logits-vs-oracle check at one large batch (the fp64 oracle forward on the host)."""
import importlib, os, sys, torch, numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
pkg = importlib.import_module("visual-question-answering_b200")
import hiecoattn_oracle as O
syn = pkg.synthetic
d, vocab, K, mlp = 512, 10000, 3001, 1024
p = syn.make_params(d, vocab, K, mlp, seed=0)
net = pkg.HieCoAttnHotPath(vocab, d, K, mlp)
net.load_state_dict({k: torch.from_numpy(v) for k, v in p.items()}, strict=False)
net.cuda().eval()


def timed(fn, it):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(it):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / it


print("| regions | T | batch | eager ms | graph ms | graph samples/s | check |")
print("|---:|---:|---:|---:|---:|---:|---|")
for N in (196, 576):
    for T in (26, 64):
        for B in (1, 2, 4, 8, 16, 32, 64, 128, 256, 512, 1024, 2048, 4096):
            x = syn.make_inputs(B, N, T, d, vocab, K, seed=B + N + T, dist="D2")
            feats, tok = torch.from_numpy(x["feats"]).cuda(), torch.from_numpy(x["tokens"]).cuda()
            ql = pkg.QuestionLens(torch.from_numpy(x["lens"]), "cuda")
            it = 20 if B <= 256 else (6 if B <= 1024 else 3)
            with torch.no_grad():
                eager = timed(lambda: net(feats, tok, ql), it)
            graph = None
            if B <= 1024:
                sess = pkg.InferenceSession(net, B, N, T)
                graph = timed(lambda: sess(feats, tok, ql), it)
                logits = sess(feats, tok, ql).clone()
                del sess
            else:
                with torch.no_grad():
                    logits = net(feats, tok, ql)
            check = "finite" if bool(torch.isfinite(logits).all()) else "NOT FINITE"
            if B == 256:                                 # one large batch per (N, T) against the fp64 oracle
                ref = O.hiecoattn_forward({k: v.astype(np.float64) for k, v in p.items()}, x["feats"].astype(np.float64), x["tokens"], x["lens"])
                got = logits.cpu().numpy()
                rel = float(np.linalg.norm(got - ref) / np.linalg.norm(ref))
                check = f"logits rel err vs fp64 oracle {rel:.1e}, argmax agreement {(got.argmax(1) == ref.argmax(1)).mean():.3f}"
            best = graph if graph is not None else eager
            print(f"| {N} | {T} | {B} | {eager:.3f} | {'-' if graph is None else f'{graph:.3f}'} | {B / best * 1e3:,.0f} | {check} |", flush=True)
            del feats, logits
            torch.cuda.empty_cache()
