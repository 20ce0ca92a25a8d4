"""Generate the golden fixtures under tests/golden/ by running the UNMODIFIED reference.

TEST INFRASTRUCTURE ONLY (see oracle/hiecoattn_oracle.py header).  Runs in the build container only:
it imports /root/reference/model.py (read-only), which does not exist on the GPU box.  The fixtures
it writes are committed, so nothing at test / smoke / bench time needs the reference tree.

    python oracle/make_golden.py            # rewrites tests/golden/*.npz

What is recorded (reference modules: model.py:246-434; loss: main.py:179,214):
  * ``small_*.npz``   tiny shapes, fp64 AND fp32 runs of the reference; all weights, inputs, every
                      intermediate the oracle exposes, logits, loss, and every gradient.
  * ``d512_*.npz``    the real widths (d=512, N=196, T=26, K=1001, small batch).  Weights and inputs
                      come from the numpy generators in visual-question-answering_b200/synthetic.py
                      (seeded, portable) so only outputs are stored: logits, loss, phrase max-pool
                      indices, attended features, per-tensor gradient digests (sum, L2 norm, 64 strided
                      samples).
"""
from __future__ import annotations

import importlib
import os
import sys
import warnings

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference")
warnings.filterwarnings("ignore")

import model as ref  # noqa: E402  (the reference, unmodified)

syn = importlib.import_module("visual-question-answering_b200.synthetic")


def build_reference(p, d, vocab, K, mlp_dim, dtype):
    qe = ref.QuestionCoAttentionEncoder(vocab, d, d)
    ca = ref.ParallelCoAttention(d)
    ml = ref.MLPClassifier(d, mlp_dim, K)
    mods = {"question_encoder": qe, "co_attention": ca, "mlp_classify": ml}
    for prefix, m in mods.items():
        sd = {k[len(prefix) + 1:]: torch.from_numpy(np.ascontiguousarray(v)) for k, v in p.items() if k.startswith(prefix + ".")}
        m.to(dtype)                                   # convert first: load_state_dict copies INTO the params' dtype
        m.load_state_dict(sd, strict=True)
    return qe, ca, ml


def run_reference(p, x, d, vocab, K, mlp_dim, dtype, feats_grad=False):
    qe, ca, ml = build_reference(p, d, vocab, K, mlp_dim, dtype)
    feats = torch.from_numpy(x["feats"]).to(dtype).requires_grad_(feats_grad)
    tokens = torch.from_numpy(x["tokens"])
    lens = torch.from_numpy(x["lens"])
    labels = torch.from_numpy(x["labels"])
    word, phrase, sent = qe(tokens, lens)                         # model.py:173
    for t in (word, phrase, sent):
        t.retain_grad()
    vh, qh = ca(feats, [word, phrase, sent])                      # model.py:182
    logits = ml(vh, qh)                                           # model.py:185
    loss = torch.nn.CrossEntropyLoss()(logits, labels)            # main.py:179,214
    loss.backward()
    # raw (unmasked) phrase conv output + indices straight from the reference sub-module
    with torch.no_grad():
        raw = qe.phrase_conv_pool(qe.word_embedding(tokens))
    out = dict(logits=logits, loss=loss, word=word, phrase=phrase, sent=sent, phrase_raw=raw,
               vhat=torch.stack(vh), qhat=torch.stack(qh))
    out = {k: v.detach().numpy() for k, v in out.items()}
    grads = {}
    for prefix, m in (("question_encoder", qe), ("co_attention", ca), ("mlp_classify", ml)):
        for n, prm in m.named_parameters():
            key = f"{prefix}.{n}"
            if prm.grad is None:
                assert key.startswith("co_attention.W_b"), key    # the dead affinity layer (model.py:347)
                continue
            grads[key] = prm.grad.numpy()
    out["grads"] = grads
    if feats_grad:
        out["dfeats"] = feats.grad.numpy()
    return out


def phrase_idx_from_reference(p, x, d, vocab, dtype):
    """Max-pool argmax as the reference's MaxPool2d computes it (return_indices on the same tensor)."""
    qe = ref.QuestionCoAttentionEncoder(vocab, d, d)
    sd = {k[len("question_encoder."):]: torch.from_numpy(v) for k, v in p.items() if k.startswith("question_encoder.")}
    qe.to(dtype)
    qe.load_state_dict(sd)
    pc = qe.phrase_conv_pool
    with torch.no_grad():
        e = qe.word_embedding(torch.from_numpy(x["tokens"]))
        B, T, E = e.shape
        xq = e.permute(0, 2, 1)
        cat = torch.cat([pc.conv_unigram(xq), pc.conv_bigram(xq), pc.conv_trigram(xq)], dim=1)   # model.py:319-324
        cat = cat.permute(0, 2, 1).unsqueeze(3).reshape(B, T, E, 3)                               # model.py:327-329
        _, idx = torch.nn.functional.max_pool2d(cat, (1, 3), return_indices=True)                 # model.py:311,332
        # MaxPool2d indices are flat over the pooled (H=E, W=3) plane -> j = idx % 3
        return (idx.squeeze(3) % 3).to(torch.uint8).numpy()


def digest(a):
    a = np.asarray(a, np.float64).reshape(-1)
    stride = max(1, a.size // 64)
    return np.concatenate([[a.sum(), np.sqrt((a * a).sum())], a[::stride][:64]])


def main():
    gold = os.path.join(ROOT, "tests", "golden")
    os.makedirs(gold, exist_ok=True)

    # ---- tiny configs, everything stored ----------------------------------------------------------
    small = [dict(name="small_a", B=3, N=10, T=7, d=32, vocab=20, K=11, mlp_dim=48, seed=1, dist="D1", min_len=1),
             dict(name="small_b", B=5, N=12, T=4, d=64, vocab=30, K=7, mlp_dim=32, seed=2, dist="D2", min_len=1),
             dict(name="small_c", B=2, N=5, T=1, d=32, vocab=9, K=3, mlp_dim=16, seed=3, dist="D1", min_len=1)]
    for c in small:
        p = syn.make_params(c["d"], c["vocab"], c["K"], c["mlp_dim"], seed=10 + c["seed"], dtype=np.float64)
        x = syn.make_inputs(c["B"], c["N"], c["T"], c["d"], c["vocab"], c["K"], seed=c["seed"], dist=c["dist"],
                            min_len=c["min_len"], dtype=np.float64)
        r64 = run_reference(p, x, c["d"], c["vocab"], c["K"], c["mlp_dim"], torch.float64, feats_grad=True)
        p32 = {k: v.astype(np.float32) for k, v in p.items()}
        x32 = dict(x, feats=x["feats"].astype(np.float32))
        r32 = run_reference(p32, x32, c["d"], c["vocab"], c["K"], c["mlp_dim"], torch.float32, feats_grad=True)
        idx = phrase_idx_from_reference(p, x, c["d"], c["vocab"], torch.float64)
        blob = {f"cfg.{k}": np.asarray(v) for k, v in c.items() if k != "name" and k != "dist"}
        blob["cfg.dist"] = np.asarray(c["dist"])
        blob.update({f"p.{k}": v for k, v in p.items()})
        blob.update({f"x.{k}": v for k, v in x.items()})
        blob["idx"] = idx
        for tag, r in (("f64", r64), ("f32", r32)):
            for k, v in r.items():
                if k == "grads":
                    blob.update({f"{tag}.grad.{n}": g for n, g in v.items()})
                else:
                    blob[f"{tag}.{k}"] = v
        np.savez_compressed(os.path.join(gold, c["name"] + ".npz"), **blob)
        print("wrote", c["name"], "loss64", float(r64["loss"]), "loss32", float(r32["loss"]))

    # ---- real widths, digests only ----------------------------------------------------------------
    big = [dict(name="d512_D1", B=4, N=196, T=26, d=512, vocab=10000, K=1001, mlp_dim=1024, seed=1, dist="D1", min_len=3),
           dict(name="d512_D2", B=6, N=196, T=26, d=512, vocab=10000, K=1001, mlp_dim=1024, seed=5, dist="D2", min_len=1)]
    for c in big:
        p = syn.make_params(c["d"], c["vocab"], c["K"], c["mlp_dim"], seed=0, dtype=np.float32)
        x = syn.make_inputs(c["B"], c["N"], c["T"], c["d"], c["vocab"], c["K"], seed=c["seed"], dist=c["dist"], min_len=c["min_len"])
        p64 = {k: v.astype(np.float64) for k, v in p.items()}
        x64 = dict(x, feats=x["feats"].astype(np.float64))
        r64 = run_reference(p64, x64, c["d"], c["vocab"], c["K"], c["mlp_dim"], torch.float64, feats_grad=True)
        r32 = run_reference(p, x, c["d"], c["vocab"], c["K"], c["mlp_dim"], torch.float32, feats_grad=True)
        idx64 = phrase_idx_from_reference(p64, x64, c["d"], c["vocab"], torch.float64)
        idx32 = phrase_idx_from_reference(p, x, c["d"], c["vocab"], torch.float32)
        blob = {f"cfg.{k}": np.asarray(v) for k, v in c.items() if k != "name"}
        blob["idx64"], blob["idx32"] = idx64, idx32
        for tag, r in (("f64", r64), ("f32", r32)):
            blob[f"{tag}.logits"] = r["logits"]
            blob[f"{tag}.loss"] = r["loss"]
            blob[f"{tag}.vhat"] = r["vhat"].astype(np.float32)
            blob[f"{tag}.qhat"] = r["qhat"].astype(np.float32)
            blob[f"{tag}.dfeats.digest"] = digest(r["dfeats"])
            for n, g in r["grads"].items():
                blob[f"{tag}.grad.{n}.digest"] = digest(g)
        np.savez_compressed(os.path.join(gold, c["name"] + ".npz"), **blob)
        print("wrote", c["name"], "loss64", float(r64["loss"]), "idx flips f32 vs f64:", int((idx64 != idx32).sum()))


def record_state_dict_keys():
    """Key names and shapes of the reference modules' state_dicts (the checkpoint-compat contract, main.py:168-176)."""
    import json
    d, vocab, K, mlp = 512, 10000, 1001, 1024
    out = {}
    for prefix, m in (("question_encoder", ref.QuestionCoAttentionEncoder(vocab, d, d)), ("co_attention", ref.ParallelCoAttention(d)),
                      ("mlp_classify", ref.MLPClassifier(d, mlp, K))):
        for k, v in m.state_dict().items():
            out[f"{prefix}.{k}"] = list(v.shape)
    base_q = ref.QuestionBaselineEncoder(vocab, 300, 1024)
    for k, v in base_q.state_dict().items():
        out[f"baseline.question_encoder.{k}"] = list(v.shape)
    json.dump(out, open(os.path.join(ROOT, "tests", "golden", "state_dict_keys.json"), "w"), indent=1, sort_keys=True)
    print("wrote state_dict_keys.json", len(out))


WRAPPER_CFG = dict(B=2, img=448, T=9, vocab=200, K=21, mlp_dim=1024, d=512, vgg_seed=1234, seed=7)


def make_vgg_weights_file(path, seed):
    """Random-init vgg11_bn state_dict written the way ``--vgg_wts_path`` expects it (main.py:66, model.py:229-234).  torch's CPU
    generator is deterministic for a given version, so the GPU-side test regenerates the same file from the same seed (and
    checks the digest stored in the fixture before trusting it)."""
    import torchvision
    torch.manual_seed(seed)
    vgg = torchvision.models.vgg11_bn(weights=None)
    torch.save(vgg.state_dict(), path)
    sd = vgg.state_dict()
    return np.asarray([float(sd["features.0.weight"].double().sum()), float(sd["features.25.weight"].double().norm())])


def wrapper_fixture():
    """The FULL reference wrapper (model.py:157-187): random-init VGG11-bn trunk from a generated weights file -> permuted
    [B,196,512] view -> question encoder -> co-attention -> MLP, fp32 on CPU, eval mode (BatchNorm on its running statistics).
    Stores logits, loss, a digest of the image features, the gradient digests of every trainable tensor, and the full list of
    state_dict keys / shapes (85 keys: 56 of them VGG's)."""
    import tempfile
    c = WRAPPER_CFG
    path = os.path.join(tempfile.mkdtemp(), "vgg11_bn_random.pth")
    vgg_digest = make_vgg_weights_file(path, c["vgg_seed"])
    net = ref.HierarchicalCoAttentionNet(dict(vocab_size=c["vocab"], word_emb_dim=c["d"], hidden_dim=c["d"]),
                                         dict(is_trainable=False, weights_path=path), K=c["K"], mlp_dim=c["mlp_dim"])
    p = syn.make_params(c["d"], c["vocab"], c["K"], c["mlp_dim"], seed=c["seed"])
    missing, unexpected = net.load_state_dict({k: torch.from_numpy(v) for k, v in p.items()}, strict=False)
    assert not unexpected and all(k.startswith("image_encoder.") for k in missing)
    net.eval()
    x = syn.make_inputs(c["B"], 196, c["T"], c["d"], c["vocab"], c["K"], seed=c["seed"], min_len=1)
    rng = np.random.RandomState(c["seed"])
    images = rng.standard_normal((c["B"], 3, c["img"], c["img"])).astype(np.float32)
    feats = net.image_encoder(torch.from_numpy(images))
    assert tuple(feats.shape) == (c["B"], 196, 512) and feats.stride() == (196 * 512, 1, 196) and not feats.requires_grad
    logits = net(torch.from_numpy(images), torch.from_numpy(x["tokens"]), torch.from_numpy(x["lens"]))     # main.py:211
    loss = torch.nn.CrossEntropyLoss()(logits, torch.from_numpy(x["labels"]))                               # main.py:179,214
    loss.backward()
    blob = {f"cfg.{k}": np.asarray(v) for k, v in c.items()}
    blob["vgg_digest"] = vgg_digest
    blob["logits"] = logits.detach().numpy()
    blob["loss"] = loss.detach().numpy()
    blob["feats.digest"] = digest(feats.detach().numpy())
    for n, prm in net.named_parameters():
        if prm.grad is not None:
            blob[f"grad.{n}.digest"] = digest(prm.grad.numpy())
        else:
            assert n.startswith("co_attention.W_b") or n.startswith("image_encoder."), n
    import json
    blob["state_dict"] = np.asarray(json.dumps({k: list(v.shape) for k, v in net.state_dict().items()}))
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "wrapper_448.npz"), **blob)
    print("wrote wrapper_448", "loss", float(loss), "keys", len(net.state_dict()))


if __name__ == "__main__":
    if "--wrapper-only" not in sys.argv:
        main()
        record_state_dict_keys()
    wrapper_fixture()
