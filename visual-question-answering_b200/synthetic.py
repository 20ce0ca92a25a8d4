"""Seeded synthetic inputs and weights for the Hierarchical Co-Attention path.

numpy-only so that the very same arrays can be regenerated on any box (the build container, the GPU
box) without depending on torch's RNG streams.  Conventions restated from the reference:

  * tokens int64 padded with 0 (<PAD>) to a fixed T, valid ids >= 1   (dataloader.py:58-65, utils.py:18-30,106)
  * lens int64 >= 1, batch sorted by length, descending                (utils.py:33-45, main.py:202)
  * labels in [0, K) with K = num_cls + 1                              (main.py:155)
  * image features [B, N, d] fp32, frozen VGG => requires_grad False    (model.py:205-219, main.py:67)

Weight scales follow PyTorch's default initialisers, which is what the reference uses
(model.py:263-269, 306-308, 347-354, 409-412): Linear/Conv1d U(+-1/sqrt(fan_in)), Embedding N(0,1)
with row 0 zeroed, LSTM U(+-1/sqrt(hidden)).
"""
from __future__ import annotations

import numpy as np


def make_inputs(B, N=196, T=26, d=512, vocab=10000, K=1001, seed=1, dist="D1", min_len=3, dtype=np.float32):
    """Returns dict(feats [B,N,d], tokens [B,T] i64, lens [B] i64 (sorted desc), labels [B] i64)."""
    rng = np.random.RandomState(seed)
    lo = min(min_len, T)
    lens = np.sort(rng.randint(lo, T + 1, size=B))[::-1].astype(np.int64).copy()
    tokens = np.zeros((B, T), np.int64)
    for b in range(B):
        tokens[b, : lens[b]] = rng.randint(1, vocab, size=lens[b])
    labels = rng.randint(0, K, size=B).astype(np.int64)
    feats = rng.standard_normal((B, N, d)).astype(dtype)
    if dist == "D2":                      # VGG-like: post-ReLU / max-pool features are non-negative
        feats = np.maximum(feats, 0)
    elif dist != "D1":
        raise ValueError(dist)
    return dict(feats=feats, tokens=tokens, lens=lens, labels=labels)


def make_params(d=512, vocab=10000, K=1001, mlp_dim=1024, seed=0, dtype=np.float32):
    """Random-init weights keyed by the reference's state_dict names."""
    rng = np.random.RandomState(seed)

    def U(shape, fan_in):
        b = 1.0 / np.sqrt(fan_in)
        return rng.uniform(-b, b, size=shape).astype(dtype)

    p = {}
    emb = rng.standard_normal((vocab, d)).astype(dtype)
    emb[0] = 0
    p["question_encoder.word_embedding.weight"] = emb
    for name, k in (("unigram", 1), ("bigram", 2), ("trigram", 3)):
        p[f"question_encoder.phrase_conv_pool.conv_{name}.1.weight"] = U((d, d, k), d * k)
        p[f"question_encoder.phrase_conv_pool.conv_{name}.1.bias"] = U((d,), d * k)
    p["question_encoder.sentence_lstm.weight_ih_l0"] = U((4 * d, d), d)
    p["question_encoder.sentence_lstm.weight_hh_l0"] = U((4 * d, d), d)
    p["question_encoder.sentence_lstm.bias_ih_l0"] = U((4 * d,), d)
    p["question_encoder.sentence_lstm.bias_hh_l0"] = U((4 * d,), d)
    for n in ("W_b", "W_v", "W_q"):
        p[f"co_attention.{n}.weight"] = U((d, d), d)
        p[f"co_attention.{n}.bias"] = U((d,), d)
    for n in ("w_v", "w_q"):
        p[f"co_attention.{n}.weight"] = U((1, d), d)
        p[f"co_attention.{n}.bias"] = U((1,), d)
    for n, (o, i) in (("W_w", (d, d)), ("W_p", (d, 2 * d)), ("W_s", (mlp_dim, 2 * d)), ("W_h", (K, mlp_dim))):
        p[f"mlp_classify.{n}.weight"] = U((o, i), i)
        p[f"mlp_classify.{n}.bias"] = U((o,), i)
    return p
