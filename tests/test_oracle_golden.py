"""Pins oracle/hiecoattn_oracle.py against outputs of the reference itself (tests/golden/*.npz,
written by oracle/make_golden.py from /root/reference/model.py).  CPU only."""
import glob
import os

import numpy as np
import pytest

import hiecoattn_oracle as O
from conftest import GOLDEN


def rel(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300)


def load(name):
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    return {k: z[k] for k in z.files}


ZERO_BIASES = ("co_attention.w_v.bias", "co_attention.w_q.bias")   # analytically 0 (softmax shift invariance)


@pytest.mark.parametrize("name", ["small_a", "small_b", "small_c"])
@pytest.mark.parametrize("tag,tol", [("f64", 1e-11), ("f32", 2e-5)])
def test_small_everything(name, tag, tol):
    g = load(name)
    dt = np.float64 if tag == "f64" else np.float32
    p = {k[2:]: v.astype(dt) for k, v in g.items() if k.startswith("p.")}
    x = {k[2:]: v for k, v in g.items() if k.startswith("x.")}
    out = O.hiecoattn_step(p, x["feats"].astype(dt), x["tokens"], x["lens"], x["labels"], need_dfeats=True)
    c = out["cache"]
    assert rel(out["logits"], g[f"{tag}.logits"]) < tol
    assert abs(float(out["loss"]) - float(g[f"{tag}.loss"])) < tol * 10
    for mine, key in ((c["word"], "word"), (c["phrase"], "phrase"), (c["sent"], "sent"), (c["phrase_raw"], "phrase_raw"),
                      (np.stack(c["vhats"]), "vhat"), (np.stack(c["qhats"]), "qhat"), (out["dfeats"], "dfeats")):
        assert rel(mine, g[f"{tag}.{key}"]) < tol, key
    if tag == "f64":
        assert np.array_equal(out["idx"], g["idx"])          # max-pool argmax, bit exact
    seen = 0
    for k in O.PARAM_KEYS:
        ref = g[f"{tag}.grad.{k}"]
        mine = out["grads"][k]
        assert mine.shape == ref.shape, k
        if k in ZERO_BIASES:
            assert np.abs(mine).max() < 1e-6 and np.abs(ref).max() < 1e-6
        else:
            assert rel(mine, ref) < tol * 5, (k, rel(mine, ref))
        seen += 1
    assert seen == 27
    assert not any(k.startswith(f"{tag}.grad.co_attention.W_b") for k in g)   # F1: W_b gets no gradient


def _digest(a):
    a = np.asarray(a, np.float64).reshape(-1)
    stride = max(1, a.size // 64)
    return np.concatenate([[a.sum(), np.sqrt((a * a).sum())], a[::stride][:64]])


@pytest.mark.parametrize("name", ["d512_D1", "d512_D2"])
def test_real_widths_digests(name, syn):
    g = load(name)
    c = {k[4:]: g[k].item() for k in g if k.startswith("cfg.")}
    p = syn.make_params(c["d"], c["vocab"], c["K"], c["mlp_dim"], seed=0)
    x = syn.make_inputs(c["B"], c["N"], c["T"], c["d"], c["vocab"], c["K"], seed=c["seed"], dist=str(c["dist"]), min_len=c["min_len"])
    p64 = {k: v.astype(np.float64) for k, v in p.items()}
    out = O.hiecoattn_step(p64, x["feats"].astype(np.float64), x["tokens"], x["lens"], x["labels"], need_dfeats=True)
    assert rel(out["logits"], g["f64.logits"]) < 1e-10
    assert np.array_equal(out["idx"], g["idx64"])
    assert np.array_equal(g["idx64"], g["idx32"])
    assert rel(np.stack(out["cache"]["vhats"]), g["f64.vhat"]) < 1e-6      # stored as fp32
    assert rel(np.stack(out["cache"]["qhats"]), g["f64.qhat"]) < 1e-6
    assert rel(_digest(out["dfeats"]), g["f64.dfeats.digest"]) < 1e-9
    for k in O.PARAM_KEYS:
        d_ref = g[f"f64.grad.{k}.digest"]
        d_me = _digest(out["grads"][k])
        if k in ZERO_BIASES:
            assert np.abs(d_me).max() < 1e-9
        else:
            assert rel(d_me[1:], d_ref[1:]) < 1e-9, k      # norm + samples (the plain sum cancels heavily)
    # fp32 reference vs fp64 reference: the noise floor quoted in DESIGN.md
    assert rel(g["f32.logits"], g["f64.logits"]) < 5e-6


def test_quirk_F2_consecutive_triples():
    """model.py:324-332 pools consecutive channel triples of [uni|bi|tri], not (uni,bi,tri) of one channel."""
    rng = np.random.RandomState(0)
    E, T = 6, 4
    x = rng.standard_normal((1, T, E))
    ws = [rng.standard_normal((E, E, k)) for k in (1, 2, 3)]
    bs = [rng.standard_normal(E) for _ in range(3)]
    out, idx = O.phrase_conv_pool_fwd(x, ws[0], bs[0], ws[1], bs[1], ws[2], bs[2])
    cat = np.tanh(O.phrase_conv_preact(x, ws[0], bs[0], ws[1], bs[1], ws[2], bs[2]))
    assert np.array_equal(out[0, :, 0], cat[0, :, 0:3].max(1))          # channel 0 sees only unigram 0..2
    assert np.array_equal(out[0, :, 2], cat[0, :, 6:9].max(1))          # channel 2 sees bigram 0..2
    paper = np.stack([cat[..., :E], cat[..., E:2 * E], cat[..., 2 * E:]], 3).max(3)
    assert not np.allclose(out, paper)


def test_quirk_F3_F4_padding():
    """Pad rows of Q are exactly zero at all three levels yet still receive attention mass (model.py:388)."""
    g = load("small_a")
    p = {k[2:]: v for k, v in g.items() if k.startswith("p.")}
    x = {k[2:]: v for k, v in g.items() if k.startswith("x.")}
    _, c = O.hiecoattn_forward(p, x["feats"], x["tokens"], x["lens"], want_cache=True)
    m = ~O.valid_mask(x["lens"], x["tokens"].shape[1])
    assert m.any()
    for q in (c["word"], c["phrase"], c["sent"]):
        assert np.abs(q[m]).max() == 0.0
    for cache in c["ca_caches"]:
        assert (cache["aq"][m] > 0).all()


def test_tie_rule_first_index_wins():
    x = np.zeros((1, 1, 3))
    w = [np.zeros((3, 3, k)) for k in (1, 2, 3)]
    b = [np.zeros(3)] * 3
    out, idx = O.phrase_conv_pool_fwd(x, w[0], b[0], w[1], b[1], w[2], b[2])
    assert (idx == 0).all()
