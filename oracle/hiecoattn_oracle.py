"""CPU oracle for the Hierarchical Co-Attention hot path (numpy, explicit forward AND backward).

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is part of the product path: only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may import
it, and only as the checker / the CPU baseline.  The product (``visual-question-answering_b200``)
fails loudly when its CUDA extension is missing; it never routes through this file.

What it restates (all citations relative to the reference tree, /root/reference):

  * ``embedding``            model.py:263,282   (nn.Embedding, padding_idx=0)
  * ``phrase_conv_pool``     model.py:304-334   (uni/bi/tri Conv1d + tanh, concat, reshape, MaxPool2d((1,3)))
  * ``lstm``                 model.py:269,287-296 (pack -> nn.LSTM -> pad, zero-filled pads)
  * ``coattn``               model.py:342-397   (ParallelCoAttention, 3 levels, shared weights, W_b unused)
  * ``mlp``                  model.py:406-434   (MLPClassifier)
  * ``cross_entropy``        main.py:179,214    (nn.CrossEntropyLoss, mean)
  * ``hiecoattn_step``       model.py:171-187 + main.py:211-220 (forward + loss + backward)

The arithmetic of the reference lives in PyTorch/ATen (requirements.txt:6 pins torch==1.2.0; the
installed build is 2.11), whose source is not in the reference tree.  The algorithms restated here
are the published definitions of those ATen ops; every backward formula is written out by hand
(SURVEY.md section 3.3/3.4), so agreement with the reference's autograd is a real check of both.

Parity pinning: the reference has no tests, golden vectors or fixtures of its own for this path
(SURVEY.md section 4), so the oracle is pinned against outputs of the reference ITSELF, generated in
the build container by ``oracle/make_golden.py`` (which imports /root/reference/model.py, runs it in
fp64 and fp32 on CPU, and writes ``tests/golden/*.npz``).  ``tests/test_oracle_golden.py`` checks this
file against those fixtures on every run.

All functions are dtype-generic: pass float64 arrays for a noise-free check, float32 to mimic the
reference's working precision.
"""
from __future__ import annotations

import numpy as np

# --------------------------------------------------------------------------------------------
# helpers


def _sigmoid(x):
    return 1.0 / (1.0 + np.exp(-x))


def _softmax(x, axis):
    m = x.max(axis=axis, keepdims=True)
    e = np.exp(x - m)
    return e / e.sum(axis=axis, keepdims=True)


def valid_mask(lens, T):
    """[B,T] bool, True where t < len_b  (pack/pad semantics of model.py:287-296)."""
    lens = np.asarray(lens).reshape(-1)
    return np.arange(T)[None, :] < lens[:, None]


# --------------------------------------------------------------------------------------------
# embedding  (model.py:263, 282)


def embedding_fwd(tokens, table):
    """tokens [B,T] int64, table [V,E] -> [B,T,E].  Row 0 of the table is the pad row."""
    return table[np.asarray(tokens)]


def embedding_bwd(tokens, dout, vocab_size):
    """Scatter-add of dout rows into a [V,E] gradient; padding_idx=0 receives zero gradient."""
    tokens = np.asarray(tokens).reshape(-1)
    d = dout.reshape(tokens.shape[0], -1)
    g = np.zeros((vocab_size, d.shape[1]), dtype=dout.dtype)
    np.add.at(g, tokens, d)
    g[0] = 0
    return g


# --------------------------------------------------------------------------------------------
# PhraseConvPool  (model.py:304-334)


def _shift(x, s):
    """y[:, t] = x[:, t+s] with zeros outside [0,T)  (ConstantPad1d, model.py:306-308)."""
    y = np.zeros_like(x)
    T = x.shape[1]
    if s == 0:
        y[:] = x
    elif s < 0:
        y[:, -s:] = x[:, : T + s]
    else:
        y[:, : T - s] = x[:, s:]
    return y


def phrase_conv_preact(x, w1, b1, w2, b2, w3, b3):
    """Pre-activations of the three convs, concatenated: [B,T,3E] (model.py:319-327).

    Conv1d weights are [C_out, C_in, k] cross-correlation kernels.
      uni[t] = W1[:,:,0] x[t] + b1                               pad (0,0)
      bi [t] = W2[:,:,0] x[t-1] + W2[:,:,1] x[t] + b2            pad (1,0)
      tri[t] = W3[:,:,0] x[t-1] + W3[:,:,1] x[t] + W3[:,:,2] x[t+1] + b3   pad (1,1)
    """
    xm, xp = _shift(x, -1), _shift(x, +1)
    uni = x @ w1[:, :, 0].T + b1
    bi = xm @ w2[:, :, 0].T + x @ w2[:, :, 1].T + b2
    tri = xm @ w3[:, :, 0].T + x @ w3[:, :, 1].T + xp @ w3[:, :, 2].T + b3
    return np.concatenate([uni, bi, tri], axis=2)


def phrase_conv_pool_fwd(x, w1, b1, w2, b2, w3, b3):
    """x [B,T,E] -> out [B,T,E], idx [B,T,E] uint8.

    out[b,t,e] = max_j cat[b,t,3e+j] over CONSECUTIVE channel triples of the concatenation
    (the reshape at model.py:329 regroups the 3E axis, it does not interleave uni/bi/tri);
    idx = first j attaining the max (MaxPool2d tie rule).
    """
    B, T, E = x.shape
    cat = np.tanh(phrase_conv_preact(x, w1, b1, w2, b2, w3, b3))
    grp = cat.reshape(B, T, E, 3)
    idx = grp.argmax(axis=3).astype(np.uint8)          # numpy argmax returns the first maximum
    out = np.take_along_axis(grp, idx[..., None].astype(np.int64), axis=3)[..., 0]
    return out, idx


def phrase_conv_pool_bwd(x, w1, w2, w3, out, idx, dout):
    """Backward of phrase_conv_pool_fwd.  Returns dx, dw1, db1, dw2, db2, dw3, db3.

    The pooled gradient goes to source channel 3e+idx only, times (1 - out^2) for the tanh.
    """
    B, T, E = x.shape
    dsel = dout * (1.0 - out * out)
    dgrp = np.zeros((B, T, E, 3), dtype=x.dtype)
    np.put_along_axis(dgrp, idx[..., None].astype(np.int64), dsel[..., None], axis=3)
    dcat = dgrp.reshape(B, T, 3 * E)
    du, dbi, dtr = dcat[..., :E], dcat[..., E:2 * E], dcat[..., 2 * E:]
    xm, xp = _shift(x, -1), _shift(x, +1)
    f = lambda a: a.reshape(-1, E)
    dw1 = (f(du).T @ f(x))[:, :, None]
    dw2 = np.stack([f(dbi).T @ f(xm), f(dbi).T @ f(x)], axis=2)
    dw3 = np.stack([f(dtr).T @ f(xm), f(dtr).T @ f(x), f(dtr).T @ f(xp)], axis=2)
    db1, db2, db3 = f(du).sum(0), f(dbi).sum(0), f(dtr).sum(0)
    # d/dx[t]: terms that read x[t] directly, x[t] as "t-1" of position t+1, x[t] as "t+1" of t-1
    dx = du @ w1[:, :, 0] + dbi @ w2[:, :, 1] + dtr @ w3[:, :, 1]
    dx += _shift(dbi @ w2[:, :, 0] + dtr @ w3[:, :, 0], +1)
    dx += _shift(dtr @ w3[:, :, 2], -1)
    return dx, dw1, db1, dw2, db2, dw3, db3


# --------------------------------------------------------------------------------------------
# LSTM over the valid prefix of each sequence  (model.py:269, 287-296)


def lstm_fwd(x, lens, w_ih, w_hh, b_ih, b_hh):
    """x [B,T,E] (only t < len_b is read) -> y [B,T,H] with zeros at t >= len_b, plus a cache.

    Gate order i,f,g,o (PyTorch); h0 = c0 = 0; each sequence stops at its own length, which is what
    pack_padded_sequence / pad_packed_sequence(total_length=T) implement.
    """
    B, T, _ = x.shape
    H = w_hh.shape[1]
    m = valid_mask(lens, T)
    h = np.zeros((B, H), x.dtype)
    c = np.zeros((B, H), x.dtype)
    y = np.zeros((B, T, H), x.dtype)
    cache = []
    for t in range(T):
        z = x[:, t] @ w_ih.T + b_ih + h @ w_hh.T + b_hh
        i, f, g, o = _sigmoid(z[:, :H]), _sigmoid(z[:, H:2 * H]), np.tanh(z[:, 2 * H:3 * H]), _sigmoid(z[:, 3 * H:])
        c_new = f * c + i * g
        h_new = o * np.tanh(c_new)
        mt = m[:, t:t + 1]
        cache.append((h, c, i, f, g, o, c_new, mt))
        h = np.where(mt, h_new, h)
        c = np.where(mt, c_new, c)
        y[:, t] = np.where(mt, h_new, 0)
    return y, cache


def lstm_bwd(x, w_ih, w_hh, cache, dy):
    """Backward through lstm_fwd.  Returns dx, dw_ih, dw_hh, db_ih, db_hh."""
    B, T, _ = x.shape
    H = w_hh.shape[1]
    dx = np.zeros_like(x)
    dw_ih = np.zeros_like(w_ih)
    dw_hh = np.zeros_like(w_hh)
    db = np.zeros((4 * H,), x.dtype)
    dh_next = np.zeros((B, H), x.dtype)
    dc_next = np.zeros((B, H), x.dtype)
    for t in reversed(range(T)):
        h_prev, c_prev, i, f, g, o, c_new, mt = cache[t]
        dh = np.where(mt, dy[:, t] + dh_next, 0)
        tc = np.tanh(c_new)
        dc = np.where(mt, dc_next, 0) + dh * o * (1 - tc * tc)
        dz = np.concatenate([dc * g * i * (1 - i), dc * c_prev * f * (1 - f), dc * i * (1 - g * g), dh * tc * o * (1 - o)], axis=1)
        dx[:, t] = dz @ w_ih
        dw_ih += dz.T @ x[:, t]
        dw_hh += dz.T @ h_prev
        db += dz.sum(0)
        # positions past the end pass the carried gradient through unchanged
        dh_next = dz @ w_hh + np.where(mt, 0, dh_next)
        dc_next = dc * f + np.where(mt, 0, dc_next)
    return dx, dw_ih, dw_hh, db.copy(), db.copy()


# --------------------------------------------------------------------------------------------
# ParallelCoAttention  (model.py:342-397)


def coattn_level_fwd(V, Q, Wv, bv, Wq, bq, wv, cv, wq, cq):
    """One level.  V [B,N,d], Q [B,T,d]; Wv/Wq [d,d] (nn.Linear: x @ W.T + b); wv/wq [d]; cv/cq scalars.

    C = tanh(Q V^T) -- W_b is declared at model.py:347 but never applied (model.py:377).
    The question softmax runs over all T positions, pads included (model.py:388).
    """
    C = np.tanh(Q @ V.transpose(0, 2, 1))                   # [B,T,N]   model.py:377
    PV = V @ Wv.T + bv                                      # [B,N,d]   model.py:380
    PQ = Q @ Wq.T + bq                                      # [B,T,d]   model.py:381
    Hv = np.tanh(PV + C.transpose(0, 2, 1) @ PQ)            # [B,N,d]   model.py:380-381
    Hq = np.tanh(PQ + C @ PV)                               # [B,T,d]   model.py:383-384
    av = _softmax(Hv @ wv + cv, axis=1)                     # [B,N]     model.py:387
    aq = _softmax(Hq @ wq + cq, axis=1)                     # [B,T]     model.py:388
    vhat = (av[..., None] * V).sum(1)                       # [B,d]     model.py:391
    qhat = (aq[..., None] * Q).sum(1)                       # [B,d]     model.py:392
    return vhat, qhat, dict(C=C, PV=PV, PQ=PQ, Hv=Hv, Hq=Hq, av=av, aq=aq)


def coattn_level_bwd(V, Q, Wv, Wq, wv, wq, cache, gv, gq):
    """Backward of one level given gv = dL/dvhat, gq = dL/dqhat."""
    C, PV, PQ, Hv, Hq, av, aq = (cache[k] for k in ("C", "PV", "PQ", "Hv", "Hq", "av", "aq"))
    dav = V @ gv[..., None]                                 # [B,N,1]
    daq = Q @ gq[..., None]
    dsv = av * (dav[..., 0] - (av * dav[..., 0]).sum(1, keepdims=True))
    dsq = aq * (daq[..., 0] - (aq * daq[..., 0]).sum(1, keepdims=True))
    dwv = np.einsum("bnd,bn->d", Hv, dsv)
    dwq = np.einsum("btd,bt->d", Hq, dsq)
    dcv, dcq = dsv.sum(), dsq.sum()                         # analytically zero
    dZv = dsv[..., None] * wv * (1 - Hv * Hv)               # [B,N,d]
    dZq = dsq[..., None] * wq * (1 - Hq * Hq)               # [B,T,d]
    Ct = C.transpose(0, 2, 1)
    dPV = dZv + Ct @ dZq
    dPQ = dZq + C @ dZv
    dC = PQ @ dZv.transpose(0, 2, 1) + dZq @ PV.transpose(0, 2, 1)
    dS = dC * (1 - C * C)
    dQ = dS @ V + dPQ @ Wq + aq[..., None] * gq[:, None, :]
    dV = dS.transpose(0, 2, 1) @ Q + dPV @ Wv + av[..., None] * gv[:, None, :]
    d = V.shape[2]
    dWv = dPV.reshape(-1, d).T @ V.reshape(-1, d)
    dWq = dPQ.reshape(-1, d).T @ Q.reshape(-1, d)
    dbv = dPV.reshape(-1, d).sum(0)
    dbq = dPQ.reshape(-1, d).sum(0)
    return dict(dV=dV, dQ=dQ, dWv=dWv, dbv=dbv, dWq=dWq, dbq=dbq, dwv=dwv, dcv=dcv, dwq=dwq, dcq=dcq)


def coattn_fwd(V, Qs, Wv, bv, Wq, bq, wv, cv, wq, cq):
    """Three levels with shared weights (model.py:372).  Returns lists vhats, qhats, caches."""
    outs = [coattn_level_fwd(V, Q, Wv, bv, Wq, bq, wv, cv, wq, cq) for Q in Qs]
    return [o[0] for o in outs], [o[1] for o in outs], [o[2] for o in outs]


def coattn_bwd(V, Qs, Wv, Wq, wv, wq, caches, gvs, gqs):
    """Weight grads sum over the three levels; dV too; dQ is per level."""
    tot = None
    dQs = []
    for Q, cache, gv, gq in zip(Qs, caches, gvs, gqs):
        g = coattn_level_bwd(V, Q, Wv, Wq, wv, wq, cache, gv, gq)
        dQs.append(g.pop("dQ"))
        if tot is None:
            tot = g
        else:
            for k in tot:
                tot[k] = tot[k] + g[k]
    tot["dQs"] = dQs
    return tot


# --------------------------------------------------------------------------------------------
# MLPClassifier  (model.py:406-434)


def mlp_fwd(vs, qs, Ww, bw, Wp, bp, Ws, bs, Wh, bh):
    x_w = qs[0] + vs[0]
    h_w = np.tanh(x_w @ Ww.T + bw)                                              # model.py:427
    x_p = np.concatenate([qs[1] + vs[1], h_w], axis=1)
    h_p = np.tanh(x_p @ Wp.T + bp)                                              # model.py:428
    x_s = np.concatenate([qs[2] + vs[2], h_p], axis=1)
    h_s = np.tanh(x_s @ Ws.T + bs)                                              # model.py:429
    logits = h_s @ Wh.T + bh                                                    # model.py:432
    return logits, dict(x_w=x_w, h_w=h_w, x_p=x_p, h_p=h_p, x_s=x_s, h_s=h_s)


def mlp_bwd(Ww, Wp, Ws, Wh, cache, dlogits):
    x_w, h_w, x_p, h_p, x_s, h_s = (cache[k] for k in ("x_w", "h_w", "x_p", "h_p", "x_s", "h_s"))
    d = Ww.shape[0]
    dWh = dlogits.T @ h_s
    dbh = dlogits.sum(0)
    dzs = (dlogits @ Wh) * (1 - h_s * h_s)
    dWs, dbs = dzs.T @ x_s, dzs.sum(0)
    dxs = dzs @ Ws
    g_s = dxs[:, :d]
    dzp = dxs[:, d:] * (1 - h_p * h_p)
    dWp, dbp = dzp.T @ x_p, dzp.sum(0)
    dxp = dzp @ Wp
    g_p = dxp[:, :d]
    dzw = dxp[:, d:] * (1 - h_w * h_w)
    dWw, dbw = dzw.T @ x_w, dzw.sum(0)
    g_w = dzw @ Ww
    # q_l + v_l: the same gradient flows to both the question and the image feature of a level
    return dict(g=[g_w, g_p, g_s], dWw=dWw, dbw=dbw, dWp=dWp, dbp=dbp, dWs=dWs, dbs=dbs, dWh=dWh, dbh=dbh)


# --------------------------------------------------------------------------------------------
# loss  (main.py:179, 214)


def cross_entropy_fwd(logits, labels):
    m = logits.max(1, keepdims=True)
    lse = m[:, 0] + np.log(np.exp(logits - m).sum(1))
    B = logits.shape[0]
    return (lse - logits[np.arange(B), labels]).mean()


def cross_entropy_bwd(logits, labels):
    B = logits.shape[0]
    p = _softmax(logits, axis=1)
    p[np.arange(B), labels] -= 1
    return p / B


# --------------------------------------------------------------------------------------------
# the whole path: question encoder -> co-attention x3 -> MLP -> CE -> backward

PARAM_KEYS = (
    "question_encoder.word_embedding.weight",
    "question_encoder.phrase_conv_pool.conv_unigram.1.weight",
    "question_encoder.phrase_conv_pool.conv_unigram.1.bias",
    "question_encoder.phrase_conv_pool.conv_bigram.1.weight",
    "question_encoder.phrase_conv_pool.conv_bigram.1.bias",
    "question_encoder.phrase_conv_pool.conv_trigram.1.weight",
    "question_encoder.phrase_conv_pool.conv_trigram.1.bias",
    "question_encoder.sentence_lstm.weight_ih_l0",
    "question_encoder.sentence_lstm.weight_hh_l0",
    "question_encoder.sentence_lstm.bias_ih_l0",
    "question_encoder.sentence_lstm.bias_hh_l0",
    "co_attention.W_v.weight", "co_attention.W_v.bias",
    "co_attention.W_q.weight", "co_attention.W_q.bias",
    "co_attention.w_v.weight", "co_attention.w_v.bias",
    "co_attention.w_q.weight", "co_attention.w_q.bias",
    "mlp_classify.W_w.weight", "mlp_classify.W_w.bias",
    "mlp_classify.W_p.weight", "mlp_classify.W_p.bias",
    "mlp_classify.W_s.weight", "mlp_classify.W_s.bias",
    "mlp_classify.W_h.weight", "mlp_classify.W_h.bias",
)  # state_dict names of model.py:160-169 minus the frozen VGG and the dead co_attention.W_b


def hiecoattn_forward(p, feats, tokens, lens, want_cache=False):
    """p: dict keyed by PARAM_KEYS; feats [B,N,d]; tokens [B,T] int64; lens [B].  Returns logits (+cache)."""
    P = lambda k: p[k]
    qe, pc = "question_encoder.", "question_encoder.phrase_conv_pool."
    T = tokens.shape[1]
    m = valid_mask(lens, T)[..., None]
    word = embedding_fwd(tokens, P(qe + "word_embedding.weight"))                       # model.py:282
    conv_w = [P(pc + f"conv_{n}.1.{k}") for n in ("unigram", "bigram", "trigram") for k in ("weight", "bias")]
    phrase_raw, idx = phrase_conv_pool_fwd(word, *conv_w)                               # model.py:284
    phrase = np.where(m, phrase_raw, 0)                                                 # model.py:287,292
    lw = [P(qe + "sentence_lstm." + k) for k in ("weight_ih_l0", "weight_hh_l0", "bias_ih_l0", "bias_hh_l0")]
    sent, lstm_cache = lstm_fwd(phrase, lens, *lw)                                      # model.py:289,295
    ca = "co_attention."
    cw = (P(ca + "W_v.weight"), P(ca + "W_v.bias"), P(ca + "W_q.weight"), P(ca + "W_q.bias"),
          P(ca + "w_v.weight").reshape(-1), P(ca + "w_v.bias").reshape(()),
          P(ca + "w_q.weight").reshape(-1), P(ca + "w_q.bias").reshape(()))
    Qs = [word, phrase, sent]
    vhats, qhats, ca_caches = coattn_fwd(feats, Qs, *cw)                                # model.py:182
    ml = "mlp_classify."
    mw = [P(ml + f"{n}.{k}") for n in ("W_w", "W_p", "W_s", "W_h") for k in ("weight", "bias")]
    logits, mlp_cache = mlp_fwd(vhats, qhats, *mw)                                      # model.py:185
    if not want_cache:
        return logits
    return logits, dict(word=word, phrase_raw=phrase_raw, idx=idx, phrase=phrase, sent=sent, m=m,
                        lstm_cache=lstm_cache, ca_caches=ca_caches, mlp_cache=mlp_cache,
                        vhats=vhats, qhats=qhats, conv_w=conv_w, lw=lw, cw=cw, mw=mw)


def hiecoattn_step(p, feats, tokens, lens, labels, need_dfeats=False):
    """Forward + mean cross-entropy + full backward.  Returns dict(loss, logits, idx, grads{key: array}, [dfeats])."""
    logits, c = hiecoattn_forward(p, feats, tokens, lens, want_cache=True)
    loss = cross_entropy_fwd(logits, labels)
    dlogits = cross_entropy_bwd(logits, labels)
    Ww, _, Wp, _, Ws, _, Wh, _ = c["mw"]
    gm = mlp_bwd(Ww, Wp, Ws, Wh, c["mlp_cache"], dlogits)
    Wv, _, Wq, _, wv, _, wq, _ = c["cw"]
    Qs = [c["word"], c["phrase"], c["sent"]]
    gc = coattn_bwd(feats, Qs, Wv, Wq, wv, wq, c["ca_caches"], gm["g"], gm["g"])
    d_word, d_phrase, d_sent = gc["dQs"]
    w_ih, w_hh = c["lw"][0], c["lw"][1]
    dx_l, dw_ih, dw_hh, db_ih, db_hh = lstm_bwd(c["phrase"], w_ih, w_hh, c["lstm_cache"], d_sent)
    d_phrase = np.where(c["m"], d_phrase + dx_l, 0)        # masked rows are constants
    w1, _, w2, _, w3, _ = c["conv_w"]
    dx_c, dw1, db1, dw2, db2, dw3, db3 = phrase_conv_pool_bwd(c["word"], w1, w2, w3, c["phrase_raw"], c["idx"], d_phrase)
    d_emb = embedding_bwd(tokens, d_word + dx_c, p[PARAM_KEYS[0]].shape[0])
    d = feats.shape[2]
    grads = dict(zip(PARAM_KEYS, (
        d_emb, dw1, db1, dw2, db2, dw3, db3, dw_ih, dw_hh, db_ih, db_hh,
        gc["dWv"], gc["dbv"], gc["dWq"], gc["dbq"],
        gc["dwv"].reshape(1, d), np.asarray(gc["dcv"]).reshape(1), gc["dwq"].reshape(1, d), np.asarray(gc["dcq"]).reshape(1),
        gm["dWw"], gm["dbw"], gm["dWp"], gm["dbp"], gm["dWs"], gm["dbs"], gm["dWh"], gm["dbh"])))
    out = dict(loss=loss, logits=logits, idx=c["idx"], grads=grads, cache=c)
    if need_dfeats:
        out["dfeats"] = gc["dV"]
    return out
