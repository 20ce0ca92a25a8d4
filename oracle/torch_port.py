"""CPU baseline port: the reference's hot path restated as plain functional PyTorch, op for op.

TEST / BENCHMARK INFRASTRUCTURE ONLY (see oracle/hiecoattn_oracle.py header).  The reference tree does not
exist on the GPU box, so ``bench.py``'s ``cpu_baseline`` leg and ``bench.py --impl reference`` time THIS
port on the box's host cores (``cpu_baseline.kind = "port"``).  It issues the same ATen calls in the same
order as the reference modules do -- including the work the reference repeats (W_v(V) and W_q(Q) are
each evaluated twice per level, model.py:380-384) -- so its cost is the reference's cost.  Gradients come
from autograd, exactly as in the reference's training loop (main.py:217-222).

``tests/test_torch_port.py`` pins it to the golden fixtures produced by the real reference.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F
from torch.nn.utils.rnn import pack_padded_sequence, pad_packed_sequence


def phrase_conv_pool(p, x, prefix="question_encoder.phrase_conv_pool."):
    """model.py:313-334"""
    B, T, E = x.shape
    xt = x.permute(0, 2, 1)
    uni = torch.tanh(F.conv1d(F.pad(xt, (0, 0)), p[prefix + "conv_unigram.1.weight"], p[prefix + "conv_unigram.1.bias"]))
    bi = torch.tanh(F.conv1d(F.pad(xt, (1, 0)), p[prefix + "conv_bigram.1.weight"], p[prefix + "conv_bigram.1.bias"]))
    tri = torch.tanh(F.conv1d(F.pad(xt, (1, 1)), p[prefix + "conv_trigram.1.weight"], p[prefix + "conv_trigram.1.bias"]))
    cat = torch.cat([uni, bi, tri], dim=1).permute(0, 2, 1).unsqueeze(3).reshape(B, T, E, 3)
    return F.max_pool2d(cat, (1, 3)).squeeze(3)


def question_encoder(p, tokens, lens, prefix="question_encoder."):
    """model.py:271-298"""
    T = tokens.shape[1]
    word = F.embedding(tokens, p[prefix + "word_embedding.weight"], padding_idx=0)
    phrase = phrase_conv_pool(p, word)
    packed = pack_padded_sequence(phrase, lens, batch_first=True)
    flat = [p[prefix + "sentence_lstm." + k] for k in ("weight_ih_l0", "weight_hh_l0", "bias_ih_l0", "bias_hh_l0")]
    H = flat[1].shape[1]
    zeros = word.new_zeros(1, int(packed.batch_sizes[0]), H)
    out, _, _ = torch._VF.lstm(packed.data, packed.batch_sizes, (zeros, zeros), flat, True, 1, 0.0, True, False)
    sent_packed = torch.nn.utils.rnn.PackedSequence(out, packed.batch_sizes, packed.sorted_indices, packed.unsorted_indices)
    phrase = pad_packed_sequence(packed, batch_first=True, total_length=T)[0]
    sent = pad_packed_sequence(sent_packed, batch_first=True, total_length=T)[0]
    return word, phrase, sent


def co_attention(p, x_img, hierarchy, prefix="co_attention."):
    """model.py:356-397 (as written: projections recomputed, W_b unused)"""
    Wv, bv = p[prefix + "W_v.weight"], p[prefix + "W_v.bias"]
    Wq, bq = p[prefix + "W_q.weight"], p[prefix + "W_q.bias"]
    wv, cv = p[prefix + "w_v.weight"], p[prefix + "w_v.bias"]
    wq, cq = p[prefix + "w_q.weight"], p[prefix + "w_q.bias"]
    img, ques = [], []
    for Q in hierarchy:
        V = x_img.permute(0, 2, 1)
        C = torch.tanh(torch.bmm(Q, V))
        V = V.permute(0, 2, 1)
        Hv = torch.tanh(F.linear(V, Wv, bv) + torch.bmm(C.transpose(2, 1), F.linear(Q, Wq, bq)))
        Hq = torch.tanh(F.linear(Q, Wq, bq) + torch.bmm(C, F.linear(V, Wv, bv)))
        av = F.softmax(F.linear(Hv, wv, cv), dim=1)
        aq = F.softmax(F.linear(Hq, wq, cq), dim=1)
        img.append(torch.sum(av * V, dim=1))
        ques.append(torch.sum(aq * Q, dim=1))
    return img, ques


def mlp_classifier(p, img, ques, prefix="mlp_classify."):
    """model.py:414-434"""
    L = lambda n, x: F.linear(x, p[prefix + n + ".weight"], p[prefix + n + ".bias"])
    h_w = torch.tanh(L("W_w", ques[0] + img[0]))
    h_p = torch.tanh(L("W_p", torch.cat([ques[1] + img[1], h_w], dim=1)))
    h_s = torch.tanh(L("W_s", torch.cat([ques[2] + img[2], h_p], dim=1)))
    return L("W_h", h_s)


def forward(p, feats, tokens, lens):
    hier = question_encoder(p, tokens, lens)
    img, ques = co_attention(p, feats, list(hier))
    return mlp_classifier(p, img, ques)


def train_step(p, feats, tokens, lens, labels, optimizer=None):
    """One reference training step (main.py:211-222): forward, mean CE, zero_grad, backward, [Adam step]."""
    logits = forward(p, feats, tokens, lens)
    loss = F.cross_entropy(logits, labels)
    if optimizer is not None:
        optimizer.zero_grad(set_to_none=True)
    else:
        for v in p.values():
            v.grad = None
    loss.backward()
    if optimizer is not None:
        optimizer.step()
    return loss, logits


def make_params(np_params, dtype=torch.float32, requires_grad=True):
    """numpy dict (synthetic.make_params) -> dict of leaf tensors; the dead co_attention.W_b is dropped from
    the trainable set exactly as it ends up gradient-less in the reference."""
    out = {}
    for k, v in np_params.items():
        t = torch.tensor(v, dtype=dtype)
        if requires_grad and not k.startswith("co_attention.W_b"):
            t.requires_grad_(True)
        out[k] = t
    return out
