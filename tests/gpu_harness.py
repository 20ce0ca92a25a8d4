"""Shared helpers for the -m gpu parity tests: run the CUDA path and the numpy oracle on the same inputs."""
import importlib

import numpy as np
import torch

import hiecoattn_oracle as O

PKG = importlib.import_module("visual-question-answering_b200")
ZERO_BIASES = ("co_attention.w_v.bias", "co_attention.w_q.bias")     # analytically zero gradients


def rel(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300))


def build_net(p, d, vocab, K, mlp_dim, device="cuda"):
    net = PKG.HieCoAttnHotPath(vocab, d, K, mlp_dim)
    sd = {k: torch.from_numpy(np.ascontiguousarray(v, dtype=np.float32)) for k, v in p.items()}
    missing, unexpected = net.load_state_dict(sd, strict=False)
    assert not missing and not unexpected, (missing, unexpected)
    return net.to(device)


def run_ours(net, x, feats_grad=False, lens_on="cpu", feats_view=None):
    """One forward + mean-CE + backward through the CUDA path.  Returns numpy results keyed like the oracle's."""
    dev = next(net.parameters()).device
    feats = torch.from_numpy(np.asarray(x["feats"], np.float32)).to(dev)
    if feats_view == "permuted":            # the layout the reference's VGG encoder hands over (model.py:217)
        feats = feats.permute(0, 2, 1).contiguous().permute(0, 2, 1)
        assert not feats.is_contiguous()
    feats.requires_grad_(feats_grad)
    tokens = torch.from_numpy(x["tokens"]).to(dev)
    labels = torch.from_numpy(x["labels"]).to(dev)
    lens = torch.from_numpy(x["lens"])
    if lens_on == "cuda":
        lens = lens.to(dev)
    elif lens_on == "both":
        lens = PKG.QuestionLens(lens, dev)
    net.zero_grad(set_to_none=True)
    hier = net.question_encoder(tokens, lens)
    for t in hier:
        t.retain_grad()
    vhat, qhat = net.co_attention.forward_stacked(feats, hier)
    logits = net.mlp_classify.forward_stacked(vhat, qhat)
    loss = torch.nn.functional.cross_entropy(logits, labels)
    loss.backward()
    torch.cuda.synchronize()
    out = dict(logits=logits.detach().cpu().numpy(), loss=float(loss), word=hier[0].detach().cpu().numpy(),
               phrase=hier[1].detach().cpu().numpy(), sent=hier[2].detach().cpu().numpy(),
               vhat=vhat.detach().cpu().numpy(), qhat=qhat.detach().cpu().numpy())
    out["grads"] = {k: v.grad.detach().cpu().numpy() for k, v in net.named_parameters() if v.grad is not None}
    out["none_grads"] = [k for k, v in net.named_parameters() if v.grad is None]
    if feats_grad:
        out["dfeats"] = feats.grad.detach().cpu().numpy()
    return out


def run_oracle(p, x, dtype=np.float64, need_dfeats=False):
    pp = {k: np.asarray(v, dtype) for k, v in p.items()}
    return O.hiecoattn_step(pp, np.asarray(x["feats"], dtype), x["tokens"], x["lens"], x["labels"], need_dfeats=need_dfeats)


def compare(ours, orc, tol=1e-3, check_dfeats=False):
    """north_star tolerances: logits and every gradient within `tol` relative (normwise), the two
    analytically-zero score biases within 1e-6 absolute.  Returns the table of errors for reporting."""
    errs = {"logits": rel(ours["logits"], orc["logits"])}
    assert errs["logits"] < tol, errs
    assert abs(ours["loss"] - float(orc["loss"])) < tol * max(1.0, abs(float(orc["loss"])))
    for k in O.PARAM_KEYS:
        g, r = ours["grads"][k], orc["grads"][k]
        assert g.shape == r.shape, (k, g.shape, r.shape)
        if k in ZERO_BIASES:
            errs[k] = float(np.abs(g).max())
            assert errs[k] < 1e-6, (k, errs[k])
        else:
            errs[k] = rel(g, r)
            assert errs[k] < tol, (k, errs[k])
    if check_dfeats:
        errs["dfeats"] = rel(ours["dfeats"], orc["dfeats"])
        assert errs["dfeats"] < tol, errs["dfeats"]
    assert sorted(ours["none_grads"]) == ["co_attention.W_b.bias", "co_attention.W_b.weight"], ours["none_grads"]
    return errs
