"""-m gpu parity tests: the CUDA path (through the C ABI / custom ops / modules) against the numpy oracle
and against the golden fixtures generated from the reference.  north_star tolerances: logits and
gradients within 1e-3 relative, phrase max-pool indices bit exact (near-tie protocol of SURVEY H1b),
argmax agreement >= 99.9 %."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from conftest import GOLDEN  # noqa: E402


def _h():
    import gpu_harness
    return gpu_harness


def load_golden(name):
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    return {k: z[k] for k in z.files}


def test_library_loaded_and_cuda_only():
    h = _h()
    assert h.PKG._lib.lib().hca_abi_version() == h.PKG._lib.ABI_VERSION
    with pytest.raises((NotImplementedError, RuntimeError)):
        h.PKG.ops.embedding(torch.zeros(2, 3, dtype=torch.long), torch.zeros(4, 8))     # CPU tensors: no fallback


@pytest.mark.parametrize("name", ["small_a", "small_b", "small_c"])
def test_small_golden_full_step(name):
    """CUDA path vs the reference's own fp64 outputs (tests/golden, from oracle/make_golden.py)."""
    h = _h()
    g = load_golden(name)
    c = {k[4:]: g[k].item() for k in g if k.startswith("cfg.")}
    p = {k[2:]: v for k, v in g.items() if k.startswith("p.")}
    x = {k[2:]: v for k, v in g.items() if k.startswith("x.")}
    net = h.build_net(p, c["d"], c["vocab"], c["K"], c["mlp_dim"])
    ours = h.run_ours(net, x, feats_grad=True)
    ref = dict(logits=g["f64.logits"], loss=g["f64.loss"], dfeats=g["f64.dfeats"],
               grads={k[len("f64.grad."):]: v for k, v in g.items() if k.startswith("f64.grad.")})
    errs = h.compare(ours, ref, tol=1e-3, check_dfeats=True)
    for key in ("word", "phrase", "sent", "vhat", "qhat"):
        assert h.rel(ours[key], g[f"f64.{key}"]) < 1e-4, key
    print(name, {k: f"{v:.1e}" for k, v in errs.items()})


@pytest.mark.parametrize("name", ["small_a", "small_b"])
def test_phrase_conv_pool_module_and_indices(name):
    h = _h()
    g = load_golden(name)
    c = {k[4:]: g[k].item() for k in g if k.startswith("cfg.")}
    p = {k[2:]: v for k, v in g.items() if k.startswith("p.")}
    net = h.build_net(p, c["d"], c["vocab"], c["K"], c["mlp_dim"])
    word = torch.from_numpy(g["f64.word"].astype(np.float32)).cuda()
    with torch.no_grad():
        out, idx = net.question_encoder.phrase_conv_pool(word, None, return_indices=True)
    assert h.rel(out.cpu().numpy(), g["f64.phrase_raw"]) < 1e-5
    assert np.array_equal(idx.cpu().numpy(), g["idx"])          # bit exact


def _near_tie_ok(p, x, idx_ours, orc, atol=4e-6):
    """Index mismatches are tolerated only where the oracle's fp64 top-2 gap (of the tanh outputs) is within a
    few fp32 ulps; returns (#mismatch, #mismatch outside the tie band)."""
    import hiecoattn_oracle as O
    bad = np.argwhere(idx_ours != orc["idx"])
    if len(bad) == 0:
        return 0, 0
    c = orc["cache"]
    w = [np.asarray(v, np.float64) for v in c["conv_w"]]
    cat = np.tanh(O.phrase_conv_preact(np.asarray(c["word"], np.float64), *w))
    B, T, E3 = cat.shape
    grp = np.sort(cat.reshape(B, T, E3 // 3, 3), axis=3)
    gap = grp[..., 2] - grp[..., 1]
    valid = O.valid_mask(x["lens"], T)
    outside = 0
    for b, t, e in bad:
        if not valid[b, t]:
            outside += 1           # masked rows must report idx 0 on both sides
        elif gap[b, t, e] > atol:
            outside += 1
    return len(bad), outside


@pytest.mark.parametrize("B,dist,seed", [(8, "D1", 1), (6, "D2", 5), (1, "D1", 7)])
def test_real_widths_full_step(B, dist, seed, syn):
    """d=512, N=196, T=26, K=1001: logits + every gradient vs the fp64 oracle; pool indices bit exact."""
    h = _h()
    d, N, T, vocab, K, mlp = 512, 196, 26, 10000, 1001, 1024
    p = syn.make_params(d, vocab, K, mlp, seed=0)
    x = syn.make_inputs(B, N, T, d, vocab, K, seed=seed, dist=dist, min_len=1)
    net = h.build_net(p, d, vocab, K, mlp)
    ours = h.run_ours(net, x, feats_grad=True, lens_on="both")
    orc = h.run_oracle(p, x, np.float64, need_dfeats=True)
    errs = h.compare(ours, orc, tol=1e-3, check_dfeats=True)
    print(f"B={B} {dist}", {k.split('.')[-2] + '.' + k.split('.')[-1] if '.' in k else k: f"{v:.1e}" for k, v in errs.items()})
    with torch.no_grad():
        word = torch.from_numpy(np.asarray(orc["cache"]["word"], np.float32)).cuda()
        lens_dev = torch.from_numpy(x["lens"]).cuda()
        _, idx = net.question_encoder.phrase_conv_pool(word, lens_dev, return_indices=True)
    idx = idx.cpu().numpy()
    orc_idx = np.where(h.O.valid_mask(x["lens"], T)[..., None], orc["idx"], 0)
    nbad, outside = _near_tie_ok(p, x, idx, dict(orc, idx=orc_idx))
    print("pool index mismatches:", nbad, "outside tie band:", outside)
    assert outside == 0 and nbad <= 2


@pytest.mark.parametrize("name", ["d512_D1", "d512_D2"])
def test_real_widths_against_reference_golden(name, syn):
    """Same widths, but against digests of the reference's own run (not the oracle)."""
    h = _h()
    g = load_golden(name)
    c = {k[4:]: g[k].item() for k in g if k.startswith("cfg.")}
    p = syn.make_params(c["d"], c["vocab"], c["K"], c["mlp_dim"], seed=0)
    x = syn.make_inputs(c["B"], c["N"], c["T"], c["d"], c["vocab"], c["K"], seed=c["seed"], dist=str(c["dist"]), min_len=c["min_len"])
    net = h.build_net(p, c["d"], c["vocab"], c["K"], c["mlp_dim"])
    ours = h.run_ours(net, x, feats_grad=True)
    assert h.rel(ours["logits"], g["f64.logits"]) < 1e-3
    assert (ours["logits"].argmax(1) == g["f64.logits"].argmax(1)).all()
    assert h.rel(ours["vhat"], g["f64.vhat"]) < 1e-3 and h.rel(ours["qhat"], g["f64.qhat"]) < 1e-3

    def digest(a):
        a = np.asarray(a, np.float64).reshape(-1)
        stride = max(1, a.size // 64)
        return np.concatenate([[np.sqrt((a * a).sum())], a[::stride][:64]])

    for k, v in ours["grads"].items():
        ref = g[f"f64.grad.{k}.digest"][1:]
        if k in h.ZERO_BIASES:
            continue
        assert h.rel(digest(v)[:1], ref[:1]) < 1e-3, k             # gradient norm
        assert np.linalg.norm(digest(v)[1:] - ref[1:]) <= 2e-3 * max(np.linalg.norm(ref[1:]), 1e-30) + 1e-9, k


def test_edge_shapes_large_grid_long_question(syn):
    """N=576 (24x24 grid), T=64, K=3001 (odd K: rows not 16-byte aligned), B=3."""
    h = _h()
    d, N, T, vocab, K, mlp = 512, 576, 64, 500, 3001, 1024
    p = syn.make_params(d, vocab, K, mlp, seed=3)
    x = syn.make_inputs(3, N, T, d, vocab, K, seed=11, dist="D2", min_len=1)
    net = h.build_net(p, d, vocab, K, mlp)
    ours = h.run_ours(net, x, feats_grad=False, lens_on="cuda")
    orc = h.run_oracle(p, x, np.float64)
    h.compare(ours, orc, tol=1e-3)


def test_len_one_and_len_T(syn):
    h = _h()
    d, N, T, vocab, K, mlp = 64, 7, 5, 50, 9, 32
    p = syn.make_params(d, vocab, K, mlp, seed=4)
    x = syn.make_inputs(4, N, T, d, vocab, K, seed=2, min_len=1)
    x["lens"][:] = [T, T, 1, 1]
    x["tokens"][2:, 1:] = 0
    x["tokens"][:2] = np.maximum(x["tokens"][:2], 1)
    net = h.build_net(p, d, vocab, K, mlp)
    ours = h.run_ours(net, x, feats_grad=True)
    orc = h.run_oracle(p, x, np.float64, need_dfeats=True)
    h.compare(ours, orc, tol=1e-3, check_dfeats=True)
    # embedding row 0 (padding_idx) gets exactly zero gradient (model.py:263)
    assert np.abs(ours["grads"]["question_encoder.word_embedding.weight"][0]).max() == 0.0
    # pad rows are exactly zero at all three levels
    m = ~h.O.valid_mask(x["lens"], T)
    for key in ("word", "phrase", "sent"):
        assert np.abs(ours[key][m]).max() == 0.0


def test_permuted_feature_view_and_lens_devices(syn):
    """x_img arrives as a non-contiguous [B,N,d] view (model.py:217); lens on CPU, on CUDA, or both."""
    h = _h()
    d, N, T, vocab, K, mlp = 128, 49, 9, 80, 21, 64
    p = syn.make_params(d, vocab, K, mlp, seed=5)
    x = syn.make_inputs(5, N, T, d, vocab, K, seed=3, min_len=1)
    net = h.build_net(p, d, vocab, K, mlp)
    orc = h.run_oracle(p, x, np.float64, need_dfeats=True)
    for lens_on in ("cpu", "cuda", "both"):
        ours = h.run_ours(net, x, feats_grad=True, lens_on=lens_on, feats_view="permuted")
        h.compare(ours, orc, tol=1e-3, check_dfeats=True)


def test_channel_major_features_at_real_widths(syn):
    """the image encoder's output is a [B, 512, 196] map viewed as [B, 196, 512] (model.py:217, strides (100352, 1, 196)).
    That view goes straight into the operand planes (no dense fp32 copy): logits, every gradient and the feature gradient against the
    fp64 oracle, and the same results as the contiguous layout (to the run-to-run noise of the fp32 atomics in the weighted sums)."""
    h = _h()
    d, N, T, vocab, K, mlp = 512, 196, 26, 10000, 1001, 1024
    p = syn.make_params(d, vocab, K, mlp, seed=0)
    x = syn.make_inputs(8, N, T, d, vocab, K, seed=12, dist="D2", min_len=1)
    net = h.build_net(p, d, vocab, K, mlp)
    orc = h.run_oracle(p, x, np.float64, need_dfeats=True)
    before = h.PKG._lib.launch_count()
    ours = h.run_ours(net, x, feats_grad=True, lens_on="both", feats_view="permuted")
    n_perm = h.PKG._lib.launch_count() - before
    h.compare(ours, orc, tol=1e-3, check_dfeats=True)
    before = h.PKG._lib.launch_count()
    dense = h.run_ours(net, x, feats_grad=True, lens_on="both")
    n_dense = h.PKG._lib.launch_count() - before
    assert n_perm == n_dense                       # one plane-split launch either way: no extra gather pass for the permuted view
    assert h.rel(ours["logits"], dense["logits"]) < 1e-5 and h.rel(ours["vhat"], dense["vhat"]) < 1e-5


def test_misuse_raises_like_the_reference(syn):
    h = _h()
    d, N, T, vocab, K, mlp = 64, 7, 5, 50, 9, 32
    net = h.build_net(syn.make_params(d, vocab, K, mlp), d, vocab, K, mlp)
    tokens = torch.ones(3, T, dtype=torch.long, device="cuda")
    with pytest.raises(RuntimeError):
        net.question_encoder(tokens, torch.tensor([2, 5, 3]))          # unsorted lengths (utils.py:33-45)
    with pytest.raises(RuntimeError):
        net.question_encoder(tokens, torch.tensor([3, 2, 0]))          # zero length


def test_eval_no_grad_forward_matches_training_forward(syn):
    """Validation path (main.py:301-327): eval() + no_grad(), no autograd graph."""
    h = _h()
    d, N, T, vocab, K, mlp = 512, 196, 26, 1000, 1001, 1024
    p = syn.make_params(d, vocab, K, mlp, seed=0)
    x = syn.make_inputs(16, N, T, d, vocab, K, seed=9)
    net = h.build_net(p, d, vocab, K, mlp).eval()
    feats = torch.from_numpy(x["feats"]).cuda()
    tokens = torch.from_numpy(x["tokens"]).cuda()
    lens = torch.from_numpy(x["lens"])
    with torch.no_grad():
        a = net(feats, tokens, lens)
    b = net(feats, tokens, lens)
    assert not a.requires_grad and b.requires_grad
    assert torch.allclose(a, b.detach(), rtol=1e-5, atol=1e-6)      # fp32 atomics: summation order may differ run to run
    import hiecoattn_oracle as O
    ref = O.hiecoattn_forward({k: v.astype(np.float64) for k, v in p.items()}, x["feats"].astype(np.float64), x["tokens"], x["lens"])
    assert h.rel(a.cpu().numpy(), ref) < 1e-3
    assert (a.argmax(1).cpu().numpy() == ref.argmax(1)).all()


def test_module_list_api_matches_reference_signatures(syn):
    """ParallelCoAttention / MLPClassifier called the way the reference wrapper calls them (model.py:182-185)."""
    h = _h()
    d, N, T, vocab, K, mlp = 64, 12, 6, 50, 9, 32
    p = syn.make_params(d, vocab, K, mlp, seed=6)
    x = syn.make_inputs(4, N, T, d, vocab, K, seed=4, min_len=1)
    net = h.build_net(p, d, vocab, K, mlp)
    feats = torch.from_numpy(x["feats"]).cuda()
    hier = net.question_encoder(torch.from_numpy(x["tokens"]).cuda(), torch.from_numpy(x["lens"]))
    img_l, q_l = net.co_attention(feats, list(hier))
    assert isinstance(img_l, list) and len(img_l) == 3 and img_l[0].shape == (4, d)
    logits = net.mlp_classify(img_l, q_l)
    logits.sum().backward()
    ref = h.O.hiecoattn_forward({k: v.astype(np.float64) for k, v in p.items()}, x["feats"].astype(np.float64), x["tokens"], x["lens"])
    assert h.rel(logits.detach().cpu().numpy(), ref) < 1e-3
    assert net.co_attention.W_b.weight.grad is None and net.co_attention.W_v.weight.grad is not None


def test_headline_batch_properties(syn):
    """B=160 (BASELINE config 3): size-independent properties + argmax agreement with the fp32 oracle forward."""
    h = _h()
    d, N, T, vocab, K, mlp = 512, 196, 26, 10000, 1001, 1024
    p = syn.make_params(d, vocab, K, mlp, seed=0)
    x = syn.make_inputs(160, N, T, d, vocab, K, seed=1)
    net = h.build_net(p, d, vocab, K, mlp)
    ours = h.run_ours(net, x, lens_on="both")
    assert np.isfinite(ours["logits"]).all() and all(np.isfinite(g).all() for g in ours["grads"].values())
    # batch independence: the first 8 samples alone give the same logits (no cross-sample op in the path)
    x8 = {k: v[:8] for k, v in x.items()}
    ours8 = h.run_ours(net, x8, lens_on="both")
    assert h.rel(ours8["logits"], ours["logits"][:8]) < 1e-5
    # attention weights are a softmax: attended features are convex combinations of the rows
    f = x["feats"]
    assert (ours["vhat"] <= f.max(1)[None] + 1e-4).all() and (ours["vhat"] >= f.min(1)[None] - 1e-4).all()
    ref = h.O.hiecoattn_forward({k: v.astype(np.float64) for k, v in p.items()}, x["feats"].astype(np.float64), x["tokens"], x["lens"])
    assert h.rel(ours["logits"], ref) < 1e-3
    agree = (ours["logits"].argmax(1) == ref.argmax(1)).mean()
    print("argmax agreement", agree)
    assert agree >= 0.999


def test_opcheck_schemas():
    """torch.library hygiene: schema, fake kernels and autograd registration of every custom op."""
    h = _h()
    ops = h.PKG.ops
    dev = "cuda"
    g = torch.Generator(device="cpu").manual_seed(0)
    r = lambda *s: torch.randn(*s, generator=g).to(dev)
    B, N, T, d, K, m = 2, 5, 4, 32, 7, 16
    tokens = torch.randint(0, 10, (B, T), generator=g).to(dev)
    kinds = ("test_schema", "test_faketensor", "test_autograd_registration")
    torch.library.opcheck(ops.embedding, (tokens, r(10, d).requires_grad_()), test_utils=kinds)
    conv = [r(d, d, 1), r(d), r(d, d, 2), r(d), r(d, d, 3), r(d)]
    torch.library.opcheck(ops.phrase_conv_pool, (r(B, T, d).requires_grad_(), *[c.requires_grad_() for c in conv], None), test_utils=kinds)
    ca = [r(B, N, d), r(B, T, d).requires_grad_(), r(B, T, d), r(B, T, d), r(d, d).requires_grad_(), r(d), r(d, d), r(d), r(1, d), r(1), r(1, d), r(1)]
    torch.library.opcheck(ops.coattn, tuple(ca), test_utils=kinds)
    ml = [r(3, B, d).requires_grad_(), r(3, B, d), r(d, d), r(d), r(d, 2 * d), r(d), r(m, 2 * d), r(m), r(K, m).requires_grad_(), r(K)]
    torch.library.opcheck(ops.mlp, tuple(ml), test_utils=kinds)


def test_fused_cross_entropy_matches_torch():
    """hiecoattn::cross_entropy (loss + gradient in one launch) against F.cross_entropy and its autograd, K = 1001 and odd sizes."""
    h = _h()
    ops = h.PKG.ops
    g = torch.Generator(device="cpu").manual_seed(3)
    for B, K, scale in ((160, 1001, 1.0), (7, 13, 0.5), (1, 3001, 1.0)):
        logits = (torch.randn(B, K, generator=g) * 3).cuda().requires_grad_()
        labels = torch.randint(0, K, (B,), generator=g).cuda()
        ref_in = logits.detach().double().requires_grad_()
        ref = torch.nn.functional.cross_entropy(ref_in, labels) * scale
        ref.backward()
        for _ in range(2):                      # twice: the completion counter re-arms itself
            logits.grad = None
            loss = ops.cross_entropy(logits, labels, scale)
            (loss * 2.0).backward()
            assert abs(float(loss) - float(ref)) < 1e-5 * max(1.0, abs(float(ref)))
            assert h.rel(logits.grad.cpu().numpy(), 2.0 * ref_in.grad.cpu().numpy()) < 1e-5
    crit = h.PKG.CrossEntropyLoss()
    assert abs(float(crit(logits.detach(), labels)) - float(torch.nn.functional.cross_entropy(logits.detach(), labels))) < 1e-5


def test_flat_gradient_sink_direct_write(syn):
    """dp.FlatGradAllReduce as a gradient sink: the backward kernels write into the flat buffer, param.grad stays a view of
    it, values equal the plain autograd path, and a second step does not see the first step's gradients."""
    h = _h()
    d, N, T, vocab, K, mlp = 64, 12, 6, 50, 9, 32
    p = syn.make_params(d, vocab, K, mlp, seed=6)
    xa = syn.make_inputs(4, N, T, d, vocab, K, seed=4, min_len=1)
    xb = syn.make_inputs(4, N, T, d, vocab, K, seed=5, min_len=1)
    plain = h.build_net(p, d, vocab, K, mlp)
    ref_b = h.run_ours(plain, xb)["grads"]
    net = h.build_net(p, d, vocab, K, mlp)
    red = h.PKG.dp.FlatGradAllReduce(net.named_parameters(), None, flat_params=True)
    try:
        for x in (xa, xb, xb):
            red.zero_grad()
            dev = "cuda"
            logits = net(torch.from_numpy(x["feats"]).to(dev), torch.from_numpy(x["tokens"]).to(dev), torch.from_numpy(x["lens"]))
            before = h.PKG._lib.launch_count()
            h.PKG.ops.cross_entropy(logits, torch.from_numpy(x["labels"]).to(dev)).backward()
            red.finish()
        assert len(red._direct_prev) == len(red.params)          # every trained tensor was written in place
        for n, q in net.named_parameters():
            if n.startswith("co_attention.W_b"):
                assert q.grad is None
                continue
            assert red.flat.data_ptr() <= q.grad.data_ptr() < red.flat.data_ptr() + red.flat.numel() * 4
            g = q.grad.cpu().numpy()
            if n in h.ZERO_BIASES:
                assert np.abs(g).max() < 1e-6
            else:
                assert h.rel(g, ref_b[n]) < 1e-4, n
    finally:
        red.close()
        assert red not in h.PKG.ops._SINKS


def test_inference_predict_and_attention_maps(syn):
    """Attention-map export / top-k prediction: a_v, a_q against the oracle's softmaxes, top-1 against its logits."""
    h = _h()
    d, N, T, vocab, K, mlp = 64, 12, 6, 50, 9, 32
    p = syn.make_params(d, vocab, K, mlp, seed=6)
    x = syn.make_inputs(5, N, T, d, vocab, K, seed=8, min_len=1)
    net = h.build_net(p, d, vocab, K, mlp).eval()
    feats = torch.from_numpy(x["feats"]).cuda()
    prob, idx, a_v, a_q = net.predict(feats, torch.from_numpy(x["tokens"]).cuda(), torch.from_numpy(x["lens"]), topk=3, return_attention=True)
    assert prob.shape == (5, 3) and idx.shape == (5, 3) and a_v.shape == (5, 3, N) and a_q.shape == (5, 3, T)
    assert torch.allclose(a_v.sum(-1), torch.ones_like(a_v.sum(-1)), atol=1e-5) and torch.allclose(a_q.sum(-1), torch.ones_like(a_q.sum(-1)), atol=1e-5)
    pp = {k: v.astype(np.float64) for k, v in p.items()}
    ref_logits = h.O.hiecoattn_forward(pp, x["feats"].astype(np.float64), x["tokens"], x["lens"])
    assert (idx[:, 0].cpu().numpy() == ref_logits.argmax(1)).all()
    orc = h.run_oracle(p, x, np.float64)
    for l, c in enumerate(orc["cache"]["ca_caches"]):          # per-level caches of the oracle: av [B,N], aq [B,T]
        assert h.rel(a_v[:, l].cpu().numpy(), np.asarray(c["av"]).reshape(5, N)) < 1e-4
        assert h.rel(a_q[:, l].cpu().numpy(), np.asarray(c["aq"]).reshape(5, T)) < 1e-4


def test_inference_session_replays_the_validation_forward(syn):
    """InferenceSession (CUDA-graph replay of the eval() + no_grad() forward, main.py:301-335): same logits as the eager call and as the
    fp64 oracle, for fresh inputs on every replay, with lengths given on the host, on the device, or as QuestionLens."""
    h = _h()
    d, N, T, vocab, K, mlp, B = 512, 196, 26, 1000, 1001, 1024, 24
    p = syn.make_params(d, vocab, K, mlp, seed=0)
    net = h.build_net(p, d, vocab, K, mlp).eval()
    sess = h.PKG.InferenceSession(net, B, N, T)
    p64 = {k: v.astype(np.float64) for k, v in p.items()}
    for i, how in enumerate(("cpu", "cuda", "both")):
        x = syn.make_inputs(B, N, T, d, vocab, K, seed=40 + i, dist="D2", min_len=1)
        feats, tokens = torch.from_numpy(x["feats"]).cuda(), torch.from_numpy(x["tokens"]).cuda()
        lens = torch.from_numpy(x["lens"])
        lens = lens.cuda() if how == "cuda" else (h.PKG.QuestionLens(lens, "cuda") if how == "both" else lens)
        got = sess(feats, tokens, lens).clone()
        with torch.no_grad():
            eager = net(feats, tokens, lens)
        ref = h.O.hiecoattn_forward(p64, x["feats"].astype(np.float64), x["tokens"], x["lens"])
        assert h.rel(got.cpu().numpy(), eager.cpu().numpy()) < 1e-5
        assert h.rel(got.cpu().numpy(), ref) < 1e-3 and (got.argmax(1).cpu().numpy() == ref.argmax(1)).all()
    with pytest.raises(RuntimeError):
        sess(feats, tokens, torch.tensor([3] + [5] * (B - 1)))            # unsorted lengths are still refused (model.py:287)


@pytest.mark.parametrize("seed", list(range(8)))
def test_random_ragged_shapes_full_step(seed, syn):
    """Randomly drawn small shapes (ragged tiles everywhere: N, T, K, B not multiples of anything; d, mlp multiples of 8 as the
    operand planes require): logits + every gradient + dfeats against the fp64 oracle."""
    h = _h()
    rng = np.random.RandomState(1000 + seed)
    d = int(rng.choice([8, 16, 40, 64, 96, 136]))
    mlp = int(rng.choice([8, 24, 40, 72, 136]))
    N, T, B, K, vocab = int(rng.randint(1, 71)), int(rng.randint(1, 13)), int(rng.randint(1, 10)), int(rng.randint(2, 41)), int(rng.randint(5, 60))
    p = syn.make_params(d, vocab, K, mlp, seed=seed)
    x = syn.make_inputs(B, N, T, d, vocab, K, seed=seed + 50, dist="D1" if seed % 2 else "D2", min_len=1)
    net = h.build_net(p, d, vocab, K, mlp)
    ours = h.run_ours(net, x, feats_grad=True, lens_on="both" if seed % 3 else "cuda")
    orc = h.run_oracle(p, x, np.float64, need_dfeats=True)
    errs = h.compare(ours, orc, tol=1e-3, check_dfeats=True)
    print(dict(d=d, mlp=mlp, N=N, T=T, B=B, K=K), f"worst {max(v for k, v in errs.items() if k not in h.ZERO_BIASES):.1e}")
