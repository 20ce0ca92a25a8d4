"""-m gpu: the persistent tcgen05 LSTM (csrc/lstm.cu) against the numpy oracle's lstm_fwd / lstm_bwd (reference
model.py:269,287-296: pack -> nn.LSTM -> pad, zero-filled pads)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _run(B, T, E, H, seed, sort=True, min_len=1):
    import gpu_harness as h
    import hiecoattn_oracle as O
    ops = h.PKG.ops
    rng = np.random.RandomState(seed)
    lens = rng.randint(min_len, T + 1, size=B).astype(np.int64)
    if sort:
        lens = np.sort(lens)[::-1].copy()
    x = rng.standard_normal((B, T, E)).astype(np.float32)
    for b in range(B):
        x[b, lens[b]:] = 0
    k = 1.0 / np.sqrt(H)
    w_ih = rng.uniform(-k, k, (4 * H, E)).astype(np.float32)
    w_hh = rng.uniform(-k, k, (4 * H, H)).astype(np.float32)
    b_ih = rng.uniform(-k, k, (4 * H,)).astype(np.float32)
    b_hh = rng.uniform(-k, k, (4 * H,)).astype(np.float32)
    dy = rng.standard_normal((B, T, H)).astype(np.float32)
    assert ops.lstm_supported(B, T, E, H)
    tx = torch.from_numpy(x).cuda().requires_grad_(True)
    tw = [torch.from_numpy(a).cuda().requires_grad_(True) for a in (w_ih, w_hh, b_ih, b_hh)]
    out, _ = ops.lstm(tx, torch.from_numpy(lens).cuda(), *tw)
    out.backward(torch.from_numpy(dy).cuda())
    torch.cuda.synchronize()
    f64 = lambda a: a.astype(np.float64)
    y, cache = O.lstm_fwd(f64(x), lens, f64(w_ih), f64(w_hh), f64(b_ih), f64(b_hh))
    dx, dw_ih, dw_hh, db_ih, db_hh = O.lstm_bwd(f64(x), f64(w_ih), f64(w_hh), cache, f64(dy))
    got = dict(out=out, dx=tx.grad, dw_ih=tw[0].grad, dw_hh=tw[1].grad, db_ih=tw[2].grad, db_hh=tw[3].grad)
    ref = dict(out=y, dx=dx, dw_ih=dw_ih, dw_hh=dw_hh, db_ih=db_ih, db_hh=db_hh)
    errs = {k: h.rel(got[k].detach().cpu().numpy(), ref[k]) for k in ref}
    # rows past the end are exactly zero, like pad_packed_sequence's fill
    o = out.detach().cpu().numpy()
    for b in range(B):
        assert not o[b, lens[b]:].any()
    return errs


@pytest.mark.parametrize("B,T,E,H", [(3, 7, 32, 32), (5, 4, 64, 64), (2, 1, 32, 32), (7, 9, 48, 80), (160, 26, 512, 512), (200, 26, 512, 512),
                                      (600, 12, 64, 512)])
def test_lstm_matches_oracle(B, T, E, H):
    errs = _run(B, T, E, H, seed=B * 31 + T)
    assert max(errs.values()) < 1e-3, errs
    assert errs["out"] < 1e-4, errs


def test_lstm_refuses_a_device_that_cannot_hold_a_row_tile(tmp_path):
    """All CTAs of a recurrence launch wait for each other, so they must be co-resident.  With the device pretended to hold only 8 CTAs
    (HCA_LSTM_MAX_CTAS, read once per process: hence the subprocess) H = 512 (32 CTAs per row tile) must be reported as unsupported --
    and the encoder must then take the cuDNN fallback and still match the oracle -- rather than being launched into a spin."""
    import os
    import subprocess
    import sys
    import textwrap
    from conftest import ROOT
    code = textwrap.dedent(f"""
        import sys
        sys.path.insert(0, {ROOT!r}); sys.path.insert(0, {os.path.join(ROOT, 'tests')!r}); sys.path.insert(0, {os.path.join(ROOT, 'oracle')!r})
        import numpy as np, torch
        import gpu_harness as h
        ops = h.PKG.ops
        assert not ops.lstm_supported(160, 26, 512, 512)
        assert ops.lstm_supported(16, 5, 64, 64)            # 4 CTAs per row tile forward, 4 backward: fits 8
        syn = h.PKG.synthetic
        d, N, T, vocab, K, mlp = 512, 20, 7, 60, 9, 64
        p = syn.make_params(d, vocab, K, mlp, seed=2)
        x = syn.make_inputs(5, N, T, d, vocab, K, seed=3, min_len=1)
        net = h.build_net(p, d, vocab, K, mlp)
        before = h.PKG._lib.launch_count()
        ours = h.run_ours(net, x, lens_on="both")
        orc = h.run_oracle(p, x, np.float64)
        h.compare(ours, orc, tol=1e-3)
        print("FALLBACK_OK")
    """)
    env = dict(os.environ, HCA_LSTM_MAX_CTAS="8")
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=600, env=env)
    assert r.returncode == 0 and "FALLBACK_OK" in r.stdout, (r.stdout[-1500:], r.stderr[-3000:])


def test_lstm_unsorted_lengths_and_len_T():
    errs = _run(37, 11, 64, 128, seed=5, sort=False)
    assert max(errs.values()) < 1e-3, errs
    errs = _run(9, 6, 32, 64, seed=6, min_len=6)          # every sequence has the full length
    assert max(errs.values()) < 1e-3, errs


def test_flat_adam_matches_torch_adam():
    """optim.FlatAdam (one fused kernel over flat buffers) against torch.optim.Adam, 4 steps (reference main.py:180,222)."""
    import gpu_harness as h
    pkg = h.PKG
    torch.manual_seed(0)
    net = torch.nn.Sequential(torch.nn.Linear(37, 53), torch.nn.Tanh(), torch.nn.Linear(53, 11)).cuda()
    ref = torch.nn.Sequential(torch.nn.Linear(37, 53), torch.nn.Tanh(), torch.nn.Linear(53, 11)).cuda()
    ref.load_state_dict(net.state_dict())
    red = pkg.dp.FlatGradAllReduce(net.named_parameters(), skip=(), flat_params=True)
    opt = pkg.optim.FlatAdam(red, lr=1e-3)
    ropt = torch.optim.Adam(ref.parameters(), lr=1e-3)
    x = torch.randn(16, 37, device="cuda")
    for _ in range(4):
        red.zero_grad(); ropt.zero_grad()
        net(x).square().mean().backward()
        ref(x).square().mean().backward()
        opt.step(); ropt.step()
    for (n, a), (_, b) in zip(net.named_parameters(), ref.named_parameters()):
        assert torch.allclose(a, b, rtol=2e-5, atol=1e-7), (n, float((a - b).abs().max()))
    assert int(opt.step_count) == 4
