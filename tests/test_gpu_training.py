"""-m gpu: the pieces around the hot path that the reference's training loop exercises (main.py:155-222, 301-335):

  * the full drop-in wrapper ``HierarchicalCoAttentionNet`` (random-init VGG11-bn trunk from a generated weights file, the
    permuted [B,196,512] feature view) against the reference's own outputs (tests/golden/wrapper_448.npz);
  * the reference's training / validation call sequence replayed on the drop-in module;
  * 5 real steps of ``FlatAdam`` on the real model against ``torch.optim.Adam``;
  * 2-rank data parallelism on two GPUs (skipped with fewer): averaged gradients == single-rank gradients of the concatenated
    batch, replicas identical after optimizer steps -- for the fused NVLink kernel and for the NCCL transport.
"""
import json
import os
import socket
import sys

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from conftest import GOLDEN, ROOT  # noqa: E402


def _h():
    import gpu_harness
    return gpu_harness


def _digest(a):
    a = np.asarray(a, np.float64).reshape(-1)
    stride = max(1, a.size // 64)
    return np.concatenate([[a.sum(), np.sqrt((a * a).sum())], a[::stride][:64]])


def _vgg_weights_file(tmp_path, seed, want_digest):
    """Regenerates the random-init vgg11_bn weights file oracle/make_golden.py used (torch's CPU generator is deterministic for a
    given version); skips if this torch / torchvision build initialises differently."""
    import torchvision
    torch.manual_seed(seed)
    vgg = torchvision.models.vgg11_bn(weights=None)
    sd = vgg.state_dict()
    got = np.asarray([float(sd["features.0.weight"].double().sum()), float(sd["features.25.weight"].double().norm())])
    if not np.allclose(got, want_digest, rtol=1e-9, atol=1e-12):
        pytest.skip("this torchvision build draws different initial VGG weights than the one that generated the fixture")
    path = os.path.join(str(tmp_path), "vgg11_bn_random.pth")
    torch.save(sd, path)
    return path


def _build_wrapper(tmp_path):
    h = _h()
    z = np.load(os.path.join(GOLDEN, "wrapper_448.npz"))
    c = {k[4:]: z[k].item() for k in z.files if k.startswith("cfg.")}
    path = _vgg_weights_file(tmp_path, c["vgg_seed"], z["vgg_digest"])
    syn = h.PKG.synthetic
    # the constructor call of main.py:164 with the dicts of main.py:400-416
    net = h.PKG.HierarchicalCoAttentionNet(dict(vocab_size=c["vocab"], word_emb_dim=c["d"], hidden_dim=c["d"]),
                                           dict(is_trainable=False, weights_path=path), K=c["K"])
    p = syn.make_params(c["d"], c["vocab"], c["K"], c["mlp_dim"], seed=c["seed"])
    missing, unexpected = net.load_state_dict({k: torch.from_numpy(v) for k, v in p.items()}, strict=False)
    assert not unexpected and all(k.startswith("image_encoder.") for k in missing)
    x = syn.make_inputs(c["B"], 196, c["T"], c["d"], c["vocab"], c["K"], seed=c["seed"], min_len=1)
    images = np.random.RandomState(c["seed"]).standard_normal((c["B"], 3, c["img"], c["img"])).astype(np.float32)
    return net, x, images, z, c


def test_full_wrapper_matches_the_reference(tmp_path):
    """model.py:157-187 end to end: state_dict keys / shapes of all 85 tensors, image encoder output layout, logits, loss and the
    gradient of every trainable tensor against the UNMODIFIED reference run in the build container."""
    h = _h()
    net, x, images, z, c = _build_wrapper(tmp_path)
    want_sd = json.loads(str(z["state_dict"]))
    assert {k: list(v.shape) for k, v in net.state_dict().items()} == want_sd and len(want_sd) == 85
    net = net.to("cuda").eval()
    prev = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False              # the trunk is stock cuDNN: keep its convolutions in fp32 for the comparison
    try:
        img = torch.from_numpy(images).cuda()
        feats = net.image_encoder(img)
        assert tuple(feats.shape) == (c["B"], 196, 512) and feats.stride() == (196 * 512, 1, 196) and not feats.requires_grad
        assert h.rel(_digest(feats.cpu().numpy()), z["feats.digest"]) < 1e-4
        before = h.PKG._lib.launch_count()
        logits = net(img, torch.from_numpy(x["tokens"]).cuda(), torch.from_numpy(x["lens"]).cuda())      # CUDA lens, as main.py:207 passes them
        loss = torch.nn.CrossEntropyLoss()(logits, torch.from_numpy(x["labels"]).cuda())
        loss.backward()
        torch.cuda.synchronize()
        assert h.PKG._lib.launch_count() - before > 20        # the hot path ran in this library's kernels
    finally:
        torch.backends.cudnn.allow_tf32 = prev
    assert h.rel(logits.detach().cpu().numpy(), z["logits"]) < 1e-3
    assert abs(float(loss) - float(z["loss"])) < 1e-3 * abs(float(z["loss"]))
    assert (logits.argmax(1).cpu().numpy() == z["logits"].argmax(1)).all()
    for n, prm in net.named_parameters():
        key = f"grad.{n}.digest"
        if key not in z.files:
            assert prm.grad is None, n                       # frozen VGG, dead W_b
            continue
        want = z[key]
        got = _digest(prm.grad.cpu().numpy())
        if n in h.ZERO_BIASES:
            assert np.abs(prm.grad.cpu().numpy()).max() < 1e-6
            continue
        assert abs(got[1] - want[1]) < 1e-3 * want[1], n                                       # gradient norm
        assert np.linalg.norm(got[2:] - want[2:]) <= 2e-3 * max(np.linalg.norm(want[2:]), 1e-30) + 1e-9, n


def test_reference_training_loop_call_sequence(tmp_path):
    """The calls the reference's loop makes, in its order, on the drop-in module: construct (main.py:164), .to(device) (:165),
    state_dict save / load (:168-176, :263), nn.CrossEntropyLoss + Adam(model.parameters(), lr) (:179-180), lens moved to the GPU
    (:207), forward (:211), loss (:214), zero_grad / backward / step (:217-222), then compute_validation_metrics (:301-335):
    eval(), no_grad(), argmax accuracy, F.cross_entropy."""
    h = _h()
    net, x, images, z, c = _build_wrapper(tmp_path)
    device = torch.device("cuda:0")
    model = net
    model.to(device)
    ckpt = os.path.join(str(tmp_path), "model_0.pth")
    torch.save(model.state_dict(), ckpt)
    model.load_state_dict(torch.load(ckpt))
    criterion = torch.nn.CrossEntropyLoss()
    trainable = [q for q in model.parameters() if q.requires_grad]
    optimizer = torch.optim.Adam(model.parameters(), 1e-4)
    image = torch.from_numpy(images).to(device)
    question = torch.from_numpy(x["tokens"]).to(device)
    ques_len = torch.from_numpy(x["lens"]).to(device)
    label = torch.from_numpy(x["labels"]).to(device)
    before = {n: q.detach().clone() for n, q in model.named_parameters()}
    losses = []
    model.train()
    for _ in range(3):
        label_predict = model(image, question, ques_len)
        loss = criterion(label_predict, label)
        optimizer.zero_grad()
        loss.backward()
        optimizer.step()
        losses.append(float(loss))
    assert losses[2] < losses[0]                               # three Adam steps on one batch reduce its loss
    moved = [n for n, q in model.named_parameters() if not torch.equal(q, before[n])]
    assert all(not n.startswith("image_encoder.") and not n.startswith("co_attention.W_b") for n in moved)
    # everything trainable but the dead W_b pair received an update (the two analytically-zero score biases may or may not move)
    assert len(trainable) - 4 <= len(moved) <= len(trainable) - 2
    # validation (main.py:301-335)
    model.eval()
    with torch.no_grad():
        label_logits = model(image, question, ques_len)
        label_predicted = torch.argmax(label_logits, dim=1)
        num_correct = (label == label_predicted).sum().item()
        val_loss = torch.nn.functional.cross_entropy(label_logits, label, reduction="mean")
    assert 0 <= num_correct <= len(label) and torch.isfinite(val_loss) and not label_logits.requires_grad
    # a checkpoint written now loads into a freshly constructed module (resume path, main.py:168-176)
    torch.save(model.state_dict(), ckpt)
    again, *_ = _build_wrapper(tmp_path)
    again.load_state_dict(torch.load(ckpt))
    again.to(device).eval()
    with torch.no_grad():
        assert torch.allclose(again(image, question, ques_len), label_logits, rtol=1e-5, atol=1e-6)


def test_flat_adam_five_real_steps(syn):
    """FlatGradAllReduce(flat_params) + FlatAdam on the real model for 5 steps of 2 alternating batches:
    (a) against torch.optim.Adam stepping a twin parameter set on the SAME gradients -- the flat-buffer bookkeeping (alignment
        padding, gradient sinks, W_b exclusion, device-side step counter) must reproduce it to rounding;
    (b) against the reference's algorithm trained on the CPU (oracle/torch_port.py + torch.optim.Adam, main.py:180,222):
        losses per step and the parameter displacement after 5 steps."""
    h = _h()
    import torch_port as TP
    d, N, T, vocab, K, mlp, B = 512, 196, 26, 2000, 1001, 1024, 16
    p = syn.make_params(d, vocab, K, mlp, seed=0)
    xs = [syn.make_inputs(B, N, T, d, vocab, K, seed=20 + i, dist="D2") for i in range(2)]
    net = h.build_net(p, d, vocab, K, mlp)
    red = h.PKG.dp.FlatGradAllReduce(net.named_parameters(), None, flat_params=True)
    opt = h.PKG.optim.FlatAdam(red, lr=1e-4)
    crit = h.PKG.CrossEntropyLoss()
    twin = {n: q.detach().clone().requires_grad_(True) for n, q in net.named_parameters() if not n.startswith("co_attention.W_b")}
    twin_opt = torch.optim.Adam(list(twin.values()), lr=1e-4)
    cpu_p = TP.make_params(p)
    cpu_opt = torch.optim.Adam([v for v in cpu_p.values() if v.requires_grad], lr=1e-4)
    start = {n: q.detach().cpu().clone() for n, q in net.named_parameters()}
    try:
        for step in range(5):
            x = xs[step % 2]
            dev = {k: torch.from_numpy(v).cuda() for k, v in x.items()}
            opt.zero_grad()
            loss = crit(net(dev["feats"], dev["tokens"], h.PKG.QuestionLens(torch.from_numpy(x["lens"]), "cuda")), dev["labels"])
            loss.backward()
            red.finish()
            for n, q in net.named_parameters():
                if n in twin:
                    twin[n].grad = q.grad.detach().clone()
            opt.step()
            twin_opt.step()
            cpu_loss, _ = TP.train_step(cpu_p, torch.from_numpy(x["feats"]), torch.from_numpy(x["tokens"]), torch.from_numpy(x["lens"]),
                                        torch.from_numpy(x["labels"]), cpu_opt)
            assert abs(float(loss) - float(cpu_loss)) < 2e-4 * abs(float(cpu_loss)), (step, float(loss), float(cpu_loss))
        assert int(opt.step_count) == 5
        worst_twin = worst_cpu = 0.0
        for n, q in net.named_parameters():
            if n.startswith("co_attention.W_b"):
                assert torch.equal(q.detach().cpu(), start[n])                     # never touched (grad None in the reference too)
                continue
            a = q.detach()
            assert torch.allclose(a, twin[n].detach(), rtol=2e-5, atol=2e-8), (n, float((a - twin[n]).abs().max()))
            worst_twin = max(worst_twin, float((a - twin[n]).abs().max()))
            da = (a.cpu() - start[n]).double()
            dc = (cpu_p[n].detach() - start[n]).double()
            if n in h.ZERO_BIASES:
                continue                                                            # analytically zero gradient: Adam normalises pure noise
            err = float((da - dc).norm() / dc.norm())
            worst_cpu = max(worst_cpu, err)
            assert err < 3e-2, (n, err)
        print(f"FlatAdam 5 steps: max |p - torch.optim.Adam(same grads)| {worst_twin:.2e}; worst displacement error vs CPU training {worst_cpu:.2e}")
    finally:
        red.close()


# ----------------------------------------------------------------------------------------------------------- 2-rank data parallel
def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _dp_worker(rank, world, port, fused, out, early=False):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    res = {}
    try:
        import gpu_harness as h
        syn = h.PKG.synthetic
        d, N, T, vocab, K, mlp, GB = 512, 196, 26, 1000, 1001, 1024, 16
        p = syn.make_params(d, vocab, K, mlp, seed=0)
        xs = [syn.make_inputs(GB, N, T, d, vocab, K, seed=31 + i, dist="D2") for i in range(2)]
        to = lambda x: {k: torch.from_numpy(v).to(dev) for k, v in x.items()}
        # single-rank reference on the concatenated batch: plain autograd gradients, then torch.optim.Adam
        ref = h.build_net(p, d, vocab, K, mlp, device=dev)
        ref_opt = torch.optim.Adam([q for n, q in ref.named_parameters() if not n.startswith("co_attention.W_b")], lr=1e-4)
        net = h.build_net(p, d, vocab, K, mlp, device=dev)
        red = h.PKG.dp.FlatGradAllReduce(net.named_parameters(), None, flat_params=True, fused=fused,
                                         early_split="question_encoder." if early else None)
        res["fused"] = bool(red.fused)
        res["multicast"] = bool(red.fused and red._symm.mc)
        crit = h.PKG.CrossEntropyLoss(scale=red.loss_scale)
        sl = h.PKG.dp.shard_batch(GB, rank, world)
        # (1) gradients: finish() without an optimizer attached performs the all-reduce
        xg = to(xs[0])
        ref.zero_grad(set_to_none=True)
        torch.nn.functional.cross_entropy(ref(xg["feats"], xg["tokens"], xg["lens"]), xg["labels"]).backward()
        red.zero_grad()
        crit(net(xg["feats"][sl], xg["tokens"][sl], xg["lens"][sl]), xg["labels"][sl]).backward()
        red.finish()
        torch.cuda.synchronize()
        worst = 0.0
        for (n, q), (_, r) in zip(net.named_parameters(), ref.named_parameters()):
            if n.startswith("co_attention.W_b"):
                assert q.grad is None
                continue
            if n in h.ZERO_BIASES:
                assert float(q.grad.abs().max()) < 1e-6
                continue
            worst = max(worst, float((q.grad - r.grad).norm() / r.grad.norm()))
        res["grad_err"] = worst
        # (2) three optimizer steps: replicas stay identical and follow the single-rank run
        opt = h.PKG.optim.FlatAdam(red, lr=1e-4)
        res["early_slice"] = bool(early and opt.overlap_early_slice(max_ctas=16))
        start = red.flat_p.clone()
        for step in range(3):
            xg = to(xs[step % 2])
            ref_opt.zero_grad(set_to_none=True)
            torch.nn.functional.cross_entropy(ref(xg["feats"], xg["tokens"], xg["lens"]), xg["labels"]).backward()
            ref_opt.step()
            opt.zero_grad()
            crit(net(xg["feats"][sl], xg["tokens"][sl], xg["lens"][sl]), xg["labels"][sl]).backward()
            red.finish()
            opt.step()
        torch.cuda.synchronize()
        mine = red.flat_p.clone()
        both = [torch.empty_like(mine) for _ in range(world)]
        dist.all_gather(both, mine)
        res["replicas_identical"] = bool(torch.equal(both[0], both[1]))
        worst = 0.0
        for (n, q), (_, r) in zip(net.named_parameters(), ref.named_parameters()):
            if n.startswith("co_attention.W_b") or n in h.ZERO_BIASES:
                continue
            off = red.offset_of(n)
            d0 = start[off:off + q.numel()].view_as(q)
            worst = max(worst, float(((q - d0) - (r - d0)).norm() / (r - d0).norm()))
        res["step_err"] = worst
        red.close()
        res["ok"] = True
    except Exception as e:                                   # report instead of hanging the peer in a collective
        import traceback
        res["ok"] = False
        res["error"] = f"{type(e).__name__}: {e}\n{traceback.format_exc()[-1500:]}"
    out[rank] = res
    try:
        dist.destroy_process_group()
    except Exception:
        pass


@pytest.mark.parametrize("fused,early", [(True, False), (True, True), (False, False)])
def test_two_rank_data_parallel_matches_single_rank(fused, early):
    """SURVEY section 4 item 4 on hardware: the averaged gradients of 2 ranks equal the single-rank gradients of the concatenated
    batch; after 3 optimizer steps the replicas are bit-identical and have moved like the single-rank run.  fused=True is the
    hand-written NVLink kernel (symmetric memory, multimem when the switch offers multicast), fused=False the NCCL transport; early=True
    additionally processes the classifier + co-attention slice on a side stream while the encoder's backward is still running."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (gpurun --gpus 2)")
    import torch.multiprocessing as mp
    world = 2
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_dp_worker, args=(world, _free_port(), fused, out, early), nprocs=world, join=True)
    res = dict(out)
    print("2-rank DP", "fused" if fused else "nccl", {r: {k: v for k, v in d.items() if k != "error"} for r, d in res.items()})
    for r in range(world):
        assert res[r]["ok"], res[r].get("error")
        assert res[r]["fused"] == fused and res[r]["early_slice"] == early
        assert res[r]["grad_err"] < 1e-4, res[r]
        assert res[r]["replicas_identical"]
        assert res[r]["step_err"] < 3e-2, res[r]


def test_staging_buffers_are_pinned_and_copy_asynchronously():
    """staging.staged(): library-allocated page-locked (write-combined) host tensors; the bytes survive the H2D copy, int64 included.
    (torch's own ``is_pinned()`` only knows its caching host allocator; the driver recognises the pages as locked, which is what makes
    ``copy_(non_blocking=True)`` a true asynchronous DMA.)"""
    h = _h()
    rng = np.random.RandomState(0)
    a = rng.standard_normal((33, 196, 8)).astype(np.float32)
    i = rng.randint(0, 1000, size=(33, 26)).astype(np.int64)
    for wc in (True, False):
        ta, ti = h.PKG.staging.staged(a, write_combined=wc), h.PKG.staging.staged(i, write_combined=wc)
        assert ta.shape == a.shape and ti.dtype == torch.int64 and torch.equal(ta, torch.from_numpy(a))
        da = torch.empty(a.shape, device="cuda")
        di = torch.empty(i.shape, dtype=torch.int64, device="cuda")
        da.copy_(ta, non_blocking=True)
        di.copy_(ti, non_blocking=True)
        torch.cuda.synchronize()
        assert np.array_equal(da.cpu().numpy(), a) and np.array_equal(di.cpu().numpy(), i)
