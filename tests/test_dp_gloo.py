"""Host-side logic of the data-parallel path on CPU: world_size-2 gloo processes, flat gradient buffer,
bucketed all-reduce launched from post-accumulate-grad hooks.  (The CUDA ops have no CPU kernels, so a
small stock model stands in: the bucketing / ordering / averaging code is model agnostic.)"""
import importlib
import os
import socket
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


class Toy(torch.nn.Module):
    """Parameter names mimic the real model so BACKWARD_ORDER and the W_b skip are exercised."""

    def __init__(self):
        super().__init__()
        self.question_encoder = torch.nn.ModuleDict(dict(word_embedding=torch.nn.Embedding(11, 8, padding_idx=0)))
        self.co_attention = torch.nn.ModuleDict(dict(W_b=torch.nn.Linear(8, 8), W_v=torch.nn.Linear(8, 8)))
        self.mlp_classify = torch.nn.ModuleDict(dict(W_h=torch.nn.Linear(8, 5)))

    def forward(self, tok):
        x = self.question_encoder["word_embedding"](tok).mean(1)
        return self.mlp_classify["W_h"](torch.tanh(self.co_attention["W_v"](x)))


def _worker(rank, world, port, overlap, out):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    dp = importlib.import_module("visual-question-answering_b200.dp")
    torch.manual_seed(0)
    net = Toy()
    g = torch.Generator().manual_seed(1)
    tok = torch.randint(0, 11, (8, 4), generator=g)
    lab = torch.randint(0, 5, (8,), generator=g)
    # single-process reference on the full batch
    ref = Toy()
    ref.load_state_dict(net.state_dict())
    torch.nn.functional.cross_entropy(ref(tok), lab).backward()
    red = dp.FlatGradAllReduce(net.named_parameters(), bucket_bytes=256, overlap=overlap)
    assert all(not n.startswith("co_attention.W_b") for n in red.names)
    assert red.names[0].startswith("mlp_classify.") and red.names[-1].startswith("question_encoder.word_embedding")
    assert len(red.buckets) >= 2
    sl = dp.shard_batch(8, rank, world)
    for _ in range(2):                                   # second iteration checks zero_grad / re-arming
        red.zero_grad()
        loss = torch.nn.functional.cross_entropy(net(tok[sl]), lab[sl])
        (loss * red.loss_scale).backward()
        red.finish()
    ok = True
    for (n, p), (_, q) in zip(net.named_parameters(), ref.named_parameters()):
        if n.startswith("co_attention.W_b"):
            ok &= p.grad is None
            continue
        ok &= p.grad.data_ptr() >= red.flat.data_ptr()                      # still a view into the flat buffer
        ok &= torch.allclose(p.grad, q.grad, rtol=1e-5, atol=1e-7)
    out[rank] = bool(ok)
    dist.destroy_process_group()


@pytest.mark.parametrize("overlap", [True, False])
def test_flat_grad_allreduce_world2(overlap):
    world = 2
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), overlap, out), nprocs=world, join=True)
    assert dict(out) == {0: True, 1: True}


def test_shard_batch_keeps_sorted_order():
    dp = importlib.import_module("visual-question-answering_b200.dp")
    lens = torch.arange(16, 0, -1)
    for r in range(4):
        sh = lens[dp.shard_batch(16, r, 4)]
        assert (sh[:-1] >= sh[1:]).all() and len(sh) == 4
    with pytest.raises(ValueError):
        dp.shard_batch(10, 0, 4)


def test_gradient_sink_bookkeeping_on_cpu():
    """The sink protocol of FlatGradAllReduce (ops._gbuf / ops._gret call take() / delivered()) exercised by hand on CPU tensors:
    a slot is handed out once per step, a direct write is never followed by a clearing pass, a slot that was neither written nor
    cleared is zeroed by finish() (stale gradients cannot reach the optimizer), and an accumulate into a stale slot raises."""
    dp = importlib.import_module("visual-question-answering_b200.dp")
    torch.manual_seed(0)
    net = torch.nn.Sequential(torch.nn.Linear(5, 3), torch.nn.Linear(3, 2))
    red = dp.FlatGradAllReduce(net.named_parameters(), None, skip=(), flat_params=True, direct_write=False)
    w0, b0, w1, b1 = list(net.parameters())
    # step 1: everything is cleared (no history), two tensors are written directly, two arrive through autograd
    red.zero_grad()
    assert float(red.flat.abs().sum()) == 0.0
    slot = red.take(w0)
    assert slot is not None and slot.data_ptr() == w0.grad.data_ptr() and red.take(w0) is None      # one direct write per step
    slot.fill_(1.0)
    assert red.delivered(slot) and not red.delivered(torch.zeros_like(slot))
    sb = red.take(b1)
    sb.fill_(2.0)
    assert red.delivered(sb)
    net(torch.randn(4, 5)).sum().backward(inputs=[b0, w1])          # autograd accumulates into the cleared views
    red.finish()
    assert torch.all(w0.grad == 1.0) and torch.all(b1.grad == 2.0) and float(w1.grad.abs().sum()) > 0
    assert red._direct_prev == {red._by_param[w0.data_ptr()], red._by_param[b1.data_ptr()]}
    # step 2: only the slots NOT written directly last time are cleared; the direct ones keep stale data until they are rewritten
    red.zero_grad()
    assert float(w1.grad.abs().sum()) == 0.0 and torch.all(w0.grad == 1.0)
    s2 = red.take(w0)
    s2.fill_(3.0)
    red.delivered(s2)
    red.finish()                                                     # b1 was neither rewritten nor cleared: finish() zeroes it
    assert torch.all(w0.grad == 3.0) and float(b1.grad.abs().sum()) == 0.0
    # step 3: accumulating into a slot that still holds last step's direct write is refused
    red.zero_grad()
    with pytest.raises(RuntimeError):
        net(torch.randn(4, 5)).sum().backward(inputs=[w0])
