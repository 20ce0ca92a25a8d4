"""torch.library custom ops (namespace ``hiecoattn``) over the C ABI of libhiecoattn_b200.so.

Each op is functional (no hidden mutation), has a fake kernel for shape inference, an autograd formula
that calls the matching ``*_bwd`` op, and works under ``no_grad``.  All ops are CUDA-only: there is no
CPU kernel behind them (a CPU tensor raises).  Tensors are fp32 at this boundary; callers that run
under autocast hand in fp16/bf16 and are upcast by the module wrappers (modules.py).
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch
from torch import Tensor

from . import _lib

NS = "hiecoattn"


def _ptr(t: Optional[Tensor]):
    return None if t is None else t.data_ptr()


def _stream():
    return torch.cuda.current_stream().cuda_stream


def _cuda_f32(*ts):
    for t in ts:
        if t is None:
            continue
        if not t.is_cuda:
            raise RuntimeError("hiecoattn_b200 ops run on CUDA only (no CPU fallback); got a CPU tensor")
        if t.dtype != torch.float32:
            raise RuntimeError(f"hiecoattn_b200 ops take float32 tensors; got {t.dtype}")


def _c(t: Tensor) -> Tensor:
    return t if t.is_contiguous() else t.contiguous()


def _cuda_i64(name: str, t: Optional[Tensor], like: Tensor) -> None:
    """Index tensors (tokens, lens, labels) reach raw-pointer kernels that read them as int64 on the device of ``like``."""
    if t is None:
        return
    if not t.is_cuda or t.device != like.device:
        raise RuntimeError(f"hiecoattn_b200: `{name}` must be a CUDA tensor on {like.device} (got {t.device}); there is no CPU fallback")
    if t.dtype != torch.int64:
        raise RuntimeError(f"hiecoattn_b200: `{name}` must be int64 (got {t.dtype})")


def _ws(nbytes: int, device) -> Tensor:
    return torch.empty(max(int(nbytes), 256), dtype=torch.uint8, device=device)


# ------------------------------------------------------------------------------------------------ gradient sinks
# A training loop that keeps every parameter gradient in one flat buffer (dp.FlatGradAllReduce) registers itself here.  The
# autograd formulas below then let the backward kernels write each weight gradient straight into its slot of that buffer
# (the ``*_bwd`` ops take their weight-gradient destinations as mutable arguments) and return None for those inputs: the
# gradient is already where ``param.grad`` points, so autograd has nothing to add (it used to cost one elementwise kernel
# per parameter and step, ~30 launches) and the buffer needs no clearing pass.
_SINKS: list = []


def add_grad_sink(sink) -> None:
    if sink not in _SINKS:
        _SINKS.append(sink)


def remove_grad_sink(sink) -> None:
    if sink in _SINKS:
        _SINKS.remove(sink)


def _gbuf(w: Tensor) -> Tensor:
    """Destination of the gradient of parameter tensor ``w``: its slot in a registered sink, else a fresh tensor."""
    for sink in _SINKS:
        t = sink.take(w)
        if t is not None:
            return t
    return torch.empty_like(w)


def _gret(g: Tensor):
    """What an autograd formula hands back for a weight gradient the kernel has written into ``g``."""
    for sink in _SINKS:
        if sink.delivered(g):
            return None
    return g


# ------------------------------------------------------------------------------------------------ embedding
@torch.library.custom_op(f"{NS}::embedding", mutates_args=(), device_types="cuda")
def embedding(tokens: Tensor, table: Tensor) -> Tensor:
    """nn.Embedding(padding_idx=0) forward (reference model.py:263,282).  Token ids must lie in [0, vocab): the reference raises
    on an out-of-range id; this kernel cannot raise without a host synchronisation and clamps the id into the table instead."""
    _cuda_f32(table)
    _cuda_i64("tokens", tokens, table)
    tokens, table = _c(tokens), _c(table)
    out = torch.empty(*tokens.shape, table.shape[1], dtype=torch.float32, device=table.device)
    with torch.cuda.device(table.device):
        _lib.check(_lib.lib().hca_embedding_fwd(_ptr(tokens), _ptr(table), _ptr(out), tokens.numel(), table.shape[1],
                                                table.shape[0], _stream()), "embedding_fwd")
    return out


@embedding.register_fake
def _(tokens, table):
    return table.new_empty(*tokens.shape, table.shape[1])


@torch.library.custom_op(f"{NS}::embedding_bwd", mutates_args=("dtable",), device_types="cuda")
def embedding_bwd(tokens: Tensor, dout: Tensor, dtable: Tensor) -> None:
    """dtable [vocab, E] (overwritten) = scatter-add of dout rows by token."""
    _cuda_f32(dout, dtable)
    _cuda_i64("tokens", tokens, dout)
    tokens, dout = _c(tokens), _c(dout)
    vocab, E = dtable.shape
    assert dtable.is_contiguous()
    with torch.cuda.device(dout.device):
        _lib.check(_lib.lib().hca_embedding_bwd(_ptr(tokens), _ptr(dout), _ptr(dtable), tokens.numel(), E, vocab, _stream()),
                   "embedding_bwd")


def _embedding_setup(ctx, inputs, output):
    tokens, table = inputs
    ctx.save_for_backward(tokens, table)


def _embedding_backward(ctx, dout):
    tokens, table = ctx.saved_tensors
    dtable = _gbuf(table)
    embedding_bwd(tokens, dout, dtable)
    return None, _gret(dtable)


embedding.register_autograd(_embedding_backward, setup_context=_embedding_setup)


# ------------------------------------------------------------------------------------------ phrase conv + pool
def _pcp_saved_bytes(B: int, T: int, E: int) -> int:
    al = lambda n: (n + 255) // 256 * 256
    return al(2 * B * T * 3 * E * 2) + al(2 * 3 * E * 3 * E * 2) + 256      # mirrors hca_phrase_conv_pool_saved_bytes (fake-tensor path)


@torch.library.custom_op(f"{NS}::phrase_conv_pool", mutates_args=(), device_types="cuda")
def phrase_conv_pool(x: Tensor, w1: Tensor, b1: Tensor, w2: Tensor, b2: Tensor, w3: Tensor, b3: Tensor,
                     lens: Optional[Tensor]) -> Tuple[Tensor, Tensor, Tensor]:
    """PhraseConvPool forward (reference model.py:313-334) -> (out [B,T,E], idx uint8 [B,T,E], saved).

    ``lens`` (int64, CUDA, optional) additionally zeroes rows t >= len (model.py:287-292).  ``saved`` is an opaque buffer
    (operand planes of the shifted input and of the weights) that the backward op reuses."""
    _cuda_f32(x, w1, b1, w2, b2, w3, b3)
    _cuda_i64("lens", lens, x)
    x, w1, b1, w2, b2, w3, b3 = map(_c, (x, w1, b1, w2, b2, w3, b3))
    B, T, E = x.shape
    out = torch.empty_like(x)
    idx = torch.empty(B, T, E, dtype=torch.uint8, device=x.device)
    L = _lib.lib()
    with torch.cuda.device(x.device):
        nb = L.hca_phrase_conv_pool_workspace(B, T, E)
        ws = _ws(nb, x.device)
        # the operand planes are produced by the forward anyway: leaving them in a returned buffer instead of the workspace costs
        # nothing and saves the backward three weight conversions and the im2col pass
        saved = _ws(_pcp_saved_bytes(B, T, E), x.device)
        _lib.check(L.hca_phrase_conv_pool_fwd(_ptr(x), _ptr(w1), _ptr(b1), _ptr(w2), _ptr(b2), _ptr(w3), _ptr(b3),
                                              _ptr(None if lens is None else _c(lens)), _ptr(out), _ptr(idx),
                                              _ptr(saved), saved.numel(), B, T, E, _ptr(ws), ws.numel(), _stream()), "phrase_conv_pool_fwd")
    return out, idx, saved


def phrase_conv_pool_tie_stats(saved: Tensor) -> Tuple[int, int]:
    """(near-ties found, capacity of the tie list) of the forward call that produced ``saved`` -- the bookkeeping of the exact
    argmax repair (csrc/phrase_conv_pool.cu).  found > capacity means the repair ran in its exhaustive mode.  Synchronises."""
    st = saved[-256:-248].view(torch.int32).cpu()
    return int(st[0]), int(st[1])


@phrase_conv_pool.register_fake
def _(x, w1, b1, w2, b2, w3, b3, lens):
    B, T, E = x.shape
    return (torch.empty_like(x), x.new_empty(x.shape, dtype=torch.uint8), x.new_empty(_pcp_saved_bytes(B, T, E), dtype=torch.uint8))


@torch.library.custom_op(f"{NS}::phrase_conv_pool_bwd", mutates_args=("dw1", "db1", "dw2", "db2", "dw3", "db3"), device_types="cuda")
def phrase_conv_pool_bwd(x: Tensor, w1: Tensor, w2: Tensor, w3: Tensor, out: Tensor, idx: Tensor, dout: Tensor,
                         lens: Optional[Tensor], saved: Tensor, need_dx: bool, dw1: Tensor, db1: Tensor, dw2: Tensor, db2: Tensor,
                         dw3: Tensor, db3: Tensor) -> Tensor:
    """Returns dx; the weight / bias gradients are WRITTEN into dw* / db* (contiguous, 16-byte aligned)."""
    _cuda_f32(x, w1, w2, w3, out, dout, dw1, db1, dw2, db2, dw3, db3)
    _cuda_i64("lens", lens, x)
    x, w1, w2, w3, out, idx, dout = map(_c, (x, w1, w2, w3, out, idx, dout))
    assert all(t.is_contiguous() for t in (dw1, db1, dw2, db2, dw3, db3))
    B, T, E = x.shape
    dx = torch.empty_like(x) if need_dx else x.new_empty(0)
    L = _lib.lib()
    with torch.cuda.device(x.device):
        ws = _ws(L.hca_phrase_conv_pool_workspace(B, T, E), x.device)
        have = saved.numel() >= L.hca_phrase_conv_pool_saved_bytes(B, T, E)
        _lib.check(L.hca_phrase_conv_pool_bwd(_ptr(x), _ptr(w1), _ptr(w2), _ptr(w3), _ptr(out), _ptr(idx), _ptr(dout),
                                              _ptr(None if lens is None else _c(lens)), _ptr(saved) if have else None,
                                              saved.numel() if have else 0, _ptr(dx) if need_dx else None,
                                              _ptr(dw1), _ptr(db1), _ptr(dw2), _ptr(db2), _ptr(dw3), _ptr(db3), B, T, E,
                                              _ptr(ws), ws.numel(), _stream()), "phrase_conv_pool_bwd")
    return dx


@phrase_conv_pool_bwd.register_fake
def _(x, w1, w2, w3, out, idx, dout, lens, saved, need_dx, dw1, db1, dw2, db2, dw3, db3):
    return torch.empty_like(x) if need_dx else x.new_empty(0)


def _pcp_setup(ctx, inputs, output):
    x, w1, b1, w2, b2, w3, b3, lens = inputs
    out, idx, saved = output
    ctx.save_for_backward(x, w1, b1, w2, b2, w3, b3, out, idx, lens, saved)
    ctx.set_materialize_grads(False)


def _pcp_backward(ctx, dout, _didx, _dsaved):
    x, w1, b1, w2, b2, w3, b3, out, idx, lens, saved = ctx.saved_tensors
    if dout is None:
        return (None,) * 8
    need_dx = ctx.needs_input_grad[0]
    gs = [_gbuf(t) for t in (w1, b1, w2, b2, w3, b3)]
    dx = phrase_conv_pool_bwd(x, w1, w2, w3, out, idx, dout, lens, saved, need_dx, *gs)
    return ((dx if need_dx else None), *(_gret(g) for g in gs), None)


phrase_conv_pool.register_autograd(_pcp_backward, setup_context=_pcp_setup)


# ------------------------------------------------------------------------------------------------ sentence LSTM
def lstm_supported(B: int, T: int, E: int, H: int) -> bool:
    return bool(_lib.lib().hca_lstm_supported(B, T, E, H))


@torch.library.custom_op(f"{NS}::lstm", mutates_args=(), device_types="cuda")
def lstm(x: Tensor, lens: Tensor, w_ih: Tensor, w_hh: Tensor, b_ih: Tensor, b_hh: Tensor) -> Tuple[Tensor, Tensor]:
    """pack_padded_sequence -> nn.LSTM -> pad_packed_sequence (reference model.py:287-296) on the padded batch.

    x [B,T,E], lens int64 [B] on the GPU.  Returns (out [B,T,H] with rows t >= len zeroed, saved) where ``saved`` is
    the opaque buffer the backward kernels read (layout private to the library)."""
    _cuda_f32(x, w_ih, w_hh, b_ih, b_hh)
    _cuda_i64("lens", lens, x)
    x, lens, w_ih, w_hh, b_ih, b_hh = map(_c, (x, lens, w_ih, w_hh, b_ih, b_hh))
    B, T, E = x.shape
    H = w_hh.shape[1]
    out = torch.empty(B, T, H, dtype=torch.float32, device=x.device)
    L = _lib.lib()
    with torch.cuda.device(x.device):
        saved = _ws(L.hca_lstm_saved_bytes(B, T, E, H), x.device)
        ws = _ws(L.hca_lstm_workspace(B, T, E, H), x.device)
        _lib.check(L.hca_lstm_fwd(_ptr(x), _ptr(lens), _ptr(w_ih), _ptr(w_hh), _ptr(b_ih), _ptr(b_hh), _ptr(out), _ptr(saved),
                                  saved.numel(), B, T, E, H, _ptr(ws), ws.numel(), _stream()), "lstm_fwd")
    return out, saved


def _lstm_saved_bytes(B: int, T: int, E: int, H: int) -> int:
    """Mirror of ``hca_lstm_saved_bytes`` for the fake-tensor path (tests/test_modules_cpu.py checks the mirrors against the library)."""
    al = lambda n: (n + 255) // 256 * 256
    BT = B * T
    return al(BT * 4 * H * 4) + al(BT * H * 4) + al(2 * BT * (E + H) * 2) + al(2 * 4 * H * E * 2) + 256


@lstm.register_fake
def _(x, lens, w_ih, w_hh, b_ih, b_hh):
    B, T, E = x.shape
    H = w_hh.shape[1]
    return x.new_empty(B, T, H), x.new_empty(_lstm_saved_bytes(B, T, E, H), dtype=torch.uint8)


@torch.library.custom_op(f"{NS}::lstm_bwd", mutates_args=("dw_ih", "dw_hh", "db_ih", "db_hh"), device_types="cuda")
def lstm_bwd(lens: Tensor, w_ih: Tensor, w_hh: Tensor, saved: Tensor, dout: Tensor, E: int, need_dx: bool, dw_ih: Tensor,
             dw_hh: Tensor, db_ih: Tensor, db_hh: Tensor) -> Tensor:
    """Returns dx; the weight / bias gradients are WRITTEN into dw_* / db_*."""
    _cuda_f32(w_ih, w_hh, dout, dw_ih, dw_hh, db_ih, db_hh)
    _cuda_i64("lens", lens, dout)
    lens, w_ih, w_hh, dout = map(_c, (lens, w_ih, w_hh, dout))
    assert all(t.is_contiguous() for t in (dw_ih, dw_hh, db_ih, db_hh))
    B, T, H = dout.shape
    dev = dout.device
    f = lambda *s: torch.empty(*s, dtype=torch.float32, device=dev)
    dx = f(B, T, E) if need_dx else f(0)
    L = _lib.lib()
    with torch.cuda.device(dev):
        ws = _ws(L.hca_lstm_workspace(B, T, E, H), dev)
        _lib.check(L.hca_lstm_bwd(_ptr(lens), _ptr(w_ih), _ptr(w_hh), _ptr(saved), saved.numel(), _ptr(dout),
                                  _ptr(dx) if need_dx else None, _ptr(dw_ih), _ptr(dw_hh), _ptr(db_ih), _ptr(db_hh), B, T, E, H,
                                  _ptr(ws), ws.numel(), _stream()), "lstm_bwd")
    return dx


@lstm_bwd.register_fake
def _(lens, w_ih, w_hh, saved, dout, E, need_dx, dw_ih, dw_hh, db_ih, db_hh):
    B, T, H = dout.shape
    return dout.new_empty(B, T, E) if need_dx else dout.new_empty(0)


def _lstm_setup(ctx, inputs, output):
    x, lens, w_ih, w_hh, b_ih, b_hh = inputs
    out, saved = output
    ctx.save_for_backward(lens, w_ih, w_hh, b_ih, b_hh, saved)
    ctx.E = x.shape[2]
    ctx.set_materialize_grads(False)


def _lstm_backward(ctx, dout, _dsaved):
    lens, w_ih, w_hh, b_ih, b_hh, saved = ctx.saved_tensors
    if dout is None:
        return (None,) * 6
    need_dx = ctx.needs_input_grad[0]
    gs = [_gbuf(t) for t in (w_ih, w_hh, b_ih, b_hh)]
    dx = lstm_bwd(lens, w_ih, w_hh, saved, dout, ctx.E, need_dx, *gs)
    return ((dx if need_dx else None), None, *(_gret(g) for g in gs))


lstm.register_autograd(_lstm_backward, setup_context=_lstm_setup)


# ------------------------------------------------------------------------------------------------ co-attention
@torch.library.custom_op(f"{NS}::coattn", mutates_args=(), device_types="cuda")
def coattn(V: Tensor, q0: Tensor, q1: Tensor, q2: Tensor, Wv: Tensor, bv: Tensor, Wq: Tensor, bq: Tensor, wv: Tensor,
           cv: Tensor, wq: Tensor, cq: Tensor) -> Tuple[Tensor, Tensor, Tensor]:
    """ParallelCoAttention over the three question levels (reference model.py:356-397).

    Returns (vhat [3,B,d], qhat [3,B,d], saved) where ``saved`` is the opaque buffer the backward kernel reads
    (bf16 operand planes of V / Q / PV / PQ / C and the attention weights; layout private to the library).
    V may be any strided [B,N,d] view (the reference passes a permuted VGG feature map)."""
    _cuda_f32(V, q0, q1, q2, Wv, bv, Wq, bq, wv, cv, wq, cq)
    q0, q1, q2, Wv, bv, Wq, bq, wv, cv, wq, cq = map(_c, (q0, q1, q2, Wv, bv, Wq, bq, wv, cv, wq, cq))
    B, N, d = V.shape
    T = q0.shape[1]
    dev = V.device
    f = lambda *s: torch.empty(*s, dtype=torch.float32, device=dev)
    vhat, qhat = f(3, B, d), f(3, B, d)
    L = _lib.lib()
    with torch.cuda.device(dev):
        saved = _ws(L.hca_coattn_saved_bytes(B, N, T, d), dev)
        ws = _ws(L.hca_coattn_workspace(B, N, T, d, 0), dev)
        _lib.check(L.hca_coattn_fwd(_ptr(V), V.stride(0), V.stride(1), V.stride(2), _ptr(q0), _ptr(q1), _ptr(q2), _ptr(Wv), _ptr(bv),
                                    _ptr(Wq), _ptr(bq), _ptr(wv), _ptr(cv), _ptr(wq), _ptr(cq), _ptr(vhat), _ptr(qhat), _ptr(saved),
                                    saved.numel(), B, N, T, d, _ptr(ws), ws.numel(), _stream()), "coattn_fwd")
    return vhat, qhat, saved


def _coattn_saved_bytes(B: int, N: int, T: int, d: int) -> int:
    """Mirror of ``hca_coattn_saved_bytes`` for the fake-tensor path."""
    r8 = lambda x: (x + 7) // 8 * 8
    al = lambda x: (x + 255) // 256 * 256
    pl = lambda rows, cols: al(2 * rows * r8(cols) * 2)
    return 2 * pl(B * N, d) + 3 * pl(B * 3 * T, d) + pl(B * 3 * T, N) + al(B * 3 * N * 4) + al(B * 3 * T * 4) + 256


@coattn.register_fake
def _(V, q0, q1, q2, Wv, bv, Wq, bq, wv, cv, wq, cq):
    B, N, d = V.shape
    T = q0.shape[1]
    return V.new_empty(3, B, d), V.new_empty(3, B, d), V.new_empty(_coattn_saved_bytes(B, N, T, d), dtype=torch.uint8)


@torch.library.custom_op(f"{NS}::coattn_bwd", mutates_args=("dWv", "dbv", "dWq", "dbq", "dwv", "dcv", "dwq", "dcq"),
                         device_types="cuda")
def coattn_bwd(Wv: Tensor, Wq: Tensor, wv: Tensor, wq: Tensor, saved: Tensor, gv: Tensor, gq: Tensor, N: int, T: int, need_dv: bool,
               dWv: Tensor, dbv: Tensor, dWq: Tensor, dbq: Tensor, dwv: Tensor, dcv: Tensor, dwq: Tensor, dcq: Tensor
               ) -> Tuple[Tensor, Tensor]:
    """Returns (dV, dQ [3,B,T,d]); the parameter gradients are WRITTEN into dWv .. dcq."""
    _cuda_f32(Wv, Wq, wv, wq, gv, gq, dWv, dbv, dWq, dbq, dwv, dcv, dwq, dcq)
    Wv, Wq, wv, wq, gv, gq = map(_c, (Wv, Wq, wv, wq, gv, gq))
    assert all(t.is_contiguous() for t in (dWv, dbv, dWq, dbq, dwv, dcv, dwq, dcq))
    _, B, d = gv.shape
    dev = gv.device
    f = lambda *s: torch.empty(*s, dtype=torch.float32, device=dev)
    dV = f(B, N, d) if need_dv else f(0)
    dQ = f(3, B, T, d)
    L = _lib.lib()
    with torch.cuda.device(dev):
        ws = _ws(L.hca_coattn_workspace(B, N, T, d, int(need_dv)), dev)
        _lib.check(L.hca_coattn_bwd(_ptr(Wv), _ptr(Wq), _ptr(wv), _ptr(wq), _ptr(saved), saved.numel(), _ptr(gv), _ptr(gq),
                                    _ptr(dV) if need_dv else None, _ptr(dQ), _ptr(dWv), _ptr(dbv), _ptr(dWq), _ptr(dbq), _ptr(dwv),
                                    _ptr(dcv), _ptr(dwq), _ptr(dcq), B, N, T, d, _ptr(ws), ws.numel(), _stream()), "coattn_bwd")
    return dV, dQ


@coattn_bwd.register_fake
def _(Wv, Wq, wv, wq, saved, gv, gq, N, T, need_dv, dWv, dbv, dWq, dbq, dwv, dcv, dwq, dcq):
    _, B, d = gv.shape
    f = lambda *s: gv.new_empty(*s)
    return (f(B, N, d) if need_dv else f(0)), f(3, B, T, d)


def _coattn_setup(ctx, inputs, output):
    V, q0, q1, q2, Wv, bv, Wq, bq, wv, cv, wq, cq = inputs
    vhat, qhat, saved = output
    ctx.save_for_backward(Wv, bv, Wq, bq, wv, cv, wq, cq, saved)
    ctx.NT = (V.shape[1], q0.shape[1])
    ctx.set_materialize_grads(False)


def _coattn_backward(ctx, gv, gq, *_unused):
    Wv, bv, Wq, bq, wv, cv, wq, cq, saved = ctx.saved_tensors
    if gv is None and gq is None:
        return (None,) * 12
    if gv is None:
        gv = torch.zeros_like(gq)
    if gq is None:
        gq = torch.zeros_like(gv)
    need_dv = ctx.needs_input_grad[0]
    N, T = ctx.NT
    gs = [_gbuf(t) for t in (Wv, bv, Wq, bq, wv, cv, wq, cq)]
    dV, dQ = coattn_bwd(Wv, Wq, wv, wq, saved, gv, gq, N, T, need_dv, *gs)
    return ((dV if need_dv else None), dQ[0], dQ[1], dQ[2], *(_gret(g) for g in gs))


coattn.register_autograd(_coattn_backward, setup_context=_coattn_setup)


def coattn_attention_maps(saved: Tensor, B: int, N: int, T: int, d: int) -> Tuple[Tensor, Tensor]:
    """The attention weights a forward call of ``coattn`` left in its ``saved`` buffer: (a_v [B,3,N], a_q [B,3,T]) as views."""
    import ctypes as C
    av, aq = C.c_size_t(0), C.c_size_t(0)
    _lib.check(_lib.lib().hca_coattn_saved_attention(B, N, T, d, C.byref(av), C.byref(aq)), "coattn_saved_attention")
    a_v = saved[av.value:av.value + B * 3 * N * 4].view(torch.float32).view(B, 3, N)
    a_q = saved[aq.value:aq.value + B * 3 * T * 4].view(torch.float32).view(B, 3, T)
    return a_v, a_q


# -------------------------------------------------------------------------------------------------------- MLP
@torch.library.custom_op(f"{NS}::mlp", mutates_args=(), device_types="cuda")
def mlp(vhat: Tensor, qhat: Tensor, Ww: Tensor, bw: Tensor, Wp: Tensor, bp: Tensor, Ws: Tensor, bs: Tensor, Wh: Tensor,
        bh: Tensor) -> Tuple[Tensor, Tensor]:
    """MLPClassifier forward (reference model.py:414-434) on stacked [3,B,d] attended features.

    Returns logits [B,K] and ``saved``, the opaque buffer the backward kernels read (bf16 operand planes of the weights
    and of the layer inputs; layout private to the library)."""
    _cuda_f32(vhat, qhat, Ww, bw, Wp, bp, Ws, bs, Wh, bh)
    vhat, qhat, Ww, bw, Wp, bp, Ws, bs, Wh, bh = map(_c, (vhat, qhat, Ww, bw, Wp, bp, Ws, bs, Wh, bh))
    _, B, d = vhat.shape
    m, K = Ws.shape[0], Wh.shape[0]
    dev = vhat.device
    logits = torch.empty(B, K, dtype=torch.float32, device=dev)
    L = _lib.lib()
    with torch.cuda.device(dev):
        saved = _ws(L.hca_mlp_saved_bytes(B, d, m, K), dev)
        _lib.check(L.hca_mlp_fwd(_ptr(vhat), _ptr(qhat), _ptr(Ww), _ptr(bw), _ptr(Wp), _ptr(bp), _ptr(Ws), _ptr(bs), _ptr(Wh), _ptr(bh),
                                 _ptr(logits), _ptr(saved), saved.numel(), B, d, m, K, None, 0, _stream()), "mlp_fwd")
    return logits, saved


def _mlp_saved_bytes(B: int, d: int, m: int, K: int) -> int:
    r8 = lambda x: (x + 7) // 8 * 8
    pl = lambda rows, ld: (2 * rows * ld * 2 + 255) // 256 * 256
    return pl(d, d) + pl(d, 2 * d) + pl(m, 2 * d) + pl(K, r8(m)) + pl(B, d) + 2 * pl(B, 2 * d) + pl(B, r8(m)) + 256


@mlp.register_fake
def _(vhat, qhat, Ww, bw, Wp, bp, Ws, bs, Wh, bh):
    _, B, d = vhat.shape
    return vhat.new_empty(B, Wh.shape[0]), vhat.new_empty(_mlp_saved_bytes(B, d, Ws.shape[0], Wh.shape[0]), dtype=torch.uint8)


@torch.library.custom_op(f"{NS}::mlp_bwd", mutates_args=("dWw", "dbw", "dWp", "dbp", "dWs", "dbs", "dWh", "dbh"), device_types="cuda")
def mlp_bwd(dlogits: Tensor, saved: Tensor, d: int, dWw: Tensor, dbw: Tensor, dWp: Tensor, dbp: Tensor, dWs: Tensor, dbs: Tensor,
            dWh: Tensor, dbh: Tensor) -> Tensor:
    """Returns g [3,B,d] (gradient of q_l + v_l per level); the parameter gradients are WRITTEN into dWw .. dbh."""
    _cuda_f32(dlogits, dWw, dbw, dWp, dbp, dWs, dbs, dWh, dbh)
    dlogits = _c(dlogits)
    assert all(t.is_contiguous() for t in (dWw, dbw, dWp, dbp, dWs, dbs, dWh, dbh))
    B, K = dlogits.shape
    m = dWs.shape[0]
    dev = dlogits.device
    g = torch.empty(3, B, d, dtype=torch.float32, device=dev)
    L = _lib.lib()
    with torch.cuda.device(dev):
        ws = _ws(L.hca_mlp_workspace(B, d, m, K), dev)
        _lib.check(L.hca_mlp_bwd(_ptr(dlogits), _ptr(saved), saved.numel(), _ptr(g), _ptr(dWw), _ptr(dbw), _ptr(dWp), _ptr(dbp),
                                 _ptr(dWs), _ptr(dbs), _ptr(dWh), _ptr(dbh), B, d, m, K, _ptr(ws), ws.numel(), _stream()), "mlp_bwd")
    return g


@mlp_bwd.register_fake
def _(dlogits, saved, d, dWw, dbw, dWp, dbp, dWs, dbs, dWh, dbh):
    return dlogits.new_empty(3, dlogits.shape[0], d)


def _mlp_setup(ctx, inputs, output):
    vhat, qhat, Ww, bw, Wp, bp, Ws, bs, Wh, bh = inputs
    logits, saved = output
    ctx.save_for_backward(Ww, bw, Wp, bp, Ws, bs, Wh, bh, saved)
    ctx.d = vhat.shape[2]
    ctx.set_materialize_grads(False)


def _mlp_backward(ctx, dlogits, *_unused):
    *params, saved = ctx.saved_tensors
    if dlogits is None:
        return (None,) * 10
    gs = [_gbuf(t) for t in params]
    g = mlp_bwd(dlogits, saved, ctx.d, *gs)
    return (g, g, *(_gret(t) for t in gs))                    # q_l + v_l: both receive the same gradient


mlp.register_autograd(_mlp_backward, setup_context=_mlp_setup)


# ------------------------------------------------------------------------------------------------------- loss
@torch.library.custom_op(f"{NS}::cross_entropy_fwd", mutates_args=(), device_types="cuda")
def cross_entropy_fwd(logits: Tensor, labels: Tensor, scale: float) -> Tuple[Tensor, Tensor]:
    """(scale * mean CE, d/dlogits of it) in one launch (reference main.py:179,214: nn.CrossEntropyLoss() with its defaults --
    mean reduction, no class weights).  Labels must be class ids in [0, K): ``ignore_index`` is not implemented (the reference's
    loop never produces -100; main.py:208 feeds the answer ids straight from the dataset)."""
    _cuda_f32(logits)
    _cuda_i64("labels", labels, logits)
    logits, labels = _c(logits), _c(labels)
    if logits.dim() != 2 or labels.shape != (logits.shape[0],):
        raise RuntimeError(f"cross_entropy: logits [B,K] and labels [B] expected, got {tuple(logits.shape)} and {tuple(labels.shape)}")
    B, K = logits.shape
    loss = torch.empty((), dtype=torch.float32, device=logits.device)
    dlogits = torch.empty_like(logits)
    with torch.cuda.device(logits.device):
        ws = _ws(_lib.lib().hca_ce_loss_workspace(B), logits.device)
        _lib.check(_lib.lib().hca_ce_loss(_ptr(logits), K, _ptr(labels), B, K, scale, _ptr(loss), _ptr(dlogits), K, _ptr(ws),
                                          ws.numel(), _stream()), "ce_loss")
    return loss, dlogits


@cross_entropy_fwd.register_fake
def _(logits, labels, scale):
    return logits.new_empty(()), torch.empty_like(logits)


def _ce_setup(ctx, inputs, output):
    ctx.save_for_backward(output[1])
    ctx.set_materialize_grads(False)


def _ce_backward(ctx, dloss, _unused):
    (dlogits,) = ctx.saved_tensors
    return (None if dloss is None else dlogits * dloss), None, None


cross_entropy_fwd.register_autograd(_ce_backward, setup_context=_ce_setup)


def cross_entropy(logits: Tensor, labels: Tensor, scale: float = 1.0) -> Tensor:
    """scale * F.cross_entropy(logits, labels) (mean over the batch) with the gradient computed in the same launch."""
    if not logits.is_cuda:
        raise RuntimeError("hiecoattn_b200 ops run on CUDA only (no CPU fallback); got a CPU tensor")
    return cross_entropy_fwd(_f32c(logits), labels, float(scale))[0]


def _f32c(t: Tensor) -> Tensor:
    return t if t.dtype == torch.float32 else t.float()


# ------------------------------------------------------------------------------------------------ optimizer
@torch.library.custom_op(f"{NS}::adam_step", mutates_args=("p", "m", "v", "step", "coef"), device_types="cuda")
def adam_step(p: Tensor, g: Tensor, m: Tensor, v: Tensor, step: Tensor, coef: Tensor, lr: float, beta1: float, beta2: float,
              eps: float) -> None:
    """Fused Adam update of the flat fp32 parameter buffer (reference main.py:180,222); step is a device int64[1]."""
    _cuda_f32(p, g, m, v, coef)
    assert step.dtype == torch.int64 and step.is_cuda and all(t.is_contiguous() for t in (p, g, m, v))
    with torch.cuda.device(p.device):
        _lib.check(_lib.lib().hca_adam_step(_ptr(p), _ptr(g), _ptr(m), _ptr(v), p.numel(), _ptr(step), _ptr(coef), lr, beta1, beta2, eps,
                                            _stream()), "adam_step")


def adam_prep(step: Tensor, coef: Tensor, lr: float, beta1: float, beta2: float) -> None:
    """Advance the device step counter and derive the bias-correction factors (first half of ``adam_step``)."""
    with torch.cuda.device(coef.device):
        _lib.check(_lib.lib().hca_adam_prep(_ptr(step), _ptr(coef), lr, beta1, beta2, _stream()), "adam_prep")


def dp_reduce_adam(symm, begin: int, end: int, mode: int, opt=None, channel: int = 0, max_ctas: int = 0) -> None:
    """One launch of the fused data-parallel kernel (csrc/dp_fused.cu) on the current stream over elements [begin, end) of the
    symmetric flat buffers ``symm`` (dp._SymmetricBlock).  mode bit 0: Adam (``opt`` = optim.FlatAdam: moments, coefficients),
    bit 1: write the summed gradient back.  Collective over the ranks of the block; not differentiable; mutates g / p / m / v."""
    dev = symm.block.device
    m = v = coef = None
    b1 = b2 = eps = 0.0
    if mode & 1:
        m, v, coef = opt.exp_avg, opt.exp_avg_sq, opt._coef
        b1, b2, eps = opt.betas[0], opt.betas[1], opt.eps
    with torch.cuda.device(dev):
        _lib.check(_lib.lib().hca_dp_reduce_adam(symm.peers, symm.mc, symm.flags_off, symm.g_off, symm.p_off, int(begin), int(end),
                                                 symm.rank, symm.world, int(channel), int(max_ctas), _ptr(m), _ptr(v), _ptr(coef),
                                                 b1, b2, eps, int(mode), _stream()), "dp_reduce_adam")


# ---------------------------------------------------------------------------------- raw GEMM (tests / profiling)
_LAYOUTS = {"nt": 0, "nn": 1, "tn": 2}


def gemm(A: Tensor, B: Tensor, bias: Optional[Tensor] = None, layout: str = "nt", path: int = 1) -> Tensor:
    """One dense contraction through the library's kernels (not differentiable; tests, benchmarks, ncu).

    layout "nt": A[M,K] . B[N,K]^T (+bias)   "nn": A[M,K] . B[K,N] (+bias)   "tn": A[K,M]^T . B[K,N]
    path 1 = tcgen05 bf16x2 split (3 MMAs), 2 = tcgen05 bf16x3 split (6 MMAs)."""
    _cuda_f32(A, B, bias)
    A, B = _c(A), _c(B)
    if layout == "nt":
        (M, K), N = A.shape, B.shape[0]
    elif layout == "nn":
        (M, K), N = A.shape, B.shape[1]
    else:
        (K, M), N = A.shape, B.shape[1]
    D = torch.empty(M, N, dtype=torch.float32, device=A.device)
    L = _lib.lib()
    with torch.cuda.device(A.device):
        ws = _ws(L.hca_gemm_workspace(M, N, K), A.device)
        _lib.check(L.hca_gemm(_ptr(A), _ptr(B), _ptr(bias), _ptr(D), M, N, K, _LAYOUTS[layout], path, _ptr(ws), ws.numel(), _stream()),
                   "gemm")
    return D


def split_planes(x: Tensor) -> Tensor:
    """fp32 [rows, cols] -> the library's operand format: bf16 hi/lo planes [2, rows, cols] (x = hi + lo to ~2^-17)."""
    _cuda_f32(x)
    x = _c(x)
    out = torch.empty(2, x.shape[0], x.shape[1], dtype=torch.bfloat16, device=x.device)
    with torch.cuda.device(x.device):
        _lib.check(_lib.lib().hca_split_planes(_ptr(x), x.shape[0], x.shape[1], _ptr(out), _stream()), "split_planes")
    return out


def proj_planes(a_planes: Tensor, w_planes: Tensor, bias: Tensor, out_planes: Tensor) -> Tensor:
    """out planes [2, M, N] = A[M, K] . W[N, K]^T + bias: ONE launch of the projection kernel (no operand conversion)."""
    _, M, K = a_planes.shape
    N = w_planes.shape[1]
    with torch.cuda.device(a_planes.device):
        _lib.check(_lib.lib().hca_proj_planes(_ptr(a_planes), M, K, _ptr(w_planes), N, _ptr(bias), _ptr(out_planes), _stream()),
                   "proj_planes")
    return out_planes


def gemm_nt(A: Tensor, B: Tensor, bias: Optional[Tensor] = None, path: int = 1) -> Tensor:
    return gemm(A, B, bias, "nt", path)
