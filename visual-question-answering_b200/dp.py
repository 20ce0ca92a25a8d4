"""Batch-sharded data parallelism for the co-attention path: one process per GPU, NCCL all-reduce of a
flat gradient buffer over NVLink 5 / NVSwitch, issued bucket by bucket in reverse-backward order so the
transfer of the early buckets overlaps the remaining backward.

The reference has no multi-GPU support at all (a commented-out nn.DataParallel TODO, main.py:102-106);
the path shards over the batch because no op in model.py:246-434 mixes samples -- the only cross-sample
reductions are the weight-gradient sums and the mean of the loss (main.py:179).

Design:
  * every participating parameter's ``.grad`` is a VIEW into one flat fp32 buffer, laid out in the order
    gradients become ready during backward (classifier -> co-attention -> LSTM / conv -> embedding), so
    there is no gather copy before the collective and no scatter after it;
  * ``co_attention.W_b`` never receives a gradient (reference model.py:347 vs :377) and frozen VGG weights
    have ``requires_grad=False``: both are left out of the buffer (``skip`` / requires_grad);
  * post-accumulate-grad hooks count down each bucket and launch its all-reduce (SUM) as soon as it is
    complete; the mean over ranks is obtained by scaling the loss by 1/world_size before backward
    (``loss_scale``), so no extra pass over the buffer is needed;
  * ``finish()`` waits for the outstanding collectives before the optimizer step;
  * the reducer is a "gradient sink" of ops.py: the backward kernels WRITE each weight gradient straight into its slot of
    the flat buffer (no per-parameter autograd accumulation kernel, no clearing pass over the buffer).  A slot takes one
    direct write per step; a parameter used twice gets its further contributions through autograd's in-place accumulation
    into the same view.  Slots that were not written directly in the previous step are cleared by ``zero_grad()``.
Works with any backend torch.distributed offers (NCCL on the GPUs; gloo in the CPU unit tests).
"""
from __future__ import annotations

from typing import Iterable, List, Optional, Sequence, Tuple

import torch
import torch.distributed as dist

# parameter-name prefixes in the order their gradients are produced by backward
BACKWARD_ORDER = ("mlp_classify.", "co_attention.", "question_encoder.sentence_lstm.", "question_encoder.phrase_conv_pool.",
                  "question_encoder.word_embedding.")


def shard_batch(n_items: int, rank: int, world: int) -> slice:
    """Contiguous shard of a length-sorted global batch: every shard stays sorted (utils.py:33-45)."""
    per = n_items // world
    if per * world != n_items:
        raise ValueError(f"global batch {n_items} is not divisible by world size {world}")
    return slice(rank * per, (rank + 1) * per)


class FlatGradAllReduce:
    def __init__(self, named_params: Iterable[Tuple[str, torch.nn.Parameter]], process_group=None,
                 skip: Sequence[str] = ("co_attention.W_b.",), bucket_bytes: int = 16 << 20, overlap: bool = True,
                 flat_params: bool = False, align: int = 64, direct_write: bool = True):
        self.group = process_group
        self.world = dist.get_world_size(process_group) if dist.is_initialized() else 1
        self.loss_scale = 1.0 / self.world
        self.overlap = overlap
        items = [(n, p) for n, p in named_params if p.requires_grad and not any(n.startswith(s) for s in skip)]

        def order(item):
            for i, pre in enumerate(BACKWARD_ORDER):
                if item[0].startswith(pre):
                    return i
            return len(BACKWARD_ORDER)

        items.sort(key=order)                               # stable: keeps definition order inside a group
        self.names = [n for n, _ in items]
        self.params = [p for _, p in items]
        if not self.params:
            raise ValueError("no parameters to reduce")
        dev, dt = self.params[0].device, self.params[0].dtype
        # every tensor starts on an `align`-element boundary (256 bytes for fp32): vector loads, TMA and the fused optimizer
        # all want 16-byte aligned bases; the padding elements stay zero
        pad = lambda n: (n + align - 1) // align * align
        total = sum(pad(p.numel()) for p in self.params)
        self.flat = torch.zeros(total, dtype=dt, device=dev)
        # optional: the parameters themselves become views of one flat buffer too (same offsets), which is what lets the
        # optimizer step be ONE fused kernel over (flat_p, flat, m, v) -- see optim.FlatAdam
        self.flat_p = torch.zeros(total, dtype=dt, device=dev) if flat_params else None
        # carve views and buckets
        self.buckets: List[Tuple[int, int]] = []            # [start, end) element ranges of the flat buffer
        self._bucket_of: List[int] = []
        off, b_start, limit = 0, 0, max(1, bucket_bytes // self.flat.element_size())
        for p in self.params:
            n = p.numel()
            if self.flat_p is not None:
                view = self.flat_p[off:off + n].view_as(p)
                view.copy_(p.data)
                p.data = view
            p.grad = self.flat[off:off + n].view_as(p)
            off += pad(n)
            self._bucket_of.append(len(self.buckets))
            if off - b_start >= limit:
                self.buckets.append((b_start, off))
                b_start = off
        if off > b_start:
            self.buckets.append((b_start, off))
        self._bucket_of = [min(b, len(self.buckets) - 1) for b in self._bucket_of]
        self._need = [0] * len(self.buckets)
        for b in self._bucket_of:
            self._need[b] += 1
        self._left = list(self._need)
        self._work = []
        # gradient sink state (see ops._gbuf / ops._gret)
        self._views = [p.grad for p in self.params]
        self._by_param = {p.data_ptr(): i for i, p in enumerate(self.params)}
        self._by_grad = {g.data_ptr(): i for i, g in enumerate(self._views)}
        self._taken: set = set()            # slots handed to a backward kernel this step
        self._direct_prev: Optional[set] = None   # slots written directly in the previous step (None: unknown -> clear all)
        self._clean: set = set()            # slots that hold this step's data only (cleared or directly written)
        self._counted: set = set()          # parameters already counted towards their bucket this step
        self._hooks = []
        for i, p in enumerate(self.params):
            self._hooks.append(p.register_post_accumulate_grad_hook(self._make_hook(i)))
        if direct_write and self.flat.is_cuda:
            from . import ops
            ops.add_grad_sink(self)

    # ---------------------------------------------------------------------------------------------------
    def _ready(self, i):
        """Parameter i's gradient is (at least partly) in its slot: count it towards its bucket once per step."""
        if i in self._counted:
            return
        self._counted.add(i)
        b = self._bucket_of[i]
        self._left[b] -= 1
        if self._left[b] == 0 and self.world > 1 and self.overlap:
            self._launch(b)

    def _make_hook(self, i):
        def hook(param):
            # autograd has accumulated into param.grad in place (the view of the flat buffer)
            if param.grad is not None and param.grad.data_ptr() != self._views[i].data_ptr():
                raise RuntimeError(f"FlatGradAllReduce: {self.names[i]}.grad no longer points into the flat buffer "
                                   "(use zero_grad() of the reducer / FlatAdam, not set_to_none=True)")
            if i not in self._clean:
                raise RuntimeError(f"FlatGradAllReduce: a gradient was accumulated into {self.names[i]} before zero_grad() cleared its slot")
            self._ready(i)
        return hook

    def set_overlap(self, overlap: bool):
        """Switch between all-reduce launched as buckets complete (overlapped with backward) and one pass in finish()."""
        self.overlap = overlap

    # gradient-sink protocol (ops.add_grad_sink) --------------------------------------------------------
    def take(self, w: torch.Tensor):
        """Slot of parameter tensor ``w`` for a direct write by a backward kernel, or None (unknown tensor / already taken)."""
        i = self._by_param.get(w.data_ptr())
        if i is None or i in self._taken or w.shape != self.params[i].shape:
            return None
        self._taken.add(i)
        return self._views[i]

    def delivered(self, g: torch.Tensor) -> bool:
        """True if ``g`` is one of this reducer's slots (its gradient has just been written in place)."""
        i = self._by_grad.get(g.data_ptr())
        if i is None or g.shape != self._views[i].shape:
            return False
        self._clean.add(i)
        self._ready(i)
        return True

    def _launch(self, b):
        s, e = self.buckets[b]
        self._work.append(dist.all_reduce(self.flat[s:e], op=dist.ReduceOp.SUM, group=self.group, async_op=True))

    def zero_grad(self):
        """Start of a step: clears the slots that will be accumulated into rather than overwritten (all of them the first
        time; afterwards those the previous step's backward did not write directly) and re-arms the bucket counters.
        Parameters keep their ``.grad`` views."""
        n = len(self.params)
        if self._direct_prev is None or not self._direct_prev:
            self.flat.zero_()
            self._clean = set(range(n))
        else:
            self._clean = set()
            for i in range(n):
                if i not in self._direct_prev:
                    self._views[i].zero_()
                    self._clean.add(i)
        self._taken = set()
        self._counted = set()
        self._left = list(self._need)
        self._work = []

    def finish(self):
        """Call after backward, before optimizer.step(): launches what the hooks did not and waits."""
        # slots neither cleared nor written this step would carry the previous step's gradient into the optimizer
        for i in range(len(self.params)):
            if i not in self._clean:
                self._views[i].zero_()
                self._clean.add(i)
        self._direct_prev = set(self._taken)
        if self.world > 1:
            if not self.overlap:
                for b in range(len(self.buckets)):
                    self._launch(b)
            else:
                for b, left in enumerate(self._left):
                    if left > 0:                            # a parameter received no gradient this step
                        self._launch(b)
            for w in self._work:
                w.wait()
        self._work = []
        self._left = list(self._need)

    def close(self):
        """Detach from autograd and from the backward kernels: removes the hooks and the gradient-sink registration (the flat buffers
        and the ``.grad`` / ``.data`` views stay valid)."""
        for h in self._hooks:
            h.remove()
        self._hooks = []
        from . import ops
        ops.remove_grad_sink(self)

    def grad_bytes(self) -> int:
        return self.flat.numel() * self.flat.element_size()
