"""Batch-sharded data parallelism for the co-attention path: one process per GPU, flat gradient / parameter buffers, and
the gradient all-reduce over NVLink 5 / NVSwitch.

The reference has no multi-GPU support at all (a commented-out nn.DataParallel TODO, main.py:102-106);
the path shards over the batch because no op in model.py:246-434 mixes samples -- the only cross-sample
reductions are the weight-gradient sums and the mean of the loss (main.py:179).

Design:
  * every participating parameter's ``.grad`` is a VIEW into one flat fp32 buffer, laid out in the order
    gradients become ready during backward (classifier -> co-attention -> LSTM / conv -> embedding), so
    there is no gather copy before the collective and no scatter after it;
  * ``co_attention.W_b`` never receives a gradient (reference model.py:347 vs :377) and frozen VGG weights
    have ``requires_grad=False``: both are left out of the buffer (``skip`` / requires_grad);
  * the mean over ranks is obtained by scaling the loss by 1/world_size before backward (``loss_scale``), so no
    extra pass over the buffer is needed;
  * the reducer is a "gradient sink" of ops.py: the backward kernels WRITE each weight gradient straight into its slot of
    the flat buffer (no per-parameter autograd accumulation kernel, no clearing pass over the buffer).  A slot takes one
    direct write per step; a parameter used twice gets its further contributions through autograd's in-place accumulation
    into the same view.  Slots that were not written directly in the previous step are cleared by ``zero_grad()``.

Two transports for the collective:
  * ``fused`` (CUDA, world > 1, the default when symmetric memory can be set up): gradients AND parameters live in one
    symmetric-memory block per GPU (torch.distributed._symmetric_memory does the allocation and the handle exchange; it is
    plumbing only).  ``optim.FlatAdam.step()`` then runs ONE hand-written kernel (csrc/dp_fused.cu) that reduces the
    gradients across ranks in the NVSwitch (``multimem.ld_reduce``), applies Adam to this rank's slice and multicasts
    the new parameters to every rank (``multimem.st``) -- all-reduce, optimizer and broadcast in one pass, capturable in
    a CUDA graph, no NCCL call in the step.  Ranges of the buffer that are complete early (classifier + co-attention
    gradients) can be processed on a side stream while the rest of backward still runs (``reduce_adam_range``).
  * NCCL / gloo (``fused=False``; gloo in the CPU unit tests): bucketed ``dist.all_reduce`` in ``finish()``; with
    ``overlap=True`` each bucket is launched from the gradient hooks as soon as it is complete.  Overlap is OFF by default:
    the persistent LSTM kernel wants its CTAs co-resident (csrc/lstm.cu checks and refuses otherwise), and a collective
    kernel holding SMs at that moment only delays both.

The model must be on its device BEFORE the reducer is built: ``flat_params=True`` re-points ``param.data`` into the flat
buffer, and anything that re-allocates parameter storage afterwards (``model.to()``, ``nn.LSTM.flatten_parameters()``)
would silently detach them -- ``finish()`` / ``FlatAdam.step()`` check the aliasing and raise.
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
                 skip: Sequence[str] = ("co_attention.W_b.",), bucket_bytes: int = 16 << 20, overlap: bool = False,
                 flat_params: bool = False, align: int = 64, direct_write: bool = True, fused: Optional[bool] = None,
                 early_split: Optional[str] = None):
        self.group = process_group
        self.world = dist.get_world_size(process_group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(process_group) if dist.is_initialized() else 0
        self.loss_scale = 1.0 / self.world
        self.overlap = overlap
        self.fused = False
        self._symm = None
        self._optimizer = None              # a FlatAdam that has taken over the collective (fused mode)
        # early_split = a parameter-name prefix: a bucket boundary is forced in front of the first parameter carrying it, and
        # ``on_early(end)`` (if set) is called during backward as soon as every gradient in [0, end) has been written -- the hook
        # the fused transport uses to reduce the classifier / co-attention slice while the encoder's backward still runs
        self.on_early = None
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
        want_fused = fused if fused is not None else (self.world > 1 and dev.type == "cuda" and flat_params and dt == torch.float32)
        if want_fused:
            if not (self.world > 1 and dev.type == "cuda" and flat_params and dt == torch.float32):
                raise ValueError("fused=True needs world > 1, CUDA fp32 parameters and flat_params=True")
            try:
                self._symm = _SymmetricBlock(total, dev, process_group, self.rank, self.world)
            except Exception as e:                      # no symmetric memory on this system: NCCL transport
                if fused:
                    raise
                import warnings
                warnings.warn(f"FlatGradAllReduce: symmetric memory unavailable ({type(e).__name__}: {e}); using dist.all_reduce")
        if self._symm is not None:
            self.fused = True
            self.flat, self.flat_p = self._symm.g, self._symm.p
        else:
            self.flat = torch.zeros(total, dtype=dt, device=dev)
            # optional: the parameters themselves become views of one flat buffer too (same offsets), which is what lets the
            # optimizer step be ONE fused kernel over (flat_p, flat, m, v) -- see optim.FlatAdam
            self.flat_p = torch.zeros(total, dtype=dt, device=dev) if flat_params else None
        # carve views and buckets
        self.buckets: List[Tuple[int, int]] = []            # [start, end) element ranges of the flat buffer
        self._bucket_of: List[int] = []
        off, b_start, limit = 0, 0, max(1, bucket_bytes // self.flat.element_size())
        split_done = early_split is None
        for name, p in zip(self.names, self.params):
            n = p.numel()
            if not split_done and name.startswith(early_split):
                split_done = True
                if off > b_start:
                    self.buckets.append((b_start, off))
                    b_start = off
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
        self._offsets = [int((g.data_ptr() - self.flat.data_ptr()) // self.flat.element_size()) for g in self._views]
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
        if self._left[b] == 0 and b == 0 and self.on_early is not None:
            self.on_early(self.buckets[0][1])
        if self._left[b] == 0 and self.world > 1 and self.overlap and not self.fused:
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
        self.check_aliasing()
        if self.fused:
            if self._optimizer is None:                 # no fused optimizer attached: a plain all-reduce through the same kernel
                self.reduce_adam_range(0, self.flat.numel(), mode=2)
        elif self.world > 1:
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

    def check_aliasing(self):
        """Every parameter (with flat_params) and every ``.grad`` must still be the view of the flat buffers carved at construction:
        ``model.to()`` / ``_apply`` / ``nn.LSTM.flatten_parameters()`` re-allocate parameter storage, after which the optimizer would
        update a buffer no module reads.  Host-side pointer compares only (no device work, safe during CUDA-graph capture)."""
        esz = self.flat.element_size()
        for i, p in enumerate(self.params):
            if p.grad is not None and p.grad.data_ptr() != self.flat.data_ptr() + self._offsets[i] * esz:
                raise RuntimeError(f"FlatGradAllReduce: {self.names[i]}.grad no longer points into the flat gradient buffer")
            if self.flat_p is not None and p.data_ptr() != self.flat_p.data_ptr() + self._offsets[i] * esz:
                raise RuntimeError(f"FlatGradAllReduce: parameter {self.names[i]} no longer aliases the flat parameter buffer "
                                   "(was the model moved / re-flattened after the reducer was built? build the reducer last)")

    def reduce_adam_range(self, begin: int, end: int, mode: int, opt=None, channel: int = 0, max_ctas: int = 0):
        """Fused transport: one launch of csrc/dp_fused.cu over elements [begin, end) of the flat buffers on the current stream.
        mode 1 = all-reduce + Adam + parameter broadcast (``opt`` = the FlatAdam holding m / v / coef), 2 = all-reduce only."""
        if not self.fused:
            raise RuntimeError("reduce_adam_range needs the fused (symmetric-memory) transport")
        from . import ops
        ops.dp_reduce_adam(self._symm, begin, end, mode, opt, channel, max_ctas)

    def offset_of(self, prefix: str) -> int:
        """First element of the flat buffers that belongs to a parameter whose name starts with ``prefix`` (layout = backward order)."""
        for n, o in zip(self.names, self._offsets):
            if n.startswith(prefix):
                return o
        raise KeyError(prefix)

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


class _SymmetricBlock:
    """One symmetric-memory allocation per rank holding [flag words | flat gradients | flat parameters], rendezvoused over the
    process group so that every rank has every peer's block mapped (and, on NVSwitch systems, a multicast address for it)."""

    def __init__(self, total: int, device, group, rank: int, world: int):
        import ctypes as C
        import torch.distributed._symmetric_memory as symm_mem
        from . import _lib
        if world > 8:
            raise ValueError("the fused transport covers one NVSwitch domain (<= 8 GPUs)")
        flag_words = int(_lib.lib().hca_dp_flags_bytes()) // 4
        pad = lambda n: (n + 63) // 64 * 64
        self.flags_off = 0
        self.g_off = pad(flag_words) * 4
        self.p_off = self.g_off + pad(total) * 4
        nfloat = pad(flag_words) + 2 * pad(total)
        self.block = symm_mem.empty(nfloat, dtype=torch.float32, device=device)
        self.block.zero_()
        torch.cuda.synchronize(device)
        pg = group if group is not None else dist.group.WORLD
        self.handle = symm_mem.rendezvous(self.block, pg)
        ptrs = list(self.handle.buffer_ptrs)
        if len(ptrs) != world or int(ptrs[rank]) != self.block.data_ptr():
            raise RuntimeError("symmetric memory rendezvous returned unexpected buffer pointers")
        self.peers = (C.c_uint64 * world)(*[int(x) for x in ptrs])
        self.mc = int(self.handle.multicast_ptr) if getattr(self.handle, "multicast_ptr", 0) else 0
        self.rank, self.world, self.total = rank, world, total
        self.g = self.block[self.g_off // 4: self.g_off // 4 + total]
        self.p = self.block[self.p_off // 4: self.p_off // 4 + total]
        dist.barrier(group=pg)              # every rank's flag words are zero before anybody signals
