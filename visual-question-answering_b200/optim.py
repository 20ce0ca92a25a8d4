"""Optimizer step of the training loop (reference main.py:180,222: ``torch.optim.Adam(model.parameters(), lr)``) as one
fused CUDA kernel over flat buffers.

``FlatAdam`` takes the ``dp.FlatGradAllReduce`` that owns the flat gradient buffer (built with ``flat_params=True`` so
that the parameters are views of a second flat buffer with the same offsets).  The update then reads / writes four flat
fp32 arrays in a single pass (csrc/adam.cu) instead of torch's eight multi-tensor passes, keeps its step counter on the
device (CUDA-graph capturable) and touches exactly the tensors the all-reduce covers: ``co_attention.W_b`` (never used by
the reference's forward, model.py:347 vs :377, so its gradient is None and torch's Adam skips it too) and frozen VGG
weights are left alone.

With the reducer's fused transport (world > 1, symmetric memory) the step is ALSO the collective: one kernel reduces the
gradients across ranks in the NVSwitch, updates this rank's slice and multicasts the new parameters (csrc/dp_fused.cu);
the moment buffers then hold meaningful values only for this rank's slice (sharded optimizer state).
"""
from __future__ import annotations

import torch

from . import ops


class FlatAdam:
    def __init__(self, reducer, lr: float = 1e-3, betas=(0.9, 0.999), eps: float = 1e-8):
        if reducer.flat_p is None:
            raise ValueError("FlatAdam needs FlatGradAllReduce(..., flat_params=True)")
        if not reducer.flat.is_cuda:
            raise RuntimeError("FlatAdam runs on CUDA only (no CPU fallback)")
        self.reducer = reducer
        self.lr, self.betas, self.eps = float(lr), (float(betas[0]), float(betas[1])), float(eps)
        self.exp_avg = torch.zeros_like(reducer.flat)
        self.exp_avg_sq = torch.zeros_like(reducer.flat)
        self.step_count = torch.zeros(1, dtype=torch.int64, device=reducer.flat.device)
        self._coef = torch.zeros(2, dtype=torch.float32, device=reducer.flat.device)
        self._side = None
        self._early_ctas = 16
        if reducer.fused:
            reducer._optimizer = self       # finish() leaves the collective to step()

    def overlap_early_slice(self, max_ctas: int = 16):
        """Fused transport: process the slice of the flat buffers in front of the reducer's ``early_split`` boundary (the classifier and
        co-attention gradients: complete long before backward ends) on a side stream while the rest of backward is still running, with a
        grid of at most ``max_ctas`` CTAs (the persistent LSTM kernel leaves 20 of the 148 SMs free).  ``step()`` waits for that stream and
        covers the rest.  A no-op on one GPU / on the NCCL transport.  Capturable in a CUDA graph (the side stream becomes a branch)."""
        r = self.reducer
        if not r.fused or len(r.buckets) < 2:
            return False
        self._side = torch.cuda.Stream(r.flat.device)
        self._early_ctas = int(max_ctas)

        def on_early(end):
            self._side.wait_stream(torch.cuda.current_stream(r.flat.device))
            self.step_early(end, self._side, self._early_ctas)

        r.on_early = on_early
        return True

    @torch.no_grad()
    def step(self):
        r = self.reducer
        r.check_aliasing()
        if r.fused:
            if self._side is not None:
                torch.cuda.current_stream(r.flat.device).wait_stream(self._side)
            done = self._early_end          # elements [0, done) were already reduced + updated by step_early() this step
            self._early_end = 0
            if done == 0:
                ops.adam_prep(self.step_count, self._coef, self.lr, self.betas[0], self.betas[1])
            if done < r.flat.numel():
                r.reduce_adam_range(done, r.flat.numel(), 1, self)
            return
        ops.adam_step(r.flat_p, r.flat, self.exp_avg, self.exp_avg_sq, self.step_count, self._coef, self.lr, self.betas[0],
                      self.betas[1], self.eps)

    _early_end = 0

    @torch.no_grad()
    def step_early(self, end: int, stream, max_ctas: int = 16):
        """Fused transport only: reduce + update elements [0, end) of the flat buffers NOW on ``stream`` (a side stream that has
        waited for the kernels producing those gradients), with a grid capped at ``max_ctas`` so that it shares the GPU with the
        rest of backward.  ``step()`` then covers [end, n).  The step counter advances here."""
        r = self.reducer
        if not r.fused or end <= 0:
            return
        with torch.cuda.stream(stream):
            ops.adam_prep(self.step_count, self._coef, self.lr, self.betas[0], self.betas[1])
            r.reduce_adam_range(0, end, 1, self, channel=1, max_ctas=max_ctas)
        self._early_end = end

    def zero_grad(self, set_to_none: bool = False):
        self.reducer.zero_grad()
