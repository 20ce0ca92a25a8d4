"""Host staging buffers for a step's inputs: page-locked (optionally write-combined) memory the loader fills and the GPU reads.

The reference copies pageable tensors with ``.to(device)`` every step (main.py:205-208).  ``staged(array)`` returns a CPU tensor
that lives in memory allocated by the library (``hca_pinned_alloc``), so ``dev.copy_(host, non_blocking=True)`` is a true
asynchronous DMA that overlaps the previous step; with ``write_combined=True`` the pages are write-combined -- the CPU only writes
a staging buffer, and the device's reads of such pages do not have to snoop the CPU caches, which matters when eight GPUs pull
their shards from one host at the same time.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _lib


class _Block:
    def __init__(self, nbytes: int, write_combined: bool):
        self.ptr = C.c_void_p()
        _lib.check(_lib.lib().hca_pinned_alloc(int(nbytes), int(write_combined), C.byref(self.ptr)), "pinned_alloc")
        self.nbytes = int(nbytes)

    def __del__(self):
        try:
            if self.ptr:
                _lib.lib().hca_pinned_free(self.ptr)
                self.ptr = C.c_void_p()
        except Exception:
            pass


def staged(array: np.ndarray, write_combined: bool = True) -> torch.Tensor:
    """A page-locked CPU tensor holding a copy of ``array`` (the block stays alive as long as the tensor does)."""
    array = np.ascontiguousarray(array)
    blk = _Block(max(array.nbytes, 16), write_combined)
    buf = (C.c_char * blk.nbytes).from_address(blk.ptr.value)
    t = torch.frombuffer(buf, dtype=torch.from_numpy(array[:0]).dtype, count=array.size).view(array.shape)
    t._hca_block = blk                       # keeps the allocation alive
    np.frombuffer(buf, dtype=array.dtype, count=array.size).reshape(array.shape)[...] = array       # sequential CPU writes only
    return t
