"""Pinned host tensors for HostPipeline (dxi_host_alloc): like torch.empty(...).pin_memory(), plus the write-combined flavour
for input batches (the host fills them once, the device reads them over PCIe; never read them back on the host: that is slow)."""
import ctypes

import numpy as np
import torch

from . import _lib


class _Owner:
    def __init__(self, ptr):
        self.ptr = ptr

    def __del__(self):
        try:
            if self.ptr:
                _lib.load().dxi_host_free(ctypes.c_void_p(self.ptr))
                self.ptr = 0
        except Exception:
            pass


_owners = {}


def pinned_empty(shape, dtype=torch.int16, write_combined=False):
    """A pinned host tensor of the given shape; freed when the tensor (and every view of its storage) is gone."""
    n = int(np.prod(shape))
    nbytes = max(1, n * torch.empty((), dtype=dtype).element_size())
    p = ctypes.c_void_p(0)
    _lib.check(_lib.load().dxi_host_alloc(ctypes.byref(p), nbytes, int(bool(write_combined))))
    owner = _Owner(p.value)
    buf = (ctypes.c_char * nbytes).from_address(p.value)
    buf._dxi_owner = owner            # the ctypes array keeps the allocation alive for as long as torch holds the buffer
    t = torch.frombuffer(buf, dtype=dtype, count=n).view(*shape)
    return t


def pinned_copy(array, write_combined=False):
    """numpy array / tensor -> pinned host tensor with the same contents."""
    src = torch.as_tensor(array)
    t = pinned_empty(tuple(src.shape), src.dtype, write_combined)
    t.copy_(src)
    return t
