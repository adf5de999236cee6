"""Network handle wrapper around the dxi_net_* C ABI (device weights, workspace, forward)."""
import ctypes

import numpy as np
import torch

from .. import _lib
from .._tensor import to_dev, ret, device


class DeviceNetwork:
    """A network living in libdeepxi_b200.so.  `outp` mirrors the attribute the reference reads
    (selector.py returns an object whose `.outp` is the Keras output tensor); here it is the callable
    that maps an input batch [B, T, n_feat] to x_bar [B, T, n_outp]."""

    kind = None

    def __init__(self, cfg, precision):
        if precision not in _lib.PRECISIONS:
            raise ValueError('precision must be one of %s' % sorted(_lib.PRECISIONS))
        self.precision = precision
        cfg.precision = _lib.PRECISIONS[precision]
        self._cfg = cfg
        self._h = ctypes.c_void_p(0)
        self._ws = {}          # one workspace per (device, stream): concurrent streams must not share scratch
        self._loaded = False
        self.n_feat, self.n_outp = cfg.n_feat, cfg.n_outp
        lib = _lib.load()
        device()
        _lib.check(lib.dxi_net_create(ctypes.byref(self._h), _lib.NET_KINDS[self.kind], ctypes.byref(cfg)), value_error=True)
        self.outp = self.__call__

    def load_weights(self, weights):
        """weights: {'layer_with_weights-<i>/<var>': ndarray} (deepxi_b200.weights / tfbundle)."""
        lib = _lib.load()
        for name, arr in weights.items():
            a = np.ascontiguousarray(np.asarray(arr, np.float32))
            shape = (ctypes.c_int64 * a.ndim)(*a.shape)
            _lib.check(lib.dxi_net_load(self._h, name.encode(), a.ctypes.data_as(ctypes.c_void_p), shape, a.ndim),
                       value_error=True)
        _lib.check(lib.dxi_net_finalize(self._h, _lib.stream_ptr()))
        self._loaded = True
        return self

    def workspace_bytes(self, B, T):
        return int(_lib.load().dxi_net_workspace_bytes(self._h, B, T))

    def __call__(self, inp):
        if not self._loaded:
            raise RuntimeError('network weights have not been loaded')
        x, was_np = to_dev(inp, torch.float32)
        squeeze = x.dim() == 2
        if squeeze:
            x = x[None]
        if x.dim() != 3 or x.shape[-1] != self.n_feat:
            raise ValueError('input must be [B, T, %d]' % self.n_feat)
        B, T, _ = x.shape
        out = torch.empty((B, T, self.n_outp), dtype=torch.float32, device=x.device)
        if B and T:
            need = self.workspace_bytes(B, T)
            key = (str(x.device), torch.cuda.current_stream(x.device).cuda_stream)
            ws = self._ws.get(key)
            if ws is None or ws.numel() < need:
                ws = self._ws[key] = torch.empty(need, dtype=torch.uint8, device=x.device)
            with torch.cuda.device(x.device):      # the C side launches on the current device; the handle checks it is its own
                _lib.check(_lib.load().dxi_net_forward(self._h, _lib.ptr(x), B, T, _lib.ptr(out), _lib.ptr(ws), ws.numel(),
                                                       _lib.stream_ptr(x.device)))
        return ret(out[0] if squeeze else out, was_np)

    predict = __call__

    def __del__(self):
        try:
            if self._h:
                _lib.load().dxi_net_destroy(self._h)
                self._h = ctypes.c_void_p(0)
        except Exception:
            pass
