"""Mirror of deepxi/map.py: map_selector (:15-42) and NormalCDF (:352-402) with map_type 'DBNormalCDF'.

The other map classes of the reference (Linear, Clip, Logistic, Standardise, MinMaxScaling, LaplaceCDF,
TruncatedLaplaceCDF, UniformCDF, Square) serve no committed model and are out of scope; selecting one
raises NotImplementedError, an unknown name raises ValueError("Invalid map_type.") as in map.py:42.
"""
import numpy as np
import torch

from . import _lib
from ._tensor import to_dev, ret, device

_OTHER = ('Linear', 'DB', 'Clip', 'Logistic', 'Standardise', 'MinMaxScaling', 'TruncatedLaplaceCDF', 'LaplaceCDF',
          'UniformCDF', 'Square')


def map_selector(map_type, params=None):
    if map_type is not None and 'NormalCDF' in map_type:
        if map_type != 'DBNormalCDF':
            raise NotImplementedError("only 'DBNormalCDF' (the map of the committed models, run.sh:28,116,162) is built")
        return NormalCDF(map_type, params)
    if map_type in ('Linear', 'DB') or any(k in str(map_type) for k in _OTHER[2:]):
        raise NotImplementedError('map_type %r is not used by any committed model: out of scope' % (map_type,))
    raise ValueError('Invalid map_type.')


class Map:
    """Base map class (map.py:44-95)."""

    def __init__(self, map_type, params=None):
        self.map_type = map_type
        self.params = params

    def stats(self, x):
        pass


class NormalCDF(Map):
    """Normal CDF map of the a priori SNR in dB (map.py:352-402)."""

    def __init__(self, map_type, params=None):
        super().__init__(map_type, params)
        self.mu = None
        self.sigma = None
        self._dev = {}

    def set_stats(self, mu, sigma):
        self.mu = np.ascontiguousarray(np.asarray(mu, np.float32).reshape(-1))
        self.sigma = np.ascontiguousarray(np.asarray(sigma, np.float32).reshape(-1))
        self._dev = {}
        return self

    def _stats_dev(self, dev):
        if self.mu is None:
            raise RuntimeError('NormalCDF statistics (mu, sigma) have not been set')
        key = str(dev)
        if key not in self._dev:
            self._dev[key] = (torch.from_numpy(self.mu).to(dev), torch.from_numpy(self.sigma).to(dev))
        return self._dev[key]

    def map(self, x):
        """xi -> xi_bar = Phi((10 log10 max(xi,1e-12) - mu) / sigma) (map.py:356-371, :62-73)."""
        x, was_np = to_dev(x, torch.float32)
        mu, sigma = self._stats_dev(x.device)
        if x.shape[-1] != mu.numel():
            raise ValueError('last dimension must be %d' % mu.numel())
        out = torch.empty_like(x)
        if x.numel():
            _lib.check(_lib.load().dxi_cdf_map(_lib.ptr(x), _lib.ptr(mu), _lib.ptr(sigma), x.numel() // mu.numel(),
                                               mu.numel(), _lib.ptr(out), _lib.stream_ptr(x.device)))
        return ret(out, was_np)

    def inverse(self, x_bar):
        """xi_bar -> xi = 10^((sigma sqrt(2) erfinv(2 xi_bar - 1) + mu)/10) (map.py:373-390)."""
        x_bar, was_np = to_dev(x_bar, torch.float32)
        mu, sigma = self._stats_dev(x_bar.device)
        if x_bar.shape[-1] != mu.numel():
            raise ValueError('last dimension must be %d' % mu.numel())
        out = torch.empty_like(x_bar)
        if x_bar.numel():
            _lib.check(_lib.load().dxi_map_gain(_lib.ptr(x_bar), _lib.ptr(mu), _lib.ptr(sigma),
                                                x_bar.numel() // mu.numel(), mu.numel(), 0, _lib.ptr(out), None, None,
                                                _lib.stream_ptr(x_bar.device)))
        return ret(out, was_np)

    def stats(self, x):
        """Per-bin mean / population std of 10 log10 max(xi, 1e-12) (map.py:392-402).

        Training-target side; torch reductions are used here, this is not the hot path."""
        x, _ = to_dev(x, torch.float32)
        xdb = 10.0 * torch.log10(torch.clamp(x, min=1e-12))
        self.set_stats(xdb.mean(dim=0).cpu().numpy(), xdb.std(dim=0, unbiased=False).cpu().numpy())
