"""Mirror of deepxi/gain.py: gfunc(xi, gamma, gtype, cdm) (:168-191) and the named gain functions.

Element-wise, any shape; numpy in -> numpy out, torch in -> CUDA torch out.  Unknown gtype raises
ValueError('Invalid gain function type.') as gain.py:190 does.
"""
import torch

from . import _lib
from ._tensor import to_dev, ret

_NEEDS_GAMMA = ('mmse-lsa', 'mmse-stsa', 'deepmmse')


def gfunc(xi, gamma=None, gtype=None, cdm=None):
    code = _lib.gtype_code(gtype)
    xi, was_np = to_dev(xi, torch.float32)
    g = None
    if gtype in _NEEDS_GAMMA:
        if gamma is None:
            raise ValueError('%s needs the a posteriori SNR gamma' % gtype)
        g, _ = to_dev(gamma, torch.float32)
        if g.shape != xi.shape:
            g = g.expand_as(xi).contiguous()
    G = torch.empty_like(xi)
    if xi.numel():
        _lib.check(_lib.load().dxi_gfunc(_lib.ptr(xi), _lib.ptr(g, allow_none=True), xi.numel(), code, _lib.ptr(G),
                                         _lib.stream_ptr(xi.device)), value_error=True)
    return ret(G, was_np)


def mmse_stsa(xi, gamma): return gfunc(xi, gamma, 'mmse-stsa')
def mmse_lsa(xi, gamma): return gfunc(xi, gamma, 'mmse-lsa')
def wf(xi): return gfunc(xi, None, 'wf')
def srwf(xi): return gfunc(xi, None, 'srwf')
def cwf(xi): return gfunc(xi, None, 'cwf')
def irm(xi): return gfunc(xi, None, 'irm')
def ibm(xi): return gfunc(xi, None, 'ibm')
def deepmmse(xi, gamma): return gfunc(xi, gamma, 'deepmmse')


def dgwf(xi, cdm):
    raise NotImplementedError("'dgwf' needs the constructive-deconstructive mask of the STDCT target (out of scope)")
