"""Oracle: gain functions (test infrastructure; see oracle/__init__.py).

Restates deepxi/gain.py:13-191 in numpy float32 in the reference's operation order.
scipy.special.exp1 is what the reference itself calls (gain.py:8,67); tf.math.bessel_i0/i1 in
float32 are restated as the float64 cephes value rounded to float32 (overflowing to +inf), which is
what places the Inf/NaN -> Wiener fallback of gain.py:42-44.
"""
import numpy as np
from scipy import special as spsp

f32 = np.float32
GTYPES = ('mmse-lsa', 'mmse-stsa', 'wf', 'srwf', 'cwf', 'irm', 'ibm', 'deepmmse')


def _f32(x):
    return np.asarray(x, f32)


def wf(xi):
    """gain.py:71-81."""
    xi = _f32(xi)
    with np.errstate(all='ignore'):
        return (xi / (xi + f32(1.0))).astype(f32)


def srwf(xi):
    """gain.py:83-93."""
    with np.errstate(all='ignore'):
        return np.sqrt(wf(xi)).astype(f32)


def cwf(xi):
    """gain.py:95-105: wf(sqrt(xi))."""
    with np.errstate(all='ignore'):
        return wf(np.sqrt(_f32(xi)))


def irm(xi):
    """gain.py:129-139."""
    return srwf(xi)


def ibm(xi):
    """gain.py:141-151: float(xi > 1)."""
    return (_f32(xi) > f32(1.0)).astype(f32)


def deepmmse(xi, gamma):
    """gain.py:154-166."""
    xi, gamma = _f32(xi), _f32(gamma)
    with np.errstate(all='ignore'):
        return (f32(1.0) / (f32(1.0) + xi) + xi / (gamma * (f32(1.0) + xi))).astype(f32)


def mmse_lsa(xi, gamma):
    """gain.py:47-69."""
    xi = np.maximum(_f32(xi), f32(1e-12))
    gamma = np.maximum(_f32(gamma), f32(1e-12))
    with np.errstate(all='ignore'):
        v_1 = xi / (f32(1.0) + xi)
        nu = v_1 * gamma
        v_2 = spsp.exp1(nu).astype(f32)
        return (v_1 * np.exp(f32(0.5) * v_2)).astype(f32)


def _bessel_f32(fn, x):
    with np.errstate(over='ignore'):
        return fn(x.astype(np.float64)).astype(f32)


def mmse_stsa(xi, gamma):
    """gain.py:13-45 incl. the NaN/Inf -> Wiener fallback (:42-44)."""
    xi = np.maximum(_f32(xi), f32(1e-12))
    gamma = np.maximum(_f32(gamma), f32(1e-12))
    with np.errstate(all='ignore'):
        nu = xi * (gamma / (f32(1.0) + xi))
        a = (np.sqrt(f32(np.pi)) / f32(2.0)) * (np.sqrt(nu) / gamma)
        b = a * np.exp(-nu / f32(2.0))
        c = (f32(1.0) + nu) * _bessel_f32(spsp.i0, nu / f32(2.0)) + nu * _bessel_f32(spsp.i1, nu / f32(2.0))
        G = (b * c).astype(f32)
        bad = np.isnan(G) | np.isinf(G)
        return np.where(bad, wf(xi), G).astype(f32)


def mmse_stsa_exact(xi, gamma):
    """Scaled-Bessel float64 form (no overflow); used to bound the f32 formula's error."""
    xi = np.maximum(np.asarray(xi, np.float64), 1e-12)
    gamma = np.maximum(np.asarray(gamma, np.float64), 1e-12)
    nu = xi * gamma / (1.0 + xi)
    return (np.sqrt(np.pi) / 2.0) * (np.sqrt(nu) / gamma) * ((1.0 + nu) * spsp.i0e(nu / 2) + nu * spsp.i1e(nu / 2))


def gfunc(xi, gamma=None, gtype=None, cdm=None):
    """gain.py:168-191."""
    if gtype == 'mmse-lsa': return mmse_lsa(xi, gamma)
    elif gtype == 'mmse-stsa': return mmse_stsa(xi, gamma)
    elif gtype == 'wf': return wf(xi)
    elif gtype == 'srwf': return srwf(xi)
    elif gtype == 'cwf': return cwf(xi)
    elif gtype == 'irm': return irm(xi)
    elif gtype == 'ibm': return ibm(xi)
    elif gtype == 'deepmmse': return deepmmse(xi, gamma)
    elif gtype == 'dgwf': raise NotImplementedError('dgwf needs the STDCT cdm target (out of scope, SURVEY 2)')
    else: raise ValueError('Invalid gain function type.')
