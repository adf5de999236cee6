"""Oracle: the whole inference path on the CPU (test infrastructure; see oracle/__init__.py).

Restates MagXi.enhanced_speech / xi_hat / gamma_hat (deepxi/inp_tgt.py:198-240) and the
DeepXi.infer out_type switch (deepxi/model.py:264-332) for one batch, returning arrays instead
of writing files.  Also the CPU arm that bench.py times ("port" baseline).
"""
import numpy as np
from . import sig, cdfmap, gain, tcn, attention


def xi_hat(xi_bar_hat, mu, sigma):
    """inp_tgt.py:216-227."""
    return cdfmap.normal_cdf_inverse(xi_bar_hat, mu, sigma)


def gamma_hat(xi_bar_hat, mu, sigma):
    """inp_tgt.py:229-240."""
    return (xi_hat(xi_bar_hat, mu, sigma) + np.float32(1.0)).astype(np.float32)


def enhanced_speech(x_STMS, x_STPS, xi_bar_hat, gtype, mu, sigma):
    """inp_tgt.py:198-214."""
    xi = xi_hat(xi_bar_hat, mu, sigma)
    gam = (xi + np.float32(1.0)).astype(np.float32)
    y_STMS = (np.asarray(x_STMS, np.float32) * gain.gfunc(xi, gam, gtype)).astype(np.float32)
    return sig.polar_synthesis(y_STMS, x_STPS)


def infer(x_batch, x_len, weights, mu, sigma, network='ResNetV2', padding='causal', out_type='y',
          gtype='mmse-lsa', mask_mode='none'):
    """model.py:224-332 for one batch: list of per-utterance outputs."""
    inp, sup, nfr = sig.observation_batch(x_batch, x_len)
    if network == 'ResNetV2':
        xbar = tcn.resnetv2_forward(inp, weights, padding=padding)
    elif network == 'MHANetV3':
        xbar = attention.mhanetv3_forward(inp, weights, mask_mode=mask_mode)
    else:
        raise ValueError('Invalid network type.')
    outs = []
    for i in range(len(x_len)):
        m, p, t = inp[i, :nfr[i]], sup[i, :nfr[i]], xbar[i, :nfr[i]]
        if out_type == 'y':
            outs.append(enhanced_speech(m, p, t, gtype, mu, sigma))
        elif out_type == 'xi_hat':
            outs.append(xi_hat(t, mu, sigma))
        elif out_type == 'gamma_hat':
            outs.append(gamma_hat(t, mu, sigma))
        elif out_type == 'gain':
            xi = xi_hat(t, mu, sigma)
            outs.append(gain.gfunc(xi, xi + np.float32(1.0), gtype))
        elif out_type == 'ibm_hat':
            outs.append(np.greater(xi_hat(t, mu, sigma), 1.0))
        elif out_type == 'deepmmse':
            xi = xi_hat(t, mu, sigma)
            outs.append(np.square(m) * gain.gfunc(xi, xi + np.float32(1.0), 'deepmmse'))
        else:
            raise ValueError('Invalid output type.')
    return outs
