"""The device math headers (fft.cuh, gain_math.cuh) compiled for the HOST with g++ through
tests/host/kernels_host.cpp and checked against the oracle: FFT pass indexing, split / merge steps,
atan2 polynomial, E1, gain formulas and the inverse-map chain are exactly the code the kernels run."""
import ctypes
import os
import subprocess

import numpy as np
import pytest
from scipy import special as spsp

from oracle import sig, gain, cdfmap

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
fp = ctypes.POINTER(ctypes.c_float)


@pytest.fixture(scope='module')
def hostlib(tmp_path_factory):
    out = str(tmp_path_factory.mktemp('hostlib') / 'libkernels_host.so')
    src = os.path.join(ROOT, 'tests', 'host', 'kernels_host.cpp')
    subprocess.run(['g++', '-std=c++17', '-O1', '-ffp-contract=off', '-shared', '-fPIC', '-w',
                    '-I/usr/local/cuda/include', src, '-o', out], check=True)
    lib = ctypes.CDLL(out)
    cb = ctypes.CFUNCTYPE(ctypes.c_double, ctypes.c_double)(lambda x: float(spsp.erfinv(x)))
    lib.host_set_erfinv(cb)
    lib._keep = cb
    lib.host_e1.restype = ctypes.c_float
    lib.host_e1.argtypes = [ctypes.c_float]
    return lib


def P(a):
    return a.ctypes.data_as(fp)


def test_stft_frame_math(hostlib):
    rng = np.random.default_rng(0)
    for scale in (0.1, 1e-3):
        x = (rng.standard_normal(512) * scale).astype(np.float32)
        mag, pha = np.zeros(257, np.float32), np.zeros(257, np.float32)
        hostlib.host_stft_frame(P(x), P(mag), P(pha))
        m_ref, p_ref = sig.polar_analysis(x)
        assert np.abs(mag - m_ref[0]).max() <= 1e-5 * m_ref[0].max()          # north-star tolerance
        d = np.angle(np.exp(1j * (pha.astype(np.float64) - p_ref[0])))
        assert (np.abs(d) * m_ref[0]).max() <= 1e-5 * m_ref[0].max()
    z = np.zeros(512, np.float32)
    hostlib.host_stft_frame(P(z), P(mag), P(pha))
    assert not mag.any() and not pha.any()


def test_istft_frame_math(hostlib):
    rng = np.random.default_rng(1)
    x = (rng.standard_normal(512) * 0.1).astype(np.float32)
    m, p = sig.polar_analysis(x)
    out = np.zeros(512, np.float32)
    hostlib.host_istft_frame(P(np.ascontiguousarray(m[0])), P(np.ascontiguousarray(p[0])), P(out))
    ref = np.fft.irfft(m[0] * np.exp(1j * p[0]), 512).astype(np.float32) * sig.synthesis_window(512, 256)
    assert np.abs(out - ref).max() < 2e-7 * max(1.0, np.abs(ref).max() / 0.1)


def test_expint_e1(hostlib):
    x = (10.0 ** np.linspace(-12, 3, 3000)).astype(np.float32)
    e = np.array([hostlib.host_e1(float(v)) for v in x])
    assert np.abs(e - spsp.exp1(x.astype(np.float64))).max() < 3e-6


@pytest.mark.parametrize('gtype,code', [('mmse-lsa', 0), ('mmse-stsa', 1), ('wf', 2), ('srwf', 3), ('cwf', 4), ('irm', 5),
                                        ('ibm', 6), ('deepmmse', 7)])
def test_gain_functions(hostlib, gtype, code):
    rng = np.random.default_rng(2)
    xi = (10.0 ** rng.uniform(-12, 4, 20000)).astype(np.float32)
    gam = (xi + np.float32(1.0)).astype(np.float32)
    G = np.zeros_like(xi)
    hostlib.host_gfunc(P(xi), P(gam), len(xi), code, P(G))
    ref = gain.gfunc(xi, gam, gtype)
    if gtype == 'ibm':
        assert np.array_equal(G, ref)
    elif gtype == 'mmse-stsa':
        nu = xi.astype(np.float64)            # gamma = xi + 1  =>  nu = xi
        edge = (nu > 170) & (nu < 180)        # the f32 overflow -> Wiener switch sits here (-0.14 % step)
        assert np.allclose(G[~edge], ref[~edge], rtol=2e-6)
        assert np.allclose(G[edge], ref[edge], rtol=2e-3)
    else:
        assert np.allclose(G, ref, rtol=3e-6, atol=1e-37)


def test_inverse_map_chain_and_ibm_exactness(hostlib, xi_stats):
    mu, sg = xi_stats['resnet-1.1c/mu'], xi_stats['resnet-1.1c/sigma']
    rng = np.random.default_rng(3)
    xb = rng.uniform(1e-6, 1 - 1e-6, (200, 257)).astype(np.float32)
    # the labelled edge set: per-bin IBM threshold +- {0,1,2} ulp, and the saturating ends
    thr = cdfmap.ibm_threshold(mu, sg).astype(np.float32)
    rows = [thr]
    for k in (1, 2):
        up, dn = thr.copy(), thr.copy()
        for _ in range(k):
            up, dn = np.nextafter(up, np.float32(2)), np.nextafter(dn, np.float32(-1))
        rows += [up, dn]
    xb = np.vstack([xb] + rows + [np.full(257, 2.0 ** -24, np.float32), np.full(257, 1 - 2.0 ** -24, np.float32),
                                  np.full(257, 0.5, np.float32)]).astype(np.float32)
    xi = np.zeros_like(xb)
    hostlib.host_xi_from_xbar(P(xb), P(mu), P(sg), xb.shape[0], 257, P(xi))
    ref = cdfmap.normal_cdf_inverse(xb, mu, sg)
    assert np.allclose(xi, ref, rtol=3e-5)
    assert np.array_equal(xi > 1.0, ref > 1.0)            # IBM bit-exact, including the +-ulp edge rows
    back = np.zeros_like(xb)
    hostlib.host_xbar_from_xi(P(ref), P(mu), P(sg), xb.shape[0], 257, P(back))
    assert np.abs(back - cdfmap.normal_cdf_map(ref, mu, sg)).max() < 3e-7


def test_fast_lsa_path_math(hostlib, xi_stats):
    """The MUFU-style restatement behind the fused enhancement kernel (gain_math.cuh: erfinv_fast, expint_e1_fast,
    lsa_gain_from_xbar_fast) against scipy and the oracle chain x_bar -> xi_hat -> G_LSA(xi_hat, xi_hat + 1)."""
    hostlib.host_e1_fast.restype = ctypes.c_float
    hostlib.host_e1_fast.argtypes = [ctypes.c_float]
    hostlib.host_erfinv_fast.restype = ctypes.c_float
    hostlib.host_erfinv_fast.argtypes = [ctypes.c_float]
    x = (10.0 ** np.linspace(-12, 3, 4000)).astype(np.float32)
    e = np.array([hostlib.host_e1_fast(float(v)) for v in x], dtype=np.float64)
    ref = spsp.exp1(x.astype(np.float64))
    sel = (x > 1e-3) & (x < 10)      # beyond 10, E1 < 5e-6: only the absolute error matters for exp(E1 / 2)
    assert np.abs(e - ref).max() < 3e-6 and (np.abs(e - ref) <= 1e-6 * ref)[sel].all()
    a = np.concatenate([np.linspace(-1, 1, 20001)[1:-1], 1 - 2.0 ** -np.arange(2, 25), -1 + 2.0 ** -np.arange(2, 25)]).astype(np.float32)
    u = np.array([hostlib.host_erfinv_fast(float(v)) for v in a], dtype=np.float64)
    uref = spsp.erfinv(a.astype(np.float64))
    assert np.abs(u - uref).max() < 3e-6
    mu, sg = xi_stats['resnet-1.1c/mu'], xi_stats['resnet-1.1c/sigma']
    rng = np.random.default_rng(5)
    xb = np.vstack([rng.uniform(1e-6, 1 - 1e-6, (300, 257)), rng.beta(0.3, 0.3, (300, 257)),
                    np.full((1, 257), 2.0 ** -24), np.full((1, 257), 1 - 2.0 ** -24), np.zeros((1, 257))]).astype(np.float32)
    G = np.zeros_like(xb)
    hostlib.host_lsa_from_xbar_fast(P(xb), P(mu), P(sg), xb.shape[0], 257, P(G))
    xi = cdfmap.normal_cdf_inverse(xb, mu, sg)
    Gref = gain.gfunc(xi, (xi + np.float32(1.0)).astype(np.float32), 'mmse-lsa')
    ok = np.isfinite(Gref)
    assert ok.mean() > 0.99
    assert np.allclose(G[ok], Gref[ok], rtol=2e-5, atol=1e-30)      # > 90 dB on the waveform
    # the table form the kernel runs (piecewise cubics of log2 G over xi_hat in dB): the same bound
    Gt = np.zeros_like(xb)
    hostlib.host_lsa_from_xbar_tab(P(xb), P(mu), P(sg), xb.shape[0], 257, P(Gt))
    assert np.allclose(Gt[ok], Gref[ok], rtol=2e-5, atol=1e-30)
    assert np.abs(Gt[ok] / np.maximum(G[ok], 1e-30) - 1).max() < 1e-5
