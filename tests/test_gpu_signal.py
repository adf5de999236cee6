"""GPU parity tests (through the C ABI) of analysis / synthesis / map / gain against the oracle."""
import os

import numpy as np
import pytest
import torch

from oracle import sig as osig, gain as ogain, cdfmap, wavio, pipeline
from deepxi_b200 import synth, _lib
from deepxi_b200.sig import AnalysisSynthesis
from deepxi_b200.inp_tgt import inp_tgt_selector
from deepxi_b200 import gain as dgain

pytestmark = pytest.mark.gpu
GTYPES = ['mmse-lsa', 'mmse-stsa', 'wf', 'srwf', 'cwf', 'irm', 'ibm', 'deepmmse']


def _magxi(xi_stats, ver='resnet-1.1c'):
    it = inp_tgt_selector('MagXi', 512, 256, 512, 16000, map_type='DBNormalCDF', map_params=None)
    return it.set_stats(xi_stats[ver + '/mu'], xi_stats[ver + '/sigma'])


def _phase_err(p, p_ref, m_ref):
    d = np.angle(np.exp(1j * (p.astype(np.float64) - p_ref)))
    return (np.abs(d) * m_ref).max(axis=-1) / np.maximum(m_ref.max(axis=-1), 1e-30)


def test_library_loaded_and_device_ok():
    lib = _lib.load()
    assert lib.dxi_device_check() == 0, lib.dxi_last_error()


@pytest.mark.parametrize('L', [1, 255, 256, 257, 511, 512, 513, 4000, 39088, 64000])
def test_stft_float_matches_oracle(L):
    rng = np.random.default_rng(L)
    x = (rng.standard_normal(L) * 0.1).astype(np.float32)
    a = AnalysisSynthesis(512, 256, 512, 16000)
    mag, pha = a.polar_analysis(x)
    m_ref, p_ref = osig.polar_analysis(x)
    assert mag.shape == m_ref.shape == (-(-L // 256), 257)
    # north-star tolerance: 1e-5 relative, |delta| <= 1e-5 * max_k |X[l,k]| per frame
    assert (np.abs(mag - m_ref).max(axis=-1) <= 1e-5 * m_ref.max(axis=-1)).all()
    assert (_phase_err(pha, p_ref, m_ref) <= 1e-5).all()


def test_stft_int16_batch_ragged(xi_stats):
    it = _magxi(xi_stats)
    lens = [5000, 1, 4096, 3333, 256, 2049]
    x = synth.noisy_speech(len(lens), 5000, seed=11)
    inp, pha, nfr = it.observation_batch(x, lens)
    r_inp, r_pha, r_nfr = osig.observation_batch(x, lens)
    assert nfr == r_nfr and tuple(inp.shape) == r_inp.shape
    inp, pha = inp.cpu().numpy(), pha.cpu().numpy()
    assert (np.abs(inp - r_inp).max(axis=-1) <= 1e-5 * np.maximum(r_inp.max(axis=-1), 1e-30)).all()
    assert (_phase_err(pha, r_pha, r_inp) <= 1e-5).all()
    for i, n in enumerate(nfr):                       # frames beyond n_frames are zero padded (model.py:2246-2253)
        assert not inp[i, n:].any() and not pha[i, n:].any()
    m1, p1 = it.observation(x[0])                     # single-utterance API (inp_tgt.py:87-101)
    assert np.array_equal(m1, inp[0]) and np.array_equal(p1, pha[0])


def test_stft_unaligned_and_float_strides():
    a = AnalysisSynthesis(512, 256, 512, 16000)
    rng = np.random.default_rng(5)
    x = (rng.standard_normal((3, 1001)) * 0.05).astype(np.float32)      # row stride not a multiple of 16 bytes
    mag, _ = a.polar_analysis(x)
    m_ref, _ = osig.polar_analysis(x)
    assert np.abs(mag - m_ref).max() <= 1e-5 * m_ref.max()


def test_analysis_synthesis_round_trip_and_linearity():
    a = AnalysisSynthesis(512, 256, 512, 16000)
    x = osig.normalise(synth.noisy_speech(4, 16000, seed=12))
    mag, pha = a.polar_analysis(x)
    y = a.polar_synthesis(mag, pha)
    assert y.shape == (4, (63 + 1) * 256)
    assert np.abs(y[:, 256:16000] - x[:, 256:]).max() < 1e-6
    y_ref = osig.polar_synthesis(*osig.polar_analysis(x))
    assert np.abs(y - y_ref).max() < 1e-6
    m2, _ = a.polar_analysis(2.0 * x)                # |STFT| is homogeneous
    assert np.abs(m2 - 2.0 * mag).max() <= 2e-6 * mag.max()


@pytest.mark.parametrize('gtype', GTYPES)
def test_gfunc_matches_oracle(gtype):
    rng = np.random.default_rng(21)
    xi = (10.0 ** rng.uniform(-12, 4, (97, 257))).astype(np.float32)
    gam = (xi * rng.uniform(0.5, 2.0, xi.shape) + 1).astype(np.float32)
    G = dgain.gfunc(xi, gam, gtype)
    ref = ogain.gfunc(xi, gam, gtype)
    if gtype == 'ibm':
        assert np.array_equal(G, ref)
    elif gtype == 'mmse-stsa':
        nu = xi.astype(np.float64) / (1 + xi) * gam
        edge = (nu > 168) & (nu < 182)                # the f32 overflow -> Wiener switch (-0.14 % step) sits here
        assert np.allclose(G[~edge], ref[~edge], rtol=1e-5)
        assert np.allclose(G[edge], ref[edge], rtol=2e-3)
    else:
        assert np.allclose(G, ref, rtol=1e-5, atol=1e-37)
    with pytest.raises(ValueError, match='Invalid gain function type.'):
        dgain.gfunc(xi, gam, 'bogus')


def test_inverse_map_xi_hat_and_ibm_bit_exact(xi_stats):
    it = _magxi(xi_stats)
    mu, sg = xi_stats['resnet-1.1c/mu'], xi_stats['resnet-1.1c/sigma']
    rng = np.random.default_rng(22)
    xb = rng.uniform(1e-6, 1 - 1e-6, (300, 257)).astype(np.float32)
    thr = cdfmap.ibm_threshold(mu, sg).astype(np.float32)
    rows = [thr]
    for k in (1, 2):
        up, dn = thr.copy(), thr.copy()
        for _ in range(k):
            up, dn = np.nextafter(up, np.float32(2)), np.nextafter(dn, np.float32(-1))
        rows += [up, dn]
    xb = np.vstack([xb] + rows + [np.full(257, 2.0 ** -24, np.float32), np.full(257, 1 - 2.0 ** -24, np.float32)]).astype(np.float32)
    xi = it.xi_hat(xb)
    ref = cdfmap.normal_cdf_inverse(xb, mu, sg)
    assert np.allclose(xi, ref, rtol=3e-5)
    db_err = np.abs(10 * np.log10(xi.astype(np.float64)) - 10 * np.log10(ref.astype(np.float64)))
    assert db_err.max() < 2e-4                         # xi_hat within 0.1 dB: five hundred times tighter here
    assert np.array_equal(it.ibm_hat(xb), ref > 1.0)   # IBM masks bit-exact, incl. threshold +-{0,1,2} ulp
    assert np.array_equal(it.gamma_hat(xb), (ref + np.float32(1.0)).astype(np.float32)) or \
        np.allclose(it.gamma_hat(xb), ref + 1, rtol=3e-5)
    back = it.xi_map.map(ref)
    assert np.abs(back - cdfmap.normal_cdf_map(ref, mu, sg)).max() < 1e-6
    for g in GTYPES:
        G = it.gain_hat(xb, g)
        Gr = ogain.gfunc(ref, ref + np.float32(1.0), g)
        if g == 'ibm':
            assert np.array_equal(G, Gr)
        else:
            ok = np.isfinite(Gr)
            assert np.allclose(G[ok], Gr[ok], rtol=2e-3 if g == 'mmse-stsa' else 5e-5, atol=1e-30)


def test_kat_through_the_c_abi(golden_dir, xi_stats):
    """The reference's own known-answer test through the CUDA path: <= 1 LSB."""
    it = _magxi(xi_stats)
    x, _ = wavio.read_wav_int16(os.path.join(golden_dir, 'kat_noisy.wav'))
    y_ref, _ = wavio.read_wav_int16(os.path.join(golden_dir, 'kat_y_mmse-lsa_resnet-1.0c_e180.wav'))
    xi = np.load(os.path.join(golden_dir, 'kat_xi_hat_resnet-1.0c_e180.npy'))
    mag, pha = it.observation(x)
    G = dgain.gfunc(xi, xi + np.float32(1.0), 'mmse-lsa')
    lib = _lib.load()
    dm, dp, dg = (torch.from_numpy(np.ascontiguousarray(a)).cuda() for a in (mag, pha, G))
    out = torch.empty((1, 154 * 256), dtype=torch.int16, device='cuda')
    _lib.check(lib.dxi_istft(_lib.ptr(dm), _lib.ptr(dg), _lib.ptr(dp), None, 1, 153, None, _lib.ptr(out), 154 * 256,
                             _lib.stream_ptr()))
    y = out.cpu().numpy()[0]
    d = np.abs(y.astype(np.int32) - y_ref.astype(np.int32))
    assert len(y) == len(y_ref) and d.max() <= 1 and (d != 0).sum() < 600


@pytest.mark.parametrize('gtype', ['mmse-lsa', 'mmse-stsa', 'srwf', 'cwf', 'irm', 'ibm'])
def test_enhanced_speech_fused_vs_oracle(xi_stats, gtype):
    it = _magxi(xi_stats)
    mu, sg = xi_stats['resnet-1.1c/mu'], xi_stats['resnet-1.1c/sigma']
    lens = [6000, 2500, 4097]
    x = synth.noisy_speech(3, 6000, seed=13)
    inp, pha, nfr = it.observation_batch(x, lens)
    rng = np.random.default_rng(14)
    xb = rng.uniform(0.01, 0.99, tuple(inp.shape)).astype(np.float32)
    y = it.enhanced_speech(inp, pha, torch.from_numpy(xb).cuda(), gtype, n_frames=nfr).cpu().numpy()
    yi = it.enhanced_speech(inp, pha, torch.from_numpy(xb).cuda(), gtype, n_frames=nfr, int16=True).cpu().numpy()
    inp_h, pha_h = inp.cpu().numpy(), pha.cpu().numpy()
    for i, n in enumerate(nfr):
        ref = pipeline.enhanced_speech(inp_h[i, :n], pha_h[i, :n], xb[i, :n], gtype, mu, sg)
        got = y[i, :(n + 1) * 256]
        snr = 10 * np.log10(np.sum(ref.astype(np.float64) ** 2) / max(np.sum((got - ref).astype(np.float64) ** 2), 1e-30))
        assert snr > 90.0, (gtype, i, snr)               # north star: >= 40 dB
        assert not y[i, (n + 1) * 256:].any()           # beyond the utterance: silence
        di = np.abs(yi[i, :(n + 1) * 256].astype(np.int32) - wavio.float_to_int16(ref).astype(np.int32))
        assert di.max() <= 1


@pytest.mark.parametrize('int16', [False, True])
def test_enhance_output_rows_unaligned(xi_stats, int16):
    """dxi_enhance with an output row pitch that is not a multiple of four samples (and an f32 base 4 bytes off a 16-byte boundary): the
    kernel then stores one sample at a time instead of four; the samples must be the bits of the aligned call."""
    mu, sg = xi_stats['resnet-1.1c/mu'], xi_stats['resnet-1.1c/sigma']
    it = _magxi(xi_stats)
    lens = [30000, 7777, 256]
    mag, pha, nfr = it.observation_batch(synth.noisy_speech(3, 30000, seed=81), lens)
    B, T = mag.shape[0], mag.shape[1]
    xb = torch.rand((B, T, 257), device='cuda') * 0.98 + 0.01
    ref = it.enhanced_speech(mag, pha, xb, 'mmse-lsa', n_frames=nfr, int16=int16)
    n_out = (T + 1) * 256
    pitch = n_out + 3
    dt = torch.int16 if int16 else torch.float32
    buf = torch.zeros(B * pitch + 1, dtype=dt, device='cuda')
    out = buf[1:]                                   # one element off: 2 / 4 bytes past the allocation's alignment
    nf = torch.as_tensor(np.asarray(nfr, np.int32)).cuda()
    mu_d, sg_d = torch.from_numpy(np.asarray(mu, np.float32)).cuda(), torch.from_numpy(np.asarray(sg, np.float32)).cuda()
    lib = _lib.load()
    _lib.check(lib.dxi_enhance(_lib.ptr(mag), _lib.ptr(pha), _lib.ptr(xb), _lib.ptr(mu_d), _lib.ptr(sg_d), 0, _lib.ptr(nf, torch.int32), B, T,
                               None if int16 else out.data_ptr(), out.data_ptr() if int16 else None, pitch, _lib.stream_ptr(mag.device)))
    torch.cuda.synchronize()
    got = out[:B * pitch].view(B, pitch)[:, :n_out]
    assert torch.equal(got, ref)


def test_subband_ibm_matches_oracle(xi_stats):
    """xi_hat -> mel-subband a priori SNR -> mask (deepxi/model.py:323-328, sig.py:301-346)."""
    it = inp_tgt_selector('MagXi', 512, 256, 512, 16000, map_type='DBNormalCDF', map_params=None)
    rng = np.random.default_rng(11)
    xi = (10.0 ** rng.uniform(-3, 3, (3, 77, 257))).astype(np.float32)
    sub, ibm = it.subband(xi, 40)
    r_sub, r_ibm = osig.subband_ibm(xi, 40)
    assert sub.shape == r_sub.shape == (3, 77, 40) and ibm.dtype == np.bool_
    assert np.allclose(sub, r_sub, rtol=2e-6)                     # different summation order of the 257-term dot products
    sure = np.abs(r_sub - 1.0) > 1e-5                              # the mask can only differ where the sum sits on the threshold
    assert np.array_equal(ibm[sure], r_ibm[sure])
    sub24, _ = it.subband(xi[0], 24, want_mask=False)
    assert np.allclose(sub24, osig.subband_ibm(xi[0], 24)[0], rtol=2e-6)


def test_degenerate_inputs(xi_stats):
    """Empty and zero-length inputs (SURVEY 8c edge cases): an empty waveform has no frames, a zero-length utterance inside a
    batch is all padding, element-wise functions keep empty shapes."""
    a = AnalysisSynthesis(512, 256, 512, 16000)
    mag, pha = a.polar_analysis(np.zeros(0, np.float32))
    m_ref, p_ref = osig.polar_analysis(np.zeros(0, np.float32))
    assert mag.shape == pha.shape == m_ref.shape == (0, 257)
    it = _magxi(xi_stats)
    x = synth.noisy_speech(2, 600, seed=3)
    lens = [600, 0]
    inp, phs, nfr = it.observation_batch(x, lens)
    r_inp, r_pha, r_nfr = osig.observation_batch(x, lens)
    assert nfr == r_nfr == [3, 0] and tuple(inp.shape) == r_inp.shape == (2, 3, 257)
    inp_h, phs_h = inp.cpu().numpy(), phs.cpu().numpy()
    assert not inp_h[1].any() and not phs_h[1].any()
    assert (np.abs(inp_h[0] - r_inp[0]).max(axis=-1) <= 1e-5 * r_inp[0].max(axis=-1)).all()
    xbar = torch.full((2, 3, 257), 0.5, device='cuda')
    y = it.enhanced_speech(inp, phs, xbar, 'mmse-lsa', n_frames=nfr).cpu().numpy()
    assert y.shape == (2, 4 * 256) and np.isfinite(y).all() and not y[1].any() and np.abs(y[0]).max() > 0
    for g in ('mmse-lsa', 'ibm'):
        G = dgain.gfunc(np.zeros((0, 257), np.float32), np.zeros((0, 257), np.float32), g)
        assert G.shape == (0, 257)
    assert it.xi_hat(np.zeros((0, 257), np.float32)).shape == (0, 257)
