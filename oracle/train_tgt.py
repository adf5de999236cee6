"""Oracle: training-target pipeline (test infrastructure; see oracle/__init__.py).

SURVEY 8(f) row N1.  Restates in numpy float32:
  * InputTarget.add_noise / add_noise_pad / mix        deepxi/sig.py:162-284
  * InputTarget.xi / gamma                             deepxi/sig.py:110-134
  * MagXi.example, MagTgt.transfrom_stats, MagXi.stats deepxi/inp_tgt.py:114-139, :160-196
  * NormalCDF.stats                                    deepxi/map.py:392-402
The reference draws the noise offset with tf.random.uniform (sig.py:277); here the offset is an argument so that
the CUDA path and the oracle can be compared on the same draw.  Parity unpinned: the reference ships no fixture
for these functions (data/*_inp_tgt.p hold only the resulting mu / sigma of its private training set).
"""
import numpy as np

from . import cdfmap, sig

f32 = np.float32


def add_noise(s, d, s_len, d_len, snr_db, offset):
    """add_noise (sig.py:256-284) on already-normalised f32 waveforms; returns (x, d_scaled), each [s_len]."""
    s = np.asarray(s, f32)[:s_len]
    d = np.asarray(d, f32)[:d_len][offset:offset + s_len]
    snr = np.power(f32(10.0), f32(snr_db) / f32(10.0)).astype(f32)
    P_s = np.mean(np.square(s), dtype=np.float64).astype(f32)
    P_d = np.mean(np.square(d), dtype=np.float64).astype(f32)
    alpha = np.sqrt(P_s / np.maximum(P_d * snr, f32(1e-12))).astype(f32)
    d = (d * alpha).astype(f32)
    return (s + d).astype(f32), d


def mix(s_i16, d_i16, s_len, d_len, snr_db, offsets):
    """mix on a padded int16 batch (sig.py:162-187, :214-254): returns s, d, x f32 [B, max(s_len)] and n_frames."""
    B = len(s_len)
    L = int(max(s_len))
    s_o, d_o, x_o = (np.zeros((B, L), f32) for _ in range(3))
    for i in range(B):
        s = sig.normalise(s_i16[i])
        d = sig.normalise(d_i16[i])
        x, dd = add_noise(s, d, int(s_len[i]), int(d_len[i]), snr_db[i], int(offsets[i]))
        s_o[i, :s_len[i]] = s[:s_len[i]]
        d_o[i, :s_len[i]] = dd
        x_o[i, :s_len[i]] = x
    return s_o, d_o, x_o, [sig.n_frames(int(n)) for n in s_len]


def xi(S, D):
    """InputTarget.xi (sig.py:110-121): S^2 / max(D^2, 1e-12)."""
    S, D = np.asarray(S, f32), np.asarray(D, f32)
    return (np.square(S) / np.maximum(np.square(D), f32(1e-12))).astype(f32)


def gamma(X, D):
    """InputTarget.gamma (sig.py:123-134)."""
    return xi(X, D)


def example(s_i16, d_i16, s_len, d_len, snr_db, offsets, mu, sigma):
    """MagXi.example (inp_tgt.py:173-196) for a batch: (x_STMS [B,T,257], xi_bar [B,T,257], n_frames); frames beyond an
    utterance's n_frames are zero in both (the padded waveforms are zero there and the caller masks by n_frames)."""
    s, d, x, nfr = mix(s_i16, d_i16, s_len, d_len, snr_db, offsets)
    T = max(nfr)
    B = len(s_len)
    x_STMS = np.zeros((B, T, 257), f32)
    xi_bar = np.zeros((B, T, 257), f32)
    for i in range(B):
        n = nfr[i]
        S, _ = sig.polar_analysis(s[i, :s_len[i]])
        D, _ = sig.polar_analysis(d[i, :s_len[i]])
        X, _ = sig.polar_analysis(x[i, :s_len[i]])
        x_STMS[i, :n] = X
        xi_bar[i, :n] = cdfmap.normal_cdf_map(xi(S, D), mu, sigma)
    return x_STMS, xi_bar, nfr


def xi_db_moments(s, d, wav_len):
    """Per-bin (count, sum, sum of squares) in float64 of 10 log10 max(xi, 1e-12) over all frames of all utterances
    (transfrom_stats + xi + NormalCDF.stats, inp_tgt.py:114-139, :160-171; map.py:392-402)."""
    acc = np.zeros((3, 257), np.float64)
    for i in range(len(wav_len)):
        S, _ = sig.polar_analysis(np.asarray(s[i], f32)[:wav_len[i]])
        D, _ = sig.polar_analysis(np.asarray(d[i], f32)[:wav_len[i]])
        xdb = cdfmap.db(xi(S, D)).astype(np.float64)
        acc[0] += xdb.shape[0]
        acc[1] += xdb.sum(axis=0)
        acc[2] += np.square(xdb).sum(axis=0)
    return acc


def stats_from_moments(acc):
    """mu = mean, sigma = population standard deviation (tf.math.reduce_std, map.py:401-402)."""
    n = acc[0]
    mu = acc[1] / n
    var = np.maximum(acc[2] / n - mu * mu, 0.0)
    return mu.astype(f32), np.sqrt(var).astype(f32)
