"""Oracle: analysis / synthesis (test infrastructure; see oracle/__init__.py).

Restates deepxi/sig.py:26-69 (AnalysisSynthesis), :189-212 (normalise, n_frames) and the
third-party semantics they delegate to (tf.signal.stft with pad_end=True, tf.signal.inverse_stft
with tf.signal.inverse_stft_window_fn, window_ops.hamming_window(periodic=False)).
"""
import numpy as np


def hamming(N_d, dtype=np.float32):
    """window_ops.hamming_window(N_d, periodic=False) (deepxi/sig.py:38-39)."""
    n = np.arange(N_d, dtype=np.float64)
    w = 0.54 - 0.46 * np.cos(2.0 * np.pi * n / (N_d - 1))
    return w.astype(dtype)


def synthesis_window(N_d, N_s, dtype=np.float32):
    """tf.signal.inverse_stft_window_fn(N_s, hamming) as used at deepxi/sig.py:68-69.

    sw[n] = w[n] / sum_m w[(n mod N_s) + m*N_s]^2   (for N_d = 2*N_s: two terms).
    """
    w = hamming(N_d, np.float64)
    den = np.zeros(N_s, np.float64)
    for m in range(0, N_d, N_s):
        den += w[m:m + N_s] ** 2
    den = np.tile(den, N_d // N_s)
    return (w / den).astype(dtype)


def normalise(x):
    """InputTarget.normalise (deepxi/sig.py:189-199): int16 -> f32 / 32768."""
    return np.asarray(x).astype(np.float32) / np.float32(32768.0)


def n_frames(N, N_s=256):
    """InputTarget.n_frames (deepxi/sig.py:201-212): ceil(N / N_s)."""
    return int(np.ceil(np.float32(N) / np.float32(N_s)))


def frame(x, N_d, N_s):
    """tf.signal.frame(pad_end=True): T = ceil(L/N_s) frames, zero padded at the end."""
    x = np.asarray(x)
    L = x.shape[-1]
    T = -(-L // N_s)
    Lp = (T - 1) * N_s + N_d if T > 0 else 0
    pad = [(0, 0)] * (x.ndim - 1) + [(0, max(Lp - L, 0))]
    xp = np.pad(x, pad)
    idx = np.arange(N_d)[None, :] + N_s * np.arange(T)[:, None]
    return xp[..., idx]


def polar_analysis(x, N_d=512, N_s=256, K=512, dtype=np.float32):
    """AnalysisSynthesis.polar_analysis (deepxi/sig.py:43-55).

    x: f32 [L] or [B, L].  Returns (|STFT|, angle(STFT)), each [..., T, K/2+1].
    """
    x = np.asarray(x, dtype=dtype)
    fr = frame(x, N_d, N_s) * hamming(N_d, dtype)
    X = np.fft.rfft(fr, K, axis=-1)
    if dtype == np.float32:
        X = X.astype(np.complex64)
    return np.abs(X).astype(dtype), np.angle(X).astype(dtype)


def polar_synthesis(STMS, STPS, N_d=512, N_s=256, K=512, dtype=np.float32):
    """AnalysisSynthesis.polar_synthesis (deepxi/sig.py:57-69).

    Y = |Y| e^{j phase}; irfft_K; * synthesis window; overlap-add with hop N_s.
    Output length (T-1)*N_s + N_d (not trimmed to the input length).
    """
    STMS = np.asarray(STMS, dtype=dtype)
    STPS = np.asarray(STPS, dtype=dtype)
    ctype = np.complex64 if dtype == np.float32 else np.complex128
    Y = STMS.astype(ctype) * np.exp(1j * STPS.astype(ctype)).astype(ctype)
    fr = np.fft.irfft(Y, K, axis=-1)[..., :N_d].astype(dtype)
    fr = fr * synthesis_window(N_d, N_s, dtype)
    T = fr.shape[-2]
    out = np.zeros(fr.shape[:-2] + ((T - 1) * N_s + N_d,), dtype)
    for t in range(T):
        out[..., t * N_s:t * N_s + N_d] += fr[..., t, :]
    return out


def observation(x_int16, N_d=512, N_s=256, K=512):
    """MagTgt.observation (deepxi/inp_tgt.py:87-101): normalise then polar_analysis."""
    return polar_analysis(normalise(x_int16), N_d, N_s, K)


def observation_batch(x_batch, x_batch_len, N_d=512, N_s=256, K=512):
    """DeepXi.observation_batch (deepxi/model.py:2232-2254): zero-padded [B, Tmax, 257] batches."""
    B = len(x_batch_len)
    Tmax = n_frames(max(x_batch_len), N_s)
    nb = K // 2 + 1
    inp = np.zeros([B, Tmax, nb], np.float32)
    sup = np.zeros([B, Tmax, nb], np.float32)
    nfr = [n_frames(int(l), N_s) for l in x_batch_len]
    for i in range(B):
        m, p = observation(x_batch[i][:x_batch_len[i]], N_d, N_s, K)
        inp[i, :nfr[i]] = m
        sup[i, :nfr[i]] = p
    return inp, sup, nfr


def mel_filter_bank(M, K=512, f_s=16000):
    """InputTarget.mel_filter_bank / bpoint / hz_to_mel / mel_to_hz (deepxi/sig.py:301-370): triangular filters that sum to
    unity, [M, K/2+1] float32."""
    hz_to_mel = lambda f: 2595 * np.log10(1 + (f / 700))
    mel_to_hz = lambda m: 700 * ((10 ** (m / 2595)) - 1)
    nb = K // 2 + 1
    f_l, f_h = 0, f_s / 2
    bpoint = lambda m: ((2 * nb) / f_s) * mel_to_hz(hz_to_mel(f_l) + m * ((hz_to_mel(f_h) - hz_to_mel(f_l)) / (M + 1)))
    H = np.zeros([M, nb], dtype=np.float32)
    for m in range(1, M + 1):
        bl, c, bh = bpoint(m - 1), bpoint(m), bpoint(m + 1)
        for k in range(nb):
            if bl <= k <= c:
                H[m - 1, k] = (2 * (k - bl)) / ((bh - bl) * (c - bl))
            if c <= k <= bh:
                H[m - 1, k] = (2 * (bh - k)) / ((bh - bl) * (bh - c))
    return H


def subband_ibm(xi_hat, M=40, K=512, f_s=16000):
    """(xi_hat H^T, its mask > 1): deepxi/model.py:323-328."""
    sub = np.matmul(np.asarray(xi_hat, np.float32), mel_filter_bank(M, K, f_s).transpose())
    return sub, np.greater(sub, 1.0)
