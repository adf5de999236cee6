"""Mirror of deepxi/sig.py (AnalysisSynthesis :23-69, InputTarget.normalise :189-199, n_frames :201-212).

Same class / method names and argument meaning as the reference; the arithmetic runs in the fused
CUDA kernels behind dxi_stft / dxi_istft (framing + Hamming window + 512-point real FFT + |.| / angle,
and the inverse with synthesis window + overlap-add).  numpy in -> numpy out, torch in -> CUDA torch out.
"""
import math

import torch

from . import _lib
from ._tensor import to_dev, ret, device


class AnalysisSynthesis:
    """Analysis and synthesis stages of speech enhancement (sig.py:23-41)."""

    def __init__(self, N_d, N_s, K, f_s):
        if (N_d, N_s, K) != (512, 256, 512):
            raise ValueError('the CUDA STFT kernels are specialised for N_d=512, N_s=256, K=512 '
                             '(main.py:33-35 with T_d=32 ms, T_s=16 ms at 16 kHz)')
        self.N_d, self.N_s, self.K, self.f_s = N_d, N_s, K, f_s
        self.n_bins = K // 2 + 1

    # -- analysis -----------------------------------------------------------------------------
    def _stft(self, x, x_len=None):
        lib = _lib.load()
        is_i16 = x.dtype == torch.int16
        squeeze = x.dim() == 1
        if squeeze:
            x = x[None]
        if x.dim() != 2:
            raise ValueError('Waveforms are of incorrect rank.')
        B, L = x.shape
        if x_len is None:
            T = -(-L // self.N_s)
            lens = None
        else:
            lens = x_len
            T = -(-int(lens.max().item()) // self.N_s) if B else 0
        mag = torch.empty((B, T, self.n_bins), dtype=torch.float32, device=x.device)
        pha = torch.empty_like(mag)
        if B and T:
            _lib.check(lib.dxi_stft(_lib.ptr(x), int(is_i16), _lib.ptr(lens, torch.int32, allow_none=True), B, L, T,
                                    _lib.ptr(mag), _lib.ptr(pha), _lib.stream_ptr(x.device)))
        return (mag[0], pha[0]) if squeeze else (mag, pha)

    def polar_analysis(self, x):
        """Short-time magnitude and phase spectra of a float waveform [L] or [B, L] (sig.py:43-55)."""
        x, was_np = to_dev(x, torch.float32)
        mag, pha = self._stft(x)
        return ret(mag, was_np), ret(pha, was_np)

    # -- synthesis ----------------------------------------------------------------------------
    def polar_synthesis(self, STMS, STPS):
        """Waveform of length (T-1)*N_s + N_d from magnitude / phase spectra (sig.py:57-69)."""
        lib = _lib.load()
        STMS, was_np = to_dev(STMS, torch.float32)
        STPS, _ = to_dev(STPS, torch.float32)
        if STMS.shape != STPS.shape or STMS.shape[-1] != self.n_bins:
            raise ValueError('STMS / STPS must both be [..., T, %d]' % self.n_bins)
        squeeze = STMS.dim() == 2
        if squeeze:
            STMS, STPS = STMS[None], STPS[None]
        B, T, _ = STMS.shape
        n_out = (T + 1) * self.N_s
        y = torch.empty((B, n_out), dtype=torch.float32, device=STMS.device)
        if B and T:
            _lib.check(lib.dxi_istft(_lib.ptr(STMS), None, _lib.ptr(STPS), None, B, T, _lib.ptr(y), None, n_out,
                                     _lib.stream_ptr(STMS.device)))
        return ret(y[0] if squeeze else y, was_np)

    def stdct_analysis(self, x):
        raise NotImplementedError('STDCT (deepxi/dct.py) serves only the unused STDCTXiCD target: out of scope')

    def stdct_synthesis(self, STDCT):
        raise NotImplementedError('STDCT (deepxi/dct.py) serves only the unused STDCTXiCD target: out of scope')


class InputTarget(AnalysisSynthesis):
    """Computes the input and target of Deep Xi (sig.py:96-108); inference-side methods only."""

    def normalise(self, x):
        """int16 waveform -> float32 in [-1, 1) (sig.py:189-199).  (observation() fuses this into the STFT.)"""
        x, was_np = to_dev(x, torch.int16) if not (isinstance(x, torch.Tensor) and x.dtype == torch.int32) else to_dev(x, torch.int32)
        return ret(x.to(torch.float32) / 32768.0, was_np)

    def n_frames(self, N):
        """ceil(N / N_s) (sig.py:201-212)."""
        return int(math.ceil(float(N) / float(self.N_s)))

    # -- mel filter bank for the subband IBM output ---------------------------------
    def hz_to_mel(self, f):
        """sig.py:348-358."""
        import numpy as np
        return 2595 * np.log10(1 + (f / 700))

    def mel_to_hz(self, m):
        """sig.py:360-370."""
        return 700 * ((10 ** (m / 2595)) - 1)

    def bpoint(self, m, M, f_l, f_h):
        """Frequency-bin boundary point of filter m (sig.py:332-346)."""
        K = self.K // 2 + 1
        return ((2 * K) / self.f_s) * self.mel_to_hz(self.hz_to_mel(f_l) + m * ((self.hz_to_mel(f_h) - self.hz_to_mel(f_l)) / (M + 1)))

    def mel_filter_bank(self, M):
        """Triangular mel filter bank [M, K/2+1] whose filters sum to unity (sig.py:301-330); host-side, built once."""
        import numpy as np
        f_l, f_h = 0, self.f_s / 2
        K = self.K // 2 + 1
        H = np.zeros([M, K], dtype=np.float32)
        for m in range(1, M + 1):
            bl, c, bh = self.bpoint(m - 1, M, f_l, f_h), self.bpoint(m, M, f_l, f_h), self.bpoint(m + 1, M, f_l, f_h)
            for k in range(K):
                if k >= bl and k <= c:
                    H[m - 1, k] = (2 * (k - bl)) / ((bh - bl) * (c - bl))
                if k >= c and k <= bh:
                    H[m - 1, k] = (2 * (bh - k)) / ((bh - bl) * (bh - c))
        return H

    def subband(self, xi, n_filters=40, want_mask=True):
        """xi [..., 257] -> subband a priori SNR xi H^T [..., n_filters] (and its mask > 1): model.py:323-328."""
        import numpy as np
        xi, was_np = to_dev(xi, torch.float32)
        cache = self.__dict__.setdefault('_mel_cache', {})
        key = (n_filters, str(xi.device))
        if key not in cache:
            cache[key] = torch.from_numpy(np.ascontiguousarray(self.mel_filter_bank(n_filters))).to(xi.device)
        H = cache[key]
        rows = xi.numel() // self.n_bins
        sub = torch.empty(xi.shape[:-1] + (n_filters,), dtype=torch.float32, device=xi.device)
        ibm = torch.empty(sub.shape, dtype=torch.uint8, device=xi.device) if want_mask else None
        if rows:
            _lib.check(_lib.load().dxi_subband_ibm(_lib.ptr(xi), _lib.ptr(H), rows, self.n_bins, n_filters, _lib.ptr(sub),
                                                   _lib.ptr(ibm, allow_none=True), _lib.stream_ptr(xi.device)))
        return ret(sub, was_np), (ret(ibm.bool(), was_np) if want_mask else None)

    # -- training-target side ------------------------------------------------------
    def xi(self, S, D):
        """Instantaneous a priori SNR S^2 / max(D^2, 1e-12) (sig.py:110-121)."""
        S, was_np = to_dev(S, torch.float32)
        D, _ = to_dev(D, torch.float32)
        if S.shape != D.shape:
            raise ValueError('S and D must have the same shape')
        out = torch.empty_like(S)
        if S.numel():
            nb = S.shape[-1]
            _lib.check(_lib.load().dxi_xi_map(_lib.ptr(S), _lib.ptr(D), None, None, S.numel() // nb, nb, _lib.ptr(out), None,
                                              _lib.stream_ptr(S.device)))
        return ret(out, was_np)

    def gamma(self, X, D):
        """Instantaneous a posteriori SNR X^2 / max(D^2, 1e-12) (sig.py:123-134)."""
        return self.xi(X, D)

    def mix(self, s, d, s_len, d_len, snr, offsets=None):
        """Mixes clean speech and noise at the given SNR levels (sig.py:162-187; add_noise_batch / add_noise_pad /
        add_noise :214-284).  s [B, Ls], d [B, Ld] int16 (padded), s_len / d_len / snr per utterance.

        The reference draws the start of the noise section with tf.random.uniform (sig.py:277); `offsets` passes the
        draw in (None: numpy's default generator draws it the same way).  Returns s, d, x as float32 CUDA tensors
        [B, max(s_len)] (zero from s_len on) and the list of frame counts."""
        s, _ = to_dev(s, torch.int16)
        d, _ = to_dev(d, torch.int16)
        if s.dim() != 2 or d.dim() != 2 or s.shape[0] != d.shape[0]:
            raise ValueError('Waveforms are of incorrect rank.')
        B = s.shape[0]
        s_len = [int(v) for v in s_len]
        d_len = [int(v) for v in d_len]
        if len(s_len) != B or len(d_len) != B or len(snr) != B:
            raise ValueError('one s_len, d_len and snr per utterance')
        if any(dl < sl for sl, dl in zip(s_len, d_len)):
            raise ValueError('every noise recording must be at least as long as its clean-speech utterance')
        if offsets is None:
            import numpy as np
            rng = np.random.default_rng()
            offsets = [int(rng.integers(0, 1 + dl - sl)) for sl, dl in zip(s_len, d_len)]
        if any(o < 0 or o + sl > dl for o, sl, dl in zip(offsets, s_len, d_len)):
            raise ValueError('noise offset outside [0, d_len - s_len]')
        dev = s.device
        L = max(s_len) if B else 0
        i32 = lambda v: torch.tensor([int(u) for u in v], dtype=torch.int32).to(dev, non_blocking=True)
        snr_t = torch.tensor([float(u) for u in snr], dtype=torch.float32).to(dev, non_blocking=True)
        so, do, xo = (torch.empty((B, L), dtype=torch.float32, device=dev) for _ in range(3))
        if B and L:
            lib = _lib.load()
            ws = torch.empty(int(lib.dxi_mix_workspace_bytes(B)), dtype=torch.uint8, device=dev)
            sl_t, dl_t, of_t = i32(s_len), i32(d_len), i32(offsets)      # named: they must outlive the launch
            _lib.check(lib.dxi_mix(_lib.ptr(s), _lib.ptr(d), _lib.ptr(sl_t), _lib.ptr(dl_t), _lib.ptr(snr_t),
                                   _lib.ptr(of_t), B, s.shape[1], d.shape[1], _lib.ptr(so), _lib.ptr(do), _lib.ptr(xo),
                                   L, _lib.ptr(ws), _lib.stream_ptr(dev)))
        return so, do, xo, [self.n_frames(n) for n in s_len]
