"""Mirror of deepxi/inp_tgt.py: inp_tgt_selector (:24-71), MagTgt.observation (:87-101) and MagXi (:141-240).

Only the 'MagXi' input/target pair has committed models; the other pairs of the reference raise
NotImplementedError, unknown names raise ValueError("Invalid inp_tgt type.") (inp_tgt.py:71).
"""
import math

import numpy as np
import torch

from . import _lib
from ._tensor import to_dev, ret
from .map import map_selector
from .sig import InputTarget

_OTHER = ('MagGamma', 'MagMag', 'MagSMM', 'MagPhaXiPha', 'STDCTXiCD', 'MagPhonme')


def inp_tgt_selector(inp_tgt_type, N_d, N_s, K, f_s, **kwargs):
    if inp_tgt_type == 'MagXi':
        return MagXi(N_d, N_s, K, f_s, xi_map_type=kwargs['map_type'], xi_map_params=kwargs.get('map_params'))
    if inp_tgt_type == 'MagXiGamma':      # inp_tgt.py:46-50: map_type / map_params are [xi, gamma] pairs
        mt, mp = kwargs['map_type'], kwargs.get('map_params') or [None, None]
        return MagXiGamma(N_d, N_s, K, f_s, xi_map_type=mt[0], xi_map_params=mp[0], gamma_map_type=mt[1], gamma_map_params=mp[1])
    if inp_tgt_type == 'MagGain':         # inp_tgt.py:51-52
        return MagGain(N_d, N_s, K, f_s, gain=kwargs['gain'])
    if inp_tgt_type in _OTHER:
        raise NotImplementedError('%s has no committed model: out of scope (SURVEY 2)' % inp_tgt_type)
    raise ValueError('Invalid inp_tgt type.')


class MagTgt(InputTarget):
    """Magnitude-spectrum input and any target (inp_tgt.py:73-139)."""

    def observation(self, x):
        """Noisy-speech int16 waveform [L] (or [B, L]) -> (x_STMS, x_STPS) (inp_tgt.py:87-101).

        normalise (int16 / 32768) is fused into the STFT kernel."""
        was_np = not isinstance(x, torch.Tensor)
        if was_np:
            x = np.asarray(x)
            if x.dtype != np.int16:
                x = x.astype(np.int16)
        elif x.dtype != torch.int16:
            x = x.to(torch.int16)
        x, _ = to_dev(x, torch.int16)
        mag, pha = self._stft(x)
        return ret(mag, was_np), ret(pha, was_np)

    def observation_batch(self, x_batch, x_batch_len):
        """DeepXi.observation_batch (model.py:2232-2254) in one launch: zero-padded device batches
        inp [B, Tmax, 257], phase [B, Tmax, 257] and the per-utterance frame counts (host list)."""
        x, _ = to_dev(x_batch, torch.int16)
        lens_host = [int(l) for l in x_batch_len]
        key = (str(x.device), torch.cuda.current_stream(x.device).cuda_stream, tuple(lens_host))
        cache = self.__dict__.setdefault('_lens_cache', {})
        if key not in cache:                              # the lengths of a repeated batch are uploaded once per stream
            if len(cache) > 16:
                cache.clear()
            cache[key] = torch.tensor(lens_host, dtype=torch.int32).to(x.device, non_blocking=True)
        lens = cache[key]
        if x.dim() != 2 or x.shape[0] != len(lens_host):
            raise ValueError('x_batch must be [B, Lmax] with one length per row')
        Tmax = self.n_frames(max(lens_host)) if lens_host else 0
        lib = _lib.load()
        B, L = x.shape
        mag = torch.empty((B, Tmax, self.n_bins), dtype=torch.float32, device=x.device)
        pha = torch.empty_like(mag)
        if B and Tmax:
            _lib.check(lib.dxi_stft(_lib.ptr(x), 1, _lib.ptr(lens), B, L, Tmax, _lib.ptr(mag), _lib.ptr(pha),
                                    _lib.stream_ptr(x.device)))
        return mag, pha, [self.n_frames(l) for l in lens_host]


class MagXi(MagTgt):
    """Magnitude-spectrum input and mapped a priori SNR target (inp_tgt.py:141-240)."""

    def __init__(self, N_d, N_s, K, f_s, xi_map_type, xi_map_params=None):
        super().__init__(N_d, N_s, K, f_s)
        self.n_feat = math.ceil(K / 2 + 1)
        self.n_outp = self.n_feat
        self.xi_map = map_selector(xi_map_type, xi_map_params)

    def set_stats(self, mu, sigma):
        self.xi_map.set_stats(mu, sigma)
        return self

    # -- training-target side ------------------------------------------------------
    def example(self, s, d, s_len, d_len, snr, offsets=None):
        """Observation and target of a training batch (inp_tgt.py:173-196): noisy-speech magnitude spectrum x_STMS
        and mapped a priori SNR xi_bar, both [B, Tmax, 257] (rows beyond an utterance's frame count are the
        transform of silence), and the frame counts."""
        so, do, xo, nfr = self.mix(s, d, s_len, d_len, snr, offsets)
        lens = torch.tensor([int(v) for v in s_len], dtype=torch.int32).to(so.device, non_blocking=True)
        S, _ = self._stft(so, lens)
        D, _ = self._stft(do, lens)
        X, _ = self._stft(xo, lens)
        mu, sigma = self.xi_map._stats_dev(S.device)
        xi_bar = torch.empty_like(S)
        if S.numel():
            _lib.check(_lib.load().dxi_xi_map(_lib.ptr(S), _lib.ptr(D), _lib.ptr(mu), _lib.ptr(sigma), S.numel() // self.n_feat,
                                              self.n_feat, None, _lib.ptr(xi_bar), _lib.stream_ptr(S.device)))
        return X, xi_bar, nfr

    def xi_db_moments(self, s_sample, d_sample, wav_len):
        """Per-bin (count, sum, sum of squares) of 10 log10 max(xi, 1e-12) over every frame of the sample, float64
        [3, 257] on the device: the mergeable form of transfrom_stats + xi + NormalCDF.stats (inp_tgt.py:114-139,
        :160-171; map.py:392-402).  s_sample / d_sample: float32 [N, L] clean speech and scaled noise."""
        s, _ = to_dev(s_sample, torch.float32)
        d, _ = to_dev(d_sample, torch.float32)
        if s.shape != d.shape or s.dim() != 2:
            raise ValueError('s_sample and d_sample must both be [N, L]')
        lens_host = [int(v) for v in wav_len]
        lens = torch.tensor(lens_host, dtype=torch.int32).to(s.device, non_blocking=True)
        acc = torch.zeros((3, self.n_feat), dtype=torch.float64, device=s.device)
        if s.shape[0]:
            S, _ = self._stft(s, lens)
            D, _ = self._stft(d, lens)
            nfr = torch.tensor([self.n_frames(n) for n in lens_host], dtype=torch.int32).to(s.device, non_blocking=True)
            _lib.check(_lib.load().dxi_xi_db_moments(_lib.ptr(S), _lib.ptr(D), _lib.ptr(nfr), S.shape[0], S.shape[1], self.n_feat,
                                                     _lib.ptr(acc), _lib.stream_ptr(s.device)))
        return acc

    def stats(self, s_sample, d_sample, x_sample, wav_len, group=None):
        """Statistics of the a priori SNR in dB for the CDF map (inp_tgt.py:160-171).  When torch.distributed is
        initialised every rank passes ITS shard of the sample and the moments are summed over the ranks (NCCL
        all-reduce of 3 x 257 float64 on the GPU, or gloo through the host): all ranks end with the statistics of
        the whole sample -- the only collective anywhere on a Deep Xi path."""
        from .stats import allreduce_moments, stats_from_moments
        acc = allreduce_moments(self.xi_db_moments(s_sample, d_sample, wav_len), group)
        mu, sigma = stats_from_moments(acc.cpu().numpy())
        self.set_stats(mu, sigma)
        return mu, sigma

    def xi_hat(self, xi_bar_hat):
        """A priori SNR estimate (inp_tgt.py:216-227)."""
        return self.xi_map.inverse(xi_bar_hat)

    def gamma_hat(self, xi_bar_hat):
        """Maximum-likelihood a posteriori SNR estimate xi_hat + 1 (inp_tgt.py:229-240)."""
        xi = self.xi_map.inverse(xi_bar_hat)
        return xi + np.float32(1.0) if isinstance(xi, np.ndarray) else xi + 1.0

    def _map_gain(self, xi_bar_hat, gtype, want_xi=False, want_gain=False, want_ibm=False):
        xb, was_np = to_dev(xi_bar_hat, torch.float32)
        mu, sigma = self.xi_map._stats_dev(xb.device)
        code = _lib.gtype_code(gtype) if want_gain else 0
        xi = torch.empty_like(xb) if want_xi else None
        G = torch.empty_like(xb) if want_gain else None
        ibm = torch.empty(xb.shape, dtype=torch.uint8, device=xb.device) if want_ibm else None
        if xb.numel():
            _lib.check(_lib.load().dxi_map_gain(_lib.ptr(xb), _lib.ptr(mu), _lib.ptr(sigma), xb.numel() // self.n_outp,
                                                self.n_outp, code, _lib.ptr(xi, allow_none=True),
                                                _lib.ptr(G, allow_none=True), _lib.ptr(ibm, allow_none=True),
                                                _lib.stream_ptr(xb.device)), value_error=True)
        return xi, G, ibm, was_np

    def gain_hat(self, xi_bar_hat, gtype):
        """gfunc(xi_hat, xi_hat + 1, gtype): the `gain` output type named by args.py:60-64 (not implemented in
        the reference's infer)."""
        _, G, _, was_np = self._map_gain(xi_bar_hat, gtype, want_gain=True)
        return ret(G, was_np)

    def deepmmse(self, x_STMS, xi_bar_hat):
        """Noise PSD estimate of out_type 'deepmmse' (model.py:314-318): |X|^2 * gfunc(xi_hat, xi_hat + 1, 'deepmmse'), one kernel."""
        mag, was_np = to_dev(x_STMS, torch.float32)
        xb, _ = to_dev(xi_bar_hat, torch.float32)
        if mag.shape != xb.shape or mag.shape[-1] != self.n_outp:
            raise ValueError('x_STMS and xi_bar_hat must share the shape [..., T, %d]' % self.n_outp)
        mu, sigma = self.xi_map._stats_dev(xb.device)
        out = torch.empty_like(xb)
        if xb.numel():
            _lib.check(_lib.load().dxi_deepmmse(_lib.ptr(mag), _lib.ptr(xb), _lib.ptr(mu), _lib.ptr(sigma), xb.numel() // self.n_outp,
                                                self.n_outp, _lib.ptr(out), _lib.stream_ptr(xb.device)))
        return ret(out, was_np)

    def ibm_hat(self, xi_bar_hat):
        """xi_hat > 1 as bool (model.py:319-322)."""
        _, _, ibm, was_np = self._map_gain(xi_bar_hat, None, want_ibm=True)
        return ret(ibm.bool(), was_np)

    def enhanced_speech(self, x_STMS, x_STPS, xi_bar_hat, gtype, n_frames=None, int16=False):
        """Enhanced speech (inp_tgt.py:198-214): inverse map -> gamma_hat = xi_hat + 1 -> gain -> synthesis,
        fused in one kernel.  Accepts [T, 257] or batched [B, T, 257] (+ optional per-utterance n_frames)."""
        code = _lib.gtype_code(gtype)
        mag, was_np = to_dev(x_STMS, torch.float32)
        pha, _ = to_dev(x_STPS, torch.float32)
        xb, _ = to_dev(xi_bar_hat, torch.float32)
        if not (mag.shape == pha.shape == xb.shape) or mag.shape[-1] != self.n_feat:
            raise ValueError('x_STMS, x_STPS and xi_bar_hat must share the shape [..., T, %d]' % self.n_feat)
        squeeze = mag.dim() == 2
        if squeeze:
            mag, pha, xb = mag[None], pha[None], xb[None]
        B, T, _ = mag.shape
        mu, sigma = self.xi_map._stats_dev(mag.device)
        n_out = (T + 1) * self.N_s
        nf = None
        if n_frames is not None:      # the frame counts of a repeated batch are uploaded once per stream (as the lengths are)
            key = (str(mag.device), torch.cuda.current_stream(mag.device).cuda_stream, tuple(int(n) for n in n_frames))
            cache = self.__dict__.setdefault('_nf_cache', {})
            if key not in cache:
                if len(cache) > 16:
                    cache.clear()
                cache[key] = torch.as_tensor(np.asarray(n_frames, np.int32)).to(mag.device, non_blocking=True)
            nf = cache[key]
        y = torch.empty((B, n_out), dtype=torch.int16 if int16 else torch.float32, device=mag.device)
        if B and T:
            _lib.check(_lib.load().dxi_enhance(_lib.ptr(mag), _lib.ptr(pha), _lib.ptr(xb), _lib.ptr(mu), _lib.ptr(sigma),
                                               code, _lib.ptr(nf, torch.int32, allow_none=True), B, T,
                                               None if int16 else _lib.ptr(y), _lib.ptr(y) if int16 else None, n_out,
                                               _lib.stream_ptr(mag.device)), value_error=True)
        return ret(y[0] if squeeze else y, was_np)


class MagXiGamma(MagTgt):
    """Magnitude-spectrum input, mapped a priori AND a posteriori SNR target (inp_tgt.py:345-457).  No model of
    this type is committed to the reference; the target-side arithmetic runs on the same kernels as MagXi."""

    def __init__(self, N_d, N_s, K, f_s, xi_map_type, xi_map_params=None, gamma_map_type=None, gamma_map_params=None):
        super().__init__(N_d, N_s, K, f_s)
        self.n_feat = math.ceil(K / 2 + 1)
        self.n_outp = self.n_feat * 2
        self.xi_map = map_selector(xi_map_type, xi_map_params)
        self.gamma_map = map_selector(gamma_map_type if gamma_map_type is not None else xi_map_type, gamma_map_params)

    def set_stats(self, xi_stats, gamma_stats):
        self.xi_map.set_stats(*xi_stats)
        self.gamma_map.set_stats(*gamma_stats)
        return self

    def _split(self, xi_gamma_bar_hat):
        t, was_np = to_dev(xi_gamma_bar_hat, torch.float32)
        if t.shape[-1] != self.n_outp:
            raise ValueError('last dimension must be %d' % self.n_outp)
        return t[..., :self.n_feat].contiguous(), t[..., self.n_feat:].contiguous(), was_np

    def xi_hat(self, xi_gamma_bar_hat):
        """inp_tgt.py:431-443."""
        xb, _, was_np = self._split(xi_gamma_bar_hat)
        return ret(self.xi_map.inverse(xb), was_np)

    def gamma_hat(self, xi_gamma_bar_hat):
        """inp_tgt.py:445-457."""
        _, gb, was_np = self._split(xi_gamma_bar_hat)
        return ret(self.gamma_map.inverse(gb), was_np)

    def example(self, s, d, s_len, d_len, snr, offsets=None):
        """x_STMS and concat(xi_bar, gamma_bar) (inp_tgt.py:383-409)."""
        so, do, xo, nfr = self.mix(s, d, s_len, d_len, snr, offsets)
        lens = torch.tensor([int(v) for v in s_len], dtype=torch.int32).to(so.device, non_blocking=True)
        S, _ = self._stft(so, lens)
        D, _ = self._stft(do, lens)
        X, _ = self._stft(xo, lens)
        xi_bar = self.xi_map.map(self.xi(S, D))
        gamma_bar = self.gamma_map.map(self.gamma(X, D))
        return X, torch.cat([xi_bar, gamma_bar], dim=-1), nfr

    def enhanced_speech(self, x_STMS, x_STPS, xi_gamma_bar_hat, gtype):
        """inp_tgt.py:411-429: both SNRs from the network, gain, synthesis."""
        from .gain import gfunc
        mag, was_np = to_dev(x_STMS, torch.float32)
        xb, gb, _ = self._split(xi_gamma_bar_hat)
        G = gfunc(self.xi_map.inverse(xb), self.gamma_map.inverse(gb), gtype)
        return ret(_synthesis_with_gain(self, mag, x_STPS, G), was_np)


class MagGain(MagTgt):
    """Magnitude-spectrum input, gain target (inp_tgt.py:459-519)."""

    def __init__(self, N_d, N_s, K, f_s, gain):
        super().__init__(N_d, N_s, K, f_s)
        _lib.gtype_code(gain)                      # same ValueError as gfunc for unknown names
        self.n_feat = math.ceil(K / 2 + 1)
        self.n_outp = self.n_feat
        self.gain = gain

    def example(self, s, d, s_len, d_len, snr, offsets=None):
        """x_STMS and G = gfunc(xi, gamma, self.gain) of the instantaneous SNRs (inp_tgt.py:476-501)."""
        from .gain import gfunc
        so, do, xo, nfr = self.mix(s, d, s_len, d_len, snr, offsets)
        lens = torch.tensor([int(v) for v in s_len], dtype=torch.int32).to(so.device, non_blocking=True)
        S, _ = self._stft(so, lens)
        D, _ = self._stft(do, lens)
        X, _ = self._stft(xo, lens)
        return X, gfunc(self.xi(S, D), self.gamma(X, D), self.gain), nfr

    def enhanced_speech(self, x_STMS, x_STPS, G_hat, gtype=None, n_frames=None, int16=False):
        """inp_tgt.py:503-519: the network output IS the gain (thresholded at 0.5 for 'ibm').  Batched [B, T, 257] inputs take the
        per-utterance n_frames; int16=True applies the save_wav rule (utils.py:28) in the synthesis kernel."""
        mag, was_np = to_dev(x_STMS, torch.float32)
        G, _ = to_dev(G_hat, torch.float32)
        if self.gain == 'ibm':
            G = (G > 0.5).to(torch.float32)
        return ret(_synthesis_with_gain(self, mag, x_STPS, G, n_frames, int16), was_np)


def _synthesis_with_gain(it, mag, x_STPS, G, n_frames=None, int16=False):
    """(|X| G) e^{j phase} -> waveform through dxi_istft (the gain multiply is fused into the synthesis kernel)."""
    pha, _ = to_dev(x_STPS, torch.float32)
    G, _ = to_dev(G, torch.float32)
    if not (mag.shape == pha.shape == G.shape) or mag.shape[-1] != it.n_bins:
        raise ValueError('x_STMS, x_STPS and the gain must share the shape [..., T, %d]' % it.n_bins)
    squeeze = mag.dim() == 2
    if squeeze:
        mag, pha, G = mag[None], pha[None], G[None]
    B, T, _ = mag.shape
    n_out = (T + 1) * it.N_s
    y = torch.empty((B, n_out), dtype=torch.int16 if int16 else torch.float32, device=mag.device)
    nf = None
    if n_frames is not None:
        nf = torch.as_tensor(np.asarray(n_frames, np.int32)).to(mag.device, non_blocking=True)
    if B and T:
        _lib.check(_lib.load().dxi_istft(_lib.ptr(mag.contiguous()), _lib.ptr(G.contiguous()), _lib.ptr(pha.contiguous()),
                                         _lib.ptr(nf, torch.int32, allow_none=True), B, T,
                                         None if int16 else _lib.ptr(y), _lib.ptr(y) if int16 else None, n_out, _lib.stream_ptr(mag.device)))
    return y[0] if squeeze else y
