"""Seeded synthetic noisy-speech generator (there is no dataset beyond one wav in the reference tree).

Per utterance: s = white N(0,1) through a one-pole low-pass (a=0.95) x 4 Hz raised-cosine envelope,
scaled to RMS 0.05; d = white N(0,1) scaled for `snr_db`; x = clip(round((s+d)*32768)) as int16
(SURVEY 8d).  Used by tests, bench.py and smoke(); not part of the hot path.
"""
import numpy as np


def noisy_speech(n_utt, n_samples, seed=1234, snr_db=5.0, f_s=16000):
    """Returns int16 [n_utt, n_samples]."""
    rng = np.random.default_rng(seed)
    out = np.empty((n_utt, n_samples), np.int16)
    t = np.arange(n_samples) / f_s
    a = 0.95
    # one-pole low-pass as an FIR truncation (128 taps of a^n) so that generation stays vectorised
    h = a ** np.arange(128)
    for i in range(n_utt):
        s = rng.standard_normal(n_samples + 127)
        s = np.convolve(s, h, mode='valid')
        env = 0.5 * (1.0 - np.cos(2.0 * np.pi * 4.0 * t + rng.uniform(0, 2 * np.pi)))
        s = s * env
        s *= 0.05 / np.sqrt(np.mean(s ** 2))
        d = rng.standard_normal(n_samples)
        d *= np.sqrt(np.mean(s ** 2) / (10.0 ** (snr_db / 10.0))) / np.sqrt(np.mean(d ** 2))
        out[i] = np.clip(np.round((s + d) * 32768.0), -32768, 32767).astype(np.int16)
    return out
