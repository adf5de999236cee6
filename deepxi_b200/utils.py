"""Mirror of the I/O helpers of deepxi/utils.py used on the inference path: save_wav (:18-29),
read_wav (:31-50), save_mat (:52-62), read_mat (:64-75).  PCM16 wav files are read / written with the
stdlib `wave` module (the reference uses soundfile / librosa, absent here)."""
import wave

import numpy as np
from scipy.io import loadmat, savemat


def save_wav(path, wav, f_s):
    """float32 waveforms are converted as utils.py:28 does: (wav * 32768) truncated toward zero, no clipping."""
    wav = np.squeeze(np.asarray(wav))
    if wav.dtype == np.float32:
        with np.errstate(invalid='ignore'):
            wav = np.asarray(np.multiply(wav, np.float32(32768.0)), dtype=np.int16)
    elif wav.dtype != np.int16:
        raise ValueError('save_wav expects float32 or int16 samples')
    with wave.open(str(path), 'wb') as f:
        f.setnchannels(1)
        f.setsampwidth(2)
        f.setframerate(int(f_s))
        f.writeframes(wav.astype('<i2').tobytes())


def read_wav(path, rule='reference', f_s=16000):
    """Returns (int16 waveform, f_s) of a PCM16 wav file.

    rule='reference' (default) is the ACTIVE branch of deepxi/utils.py:46-49: librosa.load(path, sr=16000, mono=True, float32)
    -> audio * 32767 -> astype(int16), i.e. for a 16 kHz PCM16 file trunc((x / 32768) * 32767) in float32 -- almost every non-zero
    sample moves 1 LSB towards zero -- and the mean of the channels for a multi-channel file.  librosa's resampler is not
    available here, so a file whose rate is not `f_s` is refused instead of being silently processed at the wrong rate.
    rule='pcm' returns the stored samples unchanged (the commented-out soundfile branch, utils.py:41-45; mono only)."""
    with wave.open(str(path), 'rb') as f:
        if f.getsampwidth() != 2:
            raise ValueError('%s: only PCM16 wav files are supported' % path)
        rate, n_ch = f.getframerate(), f.getnchannels()
        wav = np.frombuffer(f.readframes(f.getnframes()), dtype='<i2').astype(np.int16)
    if rule == 'pcm':
        if n_ch != 1:
            raise ValueError('%s: only mono files are supported with rule="pcm"' % path)
        return wav, rate
    if rule != 'reference':
        raise ValueError('rule must be "reference" or "pcm"')
    if rate != f_s:
        raise ValueError('%s is sampled at %d Hz; the reference resamples to %d Hz with librosa, which is not available: '
                         'resample the file first' % (path, rate, f_s))
    audio = wav.astype(np.float32) / np.float32(32768.0)              # soundfile / librosa PCM16 -> float32
    if n_ch > 1:
        audio = audio.reshape(-1, n_ch).mean(axis=1, dtype=np.float32)      # librosa.to_mono
    return (audio * np.float32(32767.0)).astype(np.int16), f_s


def save_mat(path, data, name):
    if not path.endswith('.mat'):
        path = path + '.mat'
    savemat(path, {name: np.asarray(data)})


def read_mat(path):
    if not path.endswith('.mat'):
        path = path + '.mat'
    return loadmat(path)
