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


def read_wav(path):
    """Returns (int16 waveform, f_s) of a mono PCM16 file (upstream utils.py:43-45 semantics)."""
    with wave.open(str(path), 'rb') as f:
        if f.getsampwidth() != 2 or f.getnchannels() != 1:
            raise ValueError('%s: only mono PCM16 wav files are supported' % path)
        f_s = f.getframerate()
        wav = np.frombuffer(f.readframes(f.getnframes()), dtype='<i2').astype(np.int16)
    return wav, f_s


def save_mat(path, data, name):
    if not path.endswith('.mat'):
        path = path + '.mat'
    savemat(path, {name: np.asarray(data)})


def read_mat(path):
    if not path.endswith('.mat'):
        path = path + '.mat'
    return loadmat(path)
