"""Oracle: wav I/O rules (test infrastructure; see oracle/__init__.py).

Restates deepxi/utils.py:18-29 (save_wav: f32*32768 -> int16 by C truncation, no rounding, no
clipping) and reads PCM16 with the stdlib (the reference uses soundfile / librosa, absent here).
"""
import wave
import numpy as np


def read_wav_int16(path):
    with wave.open(str(path), 'rb') as f:
        assert f.getsampwidth() == 2 and f.getnchannels() == 1
        fs = f.getframerate()
        x = np.frombuffer(f.readframes(f.getnframes()), dtype='<i2').copy()
    return x, fs


def float_to_int16(wav):
    """utils.py:28: np.asarray(np.multiply(wav, 32768.0), dtype=np.int16)."""
    with np.errstate(invalid='ignore'):
        return np.asarray(np.multiply(np.asarray(wav, np.float32), np.float32(32768.0)), dtype=np.int16)


def write_wav_int16(path, x, fs=16000):
    with wave.open(str(path), 'wb') as f:
        f.setnchannels(1); f.setsampwidth(2); f.setframerate(fs)
        f.writeframes(np.asarray(x, '<i2').tobytes())
