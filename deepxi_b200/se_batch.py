"""Mirror of deepxi/se_batch.py Batch (:12-55): pads all wav files of a directory into an int16 matrix."""
import glob
import os

import numpy as np

from .utils import read_wav


def Batch(fdir, snr_l=[], f_s=16000):
    """Files are read with the reference's rule (utils.read_wav); a file whose sampling rate is not f_s raises."""
    fname_l, wav_l, snr_test_l = [], [], []
    for fpath in sorted(glob.glob(os.path.join(fdir, '*.wav'))):
        for snr in snr_l:
            if fpath.find('_' + str(snr) + 'dB') != -1:
                snr_test_l.append(snr)
        wav, _ = read_wav(fpath, f_s=f_s)
        wav_l.append(wav)
        fname_l.append(os.path.basename(os.path.splitext(fpath)[0]))
    if not wav_l:
        raise ValueError('no .wav files in %s' % fdir)
    maxlen = max(len(w) for w in wav_l)
    wav_np = np.zeros([len(wav_l), maxlen], np.int16)
    for i, w in enumerate(wav_l):
        wav_np[i, :len(w)] = w
    return wav_np, np.array([len(w) for w in wav_l], np.int32), np.array(snr_test_l, np.int32), fname_l
