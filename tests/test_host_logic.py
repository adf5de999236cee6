"""CPU tests of the host side: checkpoint index reader, weight shapes / parameter counts, statistics
loaders, the C-ABI library (loads, exports every symbol of include/deepxi_b200.h), API error behaviour."""
import ctypes
import os
import re

import numpy as np
import pytest

from deepxi_b200 import tfbundle, weights, stats, synth, _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_checkpoint_index_shapes_and_param_counts(golden_dir):
    for ver, n_params, n_entries, shards in (('resnet-1.1c_e199', 1949953, 744, 2), ('resnet-1.1n_e179', 1949953, 744, 2),
                                             ('mhanet-1.1c_e199', 4600321, 206, 1)):
        path = os.path.join(golden_dir, ver + '_variables.index')
        num_shards, entries = tfbundle.read_index(path)
        assert num_shards == shards and len(entries) + 1 == n_entries        # +1: the header entry
        shapes = tfbundle.keras_weight_shapes(path)
        assert sum(int(np.prod(s)) for s in shapes.values()) == n_params    # log/summary/<ver>.txt
        synth_w = weights.synthetic_resnetv2(0) if 'resnet' in ver else weights.synthetic_mhanetv3(0)
        assert {k: v.shape for k, v in synth_w.items()} == shapes


def test_checkpoint_shards_missing_is_loud(golden_dir, tmp_path):
    import shutil
    d = tmp_path / 'epoch-199' / 'variables'
    d.mkdir(parents=True)
    shutil.copyfile(os.path.join(golden_dir, 'resnet-1.1c_e199_variables.index'), d / 'variables.index')
    with pytest.raises(FileNotFoundError):
        weights.load_checkpoint(str(tmp_path), 199)


def test_tensor_bundle_round_trip(tmp_path):
    """Writes a tiny bundle by hand (index table + shard) and reads it back with crc verification."""
    import struct

    def varint(v):
        out = b''
        while True:
            b = v & 0x7f
            v >>= 7
            out += bytes([b | (0x80 if v else 0)])
            if not v:
                return out

    arr = np.arange(12, dtype=np.float32).reshape(3, 4)
    raw = arr.tobytes()
    shape = b''.join(b'\x12' + varint(len(varint(d)) + 1) + b'\x08' + varint(d) for d in arr.shape)
    entry = (b'\x08\x01' + b'\x12' + varint(len(shape)) + shape + b'\x18\x00' + b'\x20\x00' + b'\x28' + varint(len(raw))
             + b'\x35' + struct.pack('<I', tfbundle.masked_crc32c(raw)))
    header = b'\x08\x01'
    key = b'layer_with_weights-0/kernel/.ATTRIBUTES/VARIABLE_VALUE'

    def block(items):
        body = b''
        for k, v in items:                       # no prefix sharing, one restart
            body += varint(0) + varint(len(k)) + varint(len(v)) + k + v
        return body + struct.pack('<II', 0, 1)
    data = block([(b'', header), (key, entry)])
    idx = block([(b'\xff', varint(0) + varint(len(data)))])
    blob = data + b'\x00' + b'\x00' * 4
    ioff = len(blob)
    blob += idx + b'\x00' + b'\x00' * 4
    footer = varint(0) + varint(0) + varint(ioff) + varint(len(idx))
    footer += b'\x00' * (40 - len(footer)) + struct.pack('<Q', 0xdb4775248b80fb57)
    (tmp_path / 'variables.index').write_bytes(blob + footer)
    (tmp_path / 'variables.data-00000-of-00001').write_bytes(raw)
    w = tfbundle.keras_weights(str(tmp_path / 'variables'), verify_crc=True)
    assert np.array_equal(w['layer_with_weights-0/kernel'], arr)
    (tmp_path / 'variables.data-00000-of-00001').write_bytes(raw[:-4] + b'\x00\x00\x00\x01')
    with pytest.raises(IOError):
        tfbundle.keras_weights(str(tmp_path / 'variables'), verify_crc=True)


def test_stats_loaders(golden_dir):
    st = stats.load_inp_tgt_pickle(os.path.join(golden_dir, 'resnet-1.1c_inp_tgt.p'))
    assert (st['N_d'], st['N_s'], st['K'], st['f_s'], st['map_type']) == (512, 256, 512, 16000, 'DBNormalCDF')
    mu, sg = stats.packaged('resnet-1.1c')
    assert np.array_equal(mu, st['mu']) and np.array_equal(sg, st['sigma'])
    assert np.allclose(mu[:4], [4.620561, -1.3624355, -6.7945337, -8.149311])
    mu_n, _ = stats.packaged('resnet-1.1n')
    assert np.array_equal(mu, mu_n)                        # the two resnet pickles are byte-identical (F4)
    mu_m, sg_m = stats.load_stats_mat(os.path.join(golden_dir, 'stats.mat'))
    assert mu_m.shape == (257,) and np.allclose(mu_m[:2], [4.808382, -1.0482588])
    assert np.array_equal(stats.packaged('stats.mat')[0], mu_m)
    assert 0.5 < np.abs(mu_m - mu).max() < 1.0             # the stale file differs by up to 0.94 dB
    with pytest.raises(KeyError):
        stats.packaged('nope')


def test_library_exports_every_declared_symbol(built_lib):
    header = open(os.path.join(ROOT, 'include', 'deepxi_b200.h')).read()
    declared = set(re.findall(r'\b(dxi_[a-z0-9_]+)\s*\(', header))
    declared -= {'dxi_net'}
    assert declared == set(_lib.SYMBOLS)
    lib = ctypes.CDLL(built_lib)
    for name in sorted(declared):
        assert hasattr(lib, name), name
    lib.dxi_version.restype = ctypes.c_int
    assert lib.dxi_version() == 100                        # no compute call: just proves the library loads


def test_library_has_no_oracle_or_cpu_fallback():
    """The product package must not import the oracle, and must refuse to run without CUDA."""
    import subprocess, sys
    code = ("import sys; import deepxi_b200, deepxi_b200.model, deepxi_b200.sig, deepxi_b200.gain, deepxi_b200.map; "
            "assert not any(m == 'oracle' or m.startswith('oracle.') for m in sys.modules), 'oracle imported'")
    subprocess.run([sys.executable, '-c', code], check=True, cwd=ROOT)
    for dirpath, _, files in os.walk(os.path.join(ROOT, 'deepxi_b200')):
        for f in files:
            if f.endswith(('.py', '.cu', '.cuh')):
                src = open(os.path.join(dirpath, f)).read()
                assert 'import oracle' not in src and 'from oracle' not in src, f
    import torch
    if not torch.cuda.is_available():
        from deepxi_b200 import gain
        with pytest.raises(RuntimeError):
            gain.gfunc(np.ones(4, np.float32), None, 'wf')


def test_api_errors_match_reference_conventions():
    from deepxi_b200 import gain, map as dmap, inp_tgt
    from deepxi_b200.network.selector import network_selector
    with pytest.raises(ValueError, match='Invalid gain function type.'):       # gain.py:190
        gain.gfunc(np.ones(4, np.float32), None, 'bogus')
    with pytest.raises(ValueError, match='Invalid map_type.'):                 # map.py:42
        dmap.map_selector('Bogus', None)
    with pytest.raises(ValueError, match='Invalid inp_tgt type.'):             # inp_tgt.py:71
        inp_tgt.inp_tgt_selector('Bogus', 512, 256, 512, 16000, map_type='DBNormalCDF', map_params=None)
    with pytest.raises(ValueError, match='Invalid network type.'):             # selector.py:131
        network_selector('Bogus', None, 257)
    with pytest.raises(NotImplementedError):
        network_selector('ResLSTM', None, 257)
    with pytest.raises(ValueError):
        inp_tgt.inp_tgt_selector('MagXi', 400, 160, 512, 16000, map_type='DBNormalCDF', map_params=None)
    it = inp_tgt.inp_tgt_selector('MagXi', 512, 256, 512, 16000, map_type='DBNormalCDF', map_params=None)
    assert it.n_feat == it.n_outp == 257 and it.n_frames(64000) == 250 and it.n_frames(39088) == 153


def test_synthetic_inputs_are_seeded():
    a = synth.noisy_speech(2, 4000, seed=7)
    b = synth.noisy_speech(2, 4000, seed=7)
    assert a.dtype == np.int16 and np.array_equal(a, b) and np.abs(a).max() > 1000


def test_mel_filter_bank_mirror_equals_oracle():
    """Host-side mirror of InputTarget.mel_filter_bank (sig.py:301-346) against the oracle's restatement."""
    from deepxi_b200.sig import InputTarget
    from oracle import sig as osig
    it = InputTarget(512, 256, 512, 16000)
    for M in (24, 40):
        H = it.mel_filter_bank(M)
        assert H.shape == (M, 257) and H.dtype == np.float32
        assert np.array_equal(H, osig.mel_filter_bank(M))
        assert (H >= 0).all() and np.all(np.diff(H.argmax(axis=1)) > 0)      # triangular, centres increase


def test_cli_flags_mirror_run_sh():
    """deepxi_b200.args parses the flag set run.sh passes for VER=resnet-1.1c INFER=1 (run.sh:98-140, main.py:12-60) the way
    deepxi/args.py does (str_to_bool / str_to_list / read_dtype)."""
    from deepxi_b200 import args as A, main as M
    argv = ('--ver resnet-1.1c --network_type ResNetV2 --d_model 256 --n_blocks 40 --d_f 64 --k 3 --max_d_rate 16 --causal 1 '
            '--unit_type ReLU->LN->W+b --loss_fnc BinaryCrossentropy --outp_act Sigmoid --max_epochs 200 --resume_epoch 0 '
            '--test_epoch 200 --mbatch_size 8 --inp_tgt_type MagXi --map_type DBNormalCDF --sample_size 1000 --f_s 16000 --T_d 32 '
            '--T_s 16 --min_snr -10 --max_snr 20 --snr_inter 1 --out_type y --gain mmse-lsa,mmse-stsa --infer 1 --gpu 0').split()
    a = A.get_args(argv)
    assert a.infer is True and a.train is False and a.causal is True and a.val_flag is True
    assert a.test_epoch == 200 and a.gain == ['mmse-lsa', 'mmse-stsa'] and a.map_type == 'DBNormalCDF'
    assert a.map_params == [None, None] and a.out_path == 'out' and a.test_x_path == 'set/test_noisy_speech'
    assert A.str_to_list('1,2;3,4') == [[1, 2], [3, 4]] and A.str_to_list('neg_5,0.5') == [-5, 0.5] and A.read_dtype('pi') > 3.14
    a.padding = 'causal'
    kw = M.network_kwargs(a)
    assert kw == dict(d_model=256, n_blocks=40, d_f=64, k=3, max_d_rate=16, padding='causal', unit_type='ReLU->LN->W+b',
                      outp_act='Sigmoid', precision='f16x3')
    with pytest.raises(NotImplementedError):
        M.main(argv[:-4] + ['--train', '1'])


def test_checkpoint_writer_round_trip(tmp_path, golden_dir):
    """tfbundle.save_keras_weights writes model/<ver>/epoch-<n>/variables/variables.{index,data-00000-of-00001};
    the TensorFlow-free reader gets the same tensors back (per-tensor masked crc32c verified), corruption is detected, and the
    entry names / shapes equal the ones in the reference's own index file for the same architecture."""
    w = weights.synthetic_resnetv2(3)
    model_path = tmp_path / 'model' / 'resnet-1.1c'
    prefix = model_path / 'epoch-199' / 'variables' / 'variables'
    tfbundle.save_keras_weights(str(prefix), w)
    back = weights.load_checkpoint(str(model_path), 199)
    assert sorted(back) == sorted(w) and all(np.array_equal(back[k], w[k]) and back[k].dtype == np.float32 for k in w)
    ref_shapes = tfbundle.keras_weight_shapes(os.path.join(golden_dir, 'resnet-1.1c_e199_variables.index'))
    assert tfbundle.keras_weight_shapes(str(prefix) + '.index') == ref_shapes
    data = prefix.parent / 'variables.data-00000-of-00001'
    raw = bytearray(data.read_bytes())
    raw[1000] ^= 0x40
    data.write_bytes(bytes(raw))
    with pytest.raises(IOError):
        weights.load_checkpoint(str(model_path), 199)


def test_read_wav_follows_the_reference_rule(tmp_path):
    """deepxi/utils.py:46-49 (the active branch): librosa.load(sr=16000, mono=True) -> * 32767 -> astype(int16).  For a 16 kHz PCM16
    file that is trunc((x / 32768) * 32767) in float32; stereo files are averaged; another rate would be resampled by librosa,
    which is absent here, so it raises instead of being processed at the wrong rate."""
    import wave
    from deepxi_b200 import utils, se_batch
    x = np.array([0, 1, -1, 2, 1000, -1000, 32767, -32768, 12345, -23456], np.int16)
    utils.save_wav(str(tmp_path / 'a.wav'), x, 16000)
    got, fs = utils.read_wav(str(tmp_path / 'a.wav'))
    want = ((x.astype(np.float32) / np.float32(32768.0)) * np.float32(32767.0)).astype(np.int16)
    assert fs == 16000 and got.dtype == np.int16 and np.array_equal(got, want)
    assert np.array_equal(want, [0, 0, 0, 1, 999, -999, 32766, -32767, 12344, -23455])
    raw, _ = utils.read_wav(str(tmp_path / 'a.wav'), rule='pcm')
    assert np.array_equal(raw, x)
    utils.save_wav(str(tmp_path / 'b.wav'), x, 44100)
    with pytest.raises(ValueError, match='44100'):
        utils.read_wav(str(tmp_path / 'b.wav'))
    (tmp_path / 'd').mkdir()
    utils.save_wav(str(tmp_path / 'd' / 'u.wav'), x, 8000)
    with pytest.raises(ValueError, match='8000'):
        se_batch.Batch(str(tmp_path / 'd'), f_s=16000)
    with wave.open(str(tmp_path / 's.wav'), 'wb') as f:      # stereo: librosa.to_mono = mean of the channels
        f.setnchannels(2); f.setsampwidth(2); f.setframerate(16000)
        f.writeframes(np.stack([x, x[::-1]], axis=1).astype('<i2').tobytes())
    st, _ = utils.read_wav(str(tmp_path / 's.wav'))
    m = (x.astype(np.float32) / np.float32(32768.0) + x[::-1].astype(np.float32) / np.float32(32768.0)) / np.float32(2)
    assert np.array_equal(st, (m * np.float32(32767.0)).astype(np.int16))
