"""Regenerates tests/golden/ and deepxi_b200/data/ from the read-only reference tree.

Run in the build container only (needs /root/reference); the GPU box uses the committed files.
Only DATA artefacts are copied (wav / mat / checkpoint index / statistics), never reference source.

  python tests/golden/make_golden.py
"""
import os, shutil, pickle, sys
import numpy as np
from scipy.io import loadmat

REF = os.environ.get('DEEPXI_REFERENCE', '/root/reference')
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
DATA = os.path.join(ROOT, 'deepxi_b200', 'data')
NAME = 'FB_FB10_07_voice-babble_5dB'


class _Stub:
    def __init__(self, *a, **k): pass
    def __setstate__(self, s): self.__dict__.update(s)


class _Unpickler(pickle.Unpickler):
    """Loads data/<ver>_inp_tgt.p (model.py:90-93) without TensorFlow / deepxi importable."""
    def find_class(self, mod, name):
        if mod.startswith('tensorflow'):
            return (lambda x, *a, **k: x) if name == 'convert_to_tensor' else (lambda *a, **k: None)
        if mod.startswith('deepxi'):
            return type(name, (_Stub,), {})
        if mod.startswith('numpy.core'):
            mod = mod.replace('numpy.core', 'numpy._core')
        return super().find_class(mod, name)


def main():
    os.makedirs(DATA, exist_ok=True)
    cp = lambda src, dst: shutil.copyfile(os.path.join(REF, src), os.path.join(HERE, dst))
    # known-answer test
    cp('set/test_noisy_speech/%s.wav' % NAME, 'kat_noisy.wav')
    cp('out/resnet-1.0c/e180/y/mmse-lsa/%s.wav' % NAME, 'kat_y_mmse-lsa_resnet-1.0c_e180.wav')
    xi = loadmat(os.path.join(REF, 'out/resnet-1.0c/e180/xi_hat/%s.mat' % NAME))['xi_hat'].astype(np.float32)
    np.save(os.path.join(HERE, 'kat_xi_hat_resnet-1.0c_e180.npy'), xi)
    # checkpoint indices (tensor names / shapes / crc; the weight shards themselves are absent, F3)
    cp('model/resnet-1.1c/epoch-199/variables/variables.index', 'resnet-1.1c_e199_variables.index')
    cp('model/resnet-1.1n/epoch-179/variables/variables.index', 'resnet-1.1n_e179_variables.index')
    cp('model/mhanet-1.1c/epoch-199/variables/variables.index', 'mhanet-1.1c_e199_variables.index')
    # statistics (F4): the pickles the reference actually reads, and data/stats.mat
    out = {}
    for ver in ('resnet-1.1c', 'resnet-1.1n', 'mhanet-1.1c'):
        with open(os.path.join(REF, 'data/%s_inp_tgt.p' % ver), 'rb') as f:
            o = _Unpickler(f).load()
        assert (o.N_d, o.N_s, o.K, o.f_s, o.n_feat) == (512, 256, 512, 16000, 257)
        assert o.xi_map.map_type == 'DBNormalCDF'
        out[ver + '/mu'] = np.asarray(o.xi_map.mu, np.float32)
        out[ver + '/sigma'] = np.asarray(o.xi_map.sigma, np.float32)
        shutil.copyfile(os.path.join(REF, 'data/%s_inp_tgt.p' % ver), os.path.join(HERE, '%s_inp_tgt.p' % ver))
    st = loadmat(os.path.join(REF, 'data/stats.mat'))['stats']
    out['stats.mat/mu'] = np.asarray(st['mu_hat'][0, 0], np.float32).reshape(-1)
    out['stats.mat/sigma'] = np.asarray(st['sigma_hat'][0, 0], np.float32).reshape(-1)
    shutil.copyfile(os.path.join(REF, 'data/stats.mat'), os.path.join(HERE, 'stats.mat'))
    np.savez(os.path.join(DATA, 'xi_stats.npz'), **out)
    print('wrote', sorted(os.listdir(HERE)), 'and', os.path.join(DATA, 'xi_stats.npz'))


if __name__ == '__main__':
    sys.exit(main())
