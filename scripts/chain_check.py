"""Depth-first ResNetV2 path (tcn_chain.cu) against the float64 oracle on small shapes, then its time on a batch.
DXI_TCN_STAGED=1 selects the stage-per-launch path instead (A/B).  Usage: python scripts/chain_check.py [B] [seconds] [--no-oracle]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from deepxi_b200 import synth, weights, _lib
from deepxi_b200.network.selector import network_selector
from deepxi_b200.inp_tgt import inp_tgt_selector
from oracle import sig as osig, tcn as otcn, cdfmap

B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
SEC = float(sys.argv[2]) if len(sys.argv) > 2 else 10.0
kw = dict(d_model=256, n_blocks=40, d_f=64, k=3, max_d_rate=16, unit_type='ReLU->LN->W+b', outp_act='Sigmoid')
z = np.load(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'deepxi_b200', 'data', 'xi_stats.npz'))
mu, sg = z['resnet-1.1c/mu'], z['resnet-1.1c/sigma']
mode = 'staged' if os.environ.get('DXI_TCN_STAGED') else 'chain'
w = weights.synthetic_resnetv2(0)


def db_err(xbar, ref64):
    a = cdfmap.normal_cdf_inverse_db(np.asarray(xbar).astype(np.float64), mu, sg)
    b = cdfmap.normal_cdf_inverse_db(ref64, mu, sg)
    m = np.isfinite(b) & (np.abs(b) < 40)
    return np.abs(a - b)[m]


if '--no-oracle' not in sys.argv:
    for prec in ('f16x3', 'f16'):
        net = network_selector('ResNetV2', None, 257, padding='causal', precision=prec, **kw).load_weights(w)
        for lens in ([20000, 12345, 33], [64000], [256 * 129, 256 * 128, 256 * 300]):
            x = synth.noisy_speech(len(lens), max(lens), seed=31)
            inp, _, nfr = osig.observation_batch(x, lens)
            ref = otcn.resnetv2_forward(inp, w, padding='causal', dtype=torch.float64)
            t0 = time.time()
            xbar = net(inp)
            torch.cuda.synchronize()
            e = db_err(xbar, ref)
            again = net(inp)
            print('%s %s lens %s T %d: |d xi_hat| dB median %.5f p99 %.5f max %.5f  repeatable %s  (%.1f ms)' % (
                mode, prec, lens, inp.shape[1], np.median(e), np.percentile(e, 99), e.max(), bool(np.array_equal(np.asarray(xbar), np.asarray(again))),
                (time.time() - t0) * 1e3), flush=True)

net = network_selector('ResNetV2', None, 257, padding='causal', precision='f16x3', **kw).load_weights(w)
it = inp_tgt_selector('MagXi', 512, 256, 512, 16000, map_type='DBNormalCDF', map_params=None)
L = int(SEC * 16000)
x = np.tile(synth.noisy_speech(min(B, 16), L, seed=5), (-(-B // min(B, 16)), 1))[:B]
inp, _, _ = it.observation_batch(torch.from_numpy(x).cuda(), [L] * B)
for _ in range(3):
    y = net(inp)
torch.cuda.synchronize()
_lib.profile_enable(True)
for k in ('tcn_chain', 'tcn_stage', 'tcn_stem', 'tcn_head'):
    _lib.profile_read(k)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
N = 10
for _ in range(N):
    y = net(inp)
e1.record()
torch.cuda.synchronize()
prof = {k: _lib.profile_read(k)[0] / N for k in ('tcn_chain', 'tcn_stage', 'tcn_stem', 'tcn_head')}
print('%s B %d T %d: %.3f ms / forward  finite %s  kernels %s' % (mode, B, inp.shape[1], e0.elapsed_time(e1) / N, bool(torch.isfinite(y).all()),
                                                                  {k: round(v, 3) for k, v in prof.items()}), flush=True)
if len(sys.argv) > 3 and sys.argv[3].endswith('.npy'):
    np.save(sys.argv[3], y.cpu().numpy())
