import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from deepxi_b200 import weights, synth, _lib
from deepxi_b200.network.selector import network_selector
from deepxi_b200.inp_tgt import inp_tgt_selector
kw = dict(d_model=256, n_blocks=40, d_f=64, k=3, max_d_rate=16, unit_type='ReLU->LN->W+b', outp_act='Sigmoid')
net = network_selector('ResNetV2', None, 257, padding='causal', precision='f16x3', **kw).load_weights(weights.synthetic_resnetv2(0))
it = inp_tgt_selector('MagXi', 512, 256, 512, 16000, map_type='DBNormalCDF', map_params=None)
x = np.tile(synth.noisy_speech(4, 160000, seed=51), (16, 1))
inp, _, _ = it.observation_batch(torch.from_numpy(x).cuda(), [160000] * 64)
lib = _lib.load()
small = inp[:3, :300].contiguous()
B, tiles, Ts = 3, 3, 3 * 128 + 64
h_bytes = B * tiles * 128 * 256 * 4
c1_bytes = B * 2 * 8 * Ts * 16
def region(net, off, nbytes):
    ws = list(net._ws.values())[0]
    base = (ws.data_ptr() + 255) // 256 * 256 - ws.data_ptr()
    return ws[base + off: base + off + nbytes].clone()
net(inp)
for S, which in ((0, 0), (1, 1)):
    lib.dxi_debug_tcn_stop_after(S)
    net(small); net(small); clean = region(net, h_bytes + which * c1_bytes, c1_bytes).view(torch.float16).view(B, 2, 8, Ts, 8)
    lib.dxi_debug_tcn_stop_after(-1); net(inp); lib.dxi_debug_tcn_stop_after(S)
    net(small); dirty = region(net, h_bytes + which * c1_bytes, c1_bytes).view(torch.float16).view(B, 2, 8, Ts, 8)
    d = (clean.float() - dirty.float()).abs()
    bad = (d > 0).nonzero()
    print('after stage %d: c1[%d] n diff %d' % (S, which, bad.shape[0]))
    if bad.shape[0]:
        print('  utts', sorted(set(bad[:, 0].tolist())), 'planes', sorted(set(bad[:, 1].tolist())), 'units', sorted(set(bad[:, 2].tolist())),
              'rows (incl. 32 pad)', int(bad[:, 3].min()), int(bad[:, 3].max()), 'maxdiff', float(d.max()))
        print('  clean zero frac in bad rows', float((clean[1, :, :, 160:] == 0).float().mean()), 'dirty', float((dirty[1, :, :, 160:] == 0).float().mean()))
lib.dxi_debug_tcn_stop_after(-1)
