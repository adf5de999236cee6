"""A/B check of the tile-level stage dependencies (flags) against whole-launch dependencies (DXI_TCN_NO_FLAGS=1): the two
must agree bit for bit.  Usage: python scripts/flags_check.py out.npy [B] [L]; run once per mode, then `cmp` the files."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from deepxi_b200 import synth, weights
from deepxi_b200.network.selector import network_selector
from deepxi_b200.inp_tgt import inp_tgt_selector

out = sys.argv[1]
B = int(sys.argv[2]) if len(sys.argv) > 2 else 48
L = int(sys.argv[3]) if len(sys.argv) > 3 else 100000
pad = sys.argv[4] if len(sys.argv) > 4 else 'causal'
kw = dict(d_model=256, n_blocks=40, d_f=64, k=3, max_d_rate=16, unit_type='ReLU->LN->W+b', outp_act='Sigmoid')
net = network_selector('ResNetV2', None, 257, padding=pad, precision='f16x3', **kw).load_weights(weights.synthetic_resnetv2(0))
it = inp_tgt_selector('MagXi', 512, 256, 512, 16000, map_type='DBNormalCDF', map_params=None)
lens = [L - 997 * (i % 7) for i in range(B)]
x = synth.noisy_speech(B, L, seed=5)
inp, _, _ = it.observation_batch(x, lens)
res = []
for rep in range(3):
    y = net(inp)
    torch.cuda.synchronize()
    res.append(y.clone())
assert all(torch.equal(res[0], r) for r in res[1:]), 'not repeatable'
t0 = time.time()
for rep in range(5):
    y = net(inp)
torch.cuda.synchronize()
print('mode', 'NO_FLAGS' if os.environ.get('DXI_TCN_NO_FLAGS') else 'flags', 'B', B, 'T', inp.shape[1], '%.3f ms / forward' % ((time.time() - t0) / 5 * 1e3),
      'finite', bool(torch.isfinite(y).all()))
np.save(out, y.cpu().numpy())
