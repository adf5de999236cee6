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
def ws_h(net):
    ws = list(net._ws.values())[0]
    base = (ws.data_ptr() + 255) // 256 * 256 - ws.data_ptr()
    n = 9 * 128 * 256
    return ws[base:base + n * 4].view(torch.float32).clone()
net(inp)    # allocate the big workspace first so that every run below uses the same buffer
for S in (0, 1, 2, 3, 5, 10, 40):
    lib.dxi_debug_tcn_stop_after(S)
    net(small); net(small); clean = ws_h(net)          # clean: second small run
    lib.dxi_debug_tcn_stop_after(-1)
    net(inp)                                            # dirty the workspace with a big run
    lib.dxi_debug_tcn_stop_after(S)
    net(small); dirty = ws_h(net)
    d = (clean - dirty).abs()
    tiles = sorted(set((d.view(9, -1).max(dim=1).values > 0).nonzero().flatten().tolist()))
    print('stop after stage %2d: h max diff %.3e, differing tiles %s' % (S, float(d.max()), tiles))
lib.dxi_debug_tcn_stop_after(-1)
