import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from deepxi_b200 import weights, synth
from deepxi_b200.network.selector import network_selector
from deepxi_b200.inp_tgt import inp_tgt_selector
kw = dict(d_model=256, n_blocks=40, d_f=64, k=3, max_d_rate=16, unit_type='ReLU->LN->W+b', outp_act='Sigmoid')
prec = sys.argv[1] if len(sys.argv) > 1 else 'f16x3'
net = network_selector('ResNetV2', None, 257, padding='causal', precision=prec, **kw).load_weights(weights.synthetic_resnetv2(0))
it = inp_tgt_selector('MagXi', 512, 256, 512, 16000, map_type='DBNormalCDF', map_params=None)
x = synth.noisy_speech(4, 160000, seed=51)
inp, _, _ = it.observation_batch(torch.from_numpy(x).cuda(), [160000] * 4)
a = net(inp)
for B in (1, 2, 3, 4):
    for Tc in (300, 200, 257, 129, 385, 500):
        c = net(inp[:B, :Tc].contiguous()); c2 = net(inp[:B, :Tc].contiguous())
        d = (c - a[:B, :Tc]).abs(); bad = (d > 0).nonzero()
        d2 = (c - c2).abs(); bad2 = (d2 > 0).nonzero()
        info = ''
        if bad.shape[0]:
            info = ' bad utts %s t range (%d,%d)' % (sorted(set(bad[:, 0].tolist())), int(bad[:, 1].min()), int(bad[:, 1].max()))
        print('B=%d T=%d tiles=%d: repeat-equal %s, prefix-equal %s%s' % (B, Tc, B * ((Tc + 127) // 128), bad2.shape[0] == 0, bad.shape[0] == 0, info))
