import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from deepxi_b200 import weights, synth
from deepxi_b200.network.selector import network_selector
from deepxi_b200.inp_tgt import inp_tgt_selector
kw = dict(d_model=256, n_blocks=40, d_f=64, k=3, max_d_rate=16, unit_type='ReLU->LN->W+b', outp_act='Sigmoid')
net = network_selector('ResNetV2', None, 257, padding='causal', precision='f16x3', **kw).load_weights(weights.synthetic_resnetv2(0))
it = inp_tgt_selector('MagXi', 512, 256, 512, 16000, map_type='DBNormalCDF', map_params=None)
x = np.tile(synth.noisy_speech(4, 160000, seed=51), (16, 1))
inp, _, _ = it.observation_batch(torch.from_numpy(x).cuda(), [160000] * 64)
def rep(tag, c, ref):
    d = (c - ref).abs(); bad = (d > 0).nonzero()
    print(tag, 'equal' if bad.shape[0] == 0 else 'BAD utts %s t (%d,%d) maxdiff %.2e' % (sorted(set(bad[:, 0].tolist())), int(bad[:, 1].min()), int(bad[:, 1].max()), float(d.max())))
small = inp[:3, :300].contiguous()
ref_small = net(small).clone()          # before any big run
rep('small again (no big run yet)', net(small), ref_small)
a = net(inp)
for i in range(3): rep('small #%d after big' % i, net(small), ref_small)
a = net(inp); torch.cuda.synchronize()
rep('small after big + sync', net(small), ref_small)
a = net(inp)
s256 = net(inp[:3, :256].contiguous())
rep('small after big + other small', net(small), ref_small)
rep('big prefix vs small', a[:3, :300], ref_small)
