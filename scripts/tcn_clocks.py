"""Per-phase cycle breakdown of one tcn_stage_kernel launch (debug aid; uses dxi_debug_tcn_clocks)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from deepxi_b200 import weights, _lib
from deepxi_b200.network.selector import network_selector
kw = dict(d_model=256, n_blocks=40, d_f=64, k=3, max_d_rate=16, unit_type='ReLU->LN->W+b', outp_act='Sigmoid')
B, T = int(sys.argv[1]) if len(sys.argv) > 1 else 256, 625
stage = int(sys.argv[2]) if len(sys.argv) > 2 else 20
flags = int(sys.argv[3]) if len(sys.argv) > 3 else 0
net = network_selector('ResNetV2', None, 257, padding='causal', precision='f16x3', **kw).load_weights(weights.synthetic_resnetv2(0))
x = torch.rand(B, T, 257, device='cuda')
for _ in range(2): net(x)
n_tiles = B * ((T + 127) // 128)
buf = torch.zeros(n_tiles * 16, dtype=torch.int64, device='cuda')
lib = _lib.load()
lib.dxi_debug_tcn_clocks(_lib.ptr(buf), stage | (flags << 8))
net(x); torch.cuda.synchronize()
lib.dxi_debug_tcn_clocks(None, -1)
c = buf.cpu().numpy().reshape(n_tiles, 16)[:, :10]
d = np.diff(c, axis=1)
names = ['wait GEMM1', 'P1 (A2)', 'merge2', 'wait GEMM2 g0', 'P2 (h, A3)', 'merge3', 'A1 next', 'wait GEMM3', 'P3 (c1 out)']
print('flags', flags, 'tiles', n_tiles, 'stage', stage, ' total per tile: median %.0f cycles' % np.median(c[:, 9] - c[:, 0]))
for i, n in enumerate(names):
    print('%-16s median %7.0f  mean %7.0f  p90 %7.0f' % (n, np.median(d[:, i]), d[:, i].mean(), np.percentile(d[:, i], 90)))
# gap between consecutive tiles of the same CTA
g = 148
gaps = [c[t + g, 0] - c[t, 9] for t in range(0, n_tiles - g)]
print('inter-tile gap median %.0f' % np.median(gaps))
