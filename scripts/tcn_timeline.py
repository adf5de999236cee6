"""Per-CTA timeline of all 41 tcn_stage_kernel launches of one forward pass (globaltimer ns; debug aid: dxi_debug_tcn_clocks
with stage 255).  Per SM: how long between a CTA's exit and its successor's entry, the prologue, the first dependency wait,
the tile loop, and what share of the forward pass each takes."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from deepxi_b200 import weights, _lib
from deepxi_b200.network.selector import network_selector
kw = dict(d_model=256, n_blocks=40, d_f=64, k=3, max_d_rate=16, unit_type='ReLU->LN->W+b', outp_act='Sigmoid')
B, T = int(sys.argv[1]) if len(sys.argv) > 1 else 256, 625
net = network_selector('ResNetV2', None, 257, padding='causal', precision='f16x3', **kw).load_weights(weights.synthetic_resnetv2(0))
x = torch.rand(B, T, 257, device='cuda')
for _ in range(3): net(x)
n_tiles = B * ((T + 127) // 128)
grid = min(148, n_tiles)
buf = torch.zeros(41 * grid * 8, dtype=torch.int64, device='cuda')
lib = _lib.load()
lib.dxi_debug_tcn_clocks(_lib.ptr(buf), 255)
net(x); torch.cuda.synchronize()
lib.dxi_debug_tcn_clocks(None, -1)
a = buf.cpu().numpy().reshape(41, grid, 8).astype(np.int64)
t0 = a[:, :, 0].min()
entry, wready, dep, loop0, loop1, done, smid, n_r = [a[:, :, k] for k in range(8)]
print('mode', 'NO_FLAGS' if os.environ.get('DXI_TCN_NO_FLAGS') else 'flags', ' forward (first entry -> last exit) %.1f us' % ((done.max() - t0) / 1e3))
print('per CTA medians (us): prologue entry->weights %.2f | weights->first dep ok %.2f | dep ok->loop start %.2f | tile loop %.2f (per tile %.2f) | loop end->exit %.2f'
      % (np.median(wready - entry) / 1e3, np.median(dep - wready) / 1e3, np.median(loop0 - dep) / 1e3, np.median(loop1 - loop0) / 1e3,
         np.median((loop1 - loop0) / np.maximum(n_r, 1)) / 1e3, np.median(done - loop1) / 1e3))
# per SM: successive CTAs
gaps, busy = [], []
for sm in np.unique(smid):
    idx = np.argwhere(smid == sm)
    ev = sorted((entry[s, c], done[s, c], s, c) for s, c in idx)
    for (e0, d0, s0, c0), (e1, d1, s1, c1) in zip(ev[:-1], ev[1:]):
        gaps.append((e1 - d0) / 1e3)
    busy.append(sum(d - e for e, d, _, _ in ev) / 1e3)
gaps = np.array(gaps)
print('exit -> next CTA entry on the same SM: median %.2f us  mean %.2f  p90 %.2f  max %.2f   (n = %d)' % (np.median(gaps), gaps.mean(), np.percentile(gaps, 90), gaps.max(), len(gaps)))
print('per SM: CTA-resident time median %.1f us of %.1f us' % (np.median(busy), (done.max() - t0) / 1e3))
print('stage starts (first entry, us):', np.round((entry.min(axis=1)[:6] - t0) / 1e3, 1), '... stage ends (last exit):', np.round((done.max(axis=1)[:6] - t0) / 1e3, 1))
tot = (done.max() - t0) / 1e3
for name, v in (('prologue', wready - entry), ('first dep wait', dep - wready), ('first taps', loop0 - dep), ('tile loop', loop1 - loop0), ('tail', done - loop1)):
    print('  %-16s %.1f %% of SM time' % (name, 100 * v.sum() / 1e3 / (tot * len(np.unique(smid)))))
print('  %-16s %.1f %% of SM time' % ('no CTA resident', 100 * (1 - sum(busy) / (tot * len(busy)))))
