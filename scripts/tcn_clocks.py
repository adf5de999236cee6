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
call = buf.cpu().numpy().reshape(n_tiles, 16)
c = call[:, :10]
d = np.diff(c, axis=1)
names = ['wait GEMM1', 'P1 (A2)', 'merge2', 'wait GEMM2 g0', 'P2 (h, A3)', 'merge3', 'A1 next', 'wait GEMM3', 'P3 (c1 out)']
print('flags', flags, 'tiles', n_tiles, 'stage', stage, ' total per tile: median %.0f cycles' % np.median(c[:, 9] - c[:, 0]))
for i, n in enumerate(names):
    print('%-16s median %7.0f  mean %7.0f  p90 %7.0f' % (n, np.median(d[:, i]), d[:, i].mean(), np.percentile(d[:, i], 90)))
# tile ownership in time order (tcn_umma.cu tile_at): CTA c walks the full rounds ascending, or descending in a `reverse`
# stage (even stages), and takes the tile of the ragged last round last
g = min(148, n_tiles)
n_full = n_tiles // g
reverse = (stage & 1) == 0 and not os.environ.get('DXI_TCN_NO_REVERSE')
def tiles_of(cta):
    rounds = list(range(n_full))[::-1] if reverse else list(range(n_full))
    ts = [cta + r * g for r in rounds]
    if cta + n_full * g < n_tiles:
        ts.append(cta + n_full * g)
    return ts
owned = [tiles_of(cta) for cta in range(g)]
gaps = [c[ts[i + 1], 0] - c[ts[i], 9] for ts in owned for i in range(len(ts) - 1)]
print('inter-tile gap median %.0f' % np.median(gaps))

# finer stamps inside P2 (thread 0: chunks cc = 0 then 4): 3 -> [h loads issued, wait d2] 4 -> [tmem ld] 10 -> [math, STG] 11
# -> [split, tmem st, arrive] 12 -> [h loads, wait d2] 13 -> [tmem ld] 14 -> [rest of chunk 1] 5
seq = [3, 4, 10, 11, 12, 13, 14, 5]
lab = ['i0 ldg issue + wait d2', 'i0 tmem ld', 'i0 math + STG', 'i0 split + tmem st + arrive', 'i1 ldg issue + wait d2', 'i1 tmem ld', 'i1 math..arrive']
for k in range(len(seq) - 1):
    dd = call[:, seq[k + 1]] - call[:, seq[k]]
    print('  P2 %-28s median %7.0f  mean %7.0f  p90 %7.0f' % (lab[k], np.median(dd), dd.mean(), np.percentile(dd, 90)))
# per-CTA busy span (first stamp of its first tile -> last stamp of its last tile; one SM clock) and per-round tile time
spans = np.array([c[ts[-1], 9] - c[ts[0], 0] for ts in owned])
print('per-CTA busy span: median %.0f  max %.0f  min %.0f cycles (%.1f us at 1.965 GHz max)' % (np.median(spans), spans.max(), spans.min(), spans.max() / 1965.0))
for r in range(max(len(ts) for ts in owned)):
    ts = np.array([o[r] for o in owned if len(o) > r])
    print('  step %d of the CTAs: tiles %4d  median tile %6.0f  mean %6.0f' % (r, len(ts), np.median(c[ts, 9] - c[ts, 0]), (c[ts, 9] - c[ts, 0]).mean()))
