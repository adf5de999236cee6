"""Phase clocks of the depth-first ResNetV2 kernel (tuning build: DXI_DEBUG_BUILD=1 python -m deepxi_b200.build, then
DXI_LIB=deepxi_b200/libdeepxi_b200_dbg.so python scripts/chain_clocks.py [B] [seconds] [item]).  Prints, for one work item,
the average cycles between the stamps of the epilogue's thread 0 and of the MMA warp over the 40 blocks."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from deepxi_b200 import synth, weights, _lib
from deepxi_b200.network.selector import network_selector
from deepxi_b200.inp_tgt import inp_tgt_selector

B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
SEC = float(sys.argv[2]) if len(sys.argv) > 2 else 10.0
ITEM = int(sys.argv[3]) if len(sys.argv) > 3 else 300
kw = dict(d_model=256, n_blocks=40, d_f=64, k=3, max_d_rate=16, unit_type='ReLU->LN->W+b', outp_act='Sigmoid')
lib = _lib.load()
lib.dxi_debug_chain_clocks.argtypes = [ctypes.c_void_p, ctypes.c_int]
net = network_selector('ResNetV2', None, 257, padding='causal', precision=os.environ.get('DXI_CLK_PREC', 'f16x3'), **kw).load_weights(weights.synthetic_resnetv2(0))
it = inp_tgt_selector('MagXi', 512, 256, 512, 16000, map_type='DBNormalCDF', map_params=None)
L = int(SEC * 16000)
x = np.tile(synth.noisy_speech(min(B, 16), L, seed=5), (-(-B // min(B, 16)), 1))[:B]
inp, _, _ = it.observation_batch(torch.from_numpy(x).cuda(), [L] * B)
for _ in range(2):
    net(inp)
buf = torch.zeros(40 * 32, dtype=torch.int64, device='cuda')
_lib.check(lib.dxi_debug_chain_clocks(ctypes.c_void_p(buf.data_ptr()), ITEM))
net(inp)
torch.cuda.synchronize()
_lib.check(lib.dxi_debug_chain_clocks(ctypes.c_void_p(0), -1))
s = buf.cpu().numpy().reshape(40, 32).astype(np.int64)
names = {0: 'block start', 1: 'aux landed', 2: 'GEMM2 grp -> P2.0', 3: 'H chunk loaded', 4: 'P2.0 math', 5: 'P2.0 stored + arrive', 6: 'GEMM2 grp -> P2.1',
         7: 'P2.1 done', 8: 'merge 3', 9: 'dep + halo issue', 10: 'GEMM3 done', 11: 'P3 done (c1 in smem)', 12: 'GEMM1 done', 13: 'P1 done (A2 in TMEM)'}
blk = s[1:39]      # steady-state blocks
print('work item %d: block period %.0f cycles (epilogue thread 0: start to start)' % (ITEM, np.diff(s[:, 0]).mean()))
names[14] = '  (P3: merged)'; names[15] = '  (P3: c1 stored + fenced)'
prev = 0
for k in [1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 14, 15, 11, 12, 13]:
    print('  epi  %-26s +%6.0f' % (names[k], (blk[:, k] - blk[:, prev]).mean()))
    prev = k
if os.environ.get('DXI_CLK_P1'):
    b3 = s[3:39]
    pv = 12
    for k, nm in ((24, 'P1 ld done'), (25, 'P1 math 1'), (26, 'P1 merged'), (27, 'P1 converted'), (28, 'P1 tmem st done'), (13, 'P1 arrive')):
        print('       %-26s +%6.0f' % (nm, (b3[:, k] - b3[:, pv]).mean())); pv = k
print('  epi  %-26s +%6.0f' % ('-> next block start', (s[2:40, 0] - s[1:39, 13]).mean()))
mn = {16: 'W1 landed', 17: 'A3 chunk 0 ready', 23: 'A3 chunk 3 ready', 18: 'A3 chunk 7 ready', 19: 'c1 + W2 ready', 20: 'GEMM1 issued', 21: 'A2 + W3 ready', 22: 'GEMM2 issued'}
order = [16, 17, 23, 18, 19, 20, 21, 22]
prev = None
for k in order:
    if prev is None:
        print('  mma  %-26s  (block period %.0f)' % (mn[k], np.diff(s[:, 16]).mean()))
    else:
        print('  mma  %-26s +%6.0f' % (mn[k], (blk[:, k] - blk[:, prev]).mean()))
    prev = k
print('  mma  %-26s +%6.0f' % ('-> next W1 wait passed', (s[2:40, 16] - s[1:39, 22]).mean()))
# epilogue thread 0 relative to the MMA warp
print('  GEMM3 tail: A3 chunk 7 ready (mma) -> GEMM3 done (epi): %.0f' % (blk[:, 10] - blk[:, 18]).mean())
print('  GEMM1: c1 ready (mma) -> GEMM1 done (epi): %.0f' % (blk[:, 12] - blk[:, 19]).mean())
print('  GEMM2: A2 ready (mma) -> first group seen by P2.0 of the next block (epi): %.0f' % (s[2:40, 2] - s[1:39, 21]).mean())
t = s[0]
print('  tile: prologue (stats + z load + LN + TMEM store) %d, blocks %d, wait final aux %d, read-out + store %d, tile-end barrier %d  => tile %d cycles'
      % (t[25] - t[24], t[26] - t[25], t[27] - t[26], t[28] - t[27], t[29] - t[28], t[29] - t[24]))
print('  first block of the tile: start -> aux landed %d (W1 of block 0 is re-loaded after the tile-end barrier)' % (s[0][1] - s[0][0]))
f1, f2 = s[1][24:32], s[2][24:32]
if f1[0]:
    print('  first layer: round 0 %d, rounds 1-3 %d, wait aux + MMA done %d, statistics pass %d, merge + write-back pass %d'
          % (f1[1] - f1[0], f1[2] - f1[1], f1[3] - f1[2], f1[4] - f1[3], f1[5] - f1[4]))
    print('  output layer: max pass + merge %d (from tile-end stamp), convert pass + 257th column %d, wait MMAs %d, sigmoid half 0 %d, stores + half 1 %d'
          % (f2[0] - s[0][27], f2[1] - f2[0], f2[2] - f2[1], f2[3] - f2[2], s[0][28] - f2[3]))
