import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from deepxi_b200 import _lib
lib = _lib.load()
out = torch.zeros(2, dtype=torch.int64, device='cuda')
for mode, name in ((0, 'ld'), (1, 'st'), (2, 'ld+st')):
    for warps in (4, 8, 16):
        for rounds in (64,):
            _lib.check(lib.dxi_debug_tmem_bw(mode, warps, rounds, _lib.ptr(out), _lib.stream_ptr()))
            torch.cuda.synchronize()
            cyc = int(out[0])
            byts = warps * rounds * 4096 * (2 if mode == 2 else 1)
            print('%-6s warps %2d rounds %3d: %7d cycles  %.1f B/cycle/SM' % (name, warps, rounds, cyc, byts / cyc))
