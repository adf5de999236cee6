"""Prints the register <-> (lane, column) layout of tcgen05.ld.16x256b.x8 and tcgen05.st.16x128b.x8 (tuning build:
DXI_LIB=deepxi_b200/libdeepxi_b200_dbg.so python scripts/tmem_layout.py)."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from deepxi_b200 import _lib
lib = _lib.load()
lib.dxi_debug_tmem_layout.argtypes = [ctypes.c_void_p, ctypes.c_void_p]
out = torch.zeros(128 * 64 + 128 * 32, dtype=torch.int32, device='cuda')
_lib.check(lib.dxi_debug_tmem_layout(ctypes.c_void_p(out.data_ptr()), _lib.stream_ptr()))
torch.cuda.synchronize()
o = out.cpu().numpy().astype(np.int64)
ld = o[:128 * 64].reshape(128, 64)
print('ld.16x256b.x8: thread t, register i (half h = lanes 16h..16h+15 of the warp) -> (lane, column)')
for t in (0, 1, 2, 3, 4, 5, 31, 32, 33):
    print(' t=%3d half0:' % t, ' '.join('(%d,%d)' % (v // 256, v % 256) for v in ld[t, :12]), '...', ' '.join('(%d,%d)' % (v // 256, v % 256) for v in ld[t, 28:32]))
    print('       half1:', ' '.join('(%d,%d)' % (v // 256, v % 256) for v in ld[t, 32:44]))
# model: reg i of half h: rep j = i // 4, e = i % 4 -> lane 32w + 16h + (t%32)//4 + 8 (e // 2), column 8 j + 2 (t % 4) + (e % 2)
ok = True
for t in range(128):
    w, l = t // 32, t % 32
    for h in range(2):
        for i in range(32):
            j, e = i // 4, i % 4
            lane, col = 32 * w + 16 * h + l // 4 + 8 * (e // 2), 8 * j + 2 * (l % 4) + (e % 2)
            if ld[t, 32 * h + i] != lane * 256 + col:
                ok = False
print('ld model (reg 4j+e -> lane 16h + l/4 + 8(e/2), col 8j + 2(l%4) + e%2):', 'MATCHES' if ok else 'does NOT match')
st = o[128 * 64:].reshape(128, 32)
print('st.16x128b.x8 read back with 32x32b: lane L, column c (of the 32 written) <- (thread, register)')
for L in (0, 1, 8, 9, 16, 17, 31, 32):
    print(' lane %3d:' % L, ' '.join('(%d,%d)' % (v >> 8, v & 255) for v in st[L, :10]))
ok = True
for L in range(128):
    w, l = L // 32, L % 32
    h, lr = l // 16, l % 16
    for c in range(32):
        j, cc = c // 4, c % 4
        t = 32 * w + 4 * (lr % 8) + cc
        reg = 16 * h + 2 * j + (lr // 8)
        if st[L, c] != ((t << 8) | reg):
            ok = False
print('st model (lane 16h + lr, col 4j + cc <- thread 4(lr%8) + cc, reg 2j + lr/8):', 'MATCHES' if ok else 'does NOT match')
