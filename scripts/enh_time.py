"""Time of the fused enhancement kernel alone (256 x 10 s, MMSE-LSA, int16 out); DXI_LIB selects a tuning build."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from deepxi_b200 import synth, _lib
from deepxi_b200.inp_tgt import inp_tgt_selector
it = inp_tgt_selector('MagXi', 512, 256, 512, 16000, map_type='DBNormalCDF', map_params=None)
z = np.load(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'deepxi_b200', 'data', 'xi_stats.npz'))
it.set_stats(z['resnet-1.1c/mu'], z['resnet-1.1c/sigma'])
B, L = 256, 160000
x = torch.from_numpy(np.tile(synth.noisy_speech(16, L, seed=5), (16, 1))).cuda()
mag, pha, nfr = it.observation_batch(x, [L] * B)
xb = torch.rand(mag.shape, device='cuda') * 0.98 + 0.01
for _ in range(3): y = it.enhanced_speech(mag, pha, xb, 'mmse-lsa', n_frames=nfr, int16=True)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20): y = it.enhanced_speech(mag, pha, xb, 'mmse-lsa', n_frames=nfr, int16=True)
e1.record(); torch.cuda.synchronize()
print('enhance %.4f ms' % (e0.elapsed_time(e1) / 20))
