"""Config C3 of BASELINE.json as one JSON line (not the driver's bench: bench.py measures C2): MHANet-1.1c (MHANetV3, 5 blocks,
8 heads, no mask = the shipped model) + MMSE-LSA on 64 synthetic utterances x 30 s, int16 waveform in -> waveform out, inputs
resident in HBM, CUDA events.  usage: python scripts/bench_mhanet.py [precision] [steps]"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from deepxi_b200 import synth, weights, _lib
from deepxi_b200.model import DeepXi

prec = sys.argv[1] if len(sys.argv) > 1 else 'f16x3'
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
B, SECONDS, F_S = 64, 30, 16000
L = SECONDS * F_S
T = -(-L // 256)
kw = dict(d_model=256, n_blocks=5, n_heads=8, warmup_steps=40000, max_len=2048, causal=1, outp_act='Sigmoid')
dx = DeepXi(512, 256, 512, F_S, 'MagXi', 'MHANetV3', ver='mhanet-1.1c', map_type='DBNormalCDF', map_params=None, precision=prec, **kw)
dx.set_weights(weights.synthetic_mhanetv3(0))
base = synth.noisy_speech(8, L, seed=4321)
x = torch.from_numpy(np.tile(base, (B // 8, 1)).copy()).cuda()
xs = [x, x.roll(1, 0)]
lens = [L] * B
it = dx.inp_tgt


def step(i):
    inp, pha, _ = it.observation_batch(xs[i & 1], lens)
    return it.enhanced_speech(inp, pha, dx.network(inp), 'mmse-lsa')


for i in range(3):
    y = step(i)
torch.cuda.synchronize()
keys = ('stft', 'mha_gemm', 'mha_attn', 'enhance')
_lib.profile_enable(True)
for k in keys:
    _lib.profile_read(k)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
_lib.launch_count_reset()
e0.record()
for i in range(steps):
    y = step(i)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / steps
prof = {k: _lib.profile_read(k)[0] / steps for k in keys}
_lib.profile_enable(False)
assert torch.isfinite(y).all()
frames = B * T
net_ms = prof['mha_gemm'] + prof['mha_attn']
print(json.dumps({
    'metric': 'audio-seconds enhanced per second (MHANet-1.1c, MMSE-LSA)', 'value': B * SECONDS / (ms / 1e3), 'unit': 'audio-s/s',
    'n_gpus': 1, 'steps': steps, 'warmup': 3, 'ms_per_step': ms, 'dtype': prec, 'data': 'synthetic',
    'config': {'workload': 'MHANet-1.1c (MHANetV3, random-init weights of the checkpoint shapes, mask ignored as in the shipped model) + '
                           'MMSE-LSA, %d utt x %d s @16 kHz, STFT 512/256, int16 waveform in -> f32 waveform out' % (B, SECONDS),
               'frames': frames},
    'kernels_ms_per_step': prof, 'gpu_launches': _lib.launch_count(),
    'roofline': {'kernel': 'lin_umma_kernel + attn_umma_kernel (+ 2 fp32 edge layers)', 'bound': 'tensor', 'unit': 'TFLOP/s',
                 'achieved': frames * 17.73e6 / (net_ms * 1e-3) / 1e12, 'note': 'useful FLOPs (17.73 MFLOP per frame unmasked, SURVEY 8d)'}}))
