"""Latency of BASELINE.json configs[0] (one 4 s utterance, ResNet-1.1c + MMSE-LSA, batch 1) through DeepXi.infer_batch:
device-resident int16 in -> f32 waveform out, CUDA events, median of 50 calls.  DXI_TCN_NO_FLAGS=1 for the A/B."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from deepxi_b200 import synth, weights
from deepxi_b200.model import DeepXi

kw = dict(d_model=256, n_blocks=40, d_f=64, k=3, max_d_rate=16, unit_type='ReLU->LN->W+b', outp_act='Sigmoid')
dx = DeepXi(512, 256, 512, 16000, 'MagXi', 'ResNetV2', ver='resnet-1.1c', map_type='DBNormalCDF', map_params=None, padding='causal',
            precision='f16x3', **kw)
dx.set_weights(weights.synthetic_resnetv2(0))
for B, L in ((1, 64000), (8, 64000), (1, 160000)):
    x = torch.from_numpy(synth.noisy_speech(B, L, seed=7)).cuda()
    lens = [L] * B
    for _ in range(5):
        dx.infer_batch(x, lens, 'y', 'mmse-lsa')
    ts = []
    for _ in range(50):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        dx.infer_batch(x, lens, 'y', 'mmse-lsa')
        e1.record()
        e1.synchronize()
        ts.append(e0.elapsed_time(e1))
    print('%s  B %d x %.0f s: median %.3f ms per call (min %.3f) = %.0f x real time' %
          ('NO_FLAGS' if os.environ.get('DXI_TCN_NO_FLAGS') else 'flags   ', B, L / 16000, np.median(ts), min(ts), B * L / 16000 / (np.median(ts) * 1e-3)))
