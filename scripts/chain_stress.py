"""Stress test of the one-launch ResNetV2 kernel (tcn_chain.cu) over random batch shapes: ragged lengths, 1 .. 7 tiles per utterance,
fewer / more work items than SMs.  Every shape must (1) run without a trap, (2) give the same bits twice, (3) be independent of the
batch composition (an utterance alone == the same utterance inside the batch), (4) agree with the exact fp32 CUDA-core path
(tcn_f32.cu) within 0.05 dB of xi_hat.  Usage: python scripts/chain_stress.py [n_shapes] [seed]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from deepxi_b200 import weights
from deepxi_b200.network.selector import network_selector
from oracle import cdfmap

N = int(sys.argv[1]) if len(sys.argv) > 1 else 40
rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 0)
kw = dict(d_model=256, n_blocks=40, d_f=64, k=3, max_d_rate=16, unit_type='ReLU->LN->W+b', outp_act='Sigmoid')
z = np.load(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'deepxi_b200', 'data', 'xi_stats.npz'))
mu, sg = z['resnet-1.1c/mu'], z['resnet-1.1c/sigma']
w = weights.synthetic_resnetv2(4)
fast = network_selector('ResNetV2', None, 257, padding='causal', precision='f16x3', **kw).load_weights(w)
exact = network_selector('ResNetV2', None, 257, padding='causal', precision='f32', **kw).load_weights(w)
worst = 0.0
t0 = time.time()
for i in range(N):
    B = int(rng.choice([1, 2, 3, 7, 30, 61, 149, 200, 301]))
    T = int(rng.choice([1, 5, 127, 128, 129, 200, 256, 300, 511, 640, 800]))
    if B * T > 120000:
        B = max(1, 120000 // T)
    x = torch.rand(B, T, 257, device='cuda') * torch.rand(B, T, 1, device='cuda') * 3.0
    lens = rng.integers(1, T + 1, size=B)
    lens[0] = T
    for b in range(B):      # zero-padded frames are all-zero rows, as observation_batch leaves them
        x[b, lens[b]:] = 0.0
    y1 = fast(x)
    y2 = fast(x)
    torch.cuda.synchronize()
    assert torch.equal(y1, y2), ('not repeatable', B, T)
    pick = int(rng.integers(0, B))
    alone = fast(x[pick:pick + 1].contiguous())
    assert torch.equal(alone[0], y1[pick]), ('depends on the batch composition', B, T, pick)
    nb = min(B, 4)
    ref = exact(x[:nb].contiguous())
    a = cdfmap.normal_cdf_inverse_db(y1[:nb].cpu().numpy().astype(np.float64), mu, sg)
    r = cdfmap.normal_cdf_inverse_db(ref.cpu().numpy().astype(np.float64), mu, sg)
    m = np.isfinite(r) & (np.abs(r) < 40)
    err = float(np.abs(a - r)[m].max()) if m.any() else 0.0
    worst = max(worst, err)
    assert err < 0.05, ('accuracy', B, T, err)
    print('shape %2d: B %3d T %3d  max |d xi_hat| vs exact fp32 path %.4f dB' % (i, B, T, err), flush=True)
print('chain_stress: %d shapes ok, worst %.4f dB, %.1f s' % (N, worst, time.time() - t0))
