"""|d xi_hat| in dB of the f16x3 MHANetV3 path against the float64 oracle at the C3 utterance length (T = 1875), both mask modes.
DXI_ATTN_P_SPLIT=1 selects attention probabilities as fp16 hi | lo instead of one fp16 rounding (A/B)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from oracle import sig as osig, attention as oatt, cdfmap
from deepxi_b200 import synth, weights
from deepxi_b200.network.selector import network_selector
MHA_KW = dict(d_model=256, n_blocks=5, n_heads=8, warmup_steps=40000, max_len=2048, causal=1, outp_act='Sigmoid')
z = np.load(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'deepxi_b200', 'data', 'xi_stats.npz'))
mu, sg = z['mhanet-1.1c/mu'], z['mhanet-1.1c/sigma']
for seed in (0, 1):
    w = weights.synthetic_mhanetv3(seed)
    x = synth.noisy_speech(1, 480000, seed=72 + seed)
    inp, _, _ = osig.observation_batch(x, [480000])
    for mask_mode in ('none', 'causal+pad'):
        ref = oatt.mhanetv3_forward(inp, w, mask_mode=mask_mode, dtype=torch.float64)
        net = network_selector('MHANetV3', None, 257, mask_mode=mask_mode, precision='f16x3', **MHA_KW).load_weights(w)
        xb = np.asarray(net(inp))[0]
        a = cdfmap.normal_cdf_inverse_db(xb.astype(np.float64), mu, sg)
        b = cdfmap.normal_cdf_inverse_db(ref[0], mu, sg)
        m = np.isfinite(b) & (np.abs(b) < 40)
        e = np.abs(a - b)[m]
        print('weights %d mask %-10s median %.5f p99 %.5f max %.5f dB' % (seed, mask_mode, np.median(e), np.percentile(e, 99), e.max()))
