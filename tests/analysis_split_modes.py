"""Analysis helper (not collected by pytest; CPU only, ~6 minutes): xi_hat error of ResNetV2 when the GEMM operands are rounded the
way the tensor-core precision modes round them, against the fp64 oracle.  x3 = (a_hi + a_lo)(w_hi + w_lo) without the lo x lo term
(mode f16x3), x2a = (a_hi + a_lo) w_hi, x2w = a_hi (w_hi + w_lo), x1 = a_hi w_hi (mode f16); then one GEMM position at a time with
two products.  Numbers quoted in DESIGN.md 2.   python tests/analysis_split_modes.py"""
import os
import numpy as np, torch, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import tcn as otcn, sig as osig, cdfmap
from deepxi_b200 import synth, weights, stats
w = weights.synthetic_resnetv2(0)
mu, sg = stats.packaged('resnet-1.1c')
x = synth.noisy_speech(3, 20000, seed=31)
inp, _, _ = osig.observation_batch(x, [20000, 12345, 33])
ref = otcn.resnetv2_forward(inp, w, dtype=torch.float64)
f16 = lambda t: t.to(torch.float16).to(torch.float64)
def split(t):
    hi = f16(t); lo = f16(t - hi); return hi, lo
def conv_mode(x, kernel, bias, d, mode):
    # x fp64 activations, kernel fp64
    a_hi, a_lo = split(x); w_hi, w_lo = split(kernel)
    if mode == 'x3':   parts = [(a_hi, w_hi), (a_lo, w_hi), (a_hi, w_lo)]
    elif mode == 'x2a': parts = [(a_hi, w_hi), (a_lo, w_hi)]
    elif mode == 'x2w': parts = [(a_hi, w_hi), (a_hi, w_lo)]
    elif mode == 'x1': parts = [(a_hi, w_hi)]
    out = None
    for a, ww in parts:
        y = otcn.conv1d(a, ww, None, d, 'causal')
        out = y if out is None else out + y
    return (out + bias).to(torch.float32).to(torch.float64)      # fp32 accumulate / epilogue
def fwd(mode):
    g = lambda name: torch.as_tensor(np.asarray(w[name]), dtype=torch.float64)
    lw = 'layer_with_weights-%d/%s'
    xx = torch.as_tensor(inp, dtype=torch.float64)
    h = torch.relu(otcn.layer_norm(otcn.conv1d(xx, g(lw % (0, 'kernel')), g(lw % (0, 'bias'))), g(lw % (1, 'gamma'))))
    li = 2
    for d in otcn.dilation_rates(40, 16):
        y = h
        for d_u in (1, d, 1):
            y = conv_mode(otcn.layer_norm(torch.relu(y)).to(torch.float32).to(torch.float64), g(lw % (li, 'kernel')), g(lw % (li, 'bias')), d_u, mode)
            li += 1
        h = (h + y).to(torch.float32).to(torch.float64)
    z = otcn.conv1d(h, g(lw % (li, 'kernel')), g(lw % (li, 'bias')))
    return torch.sigmoid(z).numpy()
def db_err(xbar):
    a = cdfmap.normal_cdf_inverse_db(xbar.astype(np.float64), mu, sg)
    b = cdfmap.normal_cdf_inverse_db(ref, mu, sg)
    m = np.isfinite(b) & (np.abs(b) < 40)
    e = np.abs(a - b)[m]
    return np.median(e), np.percentile(e, 99), e.max()
for mode in ('x3', 'x2a', 'x2w', 'x1'):
    print(mode, 'median / p99 / max |d xi_hat| dB: %.5f %.5f %.5f' % db_err(fwd(mode)))
print('--- per-position ablation (position: 0 = 1x1 256->64, 1 = dilated 64->64, 2 = 1x1 64->256)')
def fwd_pos(modes):
    g = lambda name: torch.as_tensor(np.asarray(w[name]), dtype=torch.float64)
    lw = 'layer_with_weights-%d/%s'
    xx = torch.as_tensor(inp, dtype=torch.float64)
    h = torch.relu(otcn.layer_norm(otcn.conv1d(xx, g(lw % (0, 'kernel')), g(lw % (0, 'bias'))), g(lw % (1, 'gamma'))))
    li = 2
    for d in otcn.dilation_rates(40, 16):
        y = h
        for pos, d_u in enumerate((1, d, 1)):
            y = conv_mode(otcn.layer_norm(torch.relu(y)).to(torch.float32).to(torch.float64), g(lw % (li, 'kernel')), g(lw % (li, 'bias')), d_u, modes[pos])
            li += 1
        h = (h + y).to(torch.float32).to(torch.float64)
    z = otcn.conv1d(h, g(lw % (li, 'kernel')), g(lw % (li, 'bias')))
    return torch.sigmoid(z).numpy()
for modes in (('x2w','x3','x3'), ('x3','x2w','x3'), ('x3','x3','x2w'), ('x2a','x3','x3'), ('x3','x2a','x3'), ('x3','x3','x2a')):
    print(modes, 'median / p99 / max: %.5f %.5f %.5f' % db_err(fwd_pos(modes)))
