"""Oracle: MHANetV3 forward (test infrastructure; see oracle/__init__.py).

Restates deepxi/network/attention.py:387-442 (MHANetV3), :327-353 (MHANetV2.block),
:88-101 (feed_forward_network), :355-385 (AttentionMaskV2) and the third-party
tfa.layers.MultiHeadAttention einsum formulation (head_size 32, 8 heads, no projection bias):
    q = einsum('...NI,HIO->...NHO', x, Wq) / sqrt(32);  k, v likewise (no scaling)
    logits = einsum('...NHO,...MHO->...HNM', q, k);  [logits += -10e9*(1-mask)];  softmax
    out = einsum('...NHI,HIO->...NO', einsum('...HNM,...MHI->...NHI', coef, v), Wp)

mask_mode:
  'none'        what the shipped model computes: the 4th list element passed at attention.py:344 is
                ignored by tfa, so attention runs over all Tmax (zero-padded) frames.
  'causal+pad'  the mask AttentionMaskV2 builds (causal AND both frames non-zero), applied the way
                tfa applies a mask.  Rows of padded frames are don't-care.
"""
import numpy as np
import torch
from .tcn import layer_norm, conv1d


def mhanetv3_forward(inp, w, n_blocks=5, n_heads=8, mask_mode='none', dtype=torch.float32,
                     return_logits=False, quant=None):
    """inp [B, T, 257] (zero-padded frames are all-zero rows) -> x_bar [B, T, 257]."""
    q_ = quant if quant is not None else (lambda t, role: t)
    g = lambda name: torch.as_tensor(np.asarray(w[name]), dtype=dtype)
    lw = 'layer_with_weights-%d/%s'
    x_in = torch.as_tensor(np.asarray(inp), dtype=dtype)
    B, T, _ = x_in.shape
    x = conv1d(q_(x_in, 'a'), q_(g(lw % (0, 'kernel')), 'w'), None)
    x = torch.relu(layer_norm(x, g(lw % (1, 'gamma')), g(lw % (1, 'beta'))))
    x = x + g(lw % (2, 'embeddings'))[:T][None]
    mask = None
    if mask_mode == 'causal+pad':
        valid = (x_in != 0).any(dim=-1)                       # Masking(mask_value=0.0).compute_mask
        seq = valid[:, None, :] & valid[:, :, None]           # [B, N, M]
        causal = torch.tril(torch.ones(T, T, dtype=torch.bool))
        mask = (seq & causal[None]).to(dtype)[:, None]        # [B, 1, N, M]
    elif mask_mode != 'none':
        raise ValueError('mask_mode')
    li = 3
    for _ in range(n_blocks):
        Wq, Wk, Wv = g(lw % (li, 'query_kernel')), g(lw % (li, 'key_kernel')), g(lw % (li, 'value_kernel'))
        Wp = g(lw % (li, 'projection_kernel'))
        depth = Wq.shape[-1]
        xa = q_(x, 'a')
        q = torch.einsum('bni,hio->bnho', xa, q_(Wq, 'w'))
        k = torch.einsum('bni,hio->bnho', xa, q_(Wk, 'w'))
        v = torch.einsum('bni,hio->bnho', xa, q_(Wv, 'w'))
        q = q / torch.sqrt(torch.tensor(float(depth), dtype=dtype))
        logits = torch.einsum('bnho,bmho->bhnm', q_(q, 'a'), q_(k, 'a'))
        if mask is not None:
            logits = logits + (-10e9) * (1.0 - mask)
        coef = torch.softmax(logits, dim=-1)
        att = torch.einsum('bhnm,bmhi->bnhi', q_(coef, 'a'), q_(v, 'a'))
        mha = torch.einsum('bnhi,hio->bno', q_(att, 'a'), q_(Wp, 'w'))
        a = layer_norm(x + mha, g(lw % (li + 1, 'gamma')), g(lw % (li + 1, 'beta')))
        f = torch.relu(conv1d(q_(a, 'a'), q_(g(lw % (li + 2, 'kernel')), 'w'), g(lw % (li + 2, 'bias'))))
        f = conv1d(q_(f, 'a'), q_(g(lw % (li + 3, 'kernel')), 'w'), g(lw % (li + 3, 'bias')))
        x = layer_norm(a + f, g(lw % (li + 4, 'gamma')), g(lw % (li + 4, 'beta')))
        li += 5
    z = conv1d(q_(x, 'a'), q_(g(lw % (li, 'kernel')), 'w'), g(lw % (li, 'bias')))
    if return_logits:
        return z.numpy()
    return torch.sigmoid(z).numpy()
