"""Oracle: ResNetV2 forward (test infrastructure; see oracle/__init__.py).

Restates deepxi/network/tcn.py:116-225 (ResNetV2: feedforward :166-180, block :182-197,
unit "ReLU->LN->W+b" :218-223, dilation 2**(i % (log2(max_d_rate)+1)) :156-157, output conv +
sigmoid :158-161) on torch-CPU tensors.  Third-party semantics restated: Keras Conv1D
'causal' / 'same' padding with dilation, Keras LayerNormalization with epsilon=1e-6 (the
non-fused tf.nn.moments + tf.nn.batch_normalization path; biased variance).

Weights use the checkpoint's own names ('layer_with_weights-<i>/kernel' ...), Keras kernel
layout [k, C_in, C_out]; see deepxi_b200/weights.py for the index -> layer mapping.
"""
import numpy as np
import torch


def layer_norm(x, gamma=None, beta=None, eps=1e-6):
    """Keras LayerNormalization(axis=-1, epsilon=1e-6), non-fused op order."""
    mean = x.mean(dim=-1, keepdim=True)
    var = ((x - mean) ** 2).mean(dim=-1, keepdim=True)
    inv = torch.rsqrt(var + eps)
    if gamma is not None:
        inv = inv * gamma
    off = -mean * inv
    if beta is not None:
        off = beta + off
    return x * inv + off


def conv1d(x, kernel, bias, d_rate=1, padding='causal'):
    """Keras Conv1D on [B, T, C_in] with kernel [k, C_in, C_out].

    causal: out[t] = sum_j W[j] . x[t-(k-1-j)d]   (left zero padding)
    same  : out[t] = sum_j W[j] . x[t+(j-(k-1)/2)d] (symmetric zero padding, odd k)
    """
    k = kernel.shape[0]
    B, T, _ = x.shape
    out = None
    for j in range(k):
        shift = (k - 1 - j) * d_rate if padding == 'causal' else ((k - 1) // 2 - j) * d_rate
        # xs[t] = x[t - shift]
        if shift == 0:
            xs = x
        elif abs(shift) >= T:
            xs = torch.zeros_like(x)
        elif shift > 0:
            xs = torch.cat([x.new_zeros(B, shift, x.shape[2]), x[:, :T - shift]], dim=1)
        else:
            xs = torch.cat([x[:, -shift:], x.new_zeros(B, -shift, x.shape[2])], dim=1)
        y = xs @ kernel[j]
        out = y if out is None else out + y
    if bias is not None:
        out = out + bias
    return out


def dilation_rates(n_blocks=40, max_d_rate=16):
    """tcn.py:156-157."""
    return [int(2 ** (i % (np.log2(max_d_rate) + 1))) for i in range(n_blocks)]


def resnetv2_forward(inp, w, n_blocks=40, max_d_rate=16, padding='causal', dtype=torch.float32,
                     return_logits=False, quant=None):
    """ResNetV2 forward: inp [B, T, 257] -> x_bar [B, T, 257].

    w: dict of numpy arrays keyed 'layer_with_weights-<i>/{kernel,bias,gamma}'.
    quant: optional callable (tensor, role) -> tensor applied to every GEMM operand (activation
    after ReLU->LN, and weight); used to emulate 16-bit tensor-core operands for error budgets.
    """
    q = quant if quant is not None else (lambda t, role: t)
    g = lambda name: torch.as_tensor(np.asarray(w[name]), dtype=dtype)
    x = torch.as_tensor(np.asarray(inp), dtype=dtype)
    lw = 'layer_with_weights-%d/%s'
    # feedforward: Conv1D(d_model,1,bias) -> LN(scale only) -> ReLU   (tcn.py:166-180)
    h = conv1d(q(x, 'a'), q(g(lw % (0, 'kernel')), 'w'), g(lw % (0, 'bias')))
    h = torch.relu(layer_norm(h, gamma=g(lw % (1, 'gamma'))))
    li = 2
    for d in dilation_rates(n_blocks, max_d_rate):
        y = h
        for d_u in (1, d, 1):  # conv_1 (k=1), conv_2 (k, d_rate), conv_3 (k=1)   (tcn.py:182-197)
            y = layer_norm(torch.relu(y))  # "ReLU->LN->W+b" (tcn.py:218-223)
            y = conv1d(q(y, 'a'), q(g(lw % (li, 'kernel')), 'w'), g(lw % (li, 'bias')), d_u, padding)
            li += 1
        h = h + y
    z = conv1d(q(h, 'a'), q(g(lw % (li, 'kernel')), 'w'), g(lw % (li, 'bias')))
    if return_logits:
        return z.numpy()
    return torch.sigmoid(z).numpy()


def resnetv3_forward(inp, w, n_blocks=40, max_d_rate=16, padding='causal', dtype=torch.float32):
    """ResNetV3 (tcn.py:227-245): first layer Conv1D+b -> ReLU -> LayerNorm(center=False, scale=False); blocks and output as
    ResNetV2, layer indices one lower (the first LayerNorm has no weights)."""
    g = lambda name: torch.as_tensor(np.asarray(w[name]), dtype=dtype)
    x = torch.as_tensor(np.asarray(inp), dtype=dtype)
    lw = 'layer_with_weights-%d/%s'
    h = layer_norm(torch.relu(conv1d(x, g(lw % (0, 'kernel')), g(lw % (0, 'bias')))))
    li = 1
    for d in dilation_rates(n_blocks, max_d_rate):
        y = h
        for d_u in (1, d, 1):
            y = conv1d(layer_norm(torch.relu(y)), g(lw % (li, 'kernel')), g(lw % (li, 'bias')), d_u, padding)
            li += 1
        h = h + y
    return torch.sigmoid(conv1d(h, g(lw % (li, 'kernel')), g(lw % (li, 'bias')))).numpy()


def resnet_forward(inp, w, n_blocks=40, max_d_rate=16, padding='causal', dtype=torch.float32):
    """ResNet v1.0 (tcn.py:17-114): feedforward Conv1D(no bias) -> LN(gamma, beta) -> ReLU (:62-77); unit LN(gamma, beta) -> ReLU ->
    Conv1D (:95-114), use_bias only in conv_3 (:79-93); dilation 2**(i % (log2(max_d_rate)+1)) (:55-56)."""
    g = lambda name: torch.as_tensor(np.asarray(w[name]), dtype=dtype)
    x = torch.as_tensor(np.asarray(inp), dtype=dtype)
    lw = 'layer_with_weights-%d/%s'
    h = torch.relu(layer_norm(conv1d(x, g(lw % (0, 'kernel')), None), g(lw % (1, 'gamma')), g(lw % (1, 'beta'))))
    li = 2
    for d in dilation_rates(n_blocks, max_d_rate):
        y = h
        for j, d_u in enumerate((1, d, 1)):
            y = torch.relu(layer_norm(y, g(lw % (li, 'gamma')), g(lw % (li, 'beta'))))
            y = conv1d(y, g(lw % (li + 1, 'kernel')), g(lw % (li + 1, 'bias')) if j == 2 else None, d_u, padding)
            li += 2
        h = h + y
    return torch.sigmoid(conv1d(h, g(lw % (li, 'kernel')), g(lw % (li, 'bias')))).numpy()
