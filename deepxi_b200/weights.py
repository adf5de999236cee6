"""Weights for the a priori SNR estimators, keyed by the checkpoint's own tensor names.

Layer order of the reference checkpoints (model/<ver>/epoch-<n>/variables/variables.index):
  resnet-1.1c / resnet-1.1n  (deepxi/network/tcn.py:116-225, ResNetV2):
    layer_with_weights-0   first conv   kernel [1,257,256], bias [256]
    layer_with_weights-1   LN           gamma [256]
    layer_with_weights-(2+3i .. 4+3i), i=0..39   conv_1 [1,256,64], conv_2 [3,64,64], conv_3 [1,64,256] (+bias)
    layer_with_weights-122 output conv  kernel [1,256,257], bias [257]
  mhanet-1.1c  (deepxi/network/attention.py:387-442, MHANetV3):
    -0 conv kernel [1,257,256] (no bias); -1 LN gamma/beta; -2 embeddings [2048,256];
    per block b (base 3+5b): MHA {query,key,value}_kernel [8,256,32], projection_kernel [8,32,256];
    LN; conv [1,256,1024]+bias; conv [1,1024,256]+bias; LN;   -28 output conv [1,256,257]+bias.

The trained weight shards are absent from the reference tree (.MISSING_LARGE_BLOBS), so tests and
benchmarks use seeded synthetic weights of exactly these shapes; `load_checkpoint` reads real
shards (crc-verified) when they are supplied.
"""
import os
import numpy as np
from . import tfbundle

_LW = 'layer_with_weights-%d/%s'


def resnetv2_shapes(n_blocks=40, d_model=256, d_f=64, k=3, n_feat=257, n_outp=257):
    s = {_LW % (0, 'kernel'): (1, n_feat, d_model), _LW % (0, 'bias'): (d_model,), _LW % (1, 'gamma'): (d_model,)}
    li = 2
    for _ in range(n_blocks):
        for (kk, cin, cout) in ((1, d_model, d_f), (k, d_f, d_f), (1, d_f, d_model)):
            s[_LW % (li, 'kernel')] = (kk, cin, cout)
            s[_LW % (li, 'bias')] = (cout,)
            li += 1
    s[_LW % (li, 'kernel')] = (1, d_model, n_outp)
    s[_LW % (li, 'bias')] = (n_outp,)
    return s


def resnetv3_shapes(n_blocks=40, d_model=256, d_f=64, k=3, n_feat=257, n_outp=257):
    """ResNetV3 (tcn.py:227-245): ResNetV2 whose first layer has no LayerNorm weights, so every later index is one lower."""
    s = {_LW % (0, 'kernel'): (1, n_feat, d_model), _LW % (0, 'bias'): (d_model,)}
    li = 1
    for _ in range(n_blocks):
        for (kk, cin, cout) in ((1, d_model, d_f), (k, d_f, d_f), (1, d_f, d_model)):
            s[_LW % (li, 'kernel')] = (kk, cin, cout)
            s[_LW % (li, 'bias')] = (cout,)
            li += 1
    s[_LW % (li, 'kernel')] = (1, d_model, n_outp)
    s[_LW % (li, 'bias')] = (n_outp,)
    return s


def resnet_shapes(n_blocks=40, d_model=256, d_f=64, k=3, n_feat=257, n_outp=257):
    """ResNet v1.0 (tcn.py:17-114; layer order and sizes of log/summary/resnet-1.0c.txt: 1 975 553 parameters): first conv without
    bias, LN(gamma, beta); per block LN, conv, LN, conv, LN, conv (+bias only on the last); output conv with bias."""
    s = {_LW % (0, 'kernel'): (1, n_feat, d_model), _LW % (1, 'gamma'): (d_model,), _LW % (1, 'beta'): (d_model,)}
    li = 2
    for _ in range(n_blocks):
        for j, (kk, cin, cout) in enumerate(((1, d_model, d_f), (k, d_f, d_f), (1, d_f, d_model))):
            s[_LW % (li, 'gamma')] = (cin,)
            s[_LW % (li, 'beta')] = (cin,)
            s[_LW % (li + 1, 'kernel')] = (kk, cin, cout)
            if j == 2:
                s[_LW % (li + 1, 'bias')] = (cout,)
            li += 2
    s[_LW % (li, 'kernel')] = (1, d_model, n_outp)
    s[_LW % (li, 'bias')] = (n_outp,)
    return s


def mhanetv3_shapes(n_blocks=5, d_model=256, n_heads=8, max_len=2048, n_feat=257, n_outp=257):
    d_k, d_ff = d_model // n_heads, 4 * d_model
    s = {_LW % (0, 'kernel'): (1, n_feat, d_model), _LW % (1, 'gamma'): (d_model,), _LW % (1, 'beta'): (d_model,),
         _LW % (2, 'embeddings'): (max_len, d_model)}
    li = 3
    for _ in range(n_blocks):
        for nm in ('query_kernel', 'key_kernel', 'value_kernel'):
            s[_LW % (li, nm)] = (n_heads, d_model, d_k)
        s[_LW % (li, 'projection_kernel')] = (n_heads, d_k, d_model)
        for j in (1, 4):
            s[_LW % (li + j, 'gamma')] = (d_model,)
            s[_LW % (li + j, 'beta')] = (d_model,)
        s[_LW % (li + 2, 'kernel')] = (1, d_model, d_ff); s[_LW % (li + 2, 'bias')] = (d_ff,)
        s[_LW % (li + 3, 'kernel')] = (1, d_ff, d_model); s[_LW % (li + 3, 'bias')] = (d_model,)
        li += 5
    s[_LW % (li, 'kernel')] = (1, d_model, n_outp)
    s[_LW % (li, 'bias')] = (n_outp,)
    return s


def _synth(shapes, seed, scale_rules):
    rng = np.random.default_rng(seed)
    w = {}
    for name in sorted(shapes, key=lambda n: (int(n.split('/')[0].split('-')[-1]), n)):
        shp = shapes[name]
        var = name.split('/')[1]
        if var in ('gamma',):
            a = 1.0 + 0.1 * rng.standard_normal(shp)
        elif var in ('bias', 'beta'):
            a = 0.1 * rng.standard_normal(shp)
        elif var == 'embeddings':
            a = 0.1 * rng.standard_normal(shp)
        else:
            fan_in = int(np.prod(shp[:-1])) if var == 'kernel' else shp[-2]
            a = rng.standard_normal(shp) * np.sqrt(2.0 / fan_in)
            a *= scale_rules(name, shp)
        w[name] = a.astype(np.float32)
    return w


def synthetic_resnetv2(seed=0, **kw):
    """Seeded He-normal weights in the resnet-1.1c/1.1n checkpoint shapes (conv_3 x0.3, output x0.2 so
    that the residual stream and the output logits stay in a trained-network-like range)."""
    shapes = resnetv2_shapes(**kw)
    last = max(int(n.split('/')[0].split('-')[-1]) for n in shapes)

    def rule(name, shp):
        li = int(name.split('/')[0].split('-')[-1])
        if li == last: return 0.2
        if li >= 2 and (li - 2) % 3 == 2: return 0.3
        return 1.0
    return _synth(shapes, seed, rule)


def synthetic_resnetv3(seed=0, **kw):
    shapes = resnetv3_shapes(**kw)
    last = max(int(n.split('/')[0].split('-')[-1]) for n in shapes)
    return _synth(shapes, seed, lambda name, shp: 0.2 if int(name.split('/')[0].split('-')[-1]) == last else
                  (0.3 if (int(name.split('/')[0].split('-')[-1]) - 1) % 3 == 2 and int(name.split('/')[0].split('-')[-1]) >= 1 else 1.0))


def synthetic_resnet(seed=0, **kw):
    shapes = resnet_shapes(**kw)
    last = max(int(n.split('/')[0].split('-')[-1]) for n in shapes)

    def rule(name, shp):
        li = int(name.split('/')[0].split('-')[-1])
        if li == last: return 0.2
        if li >= 2 and (li - 2) % 6 == 5: return 0.3          # third conv of a block
        return 1.0
    return _synth(shapes, seed, rule)


def synthetic_mhanetv3(seed=0, **kw):
    shapes = mhanetv3_shapes(**kw)
    last = max(int(n.split('/')[0].split('-')[-1]) for n in shapes)

    def rule(name, shp):
        li = int(name.split('/')[0].split('-')[-1])
        if li == last: return 0.2
        if name.endswith('projection_kernel'): return 0.5
        return 1.0
    return _synth(shapes, seed, rule)


def load_checkpoint(model_path, epoch, verify_crc=True):
    """Weights of `model_path/epoch-<epoch>/variables/variables` (deepxi/model.py:279-280)."""
    prefix = os.path.join(model_path, 'epoch-%d' % epoch, 'variables', 'variables')
    return tfbundle.keras_weights(prefix, verify_crc=verify_crc)


def n_params(w):
    return int(sum(int(np.prod(np.shape(v))) for v in w.values()))
