"""Mirror of deepxi/network/attention.py MHANetV3 (:387-442): multi-head attention network with learned
positional embedding.  mask_mode 'none' reproduces the shipped model (tfa ignores the mask);
'causal+pad' applies the mask AttentionMaskV2 (:355-385) builds."""
from .. import _lib
from ._base import DeviceNetwork


class MHANetV3(DeviceNetwork):
    kind = 'MHANetV3'

    def __init__(self, inp=None, n_outp=257, d_model=256, n_blocks=5, n_heads=8, warmup_steps=40000, max_len=2048,
                 causal=True, outp_act='Sigmoid', n_feat=257, mask_mode='none', precision='f16x3'):
        if outp_act != 'Sigmoid':
            if outp_act in ('ReLU', 'Linear'):
                raise NotImplementedError('only the Sigmoid output activation of the committed models is built')
            raise ValueError('Invalid outp_act')
        if mask_mode not in _lib.MASK_MODES:
            raise ValueError("mask_mode must be 'none' or 'causal+pad'")
        cfg = _lib.NetCfg(n_feat=n_feat, n_outp=n_outp, d_model=d_model, n_blocks=n_blocks, d_f=0, k=0, max_d_rate=1,
                          padding=0, n_heads=n_heads, max_len=max_len, mask_mode=_lib.MASK_MODES[mask_mode], precision=0)
        self.max_len, self.causal, self.mask_mode = max_len, causal, mask_mode
        super().__init__(cfg, precision)
