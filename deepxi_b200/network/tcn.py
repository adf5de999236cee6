"""Mirror of deepxi/network/tcn.py: ResNetV2 (:116-225: bottleneck residual TCN with cyclic dilation, frame-wise layer
normalisation without affine parameters inside the blocks, unit "ReLU->LN->W+b"), and its siblings ResNet (:17-114, the
resnet-1.0c architecture) and ResNetV3 (:227-245), which run in the exact fp32 mode only."""
from .. import _lib
from ._base import DeviceNetwork


class ResNetV2(DeviceNetwork):
    kind = 'ResNetV2'

    def __init__(self, inp=None, n_outp=257, n_blocks=40, d_model=256, d_f=64, k=3, max_d_rate=16, padding='causal',
                 unit_type='ReLU->LN->W+b', outp_act='Sigmoid', n_feat=257, precision='f16x3'):
        if unit_type != 'ReLU->LN->W+b':
            if unit_type == 'LN->ReLU->W+b':
                raise NotImplementedError("unit_type 'LN->ReLU->W+b' is not used by the committed models")
            raise ValueError('Invalid unit_type.')
        if outp_act != 'Sigmoid':
            if outp_act in ('ReLU', 'Linear'):
                raise NotImplementedError('only the Sigmoid output activation of the committed models is built')
            raise ValueError('Invalid outp_act')
        if padding not in _lib.PADDINGS:
            raise ValueError("padding must be 'causal' or 'same'")
        cfg = _lib.NetCfg(n_feat=n_feat, n_outp=n_outp, d_model=d_model, n_blocks=n_blocks, d_f=d_f, k=k,
                          max_d_rate=max_d_rate, padding=_lib.PADDINGS[padding], n_heads=0, max_len=0, mask_mode=0,
                          precision=0)
        self.padding = padding
        super().__init__(cfg, precision)


class ResNetV3(ResNetV2):
    """ResNetV2 with the first layer Conv1D+b -> ReLU -> LayerNorm(no affine) (tcn.py:227-245).  precision 'f32' only."""
    kind = 'ResNetV3'

    def __init__(self, inp=None, n_outp=257, n_blocks=40, d_model=256, d_f=64, k=3, max_d_rate=16, padding='causal',
                 unit_type='ReLU->LN->W+b', outp_act='Sigmoid', n_feat=257, precision='f32'):
        super().__init__(inp, n_outp, n_blocks, d_model, d_f, k, max_d_rate, padding, unit_type, outp_act, n_feat, precision)


class ResNet(ResNetV2):
    """ResNet v1.0 (tcn.py:17-114): first layer Conv1D(no bias) -> LN(gamma, beta) -> ReLU, unit LN(gamma, beta) -> ReLU -> Conv1D,
    bias only in the third unit of a block.  precision 'f32' only."""
    kind = 'ResNet'

    def __init__(self, inp=None, n_outp=257, n_blocks=40, d_model=256, d_f=64, k=3, max_d_rate=16, padding='causal',
                 outp_act='Sigmoid', n_feat=257, precision='f32'):
        super().__init__(inp, n_outp, n_blocks, d_model, d_f, k, max_d_rate, padding, 'ReLU->LN->W+b', outp_act, n_feat, precision)
