"""Mirror of deepxi/network/selector.py network_selector (:8-132).

Same keyword arguments as the reference.  ResNetV2 and MHANetV3 (the committed checkpoints) run on tcgen05; ResNet and ResNetV3
run in the exact fp32 mode; the remaining networks (MHANetV2, MHANet, RDLNet, ResNetV4, ResLSTM, ResBiLSTM) raise NotImplementedError; unknown names raise
ValueError('Invalid network type.') as selector.py:131 does.  `precision` ('f32' | 'f16x3' | 'f16') and
`mask_mode` are extensions.
"""

_OTHER = ('MHANetV2', 'MHANet', 'RDLNet', 'ResNetV4', 'ResBiLSTM', 'ResLSTM')


def network_selector(network_type, inp, n_outp, **kwargs):
    extra = {k: kwargs[k] for k in ('precision', 'mask_mode', 'n_feat') if k in kwargs}
    if network_type == 'MHANetV3':
        from .attention import MHANetV3
        return MHANetV3(inp=inp, n_outp=n_outp, d_model=kwargs['d_model'], n_blocks=kwargs['n_blocks'],
                        n_heads=kwargs['n_heads'], warmup_steps=kwargs.get('warmup_steps'), max_len=kwargs['max_len'],
                        causal=kwargs['causal'], outp_act=kwargs['outp_act'], **extra)
    if network_type == 'ResNetV2':
        from .tcn import ResNetV2
        extra.pop('mask_mode', None)
        return ResNetV2(inp=inp, n_outp=n_outp, n_blocks=kwargs['n_blocks'], d_model=kwargs['d_model'],
                        d_f=kwargs['d_f'], k=kwargs['k'], max_d_rate=kwargs['max_d_rate'], padding=kwargs['padding'],
                        unit_type=kwargs['unit_type'], outp_act=kwargs['outp_act'], **extra)
    if network_type in ('ResNet', 'ResNetV3'):      # selector.py:86-104
        from .tcn import ResNet, ResNetV3
        extra.pop('mask_mode', None)
        extra.setdefault('precision', 'f32')
        common = dict(inp=inp, n_outp=n_outp, n_blocks=kwargs['n_blocks'], d_model=kwargs['d_model'], d_f=kwargs['d_f'], k=kwargs['k'],
                      max_d_rate=kwargs['max_d_rate'], padding=kwargs['padding'], outp_act=kwargs['outp_act'], **extra)
        if network_type == 'ResNet':
            return ResNet(**common)
        return ResNetV3(unit_type=kwargs.get('unit_type', 'ReLU->LN->W+b'), **common)
    if network_type in _OTHER:
        raise NotImplementedError('%s has no committed checkpoint: out of scope (SURVEY 2)' % network_type)
    raise ValueError('Invalid network type.')
