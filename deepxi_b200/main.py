"""Mirror of main.py (:12-134) for the inference path:  python -m deepxi_b200.main --ver resnet-1.1c --infer 1 ...

Takes the flags run.sh passes (deepxi_b200/args.py), reads the noisy-speech directory with se_batch.Batch (main.py:45-46),
builds DeepXi with the same keyword names (main.py:51-59) and calls DeepXi.infer (the upstream call of main.py:122-133;
this fork's infer_noisy_wav / infer_hybrid experiments are out of scope).  --train / --test need TensorFlow-side
components that are out of scope and raise NotImplementedError."""
import sys

import numpy as np

from .args import get_args
from .model import DeepXi
from .se_batch import Batch


def network_kwargs(args):
    """Constructor arguments of the selected network from the flat flag namespace (what **vars(args) does in main.py:59)."""
    if args.network_type in ('ResNetV2', 'ResNetV3', 'ResNet'):
        kw = dict(d_model=args.d_model, n_blocks=args.n_blocks, d_f=args.d_f, k=args.k, max_d_rate=args.max_d_rate,
                  padding=args.padding, unit_type=args.unit_type, outp_act=args.outp_act)
    elif args.network_type in ('MHANetV3', 'MHANetV2'):
        kw = dict(d_model=args.d_model, n_blocks=args.n_blocks, n_heads=args.n_heads, warmup_steps=args.warmup_steps,
                  max_len=args.max_len, causal=args.causal, outp_act=args.outp_act, mask_mode=args.mask_mode)
    else:
        kw = {}
    kw = {k: v for k, v in kw.items() if v is not None}
    if args.inp_tgt_type == 'MagGain':
        kw['gain'] = args.gain[0] if isinstance(args.gain, list) else args.gain      # the gain the network was trained to estimate (inp_tgt.py:51-52)
    kw['precision'] = args.precision or ('f32' if args.network_type in ('ResNet', 'ResNetV3') else 'f16x3')
    return kw


def main(argv=None):
    args = get_args(argv)
    args.padding = 'causal' if args.causal else 'same'                                       # main.py:19-20
    args.model_path = args.model_path + '/' + args.ver                                       # main.py:22
    if args.set_path != 'set':
        args.data_path = args.data_path + '/' + args.set_path.rsplit('/', 1)[-1]             # main.py:23
    N_d = int(args.f_s * args.T_d * 0.001)                                                   # main.py:28-30
    N_s = int(args.f_s * args.T_s * 0.001)
    K = int(pow(2, np.ceil(np.log2(N_d))))
    if args.train or args.test or args.spect_dist or args.prelim:
        raise NotImplementedError('--train / --test / --spect_dist / --prelim are outside the inference hot path (SURVEY 2)')
    if not args.infer:
        return 0
    test_x, test_x_len, _, test_x_base_names = Batch(args.test_x_path, f_s=args.f_s)                       # main.py:45-46
    print('Version: %s.' % args.ver)
    map_type = args.map_type[0] if isinstance(args.map_type, list) else args.map_type
    map_params = args.map_params[0] if isinstance(args.map_params, list) and args.map_params and isinstance(args.map_params[0], list) else None
    deepxi = DeepXi(N_d=N_d, N_s=N_s, K=K, f_s=args.f_s, inp_tgt_type=args.inp_tgt_type, network_type=args.network_type,
                    min_snr=args.min_snr if args.min_snr is not None else -10, max_snr=args.max_snr if args.max_snr is not None else 20,
                    snr_inter=args.snr_inter or 1, log_path=args.log_path, sample_dir=args.data_path, ver=args.ver,
                    reset_inp_tgt=args.reset_inp_tgt, map_type=map_type, map_params=map_params, **network_kwargs(args))
    if args.synthetic_weights is not None:
        from . import weights
        make = {'ResNetV2': weights.synthetic_resnetv2, 'ResNetV3': weights.synthetic_resnetv3,
                'ResNet': weights.synthetic_resnet}.get(args.network_type, weights.synthetic_mhanetv3)
        deepxi.set_weights(make(args.synthetic_weights))
    deepxi.infer(test_x=test_x, test_x_len=test_x_len, test_x_base_names=test_x_base_names, test_epoch=args.test_epoch,
                 model_path=args.model_path, out_type=args.out_type, gain=args.gain, out_path=args.out_path,
                 n_filters=args.n_filters or 40, saved_data_path=args.saved_data_path)       # main.py:122-133
    return 0


if __name__ == '__main__':
    sys.exit(main())
