"""Mirror of deepxi/args.py (:8-133): the command-line flags of main.py / run.sh, same names, types and defaults, so that
`python -m deepxi_b200.main <the flags run.sh passes>` is a drop-in for `python3 main.py ...` on the inference path.
Flags that only steer training are accepted and carried (run.sh always passes them) but have no effect here."""
import argparse
import math


def read_dtype(x):
    """args.py:11-19."""
    x = x.replace('neg_', '-')
    if x == 'pi':
        return math.pi
    if x == '-pi':
        return -math.pi
    if any(map(str.isdigit, x)):
        return float(x) if '.' in x else int(x)
    return x


def str_to_list(x):
    """args.py:21-24: 'a,b;c,d' -> [[a, b], [c, d]], 'a,b' -> [a, b], 'a' -> a."""
    if ';' in x:
        return [[read_dtype(z) for z in y.split(',')] for y in x.split(';')]
    if ',' in x:
        return [read_dtype(y) for y in x.split(',')]
    return read_dtype(x)


def str_to_bool(s):
    """args.py:26."""
    return s.lower() in ('yes', 'true', 't', '1')


def get_parser():
    p = argparse.ArgumentParser(prog='deepxi_b200.main')
    # general (args.py:32-43)
    p.add_argument('--gpu', default='0', type=str, help='GPU selection')
    p.add_argument('--ver', type=str, help='Model version')
    p.add_argument('--test_epoch', type=str_to_list, help='Epoch to test')
    p.add_argument('--train', default=False, type=str_to_bool, help='Perform training')
    p.add_argument('--infer', default=False, type=str_to_bool, help='Perform inference and save outputs')
    p.add_argument('--test', default=False, type=str_to_bool, help='Evaluate using objective measures')
    p.add_argument('--spect_dist', default=False, type=str_to_bool, help='Find spectral distortion')
    p.add_argument('--prelim', default=False, type=str_to_bool, help='Preliminary flag')
    p.add_argument('--verbose', default=False, type=str_to_bool, help='Verbose')
    p.add_argument('--network_type', type=str, help='Network type')
    p.add_argument('--inp_tgt_type', type=str, help='Input and target type')
    p.add_argument('--sd_snr_levels', default=[-5, 0, 5, 10, 15], type=str_to_list, help='SNR levels for spectral distortion')
    # training (args.py:46-57)
    p.add_argument('--mbatch_size', type=int, help='Mini-batch size')
    p.add_argument('--sample_size', type=int, help='Sample size')
    p.add_argument('--max_epochs', type=int, help='Maximum number of epochs')
    p.add_argument('--resume_epoch', type=int, help='Epoch to resume training from')
    p.add_argument('--save_model', default=False, type=str_to_bool)
    p.add_argument('--log_iter', default=False, type=str_to_bool)
    p.add_argument('--eval_example', default=False, type=str_to_bool)
    p.add_argument('--val_flag', default=True, type=str_to_bool)
    p.add_argument('--reset_inp_tgt', default=False, type=str_to_bool)
    p.add_argument('--reset_sample', default=False, type=str_to_bool)
    # inference output (args.py:65-74)
    p.add_argument('--out_type', default='y', type=str, help='Output type for testing')
    p.add_argument('--gain', type=str_to_list, help='Gain function for testing')
    # paths (args.py:77-85)
    p.add_argument('--model_path', default='model', type=str)
    p.add_argument('--set_path', default='set', type=str)
    p.add_argument('--log_path', default='log', type=str)
    p.add_argument('--data_path', default='data', type=str)
    p.add_argument('--test_x_path', default='set/test_noisy_speech', type=str)
    p.add_argument('--test_s_path', default='set/test_clean_speech', type=str)
    p.add_argument('--test_d_path', default='set/test_noise', type=str)
    p.add_argument('--out_path', default='out', type=str)
    p.add_argument('--saved_data_path', default=None, type=str)
    # features (args.py:88-94)
    p.add_argument('--min_snr', type=int)
    p.add_argument('--max_snr', type=int)
    p.add_argument('--snr_inter', type=int)
    p.add_argument('--f_s', type=int, help='Sampling frequency (Hz)')
    p.add_argument('--T_d', type=int, help='Window duration (ms)')
    p.add_argument('--T_s', type=int, help='Window shift (ms)')
    p.add_argument('--n_filters', default=None, type=int)
    # network (args.py:97-117)
    for name in ('d_in', 'd_out', 'd_model', 'n_blocks', 'n_heads', 'warmup_steps', 'max_len', 'Noutp'):
        p.add_argument('--' + name, type=int)
    for name in ('d_b', 'd_f', 'd_ff', 'k', 'max_d_rate', 'length', 'm_1'):
        p.add_argument('--' + name, default=None, type=int)
    p.add_argument('--causal', type=str_to_bool)
    p.add_argument('--centre', type=str_to_bool)
    p.add_argument('--scale', type=str_to_bool)
    p.add_argument('--unit_type', type=str)
    p.add_argument('--loss_fnc', type=str)
    p.add_argument('--outp_act', type=str)
    # map (args.py:120-121)
    p.add_argument('--map_type', type=str_to_list)
    p.add_argument('--map_params', default=[None, None], type=str_to_list)
    # deepxi_b200 only
    p.add_argument('--precision', default=None, type=str, help="arithmetic of the network: 'f16x3' (tcgen05; default for ResNetV2 / MHANetV3) | 'f32' (default and only mode of ResNet / ResNetV3) | 'f16'")
    p.add_argument('--mask_mode', default='none', type=str, help="MHANetV3: 'none' (shipped behaviour) | 'causal+pad'")
    p.add_argument('--synthetic_weights', default=None, type=int, help='seed of random weights in the checkpoint shapes (the reference tree ships no weight shards); default: load model_path')
    return p


def get_args(argv=None):
    return get_parser().parse_args(argv)
