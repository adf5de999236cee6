"""Bitwise comparison of the chained stage kernel against one-launch-per-stage (debug aid).
usage: chain_check.py [B]   -- spawns itself with different DXI_TCN_* settings and compares x_bar."""
import os, subprocess, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

def run(B, out):
    import torch
    from deepxi_b200 import weights, synth
    from deepxi_b200.network.selector import network_selector
    from deepxi_b200.inp_tgt import inp_tgt_selector
    kw = dict(d_model=256, n_blocks=40, d_f=64, k=3, max_d_rate=16, unit_type='ReLU->LN->W+b', outp_act='Sigmoid')
    net = network_selector('ResNetV2', None, 257, padding=os.environ.get('PAD', 'causal'), precision='f16x3', **kw).load_weights(weights.synthetic_resnetv2(0))
    x = np.tile(synth.noisy_speech(4, 160000, seed=51), (-(-B // 4), 1))[:B]
    it = inp_tgt_selector('MagXi', 512, 256, 512, 16000, map_type='DBNormalCDF', map_params=None)
    inp, _, _ = it.observation_batch(torch.from_numpy(x).cuda(), [160000] * B)
    y = net(inp)
    torch.cuda.synchronize()
    np.save(out, y.cpu().numpy())

if __name__ == '__main__':
    if len(sys.argv) > 2:
        run(int(sys.argv[1]), sys.argv[2]); sys.exit(0)
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
    variants = [('legacy', {'DXI_TCN_CHAIN': '0'}), ('chain', {}), ('chain again', {}), ('chain late-A1', {'DXI_TCN_DBGFLAGS': '8'})]
    outs = {}
    for name, env in variants:
        f = '/tmp/cc_%s.npy' % name.replace(' ', '_')
        e = dict(os.environ); e.update(env)
        r = subprocess.run([sys.executable, os.path.abspath(__file__), str(B), f], env=e, timeout=120)
        outs[name] = np.load(f) if r.returncode == 0 else None
    f = '/tmp/cc_ref4.npy'
    subprocess.run([sys.executable, os.path.abspath(__file__), '4', f], env=dict(os.environ), timeout=120)
    ref = np.tile(np.load(f), (-(-B // 4), 1, 1))[:B]      # utterances repeat with period 4; a 4-utterance batch has one tile per CTA
    for name, _ in variants:
        y = outs[name]
        if y is None:
            print('%-14s FAILED to run' % name); continue
        bad = np.argwhere(np.any(y != ref, axis=2))
        print('%-14s equal=%s  mismatching (utt, frame) rows: %d  max|d|=%.3g' % (name, bad.size == 0, len(bad), np.abs(y - ref).max()))
        if len(bad):
            utts = sorted(set(bad[:, 0].tolist()))
            print('   utterances:', utts[:20], ' first bad frame per utt:', [int(bad[bad[:, 0] == u][:, 1].min()) for u in utts[:20]])
