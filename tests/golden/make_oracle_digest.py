"""Regression digest of the oracle's network forwards on seeded inputs -> tests/golden/oracle_digest.json.

Not a reference pin (no reference artefact exists for the networks: the weight shards are missing from the tree):
it pins the ORACLE against accidental drift, so that a GPU parity test that starts failing can be attributed.  For every network:
shape, float64 sum, sum of squares and eight probe values of x_bar on 2 seeded utterances with the seeded synthetic weights.

    python tests/golden/make_oracle_digest.py
"""
import json, os, sys
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import numpy as np
from oracle import sig, tcn, attention
from deepxi_b200 import synth, weights


def digest(a):
    a = np.asarray(a, np.float64)
    flat = a.ravel()
    idx = np.linspace(0, flat.size - 1, 8).astype(int)
    return {'shape': list(a.shape), 'sum': float(flat.sum()), 'sumsq': float((flat * flat).sum()), 'probes': [float(v) for v in flat[idx]]}


def compute():
    inp, _, _ = sig.observation_batch(synth.noisy_speech(2, 9000, seed=77), [9000, 5000])
    out = {}
    out['ResNetV2/causal'] = digest(tcn.resnetv2_forward(inp, weights.synthetic_resnetv2(0), padding='causal'))
    out['ResNetV2/same'] = digest(tcn.resnetv2_forward(inp, weights.synthetic_resnetv2(0), padding='same'))
    out['ResNet/causal'] = digest(tcn.resnet_forward(inp, weights.synthetic_resnet(0), padding='causal'))
    out['ResNetV3/causal'] = digest(tcn.resnetv3_forward(inp, weights.synthetic_resnetv3(0), padding='causal'))
    out['MHANetV3/none'] = digest(attention.mhanetv3_forward(inp, weights.synthetic_mhanetv3(0)))
    return out


if __name__ == '__main__':
    d = compute()
    json.dump(d, open(os.path.join(HERE, 'oracle_digest.json'), 'w'), indent=1)
    print(json.dumps({k: v['sum'] for k, v in d.items()}, indent=1))
