"""N-GPU == 1-GPU, bit for bit: every rank runs shard.infer_sharded over the CUDA path on its own GPU (one
process per GPU, NCCL only to gather the per-utterance outputs), rank 0 also runs the whole batch on its GPU alone and compares.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 scripts/shard_check.py
"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist
from deepxi_b200 import synth, weights, shard
from deepxi_b200.model import DeepXi

rank, world, local = int(os.environ['RANK']), int(os.environ['WORLD_SIZE']), int(os.environ['LOCAL_RANK'])
torch.cuda.set_device(local)
dist.init_process_group('nccl', device_id=torch.device('cuda', local))
kw = dict(d_model=256, n_blocks=40, d_f=64, k=3, max_d_rate=16, unit_type='ReLU->LN->W+b', outp_act='Sigmoid')
dx = DeepXi(512, 256, 512, 16000, 'MagXi', 'ResNetV2', ver='resnet-1.1c', map_type='DBNormalCDF', map_params=None,
            padding='causal', precision='f16x3', **kw)
dx.set_weights(weights.synthetic_resnetv2(0))
B = 24
lens = [160000 - 7001 * (i % 5) for i in range(B)]          # ragged: 10 s down to 8.25 s, several tiles per utterance
x = synth.noisy_speech(B, 160000, seed=17)


def infer_fn(xs, ls):
    # Every shard is padded to the CORPUS maximum, as the reference pads a batch to its longest utterance (model.py:2246-2253):
    # the padded length is part of the arithmetic of a frame-synchronous causal network only through which frames exist.
    y, nfr = dx.infer_batch(np.ascontiguousarray(xs), ls, 'y', 'mmse-lsa', int16=True)
    y = y.cpu().numpy()
    return [y[i, :(n + 1) * 256].copy() for i, n in enumerate(nfr)]


mine = shard.infer_sharded(infer_fn, x, lens, rank, world, gather=True)
if rank == 0:
    ref = infer_fn(x, lens)
    assert sorted(mine) == list(range(B))
    bad = [i for i in range(B) if not np.array_equal(mine[i], ref[i])]
    print('shard_check: %d ranks, %d utterances, mismatching utterances: %s' % (world, B, bad), flush=True)
    assert not bad
dist.barrier(device_ids=[local])
dist.destroy_process_group()
