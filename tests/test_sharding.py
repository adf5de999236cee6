"""Multi-process (world_size 2, gloo, CPU) test of the utterance sharding used by the N-GPU path: the
union of the ranks' outputs equals the single-process result bit for bit (there is no collective on the
data path, so nothing may change)."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from deepxi_b200 import shard, synth, weights, stats

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _infer_fn():
    from oracle import pipeline, tcn
    mu, sg = stats.packaged('resnet-1.1c')
    w = weights.synthetic_resnetv2(0, n_blocks=2)
    fwd = tcn.resnetv2_forward

    def run(x, lens):
        keep = pipeline.tcn.resnetv2_forward
        pipeline.tcn.resnetv2_forward = lambda inp, ww, padding='causal': fwd(inp, ww, n_blocks=2, padding=padding)
        try:
            return pipeline.infer(x, lens, w, mu, sg, out_type='y', gtype='mmse-lsa')
        finally:
            pipeline.tcn.resnetv2_forward = keep
    return run


def _worker(rank, world, port, x, lens, q):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    torch.set_num_threads(1)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    merged = shard.infer_sharded(_infer_fn(), x, lens, rank, world, gather=True)
    # max-over-ranks timing reduction, as bench.py does it
    t = torch.tensor([float(rank + 1)])
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        q.put((merged, float(t)))
    dist.barrier()
    dist.destroy_process_group()


def test_partition_covers_everything_once():
    lens = [5, 100, 7, 64, 64, 3, 99, 1, 50]
    for world in (1, 2, 4, 8):
        parts = shard.partition(lens, world)
        allidx = np.sort(np.concatenate(parts))
        assert np.array_equal(allidx, np.arange(len(lens)))
        assert max(len(p) for p in parts) - min(len(p) for p in parts) <= 1
    assert [len(p) for p in shard.contiguous(8192, 8)] == [1024] * 8


def test_two_rank_gloo_equals_single_process():
    lens = [3000, 1200, 2500, 700, 1900]
    x = synth.noisy_speech(len(lens), 3000, seed=71)
    torch.set_num_threads(1)
    single = _infer_fn()(x, lens)
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = 29600 + os.getpid() % 300
    procs = [ctx.Process(target=_worker, args=(r, 2, port, x, lens, q)) for r in range(2)]
    for p in procs:
        p.start()
    merged, tmax = q.get(timeout=240)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert tmax == 2.0 and sorted(merged) == list(range(len(lens)))
    for i in range(len(lens)):
        assert np.array_equal(merged[i], single[i])


@pytest.mark.gpu
def test_two_gpus_equal_one_gpu_bit_for_bit():
    """shard.infer_sharded over the CUDA path on 2 GPUs (one process per GPU, NCCL only gathers the outputs) must give the
    bits of the single-GPU run.  Needs two devices: skipped on a one-GPU box (run it with `gpurun --gpus 2`, scripts/shard_check.py)."""
    import subprocess, sys, torch
    if torch.cuda.device_count() < 2:
        pytest.skip('needs 2 GPUs')
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node', '2', '--master-addr', '127.0.0.1',
                        '--master-port', '29611', os.path.join(root, 'scripts', 'shard_check.py')], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and 'mismatching utterances: []' in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]
