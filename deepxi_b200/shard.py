"""Utterance sharding for multi-GPU inference.

Every utterance is independent on the whole inference path (LayerNorm is per frame, convolutions and
attention are per utterance), so N GPUs simply process disjoint subsets: one process per GPU, weights
replicated, NO collective on the data path.  torch.distributed is only used (optionally) to gather
outputs / timings on rank 0.
"""
import numpy as np


def partition(lengths, world_size):
    """Length-bucketed round-robin: utterances sorted by decreasing length are dealt to the ranks in turn, so
    every rank gets the same count (+-1) and a similar padded length.  Returns a list of index arrays."""
    order = np.argsort(-np.asarray(lengths, np.int64), kind='stable')
    return [np.sort(order[r::world_size]) for r in range(world_size)]


def contiguous(n, world_size):
    """Contiguous n/world slices (what bench.py uses for its fixed-length synthetic corpus)."""
    edges = [(n * r) // world_size for r in range(world_size + 1)]
    return [np.arange(edges[r], edges[r + 1]) for r in range(world_size)]


def infer_sharded(infer_fn, x_batch, x_len, rank, world_size, gather=False, group=None):
    """Runs `infer_fn(x_shard, len_shard) -> list of per-utterance outputs` on this rank's shard.

    Returns {global utterance index: output}.  With gather=True the dictionaries of all ranks are merged on
    every rank with all_gather_object (host objects; the hot path itself never communicates)."""
    idx = partition(x_len, world_size)[rank]
    xs = np.asarray(x_batch)[idx]
    ls = [int(x_len[i]) for i in idx]
    outs = infer_fn(xs, ls) if len(idx) else []
    mine = {int(i): o for i, o in zip(idx, outs)}
    if not gather or world_size == 1:
        return mine
    import torch.distributed as dist
    parts = [None] * world_size
    dist.all_gather_object(parts, mine, group=group)
    merged = {}
    for p in parts:
        merged.update(p)
    return merged
