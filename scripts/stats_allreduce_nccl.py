"""N-GPU check of the one collective on a Deep Xi path: every rank computes the (count, sum, sum of squares)
moments of xi_dB for ITS shard of a training sample with the CUDA kernels, the ranks all-reduce the 3 x 257 float64
tensor over NCCL, rank 0 compares the resulting (mu, sigma) with the oracle on the whole sample.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 scripts/stats_allreduce_nccl.py
"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist


def main():
    rank, world, local = int(os.environ['RANK']), int(os.environ['WORLD_SIZE']), int(os.environ['LOCAL_RANK'])
    torch.cuda.set_device(local)
    dist.init_process_group('nccl', device_id=torch.device('cuda', local))
    from deepxi_b200 import stats
    from deepxi_b200.inp_tgt import inp_tgt_selector
    from oracle import train_tgt
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests'))
    from test_train_tgt import _corpus
    s, d, s_len, d_len, snr, off = _corpus(n=16, seed=21)
    so, do, xo, _ = train_tgt.mix(s, d, s_len, d_len, snr, off)
    mine = list(range(rank, len(s_len), world))
    it = inp_tgt_selector('MagXi', 512, 256, 512, 16000, map_type='DBNormalCDF', map_params=None)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    acc = it.xi_db_moments(so[mine], do[mine], [s_len[i] for i in mine])
    torch.cuda.synchronize(); dist.barrier(device_ids=[local])
    e0.record()
    acc = stats.allreduce_moments(acc)
    e1.record(); torch.cuda.synchronize()
    mu, sg = stats.stats_from_moments(acc.cpu().numpy())
    if rank == 0:
        r_mu, r_sg = train_tgt.stats_from_moments(train_tgt.xi_db_moments(so, do, s_len))
        ok = np.abs(mu - r_mu).max() < 2e-3 and np.abs(sg - r_sg).max() < 2e-3 and acc[0, 0].item() == sum(-(-n // 256) for n in s_len)
        print('stats all-reduce over %d GPUs (NCCL): max |d mu| %.2e dB, max |d sigma| %.2e dB, all-reduce %.3f ms -> %s'
              % (world, np.abs(mu - r_mu).max(), np.abs(sg - r_sg).max(), e0.elapsed_time(e1), 'OK' if ok else 'MISMATCH'))
        assert ok
    dist.barrier(device_ids=[local])
    dist.destroy_process_group()


if __name__ == '__main__':
    main()
