"""Diagnostic: host<->device copy bandwidth per rank, alone and concurrently (torchrun)."""
import os, time, torch, torch.distributed as dist
rank, world, lr = int(os.environ.get('RANK', 0)), int(os.environ.get('WORLD_SIZE', 1)), int(os.environ.get('LOCAL_RANK', 0))
torch.cuda.set_device(lr)
if world > 1: dist.init_process_group('nccl', device_id=torch.device('cuda', lr))
n = 82 * 1024 * 1024
h = torch.empty(n, dtype=torch.uint8).pin_memory(); d = torch.empty(n, dtype=torch.uint8, device='cuda')
def bw(fn, reps=10):
    fn(); torch.cuda.synchronize()
    t = time.perf_counter()
    for _ in range(reps): fn()
    torch.cuda.synchronize()
    return n * reps / (time.perf_counter() - t) / 1e9
def bar():
    if world > 1: dist.barrier(device_ids=[lr])
for who in list(range(world)) + [-1]:
    bar()
    if who == -1 or who == rank:
        a = bw(lambda: d.copy_(h, non_blocking=True)); b = bw(lambda: h.copy_(d, non_blocking=True))
        s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
        def both():
            with torch.cuda.stream(s1): d.copy_(h, non_blocking=True)
            with torch.cuda.stream(s2): h.copy_(d, non_blocking=True)
        c = bw(both)
        print('rank %d (%s): H2D %.1f GB/s  D2H %.1f GB/s  duplex %.1f GB/s per direction' % (rank, 'all ranks at once' if who == -1 else 'alone', a, b, c), flush=True)
    bar()
print('rank', rank, 'cpu affinity', len(os.sched_getaffinity(0)), 'cores', flush=True)
