#!/usr/bin/env python
"""bench.py -- Deep Xi inference hot path on B200: audio-seconds enhanced per second.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
         bench.py --gpus N --steps K --warmup W

Workload (BASELINE.json configs[1]): ResNet-1.1c (ResNetV2, 40 blocks, causal) + MMSE-LSA, 256 synthetic
utterances x 10 s at 16 kHz per GPU, int16 waveform in -> enhanced waveform out, STFT 512/256.
One step = one pass of the whole hot path over that batch.  Utterances are independent, so the N-GPU run
shards them with no collective on the data path ("weak" scaling: 256 utterances per GPU).

Prints ONE JSON line (rank 0).  `value`: device-resident inputs, CUDA-event timed, max over ranks.
`e2e`: the same metric through the public API (DeepXi.infer_batch) with pinned HOST buffers, H2D and
D2H copies inside the timed region.  `roofline`: the dominant kernel (tcn_stage_kernel, tcgen05) against
the measured bf16 peak of MEASURED_PEAKS.json.  `cpu_baseline`: the oracle (CPU restatement of the
reference path; TensorFlow cannot be installed here) on a bounded sample, all host cores.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

F_S = 16000
SECONDS = 10
UTTS_PER_GPU = 256
FLOP_PER_FRAME_STAGES = 40 * 2 * (256 * 64 + 3 * 64 * 64 + 64 * 256)     # the 40 residual blocks (tensor cores)
FLOP_PER_FRAME_TOTAL = 3867648                                           # whole ResNetV2
STFT_BYTES_PER_FRAME = 512 + 2 * 257 * 4                                 # int16 in, mag + phase out
TCN_STAGE_BYTES_PER_FRAME = 2 * 1024 + 256 + 256                         # per stage: h read + write (fp32), c1 write, c1 read (once)
NCU_TCN_STAGE_TRAFFIC = 385.43e6      # dram__bytes_read.sum + dram__bytes_write.sum of one tcn_stage_kernel<true> launch (231.75 + 153.68 MB), profiles/r01_prof_tcn_stage_final.csv
RES_KW = dict(d_model=256, n_blocks=40, d_f=64, k=3, max_d_rate=16, unit_type='ReLU->LN->W+b', outp_act='Sigmoid')


def measured_peaks():
    p = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return d.get('hbm_gbs', 6650.0), d.get('bf16_tflops_sustained', d.get('bf16_tflops', 1590.0)), 'measured'
    return 6650.0, 1590.0, 'fallback'


class ClockSampler:
    """Samples SM clocks / throttle reasons with nvidia-smi while the timed region runs."""
    Q = ('index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,'
         'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')

    def __init__(self, gpu_index):
        self.idx, self.lines, self.proc = gpu_index, [], None

    def start(self):
        """Started BEFORE the warm-up steps: nvidia-smi needs ~0.1 s to produce its first line, more than a whole timed region of
        ten 4.5 ms steps.  Every line is stamped on arrival; stop() keeps the ones that arrived inside the timed windows."""
        try:
            self.proc = subprocess.Popen(['nvidia-smi', '-i', str(self.idx), '--query-gpu=' + self.Q,
                                          '--format=csv,noheader,nounits', '-lms', '20'], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.lines.append((time.time(), ln.strip()))

    def stop(self, windows=()):
        """windows: (t_begin, t_end) pairs of time.time() bracketing the timed regions (device-timed steps, e2e steps)."""
        if self.proc is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        time.sleep(0.05)
        self.proc.terminate()
        inside = [ln for t, ln in self.lines if any(a <= t <= b + 0.01 for a, b in windows)]
        window = 'timed regions (device-timed steps + e2e steps)'
        if not inside:      # a timed region shorter than one sampling period: the samples of the same load just around it
            lo = min((a for a, _ in windows), default=0.0) - 0.5
            inside = [ln for t, ln in self.lines if t >= lo]
            window = 'no sample fell inside the timed regions: samples from 0.5 s before them (warm-up steps, same load) to their end'
        sm, mx, reasons = [], [], set()
        for ln in inside:
            f = [x.strip() for x in ln.split(',')]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(('hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap'), f[5:9]):
                if v.lower().startswith('active'):
                    reasons.add(name)
        busy = [s for s in sm if s > 0]
        return {'sm_mhz': float(np.median(busy)) if busy else None, 'sm_max_mhz': max(mx) if mx else None,
                'reasons': sorted(reasons), 'samples': len(sm), 'window': window}


def run_reference(args, rank, world):
    """CPU arm: the oracle port of the reference path on this box's host cores (rank 0 only)."""
    if rank != 0:
        return
    import torch
    from oracle import pipeline
    from deepxi_b200 import synth, weights, stats
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    mu, sigma = stats.packaged('resnet-1.1c')
    w = weights.synthetic_resnetv2(0)
    n_utt = args.ref_utts
    x = synth.noisy_speech(n_utt, SECONDS * F_S, seed=1234)
    lens = [SECONDS * F_S] * n_utt
    for _ in range(args.warmup):
        pipeline.infer(x[:2], lens[:2], w, mu, sigma, out_type='y', gtype='mmse-lsa')
    t0 = time.perf_counter()
    for _ in range(args.steps):
        pipeline.infer(x, lens, w, mu, sigma, out_type='y', gtype='mmse-lsa')
    dt = (time.perf_counter() - t0) / args.steps
    v = n_utt * SECONDS / dt
    sample = '%d x %d s utterances per step (ResNet-1.1c + MMSE-LSA, y out)' % (n_utt, SECONDS)
    line = {'impl': 'reference', 'metric': 'audio-seconds enhanced per second (ResNet-1.1c, MMSE-LSA)', 'value': v,
            'unit': 'audio-s/s', 'n_gpus': args.gpus, 'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': dt * 1e3,
            'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
            'config': {'workload': 'ResNet-1.1c + MMSE-LSA, %d utt x %d s @16 kHz per GPU, STFT 512/256, int16 in -> wav out'
                                   % (UTTS_PER_GPU, SECONDS), 'arm': 'CPU restatement of the reference path (oracle/, numpy + '
                       'torch-CPU + scipy); the TF2 reference cannot be installed or imported here'},
            'cpu_baseline': {'value': v, 'unit': 'audio-s/s', 'cores': cores, 'kind': 'port', 'sample': sample},
            'e2e': {'value': v, 'unit': 'audio-s/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
            'gpu_launches': 0}
    print(json.dumps(line), flush=True)


def cpu_baseline(n_utt=8):
    import torch
    from oracle import pipeline
    from deepxi_b200 import synth, weights, stats
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    mu, sigma = stats.packaged('resnet-1.1c')
    w = weights.synthetic_resnetv2(0)
    x = synth.noisy_speech(n_utt, SECONDS * F_S, seed=1234)
    lens = [SECONDS * F_S] * n_utt
    pipeline.infer(x[:1], lens[:1], w, mu, sigma)            # warm-up
    t0 = time.perf_counter()
    reps = 0
    while reps < 2 or time.perf_counter() - t0 < 10.0:
        pipeline.infer(x, lens, w, mu, sigma, out_type='y', gtype='mmse-lsa')
        reps += 1
        if time.perf_counter() - t0 > 30.0:
            break
    dt = (time.perf_counter() - t0) / reps
    return {'value': n_utt * SECONDS / dt, 'unit': 'audio-s/s', 'cores': cores, 'kind': 'port',
            'sample': '%d x %d s utterances, %d repetitions, oracle (numpy rfft + torch-CPU fp32 ResNetV2 + scipy exp1)'
                      % (n_utt, SECONDS, reps)}


def bind_to_gpu_numa_node(local_rank):
    """Pins this process (and therefore the pinned host buffers it allocates next: first touch) to the CPUs of the NUMA node
    the GPU hangs off, so that the e2e arm's host<->device copies of the N ranks do not cross the socket interconnect.
    Best effort: returns a short description, or None when the topology cannot be read."""
    try:
        import torch
        pr = torch.cuda.get_device_properties(local_rank)
        bdf = '%04x:%02x:%02x.0' % (pr.pci_domain_id, pr.pci_bus_id, pr.pci_device_id)
        with open('/sys/bus/pci/devices/%s/numa_node' % bdf) as f:
            node = int(f.read().strip())
        if node < 0:
            return None
        with open('/sys/devices/system/node/node%d/cpulist' % node) as f:
            cpus = set()
            for part in f.read().strip().split(','):
                lo, _, hi = part.partition('-')
                cpus.update(range(int(lo), int(hi or lo) + 1))
        cpus &= os.sched_getaffinity(0)
        if not cpus:
            return None
        os.sched_setaffinity(0, cpus)
        return 'numa node %d (%d cpus)' % (node, len(cpus))
    except Exception:
        return None


def run_ours(args, rank, world, local_rank):
    import torch
    import torch.distributed as dist
    from deepxi_b200 import synth, weights, _lib
    from deepxi_b200.model import DeepXi

    torch.cuda.set_device(local_rank)
    dev = torch.device('cuda', local_rank)
    numa = bind_to_gpu_numa_node(local_rank) if not os.environ.get('DXI_BENCH_NO_NUMA') else None
    if world > 1 and not dist.is_initialized():
        # NCCL prints a version banner on stdout when the communicator is created; stdout must carry exactly
        # ONE JSON line, so fd 1 points at stderr until the first collective has run.
        sys.stdout.flush()
        keep = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group('nccl', device_id=dev)
            dist.barrier(device_ids=[local_rank])
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(keep, 1)
            os.close(keep)

    def barrier():
        if world > 1:
            dist.barrier(device_ids=[local_rank])

    B, L = args.utts, SECONDS * F_S
    T = -(-L // 256)
    dx = DeepXi(512, 256, 512, F_S, 'MagXi', 'ResNetV2', ver='resnet-1.1c', map_type='DBNormalCDF', map_params=None,
                padding='causal', precision=args.precision, **RES_KW)
    dx.set_weights(weights.synthetic_resnetv2(0))
    # every rank gets its own shard of the (synthetic) corpus: utterances rank*B .. rank*B+B-1
    base = synth.noisy_speech(min(B, 32), L, seed=1234 + rank)
    x_host = torch.from_numpy(np.tile(base, (-(-B // base.shape[0]), 1))[:B].copy()).pin_memory()
    lens = [L] * B
    x_dev = [x_host.to(dev), x_host.to(dev).roll(1, 0)]          # two input buffers, alternated
    it = dx.inp_tgt

    def step(i):
        inp, pha, nfr = it.observation_batch(x_dev[i & 1], lens)
        xbar = dx.network(inp)
        return it.enhanced_speech(inp, pha, xbar, 'mmse-lsa', n_frames=None)

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    for i in range(args.warmup):
        y = step(i)
    torch.cuda.synchronize()
    _lib.profile_enable(True)
    for k in ('stft', 'tcn_stage', 'tcn_stem', 'tcn_head', 'enhance'):
        _lib.profile_read(k)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier(); torch.cuda.synchronize()
    _lib.launch_count_reset()
    win_dev0 = time.time()
    e0.record()
    for i in range(args.steps):
        y = step(i)
    e1.record()
    torch.cuda.synchronize()
    win_dev1 = time.time()
    barrier()
    launches = _lib.launch_count()
    ms = e0.elapsed_time(e1)
    prof = {k: _lib.profile_read(k) for k in ('stft', 'tcn_stage', 'tcn_stem', 'tcn_head', 'enhance')}
    _lib.profile_enable(False)
    assert torch.isfinite(y).all(), 'non-finite samples in the enhanced waveform'

    # ---- e2e: pinned host int16 in -> public API -> pinned host int16 out, copies inside the timed region.
    # HostPipeline chains H2D -> kernels -> D2H of every step over three streams so that the PCIe copies of one step overlap the
    # kernels of its neighbours; every step still copies its whole input in and its whole result out.
    from deepxi_b200.model import HostPipeline
    pipe = HostPipeline(dx, n_streams=3)
    y_hosts = [torch.empty((B, (T + 1) * 256), dtype=torch.int16).pin_memory() for _ in range(3)]
    for i in range(max(6, args.warmup)):          # two rounds per stream: the allocator pools of all streams settle
        pipe.submit(x_host, lens, y_hosts[i % 3])
    pipe.drain()
    torch.cuda.synchronize(); barrier()
    win_e2e0 = time.time()
    t0 = time.perf_counter()
    marks = []
    for i in range(args.steps):
        pipe.submit(x_host, lens, y_hosts[i % 3])
        marks.append(time.perf_counter() - t0)
    pipe.drain()
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    clocks = sampler.stop([(win_dev0, win_dev1), (win_e2e0, time.time())]) if rank == 0 else None
    if os.environ.get('DXI_BENCH_DEBUG'):
        sys.stderr.write('rank %d e2e submit marks %s total %.4f\n' % (rank, ['%.4f' % m for m in marks], e2e_s))
    barrier()
    assert int(y_hosts[0].abs().max()) > 0

    t = torch.tensor([ms, e2e_s * 1e3], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms, e2e_ms = float(t[0]), float(t[1])
    if rank != 0:
        return
    audio_s = world * B * SECONDS
    value = audio_s * args.steps / (ms / 1e3)
    hbm_peak, tf_peak, peak_src = measured_peaks()
    frames = B * T
    st_ms, st_n = prof['tcn_stage']
    st_ms_per_launch = st_ms / max(st_n, 1)
    st_per_step = st_n / max(args.steps, 1)          # 41 (one launch per stage) or the number of utterance groups (chained kernel)
    flops_per_launch = frames * FLOP_PER_FRAME_STAGES / max(st_per_step, 1)
    achieved_tf = flops_per_launch / (st_ms_per_launch * 1e-3) / 1e12 if st_n else None
    # the same launches seen from the memory side: per frame and stage the residual stream is read and written (2 x 1 KB fp32) and
    # c1 is written once (fp16 hi + lo, 256 B) and fetched from HBM once (256 B; its other two taps are L2 hits)
    st_bytes_per_launch = frames * TCN_STAGE_BYTES_PER_FRAME * 41.0 / max(st_per_step, 1)
    st_gbs = st_bytes_per_launch / (st_ms_per_launch * 1e-3) / 1e9 if st_n else None
    # dram__bytes_read + dram__bytes_write of one launch from the committed ncu --set full capture (C2 shape, one launch per stage)
    traffic = NCU_TCN_STAGE_TRAFFIC if (frames == 160000 and abs(st_per_step - 41) < 0.5 and args.precision == 'f16x3') else None
    stft_ms, stft_n = prof['stft']
    stft_gbs = frames * STFT_BYTES_PER_FRAME / (stft_ms / max(stft_n, 1) * 1e-3) / 1e9 if stft_n else None
    enh_ms, enh_n = prof['enhance']
    line = {
        'metric': 'audio-seconds enhanced per second (ResNet-1.1c, MMSE-LSA)', 'value': value, 'unit': 'audio-s/s',
        'n_gpus': world, 'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': ms / args.steps,
        'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
        'dtype': {'f16x3': 'f16x3 (fp16 hi/lo split tensor-core operands, fp32 accumulate; fp32 elsewhere)',
                  'f16': 'f16 tensor-core operands, fp32 accumulate', 'f32': 'f32'}[args.precision],
        'data': 'synthetic',
        'config': {'workload': 'ResNet-1.1c (ResNetV2 40 blocks causal, random-init weights of the checkpoint shapes) + '
                               'MMSE-LSA, %d utt x %d s @16 kHz per GPU, STFT 512/256, int16 waveform in -> f32 waveform out'
                               % (B, SECONDS),
                   'utterances_per_gpu': B, 'frames_per_gpu': frames, 'precision': args.precision,
                   'sharding': 'utterances split across ranks, no data-path collective',
                   'l2': 'per-step working set %.2f GB per GPU (two alternating input buffers) exceeds the 126 MB L2'
                         % ((B * L * 2 + 3 * frames * 257 * 4 + frames * 1024 + B * (T + 1) * 1024) / 1e9)},
        'realtime_factor_per_gpu': value / world,
        'clocks': clocks,
        'e2e': {'value': audio_s * args.steps / (e2e_ms / 1e3), 'unit': 'audio-s/s',
                'h2d_bytes_per_step': B * L * 2, 'd2h_bytes_per_step': B * (T + 1) * 256 * 2,
                'api': 'HostPipeline(DeepXi).submit(pinned int16 in, lens, pinned int16 out): H2D, DeepXi.infer_batch and D2H on three event-chained streams, 3 batches in flight',
                'host_binding': numa},
        'gpu_launches': launches,
        'roofline': {'kernel': 'tcn_stage_kernel (tcgen05 / TMEM, %g launches per step)' % st_per_step, 'bound': 'tensor',
                     'achieved': achieved_tf, 'peak': tf_peak, 'unit': 'TFLOP/s',
                     'frac': (achieved_tf / tf_peak) if achieved_tf else None, 'traffic': traffic,
                     'traffic_source': 'profiles/r01_prof_tcn_stage_final.csv (bytes per launch; algorithmic %.0f)' % st_bytes_per_launch,
                     'hbm_view': {'bound': 'hbm', 'achieved': st_gbs, 'peak': hbm_peak, 'unit': 'GB/s',
                                  'frac': (st_gbs / hbm_peak) if st_gbs else None, 'bytes_per_frame_and_stage': TCN_STAGE_BYTES_PER_FRAME,
                                  'note': 'one stage per launch, chained tile by tile through flags: the kernel streams the fp32 residual '
                                          'through HBM at loaded-latency speed; this is the bound that binds today, the tensor figure is the target'},
                     'peak_source': peak_src + ' bf16 sustained (MEASURED_PEAKS.json)',
                     'algorithmic_flop_per_launch': flops_per_launch, 'ms_per_launch': st_ms_per_launch,
                     'share_of_step': st_ms / ms if ms else None,
                     'note': 'useful FLOPs only: in f16x3 mode the tensor cores execute 3x this'},
        'kernels_ms_per_step': {k: (v[0] / args.steps) for k, v in prof.items()},
        'stft': {'bound': 'hbm', 'achieved': stft_gbs, 'peak': hbm_peak, 'unit': 'GB/s',
                 'frac': (stft_gbs / hbm_peak) if stft_gbs else None, 'bytes_per_frame': STFT_BYTES_PER_FRAME},
        'enhance_gbs': (frames * (3 * 1028 + 1024) / (enh_ms / max(enh_n, 1) * 1e-3) / 1e9) if enh_n else None,
    }
    if world == 1 and not args.no_cpu_baseline:
        line['cpu_baseline'] = cpu_baseline(args.ref_utts)
    print(json.dumps(line), flush=True)


def main():
    os.environ.setdefault('NCCL_DEBUG', 'WARN')      # keep NCCL's version banner off stdout: rank 0 prints ONE JSON line
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=10)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--precision', default='f16x3', choices=['f16x3', 'f16', 'f32'])
    ap.add_argument('--utts', type=int, default=UTTS_PER_GPU, help='utterances per GPU')
    ap.add_argument('--ref-utts', type=int, default=8, help='utterances per step of the CPU arm')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    args = ap.parse_args()
    rank = int(os.environ.get('RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    if args.impl == 'reference':
        return run_reference(args, rank, world)
    if world == 1 and args.gpus > 1:
        # not launched by torchrun: re-launch ourselves with one process per GPU
        cmd = [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node', str(args.gpus),
               '--master-addr', '127.0.0.1', '--master-port', str(29500 + os.getpid() % 1000), os.path.abspath(__file__)] + sys.argv[1:]
        return subprocess.call(cmd)
    run_ours(args, rank, world, local_rank)
    if world > 1:
        import torch.distributed as dist
        if dist.is_initialized():
            dist.destroy_process_group()


if __name__ == '__main__':
    sys.exit(main())
