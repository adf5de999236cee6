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
D2H copies inside the timed region; `e2e.pcie_ceiling_gbs` is a duplex pinned-copy measurement taken by all ranks at
once inside the run, `e2e.frac_of_pcie_bound` the e2e rate against the bound that ceiling implies.  `roofline`: the
dominant kernel (tcn_chain_kernel: the 40 residual blocks, tcgen05 / TMEM) against the measured bf16 peak of
MEASURED_PEAKS.json.  `configs`: the other BASELINE.json configs measured in the same run (C1 latency, C3 MHANet,
the 1024-utterance shard of C5) and `sustained`: the same step looped for >= 3 s with its clock record.
`cpu_baseline`: the oracle (CPU restatement of the reference path; TensorFlow cannot be installed here) on a
bounded sample, all host cores.
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
# tcn_chain_kernel (whole network in one launch), algorithmic HBM bytes per frame: |X| in (1028) + x_bar out (1028) + per block the 32
# halo rows of c1 (fp16 hi + lo) a tile writes for, and reads from, its neighbour in the utterance (2 x 8 KB per 128 frames = 128 B per
# frame and block).  The fp32 residual stream itself never leaves the SM.
TCN_CHAIN_BYTES_PER_FRAME = 1028 + 1028 + 40 * 128
# dram__bytes_read.sum + dram__bytes_write.sum of the one tcn_chain_kernel<true> launch of this workload, from the committed
# ncu --set full capture (profiles/r02_prof_tcn_chain.csv); a citation, NOT measured in the run
NCU_TCN_CHAIN_TRAFFIC = 1235.4e6      # 732.3 MB read + 503.1 MB written
MHA_KW = dict(d_model=256, n_blocks=5, n_heads=8, warmup_steps=40000, max_len=2048, causal=1, outp_act='Sigmoid')
RES_KW = dict(d_model=256, n_blocks=40, d_f=64, k=3, max_d_rate=16, unit_type='ReLU->LN->W+b', outp_act='Sigmoid')


def measured_peaks():
    p = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return d.get('hbm_gbs', 6650.0), d.get('bf16_tflops', 1590.0), d.get('bf16_tflops_sustained', d.get('bf16_tflops', 1590.0)), 'measured'
    return 6650.0, 1590.0, 1590.0, 'fallback'


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
        return self.summarise(self.lines, windows)

    @staticmethod
    def summarise(lines, windows):
        inside = [ln for t, ln in lines if any(a <= t <= b + 0.01 for a, b in windows)]
        window = 'timed regions'
        if not inside:      # a timed region shorter than one sampling period: the samples of the same load just around it
            lo = min((a for a, _ in windows), default=0.0) - 0.5
            hi = max((b for _, b in windows), default=0.0) + 0.05
            inside = [ln for t, ln in lines if lo <= t <= hi]
            window = 'no sample fell inside the timed regions: samples from 0.5 s before them (warm-up steps, same load) to their end'
        sm, mx, pw, reasons = [], [], [], set()
        for ln in inside:
            f = [x.strip() for x in ln.split(',')]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2])); pw.append(float(f[3]))
            except ValueError:
                continue
            for name, v in zip(('hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap'), f[5:9]):
                if v.lower().startswith('active'):
                    reasons.add(name)
        busy = [s for s in sm if s > 0]
        return {'sm_mhz': float(np.median(busy)) if busy else None, 'sm_min_mhz': min(busy) if busy else None, 'sm_max_mhz': max(mx) if mx else None,
                'power_w_max': max(pw) if pw else None, 'reasons': sorted(reasons), 'samples': len(sm), 'window': window}


WORKLOAD = ('ResNet-1.1c (ResNetV2 40 blocks causal, random-init weights of the checkpoint shapes) + MMSE-LSA, %d utt x %d s '
            '@16 kHz per GPU, STFT 512/256, int16 waveform in -> enhanced waveform out')


def run_reference(args, rank, world):
    """CPU arm: the oracle port of the reference path on this box's host cores (rank 0 only).  The metric is a rate (audio
    seconds per second), so each step times a BOUNDED SAMPLE of the workload: `config.sample` says how many utterances."""
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
    sample = ('%d of the %d utterances x %d s per step (the oracle is a per-utterance loop: its rate does not depend on the batch '
              'size); ResNet-1.1c + MMSE-LSA, int16 in -> waveform out' % (n_utt, args.utts, SECONDS))
    line = {'impl': 'reference', 'metric': 'audio-seconds enhanced per second (ResNet-1.1c, MMSE-LSA)', 'value': v,
            'unit': 'audio-s/s', 'n_gpus': args.gpus, 'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': dt * 1e3,
            'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
            'config': {'workload': WORKLOAD % (args.utts, SECONDS), 'sample': sample, 'utterances_timed_per_step': n_utt,
                       'arm': 'CPU restatement of the reference path (oracle/, numpy + torch-CPU + scipy); the TF2 reference cannot '
                              'be installed or imported here'},
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
    Best effort: returns a short description of what was done or why nothing could be."""
    try:
        import torch
        pr = torch.cuda.get_device_properties(local_rank)
        bdf = '%04x:%02x:%02x.0' % (pr.pci_domain_id, pr.pci_bus_id, pr.pci_device_id)
        with open('/sys/bus/pci/devices/%s/numa_node' % bdf) as f:
            node = int(f.read().strip())
        if node < 0:
            return 'not bound: the VM reports no NUMA node for GPU %s' % bdf
        with open('/sys/devices/system/node/node%d/cpulist' % node) as f:
            cpus = set()
            for part in f.read().strip().split(','):
                lo, _, hi = part.partition('-')
                cpus.update(range(int(lo), int(hi or lo) + 1))
        cpus &= os.sched_getaffinity(0)
        if not cpus:
            return 'not bound: no allowed cpu on numa node %d' % node
        os.sched_setaffinity(0, cpus)
        return 'numa node %d (%d cpus)' % (node, len(cpus))
    except Exception as e:
        return 'not bound: %s' % type(e).__name__


def pcie_ceiling(torch, x_host, y_host, dev, barrier, seconds=0.2):
    """Duplex pinned-copy rate of this rank while ALL ranks copy at once: the same two buffers the e2e arm moves, H2D on one
    stream and D2H on another, back to back for `seconds`.  Returns (h2d GB/s, d2h GB/s)."""
    s_in, s_out = torch.cuda.Stream(), torch.cuda.Stream()
    xd = torch.empty_like(x_host, device=dev)
    yd = torch.empty_like(y_host, device=dev)
    for _ in range(2):
        with torch.cuda.stream(s_in):
            xd.copy_(x_host, non_blocking=True)
        with torch.cuda.stream(s_out):
            y_host.copy_(yd, non_blocking=True)
    torch.cuda.synchronize()
    barrier()
    n = 0
    a0, a1, b0, b1 = (torch.cuda.Event(enable_timing=True) for _ in range(4))
    a0.record(s_in); b0.record(s_out)
    t0 = time.perf_counter()
    while time.perf_counter() - t0 < seconds:
        for _ in range(4):
            with torch.cuda.stream(s_in):
                xd.copy_(x_host, non_blocking=True)
            with torch.cuda.stream(s_out):
                y_host.copy_(yd, non_blocking=True)
        n += 4
        s_in.synchronize(); s_out.synchronize()
    a1.record(s_in); b1.record(s_out)
    torch.cuda.synchronize()
    return (n * x_host.numel() * x_host.element_size() / (a0.elapsed_time(a1) * 1e-3) / 1e9,
            n * y_host.numel() * y_host.element_size() / (b0.elapsed_time(b1) * 1e-3) / 1e9)


def run_ours(args, rank, world, local_rank):
    import torch
    import torch.distributed as dist
    from deepxi_b200 import synth, weights, _lib
    from deepxi_b200.model import DeepXi

    torch.cuda.set_device(local_rank)
    dev = torch.device('cuda', local_rank)
    numa = bind_to_gpu_numa_node(local_rank) if not os.environ.get('DXI_BENCH_NO_NUMA') else 'not bound: DXI_BENCH_NO_NUMA'
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
    KEYS = ('stft', 'tcn_chain', 'tcn_stage', 'tcn_stem', 'tcn_head', 'enhance')

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
    for k in KEYS:
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
    prof = {k: _lib.profile_read(k) for k in KEYS}
    _lib.profile_enable(False)
    assert torch.isfinite(y).all(), 'non-finite samples in the enhanced waveform'

    # ---- e2e: pinned host int16 in -> public API -> pinned host int16 out, copies inside the timed region.
    # HostPipeline chains H2D -> kernels -> D2H of every step over three streams so that the PCIe copies of one step overlap the
    # kernels of its neighbours; every step still copies its whole input in and its whole result out.
    from deepxi_b200.model import HostPipeline
    n_sub, n_cs = int(os.environ.get('DXI_E2E_SUB', '2')), int(os.environ.get('DXI_E2E_CS', '2'))
    pipe = HostPipeline(dx, n_streams=3, n_sub=n_sub, n_compute=n_cs)
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
    win_e2e1 = time.time()
    if os.environ.get('DXI_BENCH_DEBUG'):
        sys.stderr.write('rank %d e2e submit marks %s total %.4f\n' % (rank, ['%.4f' % m for m in marks], e2e_s))
    barrier()
    assert int(y_hosts[0].abs().max()) > 0
    # ---- what the host <-> device links of this box give when every rank copies at once (the e2e arm's ceiling)
    h2d_gbs, d2h_gbs = pcie_ceiling(torch, x_host, y_hosts[0], dev, barrier)
    barrier()
    # the same with a write-combined input buffer (what a serving loop that only fills its input batches can use)
    from deepxi_b200 import hostmem
    x_wc = hostmem.pinned_copy(x_host, write_combined=True)
    h2d_wc_gbs, d2h_wc_gbs = pcie_ceiling(torch, x_wc, y_hosts[0], dev, barrier)
    barrier()

    # ---- sustained: the same device-resident step looped for >= 3 s (the 10-step region above is a burst of ~30 ms)
    sustained = None
    win_sus = None
    if not args.no_sustained:
        n_sus = max(args.steps, int(args.sustain_seconds / max(ms / args.steps * 1e-3, 1e-4)) + 1)
        s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier(); torch.cuda.synchronize()
        ws0 = time.time()
        s0.record()
        for i in range(n_sus):
            y = step(i)
        s1.record()
        torch.cuda.synchronize()
        win_sus = (ws0, time.time())
        sustained = (n_sus, s0.elapsed_time(s1))
        barrier()

    t = torch.tensor([ms, e2e_s * 1e3, -h2d_gbs, -d2h_gbs, sustained[1] if sustained else 0.0, -h2d_wc_gbs, -d2h_wc_gbs], dtype=torch.float64, device=dev)
    tsum = torch.tensor([h2d_gbs, d2h_gbs], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
    ms, e2e_ms, h2d_min, d2h_min, sus_ms = float(t[0]), float(t[1]), -float(t[2]), -float(t[3]), float(t[4])
    h2d_wc_min, d2h_wc_min = -float(t[5]), -float(t[6])

    # ---- the other BASELINE.json configs, in the same run (single-GPU runs only: they are per-GPU figures)
    configs = extra_configs(torch, dx, it, dev, args) if (world == 1 and not args.no_extra_configs) else None
    clocks = None
    if rank == 0:
        clocks = sampler.stop([(win_dev0, win_dev1), (win_e2e0, win_e2e1)])
        if win_sus:
            clocks_sus = ClockSampler.summarise(sampler.lines, [win_sus])
    if rank != 0:
        return
    audio_s = world * B * SECONDS
    value = audio_s * args.steps / (ms / 1e3)
    hbm_peak, tf_burst, tf_sustained, peak_src = measured_peaks()
    frames = B * T
    chain = prof['tcn_chain'][1] > 0
    st_ms, st_n = prof['tcn_chain'] if chain else prof['tcn_stage']
    st_ms_per_launch = st_ms / max(st_n, 1)
    st_per_step = st_n / max(args.steps, 1)          # 1 (depth-first kernel) or 41 (one launch per stage)
    fused = chain and prof['tcn_stem'][1] == 0      # first and output layer inside the same launch: the whole network's FLOPs
    flops_per_launch = frames * (FLOP_PER_FRAME_TOTAL if fused else FLOP_PER_FRAME_STAGES) / max(st_per_step, 1)
    achieved_tf = flops_per_launch / (st_ms_per_launch * 1e-3) / 1e12 if st_n else None
    chain_bytes = frames * TCN_CHAIN_BYTES_PER_FRAME
    stft_ms, stft_n = prof['stft']
    stft_gbs = frames * STFT_BYTES_PER_FRAME / (stft_ms / max(stft_n, 1) * 1e-3) / 1e9 if stft_n else None
    enh_ms, enh_n = prof['enhance']
    enh_gbs = (frames * (3 * 1028 + 1024) / (enh_ms / max(enh_n, 1) * 1e-3) / 1e9) if enh_n else None
    e2e_value = audio_s * args.steps / (e2e_ms / 1e3)
    # e2e bound the links imply: a step moves B*L*2 bytes in and B*(T+1)*512 bytes out per rank, the two directions run concurrently
    h2d_bytes, d2h_bytes = B * L * 2, B * (T + 1) * 256 * 2
    link_s = max(h2d_bytes / (h2d_min * 1e9), d2h_bytes / (d2h_min * 1e9)) if min(h2d_min, d2h_min) > 0 else None
    pcie_bound = (world * B * SECONDS / max(link_s, ms / args.steps * 1e-3)) if link_s else None
    line = {
        'metric': 'audio-seconds enhanced per second (ResNet-1.1c, MMSE-LSA)', 'value': value, 'unit': 'audio-s/s',
        'n_gpus': world, 'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': ms / args.steps,
        'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
        'dtype': {'f16x3': 'f16x3 (fp16 hi/lo split tensor-core operands, fp32 accumulate; fp32 elsewhere)',
                  'f16': 'f16 tensor-core operands, fp32 accumulate', 'f32': 'f32'}[args.precision],
        'data': 'synthetic',
        'config': {'workload': WORKLOAD % (B, SECONDS),
                   'utterances_per_gpu': B, 'frames_per_gpu': frames, 'precision': args.precision,
                   'sharding': 'utterances split across ranks, no data-path collective',
                   'l2': 'per-step working set %.2f GB per GPU (two alternating input buffers) exceeds the 126 MB L2'
                         % ((B * L * 2 + 3 * frames * 257 * 4 + frames * 1024 + B * (T + 1) * 1024) / 1e9)},
        'realtime_factor_per_gpu': value / world,
        'clocks': clocks,
        'e2e': {'value': e2e_value, 'unit': 'audio-s/s',
                'h2d_bytes_per_step': h2d_bytes, 'd2h_bytes_per_step': d2h_bytes,
                'api': 'HostPipeline(DeepXi).submit(pinned int16 in, lens, pinned int16 out): H2D, DeepXi.infer_batch and D2H on event-chained streams, 3 batches in flight, '
                       'each batch as %d sub-batches alternating over %d compute streams' % (n_sub, n_cs),
                'host_binding': numa,
                'pcie_ceiling_gbs': {'h2d_per_gpu_min': h2d_min, 'd2h_per_gpu_min': d2h_min, 'h2d_all_gpus': float(tsum[0]), 'd2h_all_gpus': float(tsum[1]),
                                     'with_write_combined_input': {'h2d_per_gpu_min': h2d_wc_min, 'd2h_per_gpu_min': d2h_wc_min},
                                     'how': '0.2 s of back-to-back duplex copies of the same pinned buffers, all ranks at once, CUDA events'},
                'pcie_bound': pcie_bound, 'frac_of_pcie_bound': (e2e_value / pcie_bound) if pcie_bound else None,
                'note': 'bound = audio per step / max(link time of a step at the measured duplex rates, device time of a step)'},
        'gpu_launches': launches,
        'roofline': {'kernel': ('tcn_chain_kernel (first layer + 40 residual blocks + output layer in one launch, depth first, tcgen05 / TMEM, %g launch per step)' if fused else
                                'tcn_chain_kernel (the 40 residual blocks depth first, tcgen05 / TMEM, %g launch per step)' if chain else
                                'tcn_stage_kernel (tcgen05 / TMEM, %g launches per step)') % st_per_step, 'bound': 'tensor',
                     'achieved': achieved_tf, 'peak': tf_burst, 'unit': 'TFLOP/s',
                     'frac': (achieved_tf / tf_burst) if achieved_tf else None,
                     'frac_of_sustained_peak': (achieved_tf / tf_sustained) if achieved_tf else None,
                     # what the tensor cores execute: three fp16 products per useful one in f16x3 mode (fp16 hi | lo operands)
                     'executed': (achieved_tf * (3 if args.precision == 'f16x3' else 1)) if achieved_tf else None,
                     'frac_executed': (achieved_tf * (3 if args.precision == 'f16x3' else 1) / tf_burst) if achieved_tf else None,
                     'tensor_pipe_active_pct_ncu': 44.0 if (fused and frames == 160000 and args.precision == 'f16x3') else None,
                     'traffic': NCU_TCN_CHAIN_TRAFFIC if (chain and frames == 160000 and args.precision == 'f16x3') else None,
                     'traffic_source': 'profiles/r02_prof_tcn_chain.csv (dram bytes and sm__pipe_tensor_cycles_active of the launch under ncu --set full; citations, not measured in this run); '
                                       'algorithmic %.0f bytes per launch' % chain_bytes,
                     'peak_source': peak_src + ' bf16 burst (MEASURED_PEAKS.json bf16_tflops): the timed region is %.0f ms; sustained %.1f' % (ms, tf_sustained),
                     'algorithmic_flop_per_launch': flops_per_launch, 'ms_per_launch': st_ms_per_launch,
                     'share_of_step': st_ms / ms if ms else None,
                     'note': 'useful FLOPs only: in f16x3 mode the tensor cores execute 3x this (ceiling of the useful fraction: 0.33); the '
                             'kernel keeps one tile per SM and is bound by the serial chain GEMM -> epilogue -> GEMM of a block, see DESIGN.md'},
        'kernels_ms_per_step': {k: (v[0] / args.steps) for k, v in prof.items()},
        'stft': {'bound': 'hbm', 'achieved': stft_gbs, 'peak': hbm_peak, 'unit': 'GB/s',
                 'frac': (stft_gbs / hbm_peak) if stft_gbs else None, 'bytes_per_frame': STFT_BYTES_PER_FRAME},
        'enhance': {'bound': 'hbm', 'achieved': enh_gbs, 'peak': hbm_peak, 'unit': 'GB/s', 'frac': (enh_gbs / hbm_peak) if enh_gbs else None,
                    'bytes_per_frame': 3 * 1028 + 1024},
        'enhance_gbs': enh_gbs,
    }
    if sustained:
        line['sustained'] = {'value': audio_s * sustained[0] / (sus_ms / 1e3), 'unit': 'audio-s/s', 'steps': sustained[0],
                             'seconds': sus_ms / 1e3, 'ms_per_step': sus_ms / sustained[0], 'clocks': clocks_sus}
    if configs:
        line['configs'] = configs
    if world == 1 and not args.no_cpu_baseline:
        line['cpu_baseline'] = cpu_baseline(args.ref_utts)
    print(json.dumps(line), flush=True)


def extra_configs(torch, dx, it, dev, args):
    """BASELINE.json configs other than the headline one, device-resident inputs, CUDA events (each a few tens of ms)."""
    from deepxi_b200 import synth, weights
    from deepxi_b200.model import DeepXi
    out = {}

    def timed(fn, reps, warm=2):
        for _ in range(warm):
            fn()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(reps):
            fn()
        b.record()
        torch.cuda.synchronize()
        return a.elapsed_time(b) / reps

    # C1: one 4 s utterance, batch 1 (latency of the whole path, int16 in -> waveform out)
    x1 = torch.from_numpy(synth.noisy_speech(1, 4 * F_S, seed=7)).to(dev)
    ms1 = timed(lambda: dx.infer_batch(x1, [4 * F_S], 'y', 'mmse-lsa'), 20, 3)
    out['C1'] = {'workload': 'ResNet-1.1c + MMSE-LSA, 1 utterance x 4 s, batch 1', 'ms_per_call': ms1, 'realtime_factor': 4.0 / (ms1 * 1e-3)}
    # C5: the per-GPU shard of the 8192-utterance corpus at 8 GPUs (1024 utterances x 10 s in one call)
    B5, L = 1024, SECONDS * F_S
    base = synth.noisy_speech(32, L, seed=99)
    x5 = torch.from_numpy(np.tile(base, (B5 // 32, 1))).to(dev)
    lens5 = [L] * B5
    ms5 = timed(lambda: dx.infer_batch(x5, lens5, 'y', 'mmse-lsa'), 3, 1)
    out['C5_shard'] = {'workload': 'ResNet-1.1c + MMSE-LSA, 1024 utt x 10 s (one GPU\'s shard of the 8192-utterance corpus at 8 GPUs)',
                       'ms_per_step': ms5, 'value': B5 * SECONDS / (ms5 * 1e-3), 'unit': 'audio-s/s'}
    del x5
    # C3: MHANet-1.1c, 64 utterances x 30 s
    dm = DeepXi(512, 256, 512, F_S, 'MagXi', 'MHANetV3', ver='mhanet-1.1c', map_type='DBNormalCDF', map_params=None,
                precision=args.precision, **MHA_KW)
    dm.set_weights(weights.synthetic_mhanetv3(0))
    B3, L3 = 64, 30 * F_S
    x3 = torch.from_numpy(np.tile(synth.noisy_speech(16, L3, seed=98), (4, 1))).to(dev)
    lens3 = [L3] * B3
    T3 = -(-L3 // 256)
    inp3, _, _ = dm.inp_tgt.observation_batch(x3, lens3)
    ms3n = timed(lambda: dm.network(inp3), 6, 2)      # the network alone first, the whole step second: both start from the same thermal state
    ms3 = timed(lambda: dm.infer_batch(x3, lens3, 'y', 'mmse-lsa'), 6, 2)
    mha_flop = B3 * T3 * (2 * 257 * 256 + 5 * (2 * 256 * 768 + 2 * 256 * 256 + 2 * 256 * 1024 * 2 + 4 * T3 * 256) + 2 * 256 * 257)
    hbm_peak, tf_burst, tf_sustained, _ = measured_peaks()
    out['C3'] = {'workload': 'MHANet-1.1c (MHANetV3, 5 blocks, 8 heads, mask as shipped) + MMSE-LSA, 64 utt x 30 s',
                 'ms_per_step': ms3, 'value': B3 * 30 / (ms3 * 1e-3), 'unit': 'audio-s/s', 'network_ms': ms3n,
                 'roofline': {'bound': 'tensor', 'achieved': mha_flop / (ms3n * 1e-3) / 1e12, 'peak': tf_burst, 'unit': 'TFLOP/s',
                              'frac': mha_flop / (ms3n * 1e-3) / 1e12 / tf_burst, 'note': 'whole network forward, useful FLOPs (f16x3 executes 3x)'}}
    return out


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
    ap.add_argument('--no-sustained', action='store_true')
    ap.add_argument('--no-extra-configs', action='store_true')
    ap.add_argument('--sustain-seconds', type=float, default=3.0)
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
