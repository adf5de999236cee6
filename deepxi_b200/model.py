"""Mirror of deepxi/model.py: class DeepXi -- constructor (:44-111), infer (:224-332) and
observation_batch (:2232-2254).  Training, testing with PESQ/STOI and the fork's experiment methods
(infer_pho, infer_hybrid*, infer_tracking_noise*) are out of scope.

Differences from the reference, all on the host side:
  * observation -> network -> map / gain -> synthesis run batched on the GPU instead of three
    per-utterance Python loops; files are written from one device-to-host copy per batch;
  * `out_type='gain'` (named in args.py:60-64 but raising in the reference) returns
    gfunc(xi_hat, xi_hat + 1, gain);
  * `infer_batch` returns the outputs in memory (device tensors) instead of writing files.
"""
import os

import numpy as np
import torch

from . import stats as _stats
from . import weights as _weights
from .inp_tgt import inp_tgt_selector
from .network.selector import network_selector
from .utils import save_mat, save_wav, read_mat

_OUT_TYPES = ('y', 'xi_hat', 'gamma_hat', 'gain', 'ibm_hat', 'deepmmse', 'subband_ibm_hat')
_OTHER_OUT_TYPES = ('mag_hat', 'cd_hat')


class DeepXi:
    def __init__(self, N_d, N_s, K, f_s, inp_tgt_type, network_type, min_snr=-10, max_snr=20, snr_inter=1,
                 log_path=None, sample_dir=None, ver='VERSION_NAME', train_s_list=None, train_d_list=None,
                 sample_size=None, reset_inp_tgt=False, **kwargs):
        self.inp_tgt_type = inp_tgt_type
        self.network_type = network_type
        self.min_snr, self.max_snr = min_snr, max_snr
        self.snr_levels = list(range(min_snr, max_snr + 1, snr_inter))
        self.ver = ver
        if inp_tgt_type not in ('MagXi', 'MagGain'):
            # MagXiGamma (514 outputs) and the fork's other targets have no committed checkpoint; the target classes themselves
            # (inp_tgt.MagXiGamma ...) are available for the training-target side
            raise NotImplementedError('DeepXi drives the MagXi and MagGain targets; %r has no committed model' % inp_tgt_type)
        self.inp_tgt = inp_tgt_selector(inp_tgt_type, N_d, N_s, K, f_s, **kwargs)
        self._init_rest(sample_dir, ver, reset_inp_tgt, N_d, N_s, K, network_type, kwargs)

    def _init_rest(self, sample_dir, ver, reset_inp_tgt, N_d, N_s, K, network_type, kwargs):
        if self.inp_tgt_type == 'MagGain':      # the network estimates the gain itself: no CDF statistics (inp_tgt.py:471-519)
            net_kwargs = dict(kwargs)
            for k in ('map_type', 'map_params', 'stats'):
                net_kwargs.pop(k, None)
            self.network = network_selector(network_type, None, self.inp_tgt.n_outp, **net_kwargs)
            self.model = self.network
            self._weights_epoch = None
            return
        # statistics: data/<ver>_inp_tgt.p when present (model.py:90-93), else the packaged values
        p = os.path.join(sample_dir, ver + '_inp_tgt.p') if sample_dir else None
        if p and os.path.exists(p) and not reset_inp_tgt:
            st = _stats.load_inp_tgt_pickle(p)
            if (st['N_d'], st['N_s'], st['K']) != (N_d, N_s, K):
                raise ValueError('%s was computed for a different framing' % p)
            mu, sigma = st['mu'], st['sigma']
        elif 'stats' in kwargs and kwargs['stats'] is not None:
            mu, sigma = kwargs['stats']
        else:
            try:
                mu, sigma = _stats.packaged(ver)
            except KeyError:
                raise ValueError('no statistics for ver=%r: pass sample_dir with <ver>_inp_tgt.p, or stats=(mu, sigma); '
                                 'computing them needs the training set (out of scope)' % ver)
        self.inp_tgt.set_stats(mu, sigma)
        net_kwargs = dict(kwargs)
        for k in ('map_type', 'map_params', 'stats'):
            net_kwargs.pop(k, None)
        self.network = network_selector(network_type, None, self.inp_tgt.n_outp, **net_kwargs)
        self.model = self.network
        self._weights_epoch = None

    # ------------------------------------------------------------------------------------------
    def set_weights(self, weights):
        self.network.load_weights(weights)
        self._weights_epoch = 'explicit'

    def load_weights(self, model_path, epoch):
        """model.load_weights(model_path/epoch-<epoch>/variables/variables) (model.py:279-280)."""
        self.network.load_weights(_weights.load_checkpoint(model_path, epoch))
        self._weights_epoch = (model_path, epoch)

    def observation_batch(self, x_batch, x_batch_len):
        """model.py:2232-2254."""
        return self.inp_tgt.observation_batch(x_batch, x_batch_len)

    # ------------------------------------------------------------------------------------------
    def infer_batch(self, test_x, test_x_len, out_type='y', gain='mmse-lsa', int16=False, n_filters=40):
        """One batch through the hot path; returns (output device tensor [B, ...], n_frames list).

        'y' -> [B, (Tmax+1)*256] waveform (float32, or int16 with the save_wav rule); the others
        -> [B, Tmax, 257]; slices beyond n_frames[i] are padding."""
        if out_type in _OTHER_OUT_TYPES:
            raise NotImplementedError('out_type %r belongs to targets without committed models' % out_type)
        if out_type not in _OUT_TYPES:
            raise ValueError('Invalid output type.')
        it = self.inp_tgt
        inp, pha, n_frames = it.observation_batch(test_x, test_x_len)
        xbar = self.network(inp)
        if self.inp_tgt_type == 'MagGain':      # the network output IS the gain (inp_tgt.py:500-519); only the waveform is defined
            if out_type != 'y':
                raise ValueError('Invalid output type.')
            return it.enhanced_speech(inp, pha, xbar, n_frames=n_frames, int16=int16), n_frames
        if out_type == 'y':
            out = it.enhanced_speech(inp, pha, xbar, gain, n_frames=n_frames, int16=int16)
        elif out_type == 'xi_hat':
            out = it.xi_hat(xbar)
        elif out_type == 'gamma_hat':
            out = it.gamma_hat(xbar)
        elif out_type == 'gain':
            out = it.gain_hat(xbar, gain)
        elif out_type == 'ibm_hat':
            out = it.ibm_hat(xbar)
        elif out_type == 'subband_ibm_hat':      # (xi_hat H^T) > 1 with the mel filter bank (model.py:255-258, :323-328)
            out = it.subband(it.xi_hat(xbar), n_filters)[1]
        else:  # deepmmse: |X|^2 * G_deepmmse(xi_hat, xi_hat + 1)   (model.py:314-318), one kernel
            out = it.deepmmse(inp, xbar)
        return out, n_frames

    def infer(self, test_x, test_x_len, test_x_base_names, test_epoch, model_path='model', out_type='y',
              gain='mmse-lsa', out_path='out', n_filters=40, saved_data_path=None):
        """Deep Xi inference; the specified out_type is saved (model.py:224-332)."""
        out_path_base = out_path
        if saved_data_path is not None:
            # model.py:298-300 hands the .mat contents to the target's enhanced_speech as (supplementary, saved_data); only targets
            # that are out of scope here consume it (MagXi / MagGain take the phase alone)
            raise NotImplementedError('saved_data_path is only consumed by targets without committed models')
        if not isinstance(test_epoch, list): test_epoch = [test_epoch]
        if not isinstance(gain, list): gain = [gain]
        for e in test_epoch:
            if e < 1: raise ValueError('test_epoch must be greater than 0.')
            for g in gain:
                out_path = out_path_base + '/' + self.ver + '/' + 'e' + str(e)
                if out_type == 'xi_hat': out_path = out_path + '/xi_hat'
                elif out_type == 'gamma_hat': out_path = out_path + '/gamma_hat'
                elif out_type == 'y':
                    if self.inp_tgt_type in ('MagGain', 'MagMag'): out_path = out_path + '/y'      # model.py:269-271
                    else: out_path = out_path + '/y/' + g
                elif out_type == 'deepmmse': out_path = out_path + '/deepmmse'
                elif out_type == 'ibm_hat': out_path = out_path + '/ibm_hat'
                elif out_type == 'subband_ibm_hat': out_path = out_path + '/subband_ibm_hat'
                elif out_type == 'gain': out_path = out_path + '/gain/' + g
                elif out_type in _OTHER_OUT_TYPES:
                    raise NotImplementedError('out_type %r belongs to targets without committed models' % out_type)
                else: raise ValueError('Invalid output type.')
                if not os.path.exists(out_path): os.makedirs(out_path)
                if self._weights_epoch != 'explicit' and self._weights_epoch != (model_path, e - 1):
                    self.load_weights(model_path, e - 1)
                out, n_frames = self.infer_batch(test_x, test_x_len, out_type, g, int16=(out_type == 'y'), n_filters=n_filters)
                out = out.cpu().numpy()
                key = {'xi_hat': 'xi_hat', 'gamma_hat': 'gamma_hat', 'gain': 'gain', 'ibm_hat': 'ibm_hat',
                       'deepmmse': 'd_psd_hat', 'subband_ibm_hat': 'subband_ibm_hat'}.get(out_type)
                for i, base_name in enumerate(test_x_base_names):
                    if out_type == 'y':
                        # polar_synthesis length (T-1)*N_s + N_d of the un-padded utterance (inp_tgt.py:198-214)
                        n = (n_frames[i] + 1) * self.inp_tgt.N_s
                        save_wav(out_path + '/' + base_name + '.wav', out[i, :n], self.inp_tgt.f_s)
                    else:
                        save_mat(out_path + '/' + base_name + '.mat', out[i, :n_frames[i]], key)


class HostPipeline:
    """Throughput-oriented serving loop over host buffers.  One CUDA stream copies inputs host -> device, `n_compute` streams run
    the kernels, one copies results device -> host; events chain the three steps.  A submitted batch is cut into `n_sub`
    sub-batches of utterances that travel through the pipeline one after the other: the first kernels start when the first
    sub-batch has arrived (not the whole batch) and the last copy back only carries a sub-batch, so the fill and drain of the
    pipeline shrink by n_sub; consecutive sub-batches alternate between the compute streams, and because the network kernel's CTAs
    claim their work items dynamically, the next sub-batch's CTAs take over the SMs the previous one leaves idle in its ragged last
    round.  Measured on one B200 with 256 x 10 s batches (10 steps): n_sub / n_compute = 1 / 1: 904 k audio-s/s, 2 / 1: 854 k,
    2 / 2: 959 k (the default), 3 / 2: 909 k, 4 / 2: 896 k, 8 / 2: 839 k (sub-batches much smaller than ~4 rounds of tiles pay for
    their own ragged rounds; unequal splits such as 1/8, 3/4, 1/8 - a short first copy in and last copy out - lose more in the small
    sub-batches than they save: 0.89 - 0.94 of the device rate at 20 steps against 0.99 for two halves).  Inputs / outputs are pinned host tensors owned by the caller; up to `n_streams` batches are in flight.

        pipe = HostPipeline(deepxi, n_streams=3)
        for x, lens, y in batches:          # x int16 [B, L] pinned, y int16 [B, (Tmax+1)*256] pinned
            pipe.submit(x, lens, y)
        pipe.drain()                        # all outputs are now valid
    """

    def __init__(self, deepxi, n_streams=3, out_type='y', gain='mmse-lsa', n_sub=2, n_compute=2):
        self.dx, self.out_type, self.gain = deepxi, out_type, gain
        self.n_slots = max(1, int(n_streams))
        self.n_sub = max(1, int(n_sub))
        self.h2d, self.d2h = torch.cuda.Stream(), torch.cuda.Stream()
        self.computes = [torch.cuda.Stream() for _ in range(max(1, int(n_compute)))]
        self.compute = self.computes[0]
        self._i = 0
        self._c = 0
        self._done = [None] * self.n_slots       # event: slot's device -> host copy finished
        self._keep = [None] * self.n_slots

    def submit(self, x_host, x_len, out_host):
        if not (x_host.is_pinned() and out_host.is_pinned()):
            raise ValueError('HostPipeline needs pinned host tensors')
        k = self._i % self.n_slots
        self._i += 1
        if self._done[k] is not None:
            self._done[k].synchronize()     # at most n_slots batches in flight (host-side back-pressure)
        B = x_host.shape[0]
        n_sub = min(self.n_sub, max(B, 1))
        edges = [(B * q) // n_sub for q in range(n_sub + 1)]
        keep, n_frames, done = [], [], None
        for q in range(n_sub):
            a, b = edges[q], edges[q + 1]
            if a == b:
                continue
            with torch.cuda.stream(self.h2d):
                xd = x_host[a:b].to('cuda', non_blocking=True)
                ev_in = torch.cuda.Event()
                ev_in.record(self.h2d)
            cs = self.computes[self._c % len(self.computes)]
            self._c += 1
            with torch.cuda.stream(cs):
                cs.wait_event(ev_in)
                out, nfr = self.dx.infer_batch(xd, x_len[a:b], self.out_type, self.gain, int16=out_host.dtype == torch.int16)
                ev_out = torch.cuda.Event()
                ev_out.record(cs)
            with torch.cuda.stream(self.d2h):
                self.d2h.wait_event(ev_out)
                out_host[a:b].copy_(out, non_blocking=True)
                done = torch.cuda.Event()
                done.record(self.d2h)
            keep.append((xd, out))
            n_frames.extend(nfr)
        self._done[k] = done
        # The device buffers of a slot stay alive until the slot is reused, and the slot is reused only after its `done` event has
        # been synchronised (above): every stream is finished with them by then.  That is why no record_stream() is used here: it
        # would make the caching allocator defer the reuse of these blocks to an event it polls lazily, and fall back to a
        # (device-synchronising) cudaMalloc whenever the poll comes too early -- an occasional stall of the whole pipeline.
        self._keep[k] = keep
        return n_frames

    def drain(self):
        for st in [self.h2d, self.d2h] + self.computes:
            st.synchronize()

    def __del__(self):
        try:
            self.drain()      # the kept buffers must not return to the allocator while a copy still reads them
        except Exception:
            pass
