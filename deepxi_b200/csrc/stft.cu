// STFT analysis and iSTFT synthesis kernels.
//
// Analysis  (deepxi/sig.py:43-55, :189-199; model.py:2232-2254): int16/f32 waveform -> |STFT|, angle(STFT),
//           512-sample Hamming(periodic=False) frames, hop 256, pad_end, rfft-512 -> 257 bins.
// Synthesis (deepxi/sig.py:57-69; inp_tgt.py:198-214; utils.py:28): (|X| G) e^{j phase} -> irfft-512 ->
//           inverse_stft_window_fn -> overlap-add -> f32 and/or truncated int16.
//
// HBM layout: waveforms [B, stride] sample-fastest; spectra [B, Tmax, 257] bin-fastest, so the frames
// of one utterance are one contiguous run and every global access below is a flat coalesced stream.
// Each sample is read from HBM once (the 50 % frame overlap is served from shared memory) and every
// output element is written once.
#include <math.h>
#include <stdlib.h>
#include <mutex>
#include <type_traits>
#include "fft.cuh"
#include "gain_math.cuh"

namespace dxi {

__device__ float d_win[N_D];          // analysis window
__device__ float d_swin[N_D];         // synthesis window / 256 (irfft scaling folded in)
__device__ float2 d_tw256[256];       // e^{-2 pi j m / 256}
__device__ float2 d_tw512[NBINS];     // e^{-2 pi j k / 512}, k = 0..256
__device__ float4 d_lsa_tab[LSA_TAB_N];   // piecewise cubics of log2 G_LSA(xi_hat [dB]) (gain_math.cuh)

static std::mutex g_tab_mutex;
static bool g_tab_ready[64] = {};

static int ensure_tables(cudaStream_t stream) {
  int dev = 0;
  DXI_CUDA(cudaGetDevice(&dev));
  std::lock_guard<std::mutex> lock(g_tab_mutex);
  if (dev < 64 && g_tab_ready[dev]) return DXI_OK;
  static float win[N_D], swin[N_D];
  static float2 tw256[256], tw512[NBINS];
  double w[N_D];
  for (int n = 0; n < N_D; ++n) w[n] = 0.54 - 0.46 * cos(2.0 * M_PI * n / (N_D - 1));   // sig.py:38-39
  for (int n = 0; n < N_D; ++n) {
    int m = n % N_S;
    double den = w[m] * w[m] + w[m + N_S] * w[m + N_S];     // tf.signal.inverse_stft_window_fn
    win[n] = (float)w[n];
    // the reference multiplies the float32 irfft output by the float32 synthesis window; the 1/256
    // of the half-size inverse transform is exact in binary and is folded in here
    swin[n] = (float)(w[n] / den) * (1.0f / 256.0f);
  }
  for (int m = 0; m < 256; ++m) tw256[m] = make_float2((float)cos(2.0 * M_PI * m / 256.0), (float)-sin(2.0 * M_PI * m / 256.0));
  for (int k = 0; k < NBINS; ++k) tw512[k] = make_float2((float)cos(2.0 * M_PI * k / 512.0), (float)-sin(2.0 * M_PI * k / 512.0));
  DXI_CUDA(cudaMemcpyToSymbolAsync(d_win, win, sizeof(win), 0, cudaMemcpyHostToDevice, stream));
  DXI_CUDA(cudaMemcpyToSymbolAsync(d_swin, swin, sizeof(swin), 0, cudaMemcpyHostToDevice, stream));
  DXI_CUDA(cudaMemcpyToSymbolAsync(d_tw256, tw256, sizeof(tw256), 0, cudaMemcpyHostToDevice, stream));
  DXI_CUDA(cudaMemcpyToSymbolAsync(d_tw512, tw512, sizeof(tw512), 0, cudaMemcpyHostToDevice, stream));
  static float4 lsa_tab[LSA_TAB_N];
  lsa_table_build(lsa_tab);
  DXI_CUDA(cudaMemcpyToSymbolAsync(d_lsa_tab, lsa_tab, sizeof(lsa_tab), 0, cudaMemcpyHostToDevice, stream));
  DXI_CUDA(cudaStreamSynchronize(stream));   // one-time: the host tables above are static scratch
  if (dev < 64) g_tab_ready[dev] = true;
  return DXI_OK;
}

constexpr int FR = 16;   // frames transformed per CTA pass (16 threads per frame, 256 threads)

template <bool I16>
struct StftSmem {                     // 47.5 KB (int16 input) / 56 KB (f32 input): four CTAs per SM
  float2 buf[FR * FFT_FRAME_SLOTS];   // per-frame FFT exchange, then the 256 Z values
  typename std::conditional<I16, int16_t, float>::type raw[(FR + 1) * N_S];   // samples of FR overlapping frames, as they lie in HBM
  float win[N_D];                     // analysis window x 1/2 (split step) x 1/32768 (sig.py:189-199 normalisation) for int16 input
  float2 tw256[256];
};

__device__ __forceinline__ void cp_async_16(void* smem_dst, const void* gmem_src, int src_bytes) {      // bytes beyond src_bytes: zeros
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(gmem_src),
               "r"(src_bytes)
               : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

// ----------------------------------------------------------------------------------------------
// Analysis
// ----------------------------------------------------------------------------------------------
template <bool I16>
__global__ void __launch_bounds__(256, 4) stft_kernel(const void* __restrict__ wav_, const int32_t* __restrict__ lens,
                                                   int B, int64_t stride, int Tmax, int groups_per_utt,
                                                   int vec_ok, float* __restrict__ mag, float* __restrict__ phase) {
  using elem_t = typename std::conditional<I16, int16_t, float>::type;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  StftSmem<I16>& sm = *reinterpret_cast<StftSmem<I16>*>(smem_raw);
  const int tid = threadIdx.x;
  // the factor 1/2 of the split step and the int16 normalisation ride in the window: both are exact scalings by powers of two
  for (int i = tid; i < N_D; i += 256) sm.win[i] = (I16 ? 0.5f / 32768.0f : 0.5f) * d_win[i];
  sm.tw256[tid] = d_tw256[tid];
  const float2 w_pair = d_tw512[tid & 127];      // split-step twiddle of this thread's bin pair

  const int f = tid >> 4, lane16 = tid & 15;
  const int total = B * groups_per_utt;
  // g / groups_per_utt without the integer-division sequence (it ran twice per group and thread: 4 % of the kernel's instructions):
  // the quotient (an utterance index) is far below 2^24, so the float quotient is off by at most one; total < 2^30 (dxi_stft checks)
  const float inv_gpu = 1.0f / (float)groups_per_utt;
  auto utt_of = [&](int g) {
    int b = __float2int_rz((float)g * inv_gpu);
    if ((b + 1) * groups_per_utt <= g) ++b;
    if (b * groups_per_utt > g) --b;
    return b;
  };
  const elem_t* wav = reinterpret_cast<const elem_t*>(wav_);
  // The (nf + 1) * 256 samples of a group travel HBM -> shared memory as they are (cp.async, 16 bytes per request, zero fill beyond
  // the utterance); the request for the NEXT group is issued as soon as the FFT threads have taken the current samples into
  // registers, so that its latency runs under the split / magnitude / phase phase.  Each sample is read from HBM once.
  auto stage = [&](int g) {
    if (g >= total) return;
    const int b = utt_of(g), t0 = (g - b * groups_per_utt) * FR;
    const int64_t len = min(lens ? (int64_t)lens[b] : stride, stride);
    const int64_t s0 = (int64_t)t0 * N_S;
    if (s0 >= len) return;
    const elem_t* w = wav + (int64_t)b * stride + s0;
    const int n_stage = (min(FR, Tmax - t0) + 1) * N_S;
    const int left = (int)min(len - s0, (int64_t)n_stage);      // samples of the group that exist
    constexpr int PER = 16 / (int)sizeof(elem_t);                // samples per 16-byte request
    if (vec_ok) {
      for (int i = tid * PER; i < n_stage; i += 256 * PER) {
        const int nb = min(max(left - i, 0), PER) * (int)sizeof(elem_t);
        cp_async_16(&sm.raw[i], nb > 0 ? (const void*)(w + i) : (const void*)wav, nb);
      }
    } else {
      for (int i = tid; i < n_stage; i += 256) sm.raw[i] = i < left ? w[i] : (elem_t)0;
    }
  };
  stage(blockIdx.x);
  cp_async_commit();
  for (int g = blockIdx.x; g < total; g += gridDim.x) {
    const int b = utt_of(g), t0 = (g - b * groups_per_utt) * FR;
    const int nf = min(FR, Tmax - t0);
    const int64_t len = min(lens ? (int64_t)lens[b] : stride, stride);
    const bool live = (int64_t)t0 * N_S < len;      // else: frames at or beyond ceil(len/256), zeros (model.py:2246-2253 leaves them zero)
    const int64_t out0 = ((int64_t)b * Tmax + t0) * NBINS;
    cp_async_wait_all();
    __syncthreads();      // this group's samples have landed (and, first pass, the tables); the previous group's Z values are consumed
    // ---- window + 256-point complex FFT of z[m] = x[2m] + j x[2m+1]
    float2* fb = sm.buf + f * FFT_FRAME_SLOTS;
    float2 v[16];
    const bool mine = live && f < nf;
    if (mine) {
#pragma unroll
      for (int n1 = 0; n1 < 16; ++n1) {
        const int n = 32 * n1 + 2 * lane16;
        float2 x;
        if (I16) {
          const uint32_t pr = *reinterpret_cast<const uint32_t*>(&sm.raw[f * N_S + n]);
          x = make_float2((float)(int16_t)(pr & 0xffffu), (float)(int16_t)(pr >> 16));
        } else {
          x = *reinterpret_cast<const float2*>(&sm.raw[f * N_S + n]);
        }
        float2 w = *reinterpret_cast<const float2*>(&sm.win[n]);
        v[n1] = make_float2(x.x * w.x, x.y * w.y);
      }
      fft256_pass1<-1>(v, fb, sm.tw256, lane16);
    }
    __syncwarp();
    if (mine) fft256_pass2<-1>(v, fb, lane16);
    __syncwarp();
    if (mine) {
#pragma unroll
      for (int k2 = 0; k2 < 16; ++k2) fb[lane16 + 16 * k2] = v[fft16_pos(k2)];
    }
    __syncthreads();
    stage(g + gridDim.x);      // every thread has taken its samples: the buffer is free for the next group
    cp_async_commit();
    if (!live) {
      for (int i = tid; i < nf * NBINS; i += 256) { __stcs(mag + out0 + i, 0.0f); __stcs(phase + out0 + i, 0.0f); }
      continue;
    }
    // ---- split step, magnitude and phase.  Thread = bin PAIR (j, 256 - j), j = 0..127, for every second frame of the pass (the
    // two halves of the CTA take the even / the odd frames): the pair shares both exchange loads, E, O and the twiddle product,
    // the two magnitudes and phase polynomials run as packed fp32x2 operations, the twiddle and both exchange indices are loop
    // invariants and a frame's outputs are two coalesced runs (ascending from bin 0, descending from bin 256).  Bin 128 of frame f
    // (its own partner) is done by thread f afterwards.
    {
      const int j = tid & 127, ib = (256 - j) & 255;
      const float2 w = w_pair;
#pragma unroll 2
      for (int fi = tid >> 7; fi < nf; fi += 2) {
        const float2* z = sm.buf + fi * FFT_FRAME_SLOTS;
        float2 a, c, m, ph;
        rfft_split_pair(z[j], z[ib], w, a, c);
        if (j == 0) { a.y = 0.0f; c.y = 0.0f; }       // DC and Nyquist are real
        polar_pair(a, c, m, ph);
        const int64_t o = out0 + (int64_t)fi * NBINS;
        __stcs(mag + o + j, m.x);   __stcs(mag + o + 256 - j, m.y);
        __stcs(phase + o + j, ph.x); __stcs(phase + o + 256 - j, ph.y);
      }
      if (tid < nf) {      // bin 128: X = conj(Z[128]) = 2 conj(Z'[128])
        const float2 z = sm.buf[tid * FFT_FRAME_SLOTS + 128];
        const float2 X = make_float2(2.0f * z.x, -2.0f * z.y);
        const int64_t o = out0 + (int64_t)tid * NBINS + 128;
        __stcs(mag + o, sqrt_approx(fmaf(X.x, X.x, X.y * X.y)));
        __stcs(phase + o, atan2_poly(X.y, X.x));
      }
    }
  }
}

// ----------------------------------------------------------------------------------------------
// Synthesis
// ----------------------------------------------------------------------------------------------
constexpr int HOPS_PER_STRIP_MAX = 256;   // a strip of h hops recomputes one leading frame (1/h redundant work); h is chosen per launch, see istft_launch
constexpr int ISTFT_CTAS_PER_SM = 4;  // 47 KB of shared memory and 64 registers per thread: the kernel is issue / latency bound, not HBM bound

struct IstftSmem {
  float2 buf[FR * FFT_FRAME_SLOTS];   // per frame: 257 spectrum values, then the FFT exchange, then (as 512 floats) the windowed time-domain frame
  float4 lsa[LSA_TAB_N];              // MMSE-LSA gain table (MODE 2)
  float carry[N_S];                   // second half of the last frame of the previous pass
  float swin[N_D];
  float2 tw256[256];
  float2 tw512[NBINS + 1];
};

// MODE 0: gain tensor (nullable); MODE 1: fused inverse map + gain from xbar (inp_tgt.py:198-214), exact-order math, any gain
// type; MODE 2: the same for MMSE-LSA through the MUFU-based path of gain_math.cuh (lsa_gain_from_xbar_fast)
template <int MODE>
__global__ void __launch_bounds__(256, ISTFT_CTAS_PER_SM) istft_kernel(const float* __restrict__ mag, const float* __restrict__ gain_or_xbar,
                                                    const float* __restrict__ phase, const float* __restrict__ mu,
                                                    const float* __restrict__ sigma, int gtype,
                                                    const int32_t* __restrict__ n_frames, int B, int Tmax,
                                                    int strips_per_utt, int hops_per_strip, float* __restrict__ wav_f32,
                                                    int16_t* __restrict__ wav_i16, int64_t out_stride, int out_vec) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  IstftSmem& sm = *reinterpret_cast<IstftSmem*>(smem_raw);
  const int tid = threadIdx.x;
  for (int i = tid; i < N_D; i += 256) sm.swin[i] = d_swin[i];
  sm.tw256[tid] = d_tw256[tid];
  for (int i = tid; i < NBINS; i += 256) sm.tw512[i] = d_tw512[i];
  if (MODE == 2)
    for (int i = tid; i < LSA_TAB_N; i += 256) sm.lsa[i] = d_lsa_tab[i];
  __syncthreads();

  const int f = tid >> 4, lane16 = tid & 15;
  const int total = B * strips_per_utt;
  auto frame_of = [&](int fi) { return reinterpret_cast<float*>(sm.buf + fi * FFT_FRAME_SLOTS); };      // windowed frame fi (512 floats)
  for (int s = blockIdx.x; s < total; s += gridDim.x) {
    const int b = s / strips_per_utt;
    const int hs = (s - b * strips_per_utt) * hops_per_strip;      // first hop of the strip
    const int he = min(hs + hops_per_strip, Tmax + 1);              // hops hs..he-1; hop h = tail(h-1) + head(h)
    int T = n_frames ? min(n_frames[b], Tmax) : Tmax;
    const int64_t in0 = (int64_t)b * Tmax * NBINS;
    float* of = wav_f32 ? wav_f32 + (int64_t)b * out_stride : nullptr;
    int16_t* oi = wav_i16 ? wav_i16 + (int64_t)b * out_stride : nullptr;
    // frames hs-1 .. he-1 in passes of FR; the first frame of the strip only provides its tail
    for (int f0 = hs - 1; f0 < he; f0 += FR) {
      const int nf = min(FR, he - f0);
      // ---- spectrum of every frame: Y = (|X| G) e^{j phase}.  Thread = bin (0..255) for every frame of the pass (per-bin
      // statistics are loop invariants, a frame's 256 inputs are one coalesced row, two frames per iteration so that six loads
      // are in flight); the Nyquist bin of frame f is done by thread f afterwards.
      {
        float mu_k = 0.0f, sg_k = 0.0f, s2_k = 0.0f;
        auto spectrum = [&](int fi, int k, float m, float p, float gx, bool ok, float muv, float sgv, float s2v) {
          float2 Y = make_float2(0.0f, 0.0f);
          if (ok) {
            if (MODE == 0) {
              if (gain_or_xbar) m *= gx;
            } else if (MODE == 1) {
              const float xi = xi_from_xbar(gx, muv, sgv);
              m = __fmul_rn(m, gfunc_eval(gtype, xi, __fadd_rn(xi, 1.0f)));
            } else {
              m *= lsa_gain_from_xbar_tab(gx, muv, s2v, sgv, sm.lsa);
            }
            float sn, cs;
            __sincosf(p, &sn, &cs);                                        // |p| <= pi: abs error < 5e-7
            Y = make_float2(m * cs, (k == 0 || k == 256) ? 0.0f : m * sn);   // c2r ignores Im of DC / Nyquist
          }
          sm.buf[fi * FFT_FRAME_SLOTS + k] = Y;
        };
        const int k = tid;
        if (MODE != 0) { mu_k = __ldg(mu + k); sg_k = __ldg(sigma + k); s2_k = __fmul_rn(sg_k, 1.41421354f); }
        constexpr int NU = 4;      // frames per iteration: 3 NU loads in flight per thread (NU = 2 has no spills at 64 registers but is 8 % slower)
        for (int fi = 0; fi < nf; fi += NU) {
          float mv[NU], pv[NU], gv[NU];
          const int t0 = f0 + fi;
          if (fi + NU <= nf && t0 >= 0 && t0 + NU <= T) {
            // all NU frames exist (every pass but a strip's first and an utterance's last): one base index, constant row offsets, no tests
            const int64_t g0 = in0 + (int64_t)t0 * NBINS + k;
            const float* pm = mag + g0;
            const float* pp = phase + g0;
            const float* pg = (MODE != 0 || gain_or_xbar) ? gain_or_xbar + g0 : nullptr;
#pragma unroll
            for (int u = 0; u < NU; ++u) {
              mv[u] = __ldcs(pm + u * NBINS);
              pv[u] = __ldcs(pp + u * NBINS);
              gv[u] = pg ? __ldcs(pg + u * NBINS) : 0.0f;
            }
#pragma unroll
            for (int u = 0; u < NU; ++u) spectrum(fi + u, k, mv[u], pv[u], gv[u], true, mu_k, sg_k, s2_k);
            continue;
          }
          bool ok[NU];
#pragma unroll
          for (int u = 0; u < NU; ++u) {
            const int t = f0 + fi + u;
            ok[u] = fi + u < nf && t >= 0 && t < T;
            mv[u] = pv[u] = gv[u] = 0.0f;
            if (ok[u]) {
              const int64_t gi = in0 + (int64_t)t * NBINS + k;
              mv[u] = __ldcs(mag + gi);
              pv[u] = __ldcs(phase + gi);
              if (MODE != 0 || gain_or_xbar) gv[u] = __ldcs(gain_or_xbar + gi);
            }
          }
#pragma unroll
          for (int u = 0; u < NU; ++u)
            if (fi + u < nf) spectrum(fi + u, k, mv[u], pv[u], gv[u], ok[u], mu_k, sg_k, s2_k);
        }
        if (tid < nf) {      // bin 256 of frame tid
          const int t = f0 + tid;
          const bool ok1 = t >= 0 && t < T;
          float m = 0.0f, p = 0.0f, gx = 0.0f, muv = 0.0f, sgv = 0.0f;
          if (ok1) {
            const int64_t gi = in0 + (int64_t)t * NBINS + 256;
            m = __ldcs(mag + gi); p = __ldcs(phase + gi);
            if (MODE != 0 || gain_or_xbar) gx = __ldcs(gain_or_xbar + gi);
            if (MODE != 0) { muv = __ldg(mu + 256); sgv = __ldg(sigma + 256); }
          }
          spectrum(tid, 256, m, p, gx, ok1, muv, sgv, __fmul_rn(sgv, 1.41421354f));
        }
      }
      __syncthreads();
      // ---- merge step + 256-point inverse FFT
      float2* fb = sm.buf + f * FFT_FRAME_SLOTS;
      float2 v[16];
      if (f < nf) {
#pragma unroll
        for (int n1 = 0; n1 < 16; ++n1) {
          const int k = 16 * n1 + lane16;
          v[n1] = irfft_merge(fb[k], fb[256 - k], sm.tw512[k]);
        }
      }
      __syncwarp();
      if (f < nf) fft256_pass1<1>(v, fb, sm.tw256, lane16);
      __syncwarp();
      if (f < nf) fft256_pass2<1>(v, fb, lane16);
      __syncwarp();      // both frames of the warp have taken the exchange into registers: their buffers now receive the windowed frames
      if (f < nf) {
        // z[m] = x[2m] + j x[2m+1], m = lane16 + 16 k2; synthesis window (x 1/256 folded in)
        float* fr = frame_of(f);
#pragma unroll
        for (int k2 = 0; k2 < 16; ++k2) {
          const int n = 2 * (lane16 + 16 * k2);
          float2 z = v[fft16_pos(k2)];
          float2 w = *reinterpret_cast<const float2*>(&sm.swin[n]);
          *reinterpret_cast<float2*>(&fr[n]) = make_float2(z.x * w.x, z.y * w.y);
        }
      }
      __syncthreads();
      // ---- overlap-add: hop t = head(frame t) + tail(frame t-1); coalesced stores
      const int first = (f0 == hs - 1) ? 1 : 0;
      if (out_vec) {      // four samples per thread and iteration (rows 16-byte / 8-byte aligned: istft_launch checks)
        for (int i4 = tid + first * (N_S / 4); i4 < nf * (N_S / 4); i4 += 256) {
          const int fi = i4 >> 6, n = (i4 & 63) * 4;
          const float4 a = *reinterpret_cast<const float4*>(frame_of(fi) + n);
          const float4 c = fi > 0 ? *reinterpret_cast<const float4*>(frame_of(fi - 1) + N_S + n) : *reinterpret_cast<const float4*>(sm.carry + n);
          const float4 y = make_float4(a.x + c.x, a.y + c.y, a.z + c.z, a.w + c.w);
          const int64_t o = (int64_t)(f0 + fi) * N_S + n;
          if (of) __stcs(reinterpret_cast<float4*>(of + o), y);
          if (oi) {      // utils.py:28 truncation
            const short4 q = make_short4((short)__float2int_rz(y.x * 32768.0f), (short)__float2int_rz(y.y * 32768.0f),
                                         (short)__float2int_rz(y.z * 32768.0f), (short)__float2int_rz(y.w * 32768.0f));
            __stcs(reinterpret_cast<short4*>(oi + o), q);
          }
        }
      } else {
      for (int i = tid + first * N_S; i < nf * N_S; i += 256) {
        const int fi = i >> 8, n = i & 255;
        const int t = f0 + fi;
        float y = frame_of(fi)[n] + (fi > 0 ? frame_of(fi - 1)[N_S + n] : sm.carry[n]);
        const int64_t o = (int64_t)t * N_S + n;
        if (of) __stcs(of + o, y);
        if (oi) oi[o] = (int16_t)__float2int_rz(y * 32768.0f);   // utils.py:28 truncation
      }
      }
      __syncthreads();
      sm.carry[tid] = frame_of(nf - 1)[N_S + tid];
      __syncthreads();
    }
  }
}

// DXI_ENHANCE_EXACT=1 keeps the exact-order math for MMSE-LSA too (A/B against the MUFU-based path).
static bool exact_enhance() { static const bool v = [] { const char* e = getenv("DXI_ENHANCE_EXACT"); return e && *e && *e != '0'; }(); return v; }

static int istft_launch(int mode, const float* mag, const float* g_or_xbar, const float* phase, const float* mu,
                        const float* sigma, int gtype, const int32_t* n_frames, int B, int Tmax, float* wav_f32,
                        int16_t* wav_i16, int64_t out_stride, cudaStream_t st) {
  if (int rc = ensure_tables(st)) return rc;
  // Strip length: the persistent CTAs walk B * ceil((Tmax + 1) / h) strips of h hops (+ one recomputed leading frame each) in rounds;
  // h minimises rounds x (cost of a strip), i.e. the ragged last round is kept full (C2: 592 CTAs, h = 64 gave 5 rounds of which the
  // last was a third full; h = 70 gives 4 full ones).  A strip costs its frames in the per-bin spectrum phase (57 % of the kernel) and
  // whole passes of FR frames in the FFT / overlap-add phases.
  int n_sm = 148;
  { int dev = 0; cudaGetDevice(&dev); cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev); }
  const int64_t cap = (int64_t)n_sm * ISTFT_CTAS_PER_SM;
  int hops = 64;
  {
    int64_t best = -1;
    for (int h = 16; h <= HOPS_PER_STRIP_MAX; ++h) {
      const int64_t tot = (int64_t)B * ((Tmax + 1 + h - 1) / h), g = tot < cap ? tot : cap;
      const int64_t cost = ((tot + g - 1) / g) * (57 * (h + 1) + 43 * FR * ((h + 1 + FR - 1) / FR));
      if (best < 0 || cost < best) { best = cost; hops = h; }
    }
  }
  const int strips = (Tmax + 1 + hops - 1) / hops;
  const int64_t total = (int64_t)B * strips;
  const int grid = (int)(total < cap ? total : cap);
  const size_t smem = sizeof(IstftSmem);
  // four output samples per store when every row starts 16-byte (f32) / 8-byte (int16) aligned
  const int out_vec = (out_stride & 3) == 0 && (reinterpret_cast<uintptr_t>(wav_f32) & 15) == 0 && (reinterpret_cast<uintptr_t>(wav_i16) & 7) == 0;
  ProfScope prof(mode == 0 ? "istft" : "enhance", st, 1);
  if (mode == 0) {
    DXI_CUDA(cudaFuncSetAttribute(istft_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    istft_kernel<0><<<grid, 256, smem, st>>>(mag, g_or_xbar, phase, mu, sigma, gtype, n_frames, B, Tmax, strips, hops,
                                             wav_f32, wav_i16, out_stride, out_vec);
  } else if (gtype == DXI_G_MMSE_LSA && !exact_enhance()) {
    DXI_CUDA(cudaFuncSetAttribute(istft_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    istft_kernel<2><<<grid, 256, smem, st>>>(mag, g_or_xbar, phase, mu, sigma, gtype, n_frames, B, Tmax, strips, hops,
                                             wav_f32, wav_i16, out_stride, out_vec);
  } else {
    DXI_CUDA(cudaFuncSetAttribute(istft_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    istft_kernel<1><<<grid, 256, smem, st>>>(mag, g_or_xbar, phase, mu, sigma, gtype, n_frames, B, Tmax, strips, hops,
                                             wav_f32, wav_i16, out_stride, out_vec);
  }
  DXI_LAUNCHED("istft_kernel");
  return DXI_OK;
}

}  // namespace dxi

using namespace dxi;

extern "C" DXI_API int dxi_stft(const void* wav, int wav_is_i16, const int32_t* lens, int B, int64_t wav_stride, int Tmax,
                        float* mag, float* phase, void* stream) {
  if (int rc = check_device()) return rc;
  DXI_REQUIRE(wav && mag && phase, "dxi_stft: null argument");
  DXI_REQUIRE(B >= 0 && Tmax >= 0 && wav_stride >= 0, "dxi_stft: bad shape");
  if (B == 0 || Tmax == 0) return DXI_OK;
  cudaStream_t st = as_stream(stream);
  if (int rc = ensure_tables(st)) return rc;
  const int groups = (Tmax + FR - 1) / FR;
  const int64_t total = (int64_t)B * groups;
  DXI_REQUIRE(total < (1LL << 30), "dxi_stft: batch too large");
  const int grid = (int)(total < 148 * 4 ? total : 148 * 4);
  const size_t smem = wav_is_i16 ? sizeof(StftSmem<true>) : sizeof(StftSmem<false>);
  const int elem = wav_is_i16 ? 2 : 4;
  ProfScope prof("stft", st, 1);
  const int vec_ok = ((reinterpret_cast<uintptr_t>(wav) % 16) == 0 && ((wav_stride * elem) % 16) == 0) ? 1 : 0;
  if (wav_is_i16) {
    DXI_CUDA(cudaFuncSetAttribute(stft_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    stft_kernel<true><<<grid, 256, smem, st>>>(wav, lens, B, wav_stride, Tmax, groups, vec_ok, mag, phase);
  } else {
    DXI_CUDA(cudaFuncSetAttribute(stft_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    stft_kernel<false><<<grid, 256, smem, st>>>(wav, lens, B, wav_stride, Tmax, groups, vec_ok, mag, phase);
  }
  DXI_LAUNCHED("stft_kernel");
  return DXI_OK;
}

extern "C" DXI_API int dxi_istft(const float* mag, const float* gain, const float* phase, const int32_t* n_frames, int B,
                         int Tmax, float* wav_f32, int16_t* wav_i16, int64_t out_stride, void* stream) {
  if (int rc = check_device()) return rc;
  DXI_REQUIRE(mag && phase, "dxi_istft: null argument");
  DXI_REQUIRE(wav_f32 || wav_i16, "dxi_istft: no output requested");
  DXI_REQUIRE(B >= 0 && Tmax >= 0, "dxi_istft: bad shape");
  DXI_REQUIRE(out_stride >= (int64_t)(Tmax + 1) * N_S, "dxi_istft: out_stride < (Tmax+1)*256");
  if (B == 0 || Tmax == 0) return DXI_OK;
  return istft_launch(0, mag, gain, phase, nullptr, nullptr, 0, n_frames, B, Tmax, wav_f32, wav_i16, out_stride,
                      as_stream(stream));
}

extern "C" DXI_API int dxi_enhance(const float* mag, const float* phase, const float* xbar, const float* mu,
                           const float* sigma, int gtype, const int32_t* n_frames, int B, int Tmax, float* wav_f32,
                           int16_t* wav_i16, int64_t out_stride, void* stream) {
  if (int rc = check_device()) return rc;
  DXI_REQUIRE(mag && phase && xbar && mu && sigma, "dxi_enhance: null argument");
  DXI_REQUIRE(gtype >= DXI_G_MMSE_LSA && gtype <= DXI_G_DEEPMMSE, "Invalid gain function type.");
  DXI_REQUIRE(wav_f32 || wav_i16, "dxi_enhance: no output requested");
  DXI_REQUIRE(B >= 0 && Tmax >= 0, "dxi_enhance: bad shape");
  DXI_REQUIRE(out_stride >= (int64_t)(Tmax + 1) * N_S, "dxi_enhance: out_stride < (Tmax+1)*256");
  if (B == 0 || Tmax == 0) return DXI_OK;
  return istft_launch(1, mag, xbar, phase, mu, sigma, gtype, n_frames, B, Tmax, wav_f32, wav_i16, out_stride,
                      as_stream(stream));
}
