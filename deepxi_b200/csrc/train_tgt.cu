// Training-target pipeline: noisy-speech mixing, instantaneous a priori SNR, mapped target and the
// per-bin statistics of xi_dB that parameterise the CDF map.
//
//   dxi_mix            InputTarget.mix / add_noise / add_noise_pad      deepxi/sig.py:162-284
//   dxi_xi_map         InputTarget.xi + NormalCDF.map (MagXi.example)   deepxi/sig.py:110-121, inp_tgt.py:173-196, map.py:356-371
//   dxi_xi_db_moments  MagXi.stats / NormalCDF.stats                    deepxi/inp_tgt.py:160-171, map.py:392-402
//
// All three are streaming kernels (HBM-bound): every waveform sample / spectrum bin is read once.  The statistics are
// accumulated as (count, sum, sum of squares) per bin in float64, which is what lets several GPUs combine their
// shards with one 3 x 257 all-reduce (the only collective on any Deep Xi path).
#include "gain_math.cuh"

namespace dxi {

// ---- mix: pass 1, signal powers -----------------------------------------------------------------------
// grid (chunks, B); ws[b][0] += sum s^2, ws[b][1] += sum d[off : off + s_len]^2 (normalised samples, float64)
__global__ void __launch_bounds__(256) mix_power_kernel(const int16_t* __restrict__ s, const int16_t* __restrict__ d,
                                                        const int32_t* __restrict__ s_len, const int32_t* __restrict__ d_len,
                                                        const int32_t* __restrict__ offsets, int64_t s_stride, int64_t d_stride,
                                                        double* __restrict__ ws) {
  const int b = blockIdx.y;
  const int n = s_len[b];
  const int off = offsets[b];
  const int16_t* sb = s + (int64_t)b * s_stride;
  const int16_t* db = d + (int64_t)b * d_stride + off;
  float ps = 0.0f, pd = 0.0f;      // per-thread partial sums of <= a few hundred squares: float32 is ample
  for (int i = blockIdx.x * 256 + threadIdx.x; i < n; i += gridDim.x * 256) {
    const float sv = (float)sb[i] * (1.0f / 32768.0f);
    const float dv = (off + i < d_len[b]) ? (float)db[i] * (1.0f / 32768.0f) : 0.0f;
    ps = fmaf(sv, sv, ps);
    pd = fmaf(dv, dv, pd);
  }
  double a = ps, c = pd;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) { a += __shfl_xor_sync(0xffffffffu, a, o); c += __shfl_xor_sync(0xffffffffu, c, o); }
  __shared__ double red[2][8];
  if ((threadIdx.x & 31) == 0) { red[0][threadIdx.x >> 5] = a; red[1][threadIdx.x >> 5] = c; }
  __syncthreads();
  if (threadIdx.x < 2) {
    double t = 0.0;
    for (int w = 0; w < 8; ++w) t += red[threadIdx.x][w];
    atomicAdd(ws + 2 * b + threadIdx.x, t);
  }
}

// ---- mix: pass 2, scale the noise section and add ---------------------------------------------------------
__global__ void __launch_bounds__(256) mix_apply_kernel(const int16_t* __restrict__ s, const int16_t* __restrict__ d,
                                                        const int32_t* __restrict__ s_len, const int32_t* __restrict__ d_len,
                                                        const float* __restrict__ snr_db, const int32_t* __restrict__ offsets,
                                                        int64_t s_stride, int64_t d_stride, const double* __restrict__ ws,
                                                        float* __restrict__ s_out, float* __restrict__ d_out,
                                                        float* __restrict__ x_out, int64_t out_stride) {
  const int b = blockIdx.y;
  const int n = s_len[b];
  const int off = offsets[b];
  // sig.py:276-282: snr = 10^(snr/10); alpha = sqrt(P_s / max(P_d snr, 1e-12)), means over the s_len samples
  const float snr = powf(10.0f, __fdiv_rn(snr_db[b], 10.0f));
  const float P_s = (float)(ws[2 * b] / (double)max(n, 1)), P_d = (float)(ws[2 * b + 1] / (double)max(n, 1));
  const float alpha = sqrtf(__fdiv_rn(P_s, fmaxf(__fmul_rn(P_d, snr), 1e-12f)));
  const int16_t* sb = s + (int64_t)b * s_stride;
  const int16_t* db = d + (int64_t)b * d_stride + off;
  for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < out_stride; i += (int64_t)gridDim.x * 256) {
    float sv = 0.0f, dv = 0.0f;
    if (i < n) {
      sv = (float)sb[i] * (1.0f / 32768.0f);
      dv = (off + i < d_len[b]) ? __fmul_rn((float)db[i] * (1.0f / 32768.0f), alpha) : 0.0f;
    }
    const int64_t o = (int64_t)b * out_stride + i;
    if (s_out) __stcs(s_out + o, sv);
    if (d_out) __stcs(d_out + o, dv);
    if (x_out) __stcs(x_out + o, __fadd_rn(sv, dv));
  }
}

// ---- xi (+ map) ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float xi_inst(float S, float D) {      // sig.py:110-121
  return __fdiv_rn(__fmul_rn(S, S), fmaxf(__fmul_rn(D, D), 1e-12f));
}

__global__ void __launch_bounds__(256) xi_map_kernel(const float* __restrict__ S, const float* __restrict__ D,
                                                     const float* __restrict__ mu, const float* __restrict__ sigma, int64_t n,
                                                     int n_bins, float* __restrict__ xi, float* __restrict__ xi_bar) {
  for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (int64_t)gridDim.x * 256) {
    const int k = (int)(i % n_bins);
    const float v = xi_inst(__ldcs(S + i), __ldcs(D + i));
    if (xi) __stcs(xi + i, v);
    if (xi_bar) __stcs(xi_bar + i, xbar_from_xi(v, __ldg(mu + k), __ldg(sigma + k)));
  }
}

// ---- statistics ---------------------------------------------------------------------------------------------
// grid (frame chunks, B); thread k owns bin k (thread 0 also bin 256 for n_bins = 257 ...): in general bins k, k + 256, ...
constexpr int MOM_ROWS = 64;      // frames per CTA
__global__ void __launch_bounds__(256) xi_db_moments_kernel(const float* __restrict__ S, const float* __restrict__ D,
                                                            const int32_t* __restrict__ n_frames, int Tmax, int n_bins,
                                                            double* __restrict__ acc) {
  const int b = blockIdx.y;
  const int T = n_frames ? min(n_frames[b], Tmax) : Tmax;
  const int t0 = blockIdx.x * MOM_ROWS, t1 = min(t0 + MOM_ROWS, T);
  if (t0 >= t1) return;
  for (int k = threadIdx.x; k < n_bins; k += 256) {
    double s1 = 0.0, s2 = 0.0;
    for (int t = t0; t < t1; ++t) {
      const int64_t i = ((int64_t)b * Tmax + t) * n_bins + k;
      const float v = fmaxf(xi_inst(__ldcs(S + i), __ldcs(D + i)), 1e-12f);
      const double xdb = (double)__fmul_rn(10.0f, __fdiv_rn(logf(v), 2.30258512f));      // map.py:62-73 in float32
      s1 += xdb;
      s2 = fma(xdb, xdb, s2);
    }
    atomicAdd(acc + k, (double)(t1 - t0));
    atomicAdd(acc + n_bins + k, s1);
    atomicAdd(acc + 2 * n_bins + k, s2);
  }
}

}  // namespace dxi

using namespace dxi;

extern "C" DXI_API int64_t dxi_mix_workspace_bytes(int B) { return (int64_t)B * 2 * sizeof(double); }

extern "C" DXI_API int dxi_mix(const int16_t* s, const int16_t* d, const int32_t* s_len, const int32_t* d_len, const float* snr_db,
                       const int32_t* offsets, int B, int64_t s_stride, int64_t d_stride, float* s_out, float* d_out,
                       float* x_out, int64_t out_stride, void* workspace, void* stream) {
  if (int rc = check_device()) return rc;
  DXI_REQUIRE(s && d && s_len && d_len && snr_db && offsets && workspace, "dxi_mix: null argument");
  DXI_REQUIRE(s_out || d_out || x_out, "dxi_mix: no output requested");
  DXI_REQUIRE(B >= 0 && s_stride >= 0 && d_stride >= 0 && out_stride >= 0, "dxi_mix: bad shape");
  if (B == 0 || out_stride == 0) return DXI_OK;
  cudaStream_t st = as_stream(stream);
  double* ws = reinterpret_cast<double*>(workspace);
  DXI_CUDA(cudaMemsetAsync(ws, 0, (size_t)B * 2 * sizeof(double), st));
  const int chunks = (int)((s_stride + 256 * 16 - 1) / (256 * 16));
  dim3 grid(chunks < 1 ? 1 : (chunks > 64 ? 64 : chunks), B);
  ProfScope prof("mix", st, 2);
  mix_power_kernel<<<grid, 256, 0, st>>>(s, d, s_len, d_len, offsets, s_stride, d_stride, ws);
  DXI_LAUNCHED("mix_power_kernel");
  mix_apply_kernel<<<grid, 256, 0, st>>>(s, d, s_len, d_len, snr_db, offsets, s_stride, d_stride, ws, s_out, d_out, x_out,
                                         out_stride);
  DXI_LAUNCHED("mix_apply_kernel");
  return DXI_OK;
}

extern "C" DXI_API int dxi_xi_map(const float* S, const float* D, const float* mu, const float* sigma, int64_t n_rows, int n_bins,
                          float* xi, float* xi_bar, void* stream) {
  if (int rc = check_device()) return rc;
  DXI_REQUIRE(S && D, "dxi_xi_map: null argument");
  DXI_REQUIRE(xi || xi_bar, "dxi_xi_map: no output requested");
  DXI_REQUIRE(!xi_bar || (mu && sigma), "dxi_xi_map: the mapped target needs mu and sigma");
  DXI_REQUIRE(n_rows >= 0 && n_bins > 0, "dxi_xi_map: bad shape");
  const int64_t n = n_rows * n_bins;
  if (n == 0) return DXI_OK;
  cudaStream_t st = as_stream(stream);
  const int64_t blocks = (n + 255) / 256;
  ProfScope prof("xi_map", st, 1);
  xi_map_kernel<<<(int)(blocks < 148 * 8 ? blocks : 148 * 8), 256, 0, st>>>(S, D, mu, sigma, n, n_bins, xi, xi_bar);
  DXI_LAUNCHED("xi_map_kernel");
  return DXI_OK;
}

extern "C" DXI_API int dxi_xi_db_moments(const float* S, const float* D, const int32_t* n_frames, int B, int Tmax, int n_bins,
                                 double* acc, void* stream) {
  if (int rc = check_device()) return rc;
  DXI_REQUIRE(S && D && acc, "dxi_xi_db_moments: null argument");
  DXI_REQUIRE(B >= 0 && Tmax >= 0 && n_bins > 0, "dxi_xi_db_moments: bad shape");
  if (B == 0 || Tmax == 0) return DXI_OK;
  cudaStream_t st = as_stream(stream);
  dim3 grid((Tmax + MOM_ROWS - 1) / MOM_ROWS, B);
  ProfScope prof("xi_db_moments", st, 1);
  xi_db_moments_kernel<<<grid, 256, 0, st>>>(S, D, n_frames, Tmax, n_bins, acc);
  DXI_LAUNCHED("xi_db_moments_kernel");
  return DXI_OK;
}
