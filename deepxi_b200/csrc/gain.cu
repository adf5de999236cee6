// Element-wise map / gain kernels (HBM-streaming).
#include "gain_math.cuh"

namespace dxi {

// x_bar [n_rows, NB] -> xi_hat / gain / ibm.  One thread per element, 4 elements per thread in
// flight (grid-stride by whole-grid steps keeps every warp access a contiguous 128 B).
template <int UNROLL>
__global__ void __launch_bounds__(256) map_gain_kernel(const float* __restrict__ xbar, const float* __restrict__ mu,
                                                       const float* __restrict__ sigma, int64_t n, int n_bins,
                                                       int gtype, float* __restrict__ xi_out,
                                                       float* __restrict__ g_out, uint8_t* __restrict__ ibm_out,
                                                       const float* __restrict__ mag_sq = nullptr) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t base = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; base < n; base += stride * UNROLL) {
    float xb[UNROLL];
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) {
      int64_t i = base + u * stride;
      xb[u] = i < n ? __ldcs(xbar + i) : 0.5f;
    }
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) {
      int64_t i = base + u * stride;
      if (i >= n) continue;
      int k = (int)(i % n_bins);
      float xi = xi_from_xbar(xb[u], __ldg(mu + k), __ldg(sigma + k));
      if (xi_out) __stcs(xi_out + i, xi);
      if (g_out) {
        const float g = gfunc_eval(gtype, xi, __fadd_rn(xi, 1.0f));   // gamma_hat = xi_hat + 1 (inp_tgt.py:212)
        if (mag_sq) { const float m = __ldcs(mag_sq + i); __stcs(g_out + i, __fmul_rn(__fmul_rn(m, m), g)); }      // |X|^2 G (model.py:314-318)
        else __stcs(g_out + i, g);
      }
      if (ibm_out) ibm_out[i] = xi > 1.0f ? 1 : 0;
    }
  }
}

__global__ void __launch_bounds__(256) gfunc_kernel(const float* __restrict__ xi, const float* __restrict__ gamma,
                                                    int64_t n, int gtype, float* __restrict__ G) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    float x = __ldcs(xi + i);
    float g = gamma ? __ldcs(gamma + i) : 0.0f;
    __stcs(G + i, gfunc_eval(gtype, x, g));
  }
}

__global__ void __launch_bounds__(256) cdf_map_kernel(const float* __restrict__ xi, const float* __restrict__ mu,
                                                      const float* __restrict__ sigma, int64_t n, int n_bins,
                                                      float* __restrict__ xbar) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    int k = (int)(i % n_bins);
    __stcs(xbar + i, xbar_from_xi(__ldcs(xi + i), __ldg(mu + k), __ldg(sigma + k)));
  }
}

static inline int grid_for(int64_t n, int per_thread) {
  int64_t blocks = (n + 256LL * per_thread - 1) / (256LL * per_thread);
  const int64_t cap = 148LL * 8 * 4;   // multiple of the SM count x resident CTAs
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return (int)blocks;
}

static inline bool valid_gtype(int g) { return g >= DXI_G_MMSE_LSA && g <= DXI_G_DEEPMMSE; }

}  // namespace dxi

using namespace dxi;

extern "C" DXI_API int dxi_map_gain(const float* xbar, const float* mu, const float* sigma, int64_t n_rows, int n_bins,
                            int gtype, float* xi_hat, float* gain, uint8_t* ibm, void* stream) {
  if (int rc = check_device()) return rc;
  DXI_REQUIRE(xbar && mu && sigma, "dxi_map_gain: null input");
  DXI_REQUIRE(n_rows >= 0 && n_bins > 0, "dxi_map_gain: bad shape");
  DXI_REQUIRE(!gain || valid_gtype(gtype), "Invalid gain function type.");
  DXI_REQUIRE(xi_hat || gain || ibm, "dxi_map_gain: no output requested");
  int64_t n = n_rows * n_bins;
  if (n == 0) return DXI_OK;
  ProfScope prof("map_gain", as_stream(stream), 1);
  map_gain_kernel<4><<<grid_for(n, 4), 256, 0, as_stream(stream)>>>(xbar, mu, sigma, n, n_bins, gtype, xi_hat, gain, ibm);
  DXI_LAUNCHED("map_gain_kernel");
  return DXI_OK;
}

extern "C" DXI_API int dxi_deepmmse(const float* mag, const float* xbar, const float* mu, const float* sigma, int64_t n_rows, int n_bins,
                            float* d_psd, void* stream) {
  if (int rc = check_device()) return rc;
  DXI_REQUIRE(mag && xbar && mu && sigma && d_psd, "dxi_deepmmse: null argument");
  DXI_REQUIRE(n_rows >= 0 && n_bins > 0, "dxi_deepmmse: bad shape");
  int64_t n = n_rows * n_bins;
  if (n == 0) return DXI_OK;
  ProfScope prof("map_gain", as_stream(stream), 1);
  map_gain_kernel<4><<<grid_for(n, 4), 256, 0, as_stream(stream)>>>(xbar, mu, sigma, n, n_bins, DXI_G_DEEPMMSE, nullptr, d_psd, nullptr, mag);
  DXI_LAUNCHED("map_gain_kernel");
  return DXI_OK;
}

extern "C" DXI_API int dxi_gfunc(const float* xi, const float* gamma, int64_t n, int gtype, float* G, void* stream) {
  if (int rc = check_device()) return rc;
  DXI_REQUIRE(valid_gtype(gtype), "Invalid gain function type.");
  DXI_REQUIRE(xi && G && n >= 0, "dxi_gfunc: bad argument");
  bool needs_gamma = gtype == DXI_G_MMSE_LSA || gtype == DXI_G_MMSE_STSA || gtype == DXI_G_DEEPMMSE;
  DXI_REQUIRE(gamma || !needs_gamma, "dxi_gfunc: this gain function needs gamma");
  if (n == 0) return DXI_OK;
  ProfScope prof("gfunc", as_stream(stream), 1);
  gfunc_kernel<<<grid_for(n, 1), 256, 0, as_stream(stream)>>>(xi, gamma, n, gtype, G);
  DXI_LAUNCHED("gfunc_kernel");
  return DXI_OK;
}

extern "C" DXI_API int dxi_cdf_map(const float* xi, const float* mu, const float* sigma, int64_t n_rows, int n_bins,
                           float* xbar, void* stream) {
  if (int rc = check_device()) return rc;
  DXI_REQUIRE(xi && mu && sigma && xbar, "dxi_cdf_map: null argument");
  DXI_REQUIRE(n_rows >= 0 && n_bins > 0, "dxi_cdf_map: bad shape");
  int64_t n = n_rows * n_bins;
  if (n == 0) return DXI_OK;
  ProfScope prof("cdf_map", as_stream(stream), 1);
  cdf_map_kernel<<<grid_for(n, 1), 256, 0, as_stream(stream)>>>(xi, mu, sigma, n, n_bins, xbar);
  DXI_LAUNCHED("cdf_map_kernel");
  return DXI_OK;
}


// ---- subband IBM: xi_hat [rows, n_bins] x mel filter bank H [M, n_bins]^T -> subband a priori SNR and its
// mask (deepxi/model.py:323-328, deepxi/sig.py:301-346).  One warp per row: the row is read once (coalesced), every lane keeps
// ceil(n_bins / 32) values in registers and the M dot products are warp reductions; H (M x n_bins floats, 41 KB for 40 x 257)
// is read through L1.  HBM-bound: n_bins * 4 B in + M (+ 4 M) B out per row.
namespace dxi {
__global__ void __launch_bounds__(256) subband_kernel(const float* __restrict__ xi, const float* __restrict__ H, int64_t n_rows,
                                                      int n_bins, int M, float* __restrict__ xi_sub, uint8_t* __restrict__ ibm) {
  const int lane = threadIdx.x & 31;
  for (int64_t row = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5); row < n_rows; row += (int64_t)gridDim.x * 8) {
    float x[9];
#pragma unroll
    for (int i = 0; i < 9; ++i) { const int k = lane + 32 * i; x[i] = k < n_bins ? __ldcs(xi + row * n_bins + k) : 0.0f; }
    for (int m = 0; m < M; ++m) {
      float a = 0.0f;
#pragma unroll
      for (int i = 0; i < 9; ++i) { const int k = lane + 32 * i; if (k < n_bins) a = fmaf(x[i], __ldg(H + (size_t)m * n_bins + k), a); }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
      if (lane == 0) {
        if (xi_sub) xi_sub[row * M + m] = a;
        if (ibm) ibm[row * M + m] = a > 1.0f ? 1 : 0;
      }
    }
  }
}
}  // namespace dxi

extern "C" DXI_API int dxi_subband_ibm(const float* xi, const float* H, int64_t n_rows, int n_bins, int M, float* xi_sub,
                               uint8_t* ibm, void* stream) {
  if (int rc = check_device()) return rc;
  DXI_REQUIRE(xi && H && (xi_sub || ibm), "dxi_subband_ibm: null argument");
  DXI_REQUIRE(n_rows >= 0 && n_bins > 0 && n_bins <= 288 && M > 0, "dxi_subband_ibm: bad shape (n_bins <= 288)");
  if (n_rows == 0) return DXI_OK;
  const int64_t blocks = (n_rows + 7) / 8;
  ProfScope prof("subband", as_stream(stream), 1);
  dxi::subband_kernel<<<(int)(blocks < 148 * 8 ? blocks : 148 * 8), 256, 0, as_stream(stream)>>>(xi, H, n_rows, n_bins, M, xi_sub, ibm);
  DXI_LAUNCHED("subband_kernel");
  return DXI_OK;
}
