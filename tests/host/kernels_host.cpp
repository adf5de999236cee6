// Host build of the device math headers (fft.cuh, gain_math.cuh) for CPU-only tests.
// The CUDA kernels' per-frame arithmetic (radix-16 passes, exchange indexing, split/merge steps,
// atan2 polynomial, E1, gain formulas) is compiled unchanged with g++ through the shims below and
// checked against the oracle in tests/test_host_math.py.  Intrinsics that exist only on the device
// are mapped to their IEEE host equivalents.
#include <cmath>
#include <cstdint>
#include <cstring>
#include <cuda_runtime.h>

static inline float __fdividef(float a, float b) { return a / b; }
static inline float __fdiv_rn(float a, float b) { return a / b; }
static inline float __fadd_rn(float a, float b) { return a + b; }
static inline float __fsub_rn(float a, float b) { return a - b; }
static inline float __fmul_rn(float a, float b) { return a * b; }
static inline float cyl_bessel_i0f(float x) { return (float)std::cyl_bessel_i(0.0, (double)x); }
static inline float cyl_bessel_i1f(float x) { return (float)std::cyl_bessel_i(1.0, (double)x); }
// erfinv is supplied by the test through this hook (scipy's), the device uses CUDA's erfinvf
static double (*g_erfinv)(double) = nullptr;
static inline float erfinvf(float x) { return (float)g_erfinv((double)x); }
static inline double erfinv(double x) { return g_erfinv(x); }
static inline float exp10f_host(float x) { return (float)std::pow(10.0, (double)x); }
#define exp10f exp10f_host
using std::isnan;
using std::isinf;

#include "../../deepxi_b200/csrc/fft.cuh"
#include "../../deepxi_b200/csrc/gain_math.cuh"

using namespace dxi;

static float g_win[512], g_swin[512];
static float2 g_tw256[256], g_tw512[258];
static bool g_init = false;
static void init_tables() {
  if (g_init) return;
  double w[512];
  for (int n = 0; n < 512; ++n) w[n] = 0.54 - 0.46 * std::cos(2.0 * M_PI * n / 511.0);
  for (int n = 0; n < 512; ++n) {
    int m = n % 256;
    double den = w[m] * w[m] + w[m + 256] * w[m + 256];
    g_win[n] = (float)w[n];
    g_swin[n] = (float)(w[n] / den) * (1.0f / 256.0f);
  }
  for (int m = 0; m < 256; ++m) g_tw256[m] = make_float2((float)std::cos(2.0 * M_PI * m / 256.0), (float)-std::sin(2.0 * M_PI * m / 256.0));
  for (int k = 0; k < 257; ++k) g_tw512[k] = make_float2((float)std::cos(2.0 * M_PI * k / 512.0), (float)-std::sin(2.0 * M_PI * k / 512.0));
  g_init = true;
}

extern "C" {

void host_set_erfinv(double (*fn)(double)) { g_erfinv = fn; }

// One analysis frame: x[512] (already normalised) -> mag[257], phase[257]; mirrors stft_kernel.
void host_stft_frame(const float* x, float* mag, float* phase) {
  init_tables();
  float2 buf[FFT_FRAME_SLOTS];
  float2 v[16][16];
  for (int lane = 0; lane < 16; ++lane) {
    for (int n1 = 0; n1 < 16; ++n1) {
      int n = 32 * n1 + 2 * lane;
      v[lane][n1] = make_float2(x[n] * (0.5f * g_win[n]), x[n + 1] * (0.5f * g_win[n + 1]));      // half-scaled, as in the kernel
    }
    fft256_pass1<-1>(v[lane], buf, g_tw256, lane);
  }
  for (int lane = 0; lane < 16; ++lane) fft256_pass2<-1>(v[lane], buf, lane);
  for (int lane = 0; lane < 16; ++lane)
    for (int k2 = 0; k2 < 16; ++k2) buf[lane + 16 * k2] = v[lane][fft16_pos(k2)];
  for (int j = 0; j < 128; ++j) {      // bin pairs (j, 256 - j)
    float2 a, c, m, ph;
    rfft_split_pair(buf[j], buf[(256 - j) & 255], g_tw512[j], a, c);
    if (j == 0) { a.y = 0.0f; c.y = 0.0f; }
    polar_pair(a, c, m, ph);
    mag[j] = m.x; mag[256 - j] = m.y;
    phase[j] = ph.x; phase[256 - j] = ph.y;
  }
  {
    const float2 X = make_float2(2.0f * buf[128].x, -2.0f * buf[128].y);
    mag[128] = sqrt_approx(fmaf(X.x, X.x, X.y * X.y));
    phase[128] = atan2_poly(X.y, X.x);
  }
}

// One synthesis frame: mag[257], phase[257] -> windowed time frame out[512]; mirrors istft_kernel.
void host_istft_frame(const float* mag, const float* phase, float* out) {
  init_tables();
  float2 buf[FFT_FRAME_SLOTS];
  for (int k = 0; k < 257; ++k) {
    float sn = sinf(phase[k]), cs = cosf(phase[k]);
    buf[k] = make_float2(mag[k] * cs, (k == 0 || k == 256) ? 0.0f : mag[k] * sn);
  }
  float2 v[16][16];
  for (int lane = 0; lane < 16; ++lane)
    for (int n1 = 0; n1 < 16; ++n1) {
      int k = 16 * n1 + lane;
      v[lane][n1] = irfft_merge(buf[k], buf[256 - k], g_tw512[k]);
    }
  for (int lane = 0; lane < 16; ++lane) fft256_pass1<1>(v[lane], buf, g_tw256, lane);
  for (int lane = 0; lane < 16; ++lane) {
    fft256_pass2<1>(v[lane], buf, lane);
    for (int k2 = 0; k2 < 16; ++k2) {
      int n = 2 * (lane + 16 * k2);
      float2 z = v[lane][fft16_pos(k2)];
      out[n] = z.x * g_swin[n];
      out[n + 1] = z.y * g_swin[n + 1];
    }
  }
}

void host_xi_from_xbar(const float* xbar, const float* mu, const float* sigma, int n_rows, int n_bins, float* xi) {
  for (int r = 0; r < n_rows; ++r)
    for (int k = 0; k < n_bins; ++k) xi[r * n_bins + k] = xi_from_xbar(xbar[r * n_bins + k], mu[k], sigma[k]);
}

void host_xbar_from_xi(const float* xi, const float* mu, const float* sigma, int n_rows, int n_bins, float* xbar) {
  for (int r = 0; r < n_rows; ++r)
    for (int k = 0; k < n_bins; ++k) xbar[r * n_bins + k] = xbar_from_xi(xi[r * n_bins + k], mu[k], sigma[k]);
}

void host_gfunc(const float* xi, const float* gamma, int n, int gtype, float* G) {
  for (int i = 0; i < n; ++i) G[i] = gfunc_eval(gtype, xi[i], gamma ? gamma[i] : 0.0f);
}

float host_e1(float x) { return expint_e1(x); }
float host_e1_fast(float x) { return expint_e1_fast(x); }
float host_erfinv_fast(float x) { return erfinv_fast(x); }
void host_lsa_from_xbar_fast(const float* xbar, const float* mu, const float* sigma, int rows, int bins, float* G) {
  for (int r = 0; r < rows; ++r)
    for (int k = 0; k < bins; ++k)
      G[r * bins + k] = lsa_gain_from_xbar_fast(xbar[r * bins + k], mu[k], __fmul_rn(sigma[k], 1.41421354f), sigma[k]);
}

// the table form of the same chain (what the fused enhancement kernel runs)
void host_lsa_from_xbar_tab(const float* xbar, const float* mu, const float* sigma, int rows, int bins, float* G) {
  static float4 tab[LSA_TAB_N];
  static bool ready = false;
  if (!ready) { lsa_table_build(tab); ready = true; }
  for (int r = 0; r < rows; ++r)
    for (int k = 0; k < bins; ++k)
      G[r * bins + k] = lsa_gain_from_xbar_tab(xbar[r * bins + k], mu[k], __fmul_rn(sigma[k], 1.41421354f), sigma[k], tab);
}
}
