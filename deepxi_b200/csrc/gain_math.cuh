// Device math for the inverse CDF map and the gain functions.
//
// Follows the float32 operation order of the reference so that results agree with its TF/numpy
// chain to a few ulp:  deepxi/map.py:373-390 (NormalCDF.inverse), :62-85 (db / db_inverse),
// deepxi/gain.py:13-166 (mmse_stsa, mmse_lsa, wf, srwf, cwf, irm, ibm, deepmmse).
#pragma once
#include <cmath>
#include "common.cuh"

namespace dxi {

__device__ __forceinline__ float wf_gain(float xi) { return __fdiv_rn(xi, __fadd_rn(xi, 1.0f)); }

// E1(x), x > 0, float32.  Power series for x <= 2, even-form continued fraction above.
// (The reference calls scipy.special.exp1, gain.py:67; abs error here <= 2e-6 at x=1e-12 where
// E1 ~ 27 is itself only representable to 1.9e-6, <= 2e-7 for x >= 1e-3.)
__device__ __forceinline__ float expint_e1(float x) {
  if (x <= 2.0f) {
    // sum_{k=1..16} (-1)^{k+1} x^k / (k k!)
    float p = -2.9868938e-15f;                 // k=16
    p = fmaf(p, x, 5.0981123e-14f);            // k=15
    p = fmaf(p, x, -8.1933948e-13f);           // k=14
    p = fmaf(p, x, 1.2353000e-11f);            // k=13
    p = fmaf(p, x, -1.7397141e-10f);           // k=12
    p = fmaf(p, x, 2.2774621e-9f);             // k=11
    p = fmaf(p, x, -2.7557319e-8f);            // k=10
    p = fmaf(p, x, 3.0619244e-7f);             // k=9
    p = fmaf(p, x, -3.1001984e-6f);            // k=8
    p = fmaf(p, x, 2.8344671e-5f);             // k=7
    p = fmaf(p, x, -2.3148148e-4f);            // k=6
    p = fmaf(p, x, 1.6666667e-3f);             // k=5
    p = fmaf(p, x, -1.0416667e-2f);            // k=4
    p = fmaf(p, x, 5.5555556e-2f);             // k=3
    p = fmaf(p, x, -0.25f);                    // k=2
    p = fmaf(p, x, 1.0f);                      // k=1
    return fmaf(p, x, -0.57721566490153286f - logf(x));
  }
  // E1 = e^{-x} / (x+1 - 1/(x+3 - 4/(x+5 - 9/(x+7 - ...)))), 8 levels: abs error < 1.5e-7 for x >= 2
  float d = x + 17.0f;
#pragma unroll
  for (int k = 8; k >= 1; --k) d = (x + (float)(2 * k - 1)) - __fdividef((float)(k * k), d);
  return __fdividef(expf(-x), d);
}

// mmse_lsa (gain.py:47-69)
__device__ __forceinline__ float gain_mmse_lsa(float xi, float gamma) {
  xi = fmaxf(xi, 1e-12f);
  gamma = fmaxf(gamma, 1e-12f);
  float v1 = __fdiv_rn(xi, __fadd_rn(1.0f, xi));
  float nu = __fmul_rn(v1, gamma);
  float v2 = expint_e1(nu);
  return __fmul_rn(v1, expf(__fmul_rn(0.5f, v2)));
}

// mmse_stsa (gain.py:13-45): the float32 formula with unscaled Bessel functions; where it
// overflows to Inf/NaN (nu >~ 173) the reference substitutes the Wiener gain (gain.py:42-44).
__device__ __forceinline__ float gain_mmse_stsa(float xi, float gamma) {
  xi = fmaxf(xi, 1e-12f);
  gamma = fmaxf(gamma, 1e-12f);
  float nu = __fmul_rn(xi, __fdiv_rn(gamma, __fadd_rn(1.0f, xi)));
  float a = __fmul_rn(__fdiv_rn(1.7724539f, 2.0f), __fdiv_rn(sqrtf(nu), gamma));
  float b = __fmul_rn(a, expf(__fdiv_rn(-nu, 2.0f)));
  float h = __fdiv_rn(nu, 2.0f);
  float c = __fadd_rn(__fmul_rn(__fadd_rn(1.0f, nu), cyl_bessel_i0f(h)), __fmul_rn(nu, cyl_bessel_i1f(h)));
  float G = __fmul_rn(b, c);
  if (isnan(G) || isinf(G)) G = wf_gain(xi);
  return G;
}

// deepmmse (gain.py:154-166)
__device__ __forceinline__ float gain_deepmmse(float xi, float gamma) {
  float op = __fadd_rn(1.0f, xi);
  return __fadd_rn(__fdiv_rn(1.0f, op), __fdiv_rn(xi, __fmul_rn(gamma, op)));
}

// gfunc dispatch (gain.py:168-191).  gtype is warp-uniform.
__device__ __forceinline__ float gfunc_eval(int gtype, float xi, float gamma) {
  switch (gtype) {
    case DXI_G_MMSE_LSA:  return gain_mmse_lsa(xi, gamma);
    case DXI_G_MMSE_STSA: return gain_mmse_stsa(xi, gamma);
    case DXI_G_WF:        return wf_gain(xi);
    case DXI_G_SRWF:
    case DXI_G_IRM:       return sqrtf(wf_gain(xi));
    case DXI_G_CWF:       return wf_gain(sqrtf(xi));
    case DXI_G_IBM:       return xi > 1.0f ? 1.0f : 0.0f;
    case DXI_G_DEEPMMSE:  return gain_deepmmse(xi, gamma);
  }
  return 0.0f;
}

// NormalCDF.inverse for 'DBNormalCDF' (map.py:373-390): xi = 10^((sigma*sqrt(2)*erfinv(2 xbar - 1) + mu)/10).
// Within 1e-3 dB of the 0 dB threshold the chain is re-evaluated with a double-precision erfinv and
// the reference's float32 roundings, so that (xi > 1) -- the IBM -- is decided exactly as the
// float32 chain of the reference decides it.
__device__ __forceinline__ float xi_from_xbar(float xbar, float mu, float sigma) {
  float v1 = __fmul_rn(sigma, 1.41421354f);            // sigma * f32(sqrt(2))
  float v2 = __fmul_rn(2.0f, xbar);
  float arg = __fsub_rn(v2, 1.0f);
  float v3 = erfinvf(arg);
  float x = __fadd_rn(__fmul_rn(v1, v3), mu);
  if (fabsf(x) < 1e-3f) {
    v3 = (float)erfinv((double)arg);
    x = __fadd_rn(__fmul_rn(v1, v3), mu);
    return (float)pow(10.0, (double)__fdiv_rn(x, 10.0f));
  }
  return exp10f(__fdiv_rn(x, 10.0f));
}

// NormalCDF.map for 'DBNormalCDF' (map.py:356-371, :62-73)
__device__ __forceinline__ float xbar_from_xi(float xi, float mu, float sigma) {
  xi = fmaxf(xi, 1e-12f);
  float xdb = __fmul_rn(10.0f, __fdiv_rn(logf(xi), 2.30258512f));
  float v1 = __fsub_rn(xdb, mu);
  float v2 = __fmul_rn(sigma, 1.41421354f);
  float v3 = erff(__fdiv_rn(v1, v2));
  return __fmul_rn(0.5f, __fadd_rn(1.0f, v3));
}

// ---- fast path of the fused enhancement kernel: x_bar -> G_LSA(xi_hat, xi_hat + 1) ---------------------------------
// MUFU-based (ex2 / lg2 / rcp approximations, ~2^-22 relative) restatement of xi_from_xbar + gain_mmse_lsa for the one
// combination the reference's default inference uses (inp_tgt.py:198-214 with gtype 'mmse-lsa': gamma_hat = xi_hat + 1, so
// nu = xi_hat and G depends on xi_hat alone).  Relative error of G against the float64 chain < 1e-5 (dominated by the
// float32 erfinv amplified by sigma), i.e. > 100 dB on the waveform; the exact-order functions above stay the ones behind
// dxi_map_gain / dxi_gfunc and every other gain type.
#ifdef __CUDA_ARCH__
__device__ __forceinline__ float fast_ex2(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float fast_lg2(float x) { float y; asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float fast_rcp(float x) { float y; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float fast_sqrt(float x) { float y; asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
#else
static inline float fast_ex2(float x) { return exp2f(x); }
static inline float fast_lg2(float x) { return log2f(x); }
static inline float fast_rcp(float x) { return 1.0f / x; }
static inline float fast_sqrt(float x) { return sqrtf(x); }
#endif

// erfinv(x), |x| < 1 (M. Giles, "Approximating the erfinv function", single-precision polynomials in w = -ln(1 - x^2))
__device__ __forceinline__ float erfinv_fast(float x) {
  float w = -0.69314718f * fast_lg2(fmaf(-x, x, 1.0f));
  float p;
  if (w < 5.0f) {
    w -= 2.5f;
    p = 2.81022636e-08f;
    p = fmaf(p, w, 3.43273939e-07f);
    p = fmaf(p, w, -3.5233877e-06f);
    p = fmaf(p, w, -4.39150654e-06f);
    p = fmaf(p, w, 0.00021858087f);
    p = fmaf(p, w, -0.00125372503f);
    p = fmaf(p, w, -0.00417768164f);
    p = fmaf(p, w, 0.246640727f);
    p = fmaf(p, w, 1.50140941f);
  } else {
    w = fast_sqrt(w) - 3.0f;
    p = -0.000200214257f;
    p = fmaf(p, w, 0.000100950558f);
    p = fmaf(p, w, 0.00134934322f);
    p = fmaf(p, w, -0.00367342844f);
    p = fmaf(p, w, 0.00573950773f);
    p = fmaf(p, w, -0.0076224613f);
    p = fmaf(p, w, 0.00943887047f);
    p = fmaf(p, w, 1.00167406f);
    p = fmaf(p, w, 2.83297682f);
  }
  return p * x;
}

// E1(x), x >= 1e-12: power series (9 terms) up to 1, Abramowitz & Stegun 5.1.56 (|eps| < 2e-8 on x e^x E1) above
__device__ __forceinline__ float expint_e1_fast(float x) {
  if (x <= 1.0f) {
    float p = 3.0619244e-7f;                   // k=9
    p = fmaf(p, x, -3.1001984e-6f);
    p = fmaf(p, x, 2.8344671e-5f);
    p = fmaf(p, x, -2.3148148e-4f);
    p = fmaf(p, x, 1.6666667e-3f);
    p = fmaf(p, x, -1.0416667e-2f);
    p = fmaf(p, x, 5.5555556e-2f);
    p = fmaf(p, x, -0.25f);
    p = fmaf(p, x, 1.0f);
    return fmaf(p, x, fmaf(-0.69314718f, fast_lg2(x), -0.57721566f));
  }
  const float r = fast_rcp(x);
  float n = 0.2677737343f;
  n = fmaf(n, r, 8.6347608925f);
  n = fmaf(n, r, 18.0590169730f);
  n = fmaf(n, r, 8.5733287401f);
  n = fmaf(n, r, 1.0f);
  float d = 3.9584969228f;
  d = fmaf(d, r, 21.0996530827f);
  d = fmaf(d, r, 25.6329561486f);
  d = fmaf(d, r, 9.5733223454f);
  d = fmaf(d, r, 1.0f);
  return fast_ex2(-1.44269504f * x) * r * n * fast_rcp(d);
}

// s2 = sigma * f32(sqrt(2)).  Arguments on or outside the open interval (erfinv = +-inf, NaN) take the exact-order path.
__device__ __forceinline__ float lsa_gain_from_xbar_fast(float xbar, float mu, float s2, float sigma) {
  const float arg = fmaf(2.0f, xbar, -1.0f);
  if (!(fabsf(arg) < 1.0f)) {
    const float xi = xi_from_xbar(xbar, mu, sigma);
    return gain_mmse_lsa(xi, __fadd_rn(xi, 1.0f));
  }
  const float xdb = fminf(fmaf(s2, erfinv_fast(arg), mu), 300.0f);
  const float xi = fmaxf(fast_ex2(0.33219281f * xdb), 1e-12f);
  const float wf = xi * fast_rcp(1.0f + xi);
  return wf * fast_ex2(0.72134752f * expint_e1_fast(xi));
}

// ---- table form of the same gain -------------------------------------------------------------------------------------
// With gamma_hat = xi_hat + 1 the MMSE-LSA gain is a function of xi_hat alone, smooth and monotone in dB:
//   g(t) = log2 G,  G = xi / (1 + xi) * exp(E1(xi) / 2),  xi = 10^(t / 10).
// g is tabulated as piecewise cubics (Hermite data from the analytic derivative, float64 on the host) on [LSA_TAB_LO, LSA_TAB_HI] dB
// in steps of 0.5 dB: |2^cubic / G - 1| < 1e-7.  The table starts where xi_hat is clamped (1e-12 = -120 dB, gain.py:60), so the clamp is
// the clamp of the table position; above the table G = 1 - O(1e-6).  One LDS.128 + 3 FMA + ex2 replace the E1 branches, two ex2 and an rcp.
constexpr int LSA_TAB_N = 360;
constexpr float LSA_TAB_LO = -120.0f, LSA_TAB_HI = 60.0f, LSA_TAB_INV_H = 2.0f;

// Host: the LSA_TAB_N cubics c0 + c1 f + c2 f^2 + c3 f^3, f in [0, 1) the position inside the interval.
static inline void lsa_table_build(float4* tab) {
  const double h = 1.0 / LSA_TAB_INV_H, ln2 = 0.69314718055994531, ln10 = 2.3025850929940457;
  auto gval = [&](double t, double& g, double& d) {
    const double xi = pow(10.0, t / 10.0);
    const double e1 = -std::expint(-xi);
    g = (log(xi) - log1p(xi) + 0.5 * e1) / ln2;
    d = (xi * ln10 / 10.0) * (1.0 / xi - 1.0 / (1.0 + xi) - 0.5 * exp(-xi) / xi) / ln2 * h;      // dg/dt * h
  };
  for (int i = 0; i < LSA_TAB_N; ++i) {
    double g0, d0, g1, d1;
    gval(LSA_TAB_LO + i * h, g0, d0);
    gval(LSA_TAB_LO + (i + 1) * h, g1, d1);
    tab[i] = make_float4((float)g0, (float)d0, (float)(3.0 * (g1 - g0) - 2.0 * d0 - d1), (float)(2.0 * (g0 - g1) + d0 + d1));
  }
}

// The exact-order chain for the arguments the table form does not take (erfinv = +-inf, NaN: x_bar on or outside the open unit
// interval).  Out of line on purpose: it contains a double-precision erfinv / pow and would otherwise be inlined into every unrolled
// copy of the hot loop (3000 of the kernel's 4500 instructions, far beyond the instruction cache).
#ifdef __CUDACC__
#define DXI_NOINLINE __noinline__
#else
#define DXI_NOINLINE __attribute__((noinline))
#endif
static __device__ DXI_NOINLINE float lsa_gain_exact_slow(float xbar, float mu, float sigma) {
  const float xi = xi_from_xbar(xbar, mu, sigma);
  return gain_mmse_lsa(xi, __fadd_rn(xi, 1.0f));
}

// s2 = sigma * f32(sqrt(2)); tab: the LSA_TAB_N cubics (shared memory in the kernel).
__device__ __forceinline__ float lsa_gain_from_xbar_tab(float xbar, float mu, float s2, float sigma, const float4* tab) {
  const float arg = fmaf(2.0f, xbar, -1.0f);
  if (!(fabsf(arg) < 1.0f)) return lsa_gain_exact_slow(xbar, mu, sigma);
  const float t = fmaf(s2, erfinv_fast(arg), mu);                                   // xi_hat in dB
  const float pos = fminf(fmaxf(fmaf(t, LSA_TAB_INV_H, -LSA_TAB_LO * LSA_TAB_INV_H), 0.0f), (float)LSA_TAB_N - 0.001f);
  const int i = (int)pos;
  const float f = pos - (float)i;
  const float4 c = tab[i];
  return fast_ex2(fmaf(fmaf(fmaf(c.w, f, c.z), f, c.y), f, c.x));
}

}  // namespace dxi
