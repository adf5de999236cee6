// Register/shared-memory FFT building blocks for the 512-point real STFT / iSTFT kernels.
//
// A 512-point real transform is computed as one 256-point complex transform of z[m] = x[2m] + j x[2m+1]
// plus a split step.  The 256-point transform is 16 x 16: every thread owns 16 complex points in
// registers (radix-4 x radix-4 butterflies), 16 threads own one frame, one exchange through shared
// memory sits between the two passes.
#pragma once
#include "common.cuh"

namespace dxi {

#ifdef __CUDA_ARCH__      // one FADD2 per complex addition on sm_100
__device__ __forceinline__ float2 cadd(float2 a, float2 b) { return __fadd2_rn(a, b); }
__device__ __forceinline__ float2 csub(float2 a, float2 b) { return __fadd2_rn(a, make_float2(-b.x, -b.y)); }
#else
__device__ __forceinline__ float2 cadd(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ float2 csub(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }
#endif
__device__ __forceinline__ float2 cmul(float2 a, float2 b) {
  return make_float2(fmaf(a.x, b.x, -a.y * b.y), fmaf(a.x, b.y, a.y * b.x));
}

// 4-point DFT in place; SIGN = -1 forward (e^{-j..}), +1 inverse (unscaled).
template <int SIGN>
__device__ __forceinline__ void fft4(float2& a0, float2& a1, float2& a2, float2& a3) {
  float2 t0 = cadd(a0, a2), t1 = csub(a0, a2), t2 = cadd(a1, a3), d = csub(a1, a3);
  float2 t3 = SIGN < 0 ? make_float2(d.y, -d.x) : make_float2(-d.y, d.x);   // d * (SIGN j)
  a0 = cadd(t0, t2); a1 = cadd(t1, t3); a2 = csub(t0, t2); a3 = csub(t1, t3);
}

// 16-point DFT in place.  Input v[n]; output X[k] is left at v[fft16_pos(k)].
__host__ __device__ constexpr int fft16_pos(int k) { return 4 * (k & 3) + (k >> 2); }

template <int SIGN>
__device__ __forceinline__ void fft16(float2 (&v)[16]) {
#pragma unroll
  for (int n2 = 0; n2 < 4; ++n2) fft4<SIGN>(v[n2], v[4 + n2], v[8 + n2], v[12 + n2]);
  // v[4*k1 + n2] *= W16^(n2*k1)
  constexpr float C1 = 0.92387953251128674f, S1 = 0.38268343236508977f, R = 0.70710678118654752f;
  const float sg = (float)SIGN;
  v[5]  = cmul(v[5],  make_float2(C1, sg * S1));      // W^1
  v[6]  = cmul(v[6],  make_float2(R, sg * R));        // W^2
  v[7]  = cmul(v[7],  make_float2(S1, sg * C1));      // W^3
  v[9]  = cmul(v[9],  make_float2(R, sg * R));        // W^2
  v[10] = SIGN < 0 ? make_float2(v[10].y, -v[10].x) : make_float2(-v[10].y, v[10].x);   // W^4 = SIGN j
  v[11] = cmul(v[11], make_float2(-R, sg * R));       // W^6
  v[13] = cmul(v[13], make_float2(S1, sg * C1));      // W^3
  v[14] = cmul(v[14], make_float2(-R, sg * R));       // W^6
  v[15] = cmul(v[15], make_float2(-C1, -sg * S1));    // W^9
#pragma unroll
  for (int k1 = 0; k1 < 4; ++k1) fft4<SIGN>(v[4 * k1], v[4 * k1 + 1], v[4 * k1 + 2], v[4 * k1 + 3]);
}

constexpr int FFT_EXCH_STRIDE = 17;             // float2 row stride of the 16x16 exchange (conflict-free)
constexpr int FFT_FRAME_SLOTS = 16 * FFT_EXCH_STRIDE;   // 272 float2 per frame

// 256-point complex DFT of one frame by 16 cooperating threads (lane16 = 0..15, all in one warp),
// in two passes with a warp-level barrier between and after them (issued by the caller):
//   pass1: in  v[n1] = z[16*n1 + lane16]; writes the twiddled 16x16 exchange to buf
//   pass2: reads the exchange; returns with Z[lane16 + 16*k2] at v[fft16_pos(k2)].
// buf: this frame's FFT_FRAME_SLOTS float2 of shared memory; tw256: e^{-2 pi j m/256}, m = 0..255.
template <int SIGN>
__device__ __forceinline__ void fft256_pass1(float2 (&v)[16], float2* buf, const float2* tw256, int lane16) {
  fft16<SIGN>(v);
#pragma unroll
  for (int k1 = 0; k1 < 16; ++k1) {
    float2 w = tw256[lane16 * k1];
    if (SIGN > 0) w.y = -w.y;
    buf[k1 * FFT_EXCH_STRIDE + lane16] = cmul(v[fft16_pos(k1)], w);
  }
}

template <int SIGN>
__device__ __forceinline__ void fft256_pass2(float2 (&v)[16], const float2* buf, int lane16) {
#pragma unroll
  for (int n2 = 0; n2 < 16; ++n2) v[n2] = buf[lane16 * FFT_EXCH_STRIDE + n2];
  fft16<SIGN>(v);
}

// Split step of the real FFT: X[k] = E[k] + W512^k O[k] with E, O the spectra of the even / odd
// samples recovered from Z = FFT256(x[2m] + j x[2m+1]):  zk = Z[k mod 256], zn = Z[(256-k) mod 256],
// w = e^{-2 pi j k/512}.
__device__ __forceinline__ float2 rfft_split(float2 zk, float2 zn, float2 w) {
  float ex = 0.5f * (zk.x + zn.x), ey = 0.5f * (zk.y - zn.y);
  float ox = 0.5f * (zk.y + zn.y), oy = -0.5f * (zk.x - zn.x);
  return make_float2(ex + fmaf(w.x, ox, -w.y * oy), ey + fmaf(w.x, oy, w.y * ox));
}

// Merge step of the inverse real FFT: Z[k] = E[k] + j O[k], E = (X[k] + conj X[256-k])/2,
// O = conj(W512^k) (X[k] - conj X[256-k])/2;  x[2m] + j x[2m+1] = IFFT256(Z)[m].
__device__ __forceinline__ float2 irfft_merge(float2 xk, float2 xn, float2 w) {
  float ex = 0.5f * (xk.x + xn.x), ey = 0.5f * (xk.y - xn.y);
  float dx = 0.5f * (xk.x - xn.x), dy = 0.5f * (xk.y + xn.y);
  float ox = fmaf(w.x, dx, w.y * dy), oy = fmaf(w.x, dy, -w.y * dx);
  return make_float2(ex - oy, ey + ox);
}

// atan2 with |error| <= 2e-7 rad: odd minimax polynomial of degree 17 on [0,1] plus octant fix-up.
__device__ __forceinline__ float atan2_poly(float y, float x) {
  float ax = fabsf(x), ay = fabsf(y);
  float mx = fmaxf(ax, ay), mn = fminf(ax, ay);
  float t = mx > 0.0f ? __fdividef(mn, mx) : 0.0f;
  float s = t * t;
  float p = 0.0028340641874819994f;
  p = fmaf(p, s, -0.016005029901862144f);
  p = fmaf(p, s, 0.042587608098983765f);
  p = fmaf(p, s, -0.07495445758104324f);
  p = fmaf(p, s, 0.10636754333972931f);
  p = fmaf(p, s, -0.14202570915222168f);
  p = fmaf(p, s, 0.19992484152317047f);
  p = fmaf(p, s, -0.3333306610584259f);
  p = fmaf(p, s, 1.0f);
  p *= t;
  if (ay > ax) p = 1.57079637f - p;
  if (x < 0.0f) p = 3.14159274f - p;
  return copysignf(p, y);
}

// ---- bin-pair formulation of the analysis epilogue --------------------------------------------------------------------------
// Packed fp32x2 arithmetic (FFMA2 / FMUL2 on sm_100: one issue slot for two lanes); the host twin of the tests runs two scalar
// operations instead.
__device__ __forceinline__ float2 fma2(float2 a, float2 b, float2 c) {
#ifdef __CUDA_ARCH__
  return __ffma2_rn(a, b, c);
#else
  return make_float2(fmaf(a.x, b.x, c.x), fmaf(a.y, b.y, c.y));
#endif
}
__device__ __forceinline__ float2 mul2(float2 a, float2 b) {
#ifdef __CUDA_ARCH__
  return __fmul2_rn(a, b);
#else
  return make_float2(a.x * b.x, a.y * b.y);
#endif
}
__device__ __forceinline__ float sqrt_approx(float x) {      // MUFU.SQRT: <= 2 ulp, 0 -> 0
#ifdef __CUDA_ARCH__
  float r;
  asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
#else
  return sqrtf(x);
#endif
}
__device__ __forceinline__ float rcp_approx(float x) {       // MUFU.RCP
#ifdef __CUDA_ARCH__
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
#else
  return 1.0f / x;
#endif
}

// Split step for the bin pair (k, 256 - k), 0 <= k <= 128, from Z' = FFT256 of the HALF-scaled samples (the analysis window
// carries the factor 1/2, an exact scaling): a = X[k], b = X[256 - k].  With E = zk + conj(zn), O = -j (zk - conj(zn)) and
// T = W512^k O:  X[k] = E + T,  X[256 - k] = conj(E - T).  Bit-identical to rfft_split of the unscaled transform.
__device__ __forceinline__ void rfft_split_pair(float2 zk, float2 zn, float2 w, float2& a, float2& b) {
  const float ex = zk.x + zn.x, ey = zk.y - zn.y, ox = zk.y + zn.y, oy = zn.x - zk.x;
  const float tx = fmaf(w.x, ox, -w.y * oy), ty = fmaf(w.x, oy, w.y * ox);
  a = make_float2(ex + tx, ey + ty);
  b = make_float2(ex - tx, ty - ey);
}

// |a|, |b| and atan2 of both (the polynomial of atan2_poly, evaluated for the two bins at once).
__device__ __forceinline__ void polar_pair(float2 a, float2 b, float2& mag, float2& pha) {
  const float2 p2 = fma2(make_float2(a.x, b.x), make_float2(a.x, b.x), mul2(make_float2(a.y, b.y), make_float2(a.y, b.y)));
  mag = make_float2(sqrt_approx(p2.x), sqrt_approx(p2.y));
  const float aax = fabsf(a.x), aay = fabsf(a.y), abx = fabsf(b.x), aby = fabsf(b.y);
  const float2 mn = make_float2(fminf(aax, aay), fminf(abx, aby));
  const float2 r = make_float2(rcp_approx(fmaxf(fmaxf(aax, aay), 1e-30f)), rcp_approx(fmaxf(fmaxf(abx, aby), 1e-30f)));
  const float2 t = mul2(mn, r);
  const float2 s = mul2(t, t);
  auto c = [](float v) { return make_float2(v, v); };
  float2 p = c(0.0028340641874819994f);
  p = fma2(p, s, c(-0.016005029901862144f));
  p = fma2(p, s, c(0.042587608098983765f));
  p = fma2(p, s, c(-0.07495445758104324f));
  p = fma2(p, s, c(0.10636754333972931f));
  p = fma2(p, s, c(-0.14202570915222168f));
  p = fma2(p, s, c(0.19992484152317047f));
  p = fma2(p, s, c(-0.3333306610584259f));
  p = fma2(p, s, c(1.0f));
  p = mul2(p, t);
  if (aay > aax) p.x = 1.57079637f - p.x;
  if (aby > abx) p.y = 1.57079637f - p.y;
  if (a.x < 0.0f) p.x = 3.14159274f - p.x;
  if (b.x < 0.0f) p.y = 3.14159274f - p.y;
  pha = make_float2(copysignf(p.x, a.y), copysignf(p.y, b.y));
}

}  // namespace dxi
