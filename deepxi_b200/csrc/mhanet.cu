// MHANetV3 forward (deepxi/network/attention.py:387-442; block :327-353; feed-forward :88-101;
// tfa.layers.MultiHeadAttention einsum formulation) -- fp32 CUDA-core path.
//
//   x  = ReLU(LN(inp W0)) + E[t]
//   5x: qkv = x [Wq|Wk|Wv];  att = softmax((q/sqrt(32)) k^T [+ mask]) v  per head (8 x 32);
//       a = LN(x + att Wp);  x = LN(a + ReLU(a W1 + b1) W2 + b2)
//   x_bar = sigmoid(x Wo + bo)
//
// Two kernels: a tiled SGEMM with fused epilogues (bias / ReLU / sigmoid / residual + LayerNorm /
// LayerNorm + ReLU + positional embedding) and a streaming-softmax attention kernel that never
// materialises the T x T logits.  mask_mode NONE attends over all Tmax frames of the zero-padded batch,
// which is what the shipped model computes (the mask input is ignored by tfa); CAUSAL_PAD applies the mask of
// attention.py:355-385 (causal AND both frames non-zero) the way tfa applies one (logits += -1e10 (1 - mask)).
// This is the correctness-first path of round 1; the GEMMs move to tcgen05 next.
#include <math.h>
#include <stdlib.h>
#include <algorithm>
#include "net.cuh"

namespace dxi {

constexpr int GM = 64, GN = 256, GK = 16;
enum { EPI_BIAS = 0, EPI_RELU = 1, EPI_SIGMOID = 2, EPI_LN_RELU_POS = 3, EPI_RES_LN = 4 };

struct GemmArgs {
  const float* A; int lda;
  const float* W;            // [K][N] row-major
  const float* bias;         // [N] or null
  const float* res;          // [M][N] residual (EPI_RES_LN)
  const float* gamma; const float* beta;   // LayerNorm affine (EPI_LN_*, EPI_RES_LN)
  const float* pos;          // [max_len][N] positional embedding (EPI_LN_RELU_POS), row = m % T
  float* out; int ldo;
  int M, N, K, T, epi;
};

__device__ __forceinline__ float warp_sum_m(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__global__ void __launch_bounds__(256) gemm_f32_kernel(const GemmArgs g) {
  extern __shared__ __align__(16) float sm[];
  float* As = sm;                       // [GK][GM]
  float* Ws = sm + GK * GM;             // [GK][GN]
  float* O = Ws + GK * GN;              // [GM][GN + 1]
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int m0 = blockIdx.x * GM, n0 = blockIdx.y * GN;
  const int ty = tid >> 4, tx = tid & 15;
  float acc[4][16];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int c = 0; c < 16; ++c) acc[i][c] = 0.0f;
  for (int k0 = 0; k0 < g.K; k0 += GK) {
    __syncthreads();
    for (int i = tid; i < GM * GK; i += 256) {
      const int r = i >> 4, kk = i & 15;
      As[kk * GM + r] = (m0 + r < g.M && k0 + kk < g.K) ? g.A[(size_t)(m0 + r) * g.lda + k0 + kk] : 0.0f;
    }
    for (int i = tid; i < GK * GN; i += 256) {
      const int kk = i >> 8, c = i & 255;
      Ws[i] = (k0 + kk < g.K && n0 + c < g.N) ? __ldg(g.W + (size_t)(k0 + kk) * g.N + n0 + c) : 0.0f;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < GK; ++kk) {
      const float4 a4 = *reinterpret_cast<const float4*>(&As[kk * GM + ty * 4]);
      const float a[4] = {a4.x, a4.y, a4.z, a4.w};
      float w[16];
#pragma unroll
      for (int c = 0; c < 16; ++c) w[c] = Ws[kk * GN + tx + 16 * c];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int c = 0; c < 16; ++c) acc[i][c] = fmaf(a[i], w[c], acc[i][c]);
    }
  }
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int c = 0; c < 16; ++c) {
      const int col = tx + 16 * c;
      O[(ty * 4 + i) * (GN + 1) + col] = acc[i][c] + ((g.bias && n0 + col < g.N) ? __ldg(g.bias + n0 + col) : 0.0f);
    }
  __syncthreads();
  const int ncols = min(GN, g.N - n0);
  for (int r = warp; r < GM; r += 8) {
    const int m = m0 + r;
    if (m >= g.M) continue;
    float* row = O + r * (GN + 1);
    float* dst = g.out + (size_t)m * g.ldo + n0;
    if (g.epi == EPI_BIAS) {
      for (int c = lane; c < ncols; c += 32) dst[c] = row[c];
    } else if (g.epi == EPI_RELU) {
      for (int c = lane; c < ncols; c += 32) dst[c] = fmaxf(row[c], 0.0f);
    } else if (g.epi == EPI_SIGMOID) {
      for (int c = lane; c < ncols; c += 32) dst[c] = 1.0f / (1.0f + expf(-row[c]));
    } else {
      // LayerNorm over the full row (N == GN == 256), Keras non-fused op order, eps 1e-6
      if (g.epi == EPI_RES_LN) {
        const float* rs = g.res + (size_t)m * g.N;
        for (int c = lane; c < GN; c += 32) row[c] += rs[c];
      }
      float s = 0.0f;
      for (int c = lane; c < GN; c += 32) s += row[c];
      const float mean = warp_sum_m(s) * (1.0f / GN);
      float q = 0.0f;
      for (int c = lane; c < GN; c += 32) { const float d = row[c] - mean; q = fmaf(d, d, q); }
      const float rstd = rsqrtf(warp_sum_m(q) * (1.0f / GN) + 1e-6f);
      const float* pe = g.epi == EPI_LN_RELU_POS ? g.pos + (size_t)(m % g.T) * GN : nullptr;
      for (int c = lane; c < GN; c += 32) {
        const float inv = rstd * __ldg(g.gamma + c);
        float y = fmaf(row[c], inv, __ldg(g.beta + c) - mean * inv);
        if (pe) y = fmaxf(y, 0.0f) + __ldg(pe + c);
        dst[c] = y;
      }
    }
  }
}

static int launch_gemm(const GemmArgs& g, cudaStream_t st, const char* key) {
  const size_t smem = sizeof(float) * (GK * GM + GK * GN + GM * (GN + 1));
  DXI_CUDA(cudaFuncSetAttribute(gemm_f32_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));      // per device, so per call
  dim3 grid((g.M + GM - 1) / GM, (g.N + GN - 1) / GN);
  ProfScope prof(key, st, 1);
  gemm_f32_kernel<<<grid, 256, smem, st>>>(g);
  DXI_LAUNCHED("gemm_f32_kernel");
  return DXI_OK;
}

// valid[b*T + t] = any(inp[b,t,:] != 0)   (Masking(mask_value=0.0).compute_mask, attention.py:379)
__global__ void valid_mask_kernel(const float* __restrict__ inp, int rows, int nfeat, uint8_t* __restrict__ valid) {
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (row >= rows) return;
  int any = 0;
  for (int c = lane; c < nfeat; c += 32) any |= inp[(size_t)row * nfeat + c] != 0.0f;
  any = __any_sync(0xffffffffu, any);
  if (lane == 0) valid[row] = any ? 1 : 0;
}

// Streaming-softmax attention, head size 32.  One CTA: 64 queries of one (utterance, head); keys in tiles of 64.
constexpr int AQ = 64, AK = 64, HD = 32;

template <int MASK>
__global__ void __launch_bounds__(256) attn_f32_kernel(const float* __restrict__ qkv, const uint8_t* __restrict__ valid,
                                                       int T, int d_model, float* __restrict__ att) {
  __shared__ float Qs[AQ][HD + 1], Ks[AK][HD + 1], Vs[AK][HD + 1], Ps[AQ][AK + 1];
  const int tid = threadIdx.x;
  const int q0 = blockIdx.x * AQ, h = blockIdx.y, b = blockIdx.z;
  const int ld = 3 * d_model;
  const float* base = qkv + (size_t)b * T * ld;
  const float scale = rsqrtf((float)HD);                      // query / sqrt(depth)
  for (int i = tid; i < AQ * HD; i += 256) {
    const int r = i >> 5, c = i & 31;
    Qs[r][c] = (q0 + r < T) ? base[(size_t)(q0 + r) * ld + h * HD + c] * scale : 0.0f;
  }
  const int ty = tid >> 4, tx = tid & 15;          // S: rows 4ty..4ty+3, cols tx + 16 j;  O: same rows, cols tx, tx+16
  float m_run[4], l_run[4], o[4][2];
#pragma unroll
  for (int i = 0; i < 4; ++i) { m_run[i] = -INFINITY; l_run[i] = 0.0f; o[i][0] = o[i][1] = 0.0f; }
  // causal+pad: queries attend keys j <= i only; key tiles beyond the last query of this CTA are skipped.  A
  // query that is itself padding has its whole row masked (softmax of equal values, tfa semantics): those rows
  // are don't-care and are computed with the causal range only.
  const int k_end = MASK ? min(T, q0 + AQ) : T;
  for (int k0 = 0; k0 < k_end; k0 += AK) {
    __syncthreads();
    for (int i = tid; i < AK * HD; i += 256) {
      const int r = i >> 5, c = i & 31;
      const bool in = k0 + r < T;
      Ks[r][c] = in ? base[(size_t)(k0 + r) * ld + d_model + h * HD + c] : 0.0f;
      Vs[r][c] = in ? base[(size_t)(k0 + r) * ld + 2 * d_model + h * HD + c] : 0.0f;
    }
    __syncthreads();
    float s[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) s[i][j] = 0.0f;
#pragma unroll 8
    for (int c = 0; c < HD; ++c) {
      float qv[4], kv[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) qv[i] = Qs[ty * 4 + i][c];
#pragma unroll
      for (int j = 0; j < 4; ++j) kv[j] = Ks[tx + 16 * j][c];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) s[i][j] = fmaf(qv[i], kv[j], s[i][j]);
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int qi = q0 + ty * 4 + i;
      float mx = -INFINITY;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int kj = k0 + tx + 16 * j;
        if (kj >= T) s[i][j] = -INFINITY;                                 // outside the batch: not a key at all
        else if (MASK) {
          const bool keep = kj <= qi && qi < T && valid[(size_t)b * T + kj] && valid[(size_t)b * T + qi];
          if (!keep) s[i][j] = -1e10f;                                    // logits += -10e9 * (1 - mask) (absorbs s in fp32)
        }
        mx = fmaxf(mx, s[i][j]);
      }
#pragma unroll
      for (int off = 8; off > 0; off >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, off));
      const float m_new = fmaxf(m_run[i], mx);
      const float alpha = (m_run[i] == -INFINITY) ? 0.0f : expf(m_run[i] - m_new);
      float ps = 0.0f;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float pj = (s[i][j] == -INFINITY) ? 0.0f : expf(s[i][j] - m_new);
        Ps[ty * 4 + i][tx + 16 * j] = pj;
        ps += pj;
      }
#pragma unroll
      for (int off = 8; off > 0; off >>= 1) ps += __shfl_xor_sync(0xffffffffu, ps, off);
      l_run[i] = l_run[i] * alpha + ps;
      m_run[i] = m_new;
      o[i][0] *= alpha; o[i][1] *= alpha;
    }
    __syncthreads();
#pragma unroll 8
    for (int j = 0; j < AK; ++j) {
      const float v0 = Vs[j][tx], v1 = Vs[j][tx + 16];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float pj = Ps[ty * 4 + i][j];
        o[i][0] = fmaf(pj, v0, o[i][0]);
        o[i][1] = fmaf(pj, v1, o[i][1]);
      }
    }
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int qi = q0 + ty * 4 + i;
    if (qi >= T) continue;
    const float inv = 1.0f / l_run[i];
    float* dst = att + ((size_t)b * T + qi) * d_model + h * HD;
    dst[tx] = o[i][0] * inv;
    dst[tx + 16] = o[i][1] * inv;
  }
}

int64_t mhanet_workspace_bytes(const dxi_net& net, int B, int T) {
  const int64_t rows = (int64_t)B * T, d = net.cfg.d_model;
  // the feed-forward activations and the packed K / V tiles of the tensor-core attention are never live together: one region
  const int64_t f_bytes = std::max<int64_t>(sizeof(float) * rows * 4 * d, (int64_t)mhanet_umma_attention_workspace(net, B, T));
  return 512 + sizeof(float) * rows * (d /*x*/ + 3 * d /*qkv*/ + d /*att*/ + d /*a*/) + f_bytes + rows /*valid*/;
}

int mhanet_forward(const dxi_net& net, const float* mag, int B, int T, float* xbar, void* ws, size_t ws_bytes, cudaStream_t st) {
  const dxi_net_cfg& c = net.cfg;
  if (c.d_model != GN || c.d_model / c.n_heads != HD) { set_error("MHANetV3 path is built for d_model=256, head size 32"); return DXI_E_INVALID; }
  if (c.precision != DXI_PREC_F32 && c.precision != DXI_PREC_F16X3) { set_error("MHANetV3: precisions f32 and f16x3 are built"); return DXI_E_INVALID; }
  const bool tc = c.precision == DXI_PREC_F16X3;      // every GEMM (mha_umma.cu) and the attention (attn_umma.cu) on tcgen05
  if (T > c.max_len) { set_error("MHANetV3: %d frames exceed the %d rows of the positional embedding (attention.py:432)", T, c.max_len); return DXI_E_INVALID; }
  if ((int64_t)ws_bytes < mhanet_workspace_bytes(net, B, T)) { set_error("workspace too small"); return DXI_E_NOMEM; }
  const int rows = B * T, d = c.d_model;
  float* x = reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(ws) + 255) & ~(uintptr_t)255);
  float* qkv = x + (size_t)rows * d;
  float* att = qkv + (size_t)rows * 3 * d;
  float* a = att + (size_t)rows * d;
  float* f = a + (size_t)rows * d;
  const size_t f_bytes = std::max<size_t>(sizeof(float) * (size_t)rows * 4 * d, mhanet_umma_attention_workspace(net, B, T));
  uint8_t* valid = reinterpret_cast<uint8_t*>(f) + f_bytes;
  if (c.mask_mode == DXI_MASK_CAUSAL_PAD) {
    valid_mask_kernel<<<(rows + 7) / 8, 256, 0, st>>>(mag, rows, c.n_feat, valid);
    DXI_LAUNCHED("valid_mask_kernel");
  }
  auto Wt = [&](int li, const char* v) { return net.dev_tensor(li, v); };
  GemmArgs g{};
  // x = ReLU(LN(inp W0)) + E[t]
  if (tc) {
    if (int rc = mhanet_umma_linear(net, 4 * c.n_blocks, 2 /* LN, here: + ReLU + positional embedding */, mag, c.n_feat, nullptr, nullptr, Wt(1, "gamma"),
                                    Wt(1, "beta"), Wt(2, "embeddings"), T, x, d, rows, d, c.n_feat, st)) return rc;
  } else {
    g = GemmArgs{mag, c.n_feat, Wt(0, "kernel"), nullptr, nullptr, Wt(1, "gamma"), Wt(1, "beta"), Wt(2, "embeddings"), x, d, rows, d, c.n_feat, T, EPI_LN_RELU_POS};
    if (int rc = launch_gemm(g, st, "mha_gemm")) return rc;
  }
  int li = 3;
  for (int blk = 0; blk < c.n_blocks; ++blk) {
    char nm[64];
    snprintf(nm, sizeof(nm), "packed-%d/qkv", li);
    auto it = net.d_offset.find(nm);
    if (it == net.d_offset.end()) { set_error("packed QKV weights missing"); return DXI_E_STATE; }
    const bool attn_tc = tc && !(getenv("DXI_MHA_ATTN_F32") && atoi(getenv("DXI_MHA_ATTN_F32")));
    // the QKV projection writes K / V straight into the attention kernel's operand images (DXI_MHA_UNFUSED_PACK=1: fp32 K / V + a packing pass, for A/B)
    const bool fused_pack = attn_tc && !(getenv("DXI_MHA_UNFUSED_PACK") && atoi(getenv("DXI_MHA_UNFUSED_PACK")));
    if (fused_pack) {
      if (int rc = mhanet_umma_qkv(net, blk, x, B, T, qkv, f, st)) return rc;
    } else if (tc) {
      if (int rc = mhanet_umma_linear(net, 4 * blk + 0, 0 /* plain */, x, d, nullptr, nullptr, nullptr, nullptr, nullptr, T, qkv, 3 * d, rows, 3 * d, d, st)) return rc;
    } else {
      g = GemmArgs{x, d, net.d_arena + it->second, nullptr, nullptr, nullptr, nullptr, nullptr, qkv, 3 * d, rows, 3 * d, d, T, EPI_BIAS};
      if (int rc = launch_gemm(g, st, "mha_gemm")) return rc;
    }
    if (attn_tc) {
      if (int rc = mhanet_umma_attention(net, qkv, valid, B, T, att, f, fused_pack, st)) return rc;
    } else {
      dim3 grid((T + AQ - 1) / AQ, c.n_heads, B);
      ProfScope prof("mha_attn", st, 1);
      if (c.mask_mode == DXI_MASK_CAUSAL_PAD) attn_f32_kernel<1><<<grid, 256, 0, st>>>(qkv, valid, T, d, att);
      else attn_f32_kernel<0><<<grid, 256, 0, st>>>(qkv, valid, T, d, att);
      DXI_LAUNCHED("attn_f32_kernel");
    }
    // a = LN(x + att Wp)      (projection_kernel [8,32,256] is [256][256] row-major as stored)
    if (tc) {
      if (int rc = mhanet_umma_linear(net, 4 * blk + 1, 2 /* residual + LN */, att, d, nullptr, x, Wt(li + 1, "gamma"), Wt(li + 1, "beta"), nullptr, T, a, d, rows, d, d, st)) return rc;
      if (int rc = mhanet_umma_linear(net, 4 * blk + 2, 1 /* bias + ReLU */, a, d, Wt(li + 2, "bias"), nullptr, nullptr, nullptr, nullptr, T, f, 4 * d, rows, 4 * d, d, st)) return rc;      // f = ReLU(a W1 + b1)
      if (int rc = mhanet_umma_linear(net, 4 * blk + 3, 2, f, 4 * d, Wt(li + 3, "bias"), a, Wt(li + 4, "gamma"), Wt(li + 4, "beta"), nullptr, T, x, d, rows, d, 4 * d, st)) return rc;      // x = LN(a + f W2 + b2)
    } else {
    g = GemmArgs{att, d, Wt(li, "projection_kernel"), nullptr, x, Wt(li + 1, "gamma"), Wt(li + 1, "beta"), nullptr, a, d, rows, d, d, T, EPI_RES_LN};
    if (int rc = launch_gemm(g, st, "mha_gemm")) return rc;
    // f = ReLU(a W1 + b1)
    g = GemmArgs{a, d, Wt(li + 2, "kernel"), Wt(li + 2, "bias"), nullptr, nullptr, nullptr, nullptr, f, 4 * d, rows, 4 * d, d, T, EPI_RELU};
    if (int rc = launch_gemm(g, st, "mha_gemm")) return rc;
    // x = LN(a + f W2 + b2)
    g = GemmArgs{f, 4 * d, Wt(li + 3, "kernel"), Wt(li + 3, "bias"), a, Wt(li + 4, "gamma"), Wt(li + 4, "beta"), nullptr, x, d, rows, d, 4 * d, T, EPI_RES_LN};
    if (int rc = launch_gemm(g, st, "mha_gemm")) return rc;
    }
    li += 5;
  }
  if (tc) return mhanet_umma_linear(net, 4 * c.n_blocks + 1, 3 /* sigmoid */, x, d, Wt(li, "bias"), nullptr, nullptr, nullptr, nullptr, T, xbar, c.n_outp,
                                    rows, c.n_outp, d, st);
  g = GemmArgs{x, d, Wt(li, "kernel"), Wt(li, "bias"), nullptr, nullptr, nullptr, nullptr, xbar, c.n_outp, rows, c.n_outp, d, T, EPI_SIGMOID};
  return launch_gemm(g, st, "mha_gemm");
}

}  // namespace dxi
