// MHANetV3 multi-head attention on the 5th-generation tensor cores (tcgen05 + TMEM), precision DXI_PREC_F16X3.
// softmax((q / sqrt(32)) k^T [+ mask]) v per (utterance, head), head size 32 (tfa MultiHeadAttention einsum formulation,
// SURVEY F5; mask of attention.py:355-385 in mode CAUSAL_PAD), streaming over key tiles: the T x T logits never exist.
//
// One persistent CTA per SM walks (utterance, head, tile of 128 queries).  Per key tile of 128 keys:
//   S = Q K^T   : M 128 x N 128 x K 32, Q (fp16 hi | lo) in tensor memory (written once per query tile by the thread that
//                 owns the row), K tile (hi | lo) in shared memory, three MMAs per product, fp32 S in TMEM (two buffers);
//   softmax     : thread = query row; running max / sum, P = exp(S - max) split into fp16 hi | lo and written back to
//                 TMEM as the A operand of the next product; the O accumulator (128 x 32 fp32 in TMEM) is rescaled in place;
//   O += P V    : M 128 x N 32 x K 128, V^T tile (hi | lo) in shared memory.
// Warps 0-3: softmax (TMEM lane quarter = warp).  Warps 4-7: load K / V tiles (fp32 from the fused QKV activations), split
// and lay them out as K-major UMMA operands (8 x 16-byte core matrices, no swizzle; V goes through a transposing stage).
// Warp 8: MMA issue.  A CTA needs only 192 TMEM columns (it allocates 256) and 83 KB of shared memory, so TWO CTAs share an
// SM: while one waits for its 4 softmax warps the tensor core serves the other (the single-CTA version with two S buffers
// took 4.8 k cycles per key tile against 0.8 k of tensor work).
// TMEM columns (relative to the allocation): S [0,128), overwritten IN PLACE by P (chunk of 32 fp32 columns -> 16 columns of
// fp16 hi pairs + 16 of lo pairs); O [128,160); Q hi [160,176) lo [176,192).
#include <math.h>
#include "net.cuh"
#include "umma.cuh"
#include "gain_math.cuh"

namespace dxi {
using namespace umma;

constexpr int AT = 128, AHD = 32;
constexpr uint32_t AC_S = 0, AC_O = 128, AC_QHI = 160, AC_QLO = 176, A_TMEM_COLS = 256;
constexpr int AK_PART = AT * AHD * 2;                  // 8 KB: one precision part of a K tile / of a V^T tile
constexpr int A_SLOT = 4 * AK_PART;                    // K hi | K lo | V^T hi | V^T lo
constexpr int A_STAGE_LD = AHD + 1;
constexpr int A_SMEM = 1024 + 2 * A_SLOT + AT * A_STAGE_LD * 4 + 2 * AT;
constexpr int A_THREADS = 9 * 32;

struct AttnArgs {
  const float* qkv;          // [B * T][3 * d_model]
  const uint8_t* valid;      // [B * T] (mask mode only)
  float* att;                // [B * T][d_model]
  int B, T, d_model, n_heads;
};

template <int MASK>
__global__ void __launch_bounds__(A_THREADS, 2) attn_umma_kernel(const AttnArgs g) {
  extern __shared__ unsigned char smem_raw[];
  __shared__ __align__(8) uint64_t kv_full[2], kv_empty[2], s_full, q_full, p_ready, pv_done;
  __shared__ uint32_t tmem_slot;
  unsigned char* ring = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  float* vstage = reinterpret_cast<float*>(ring + 2 * A_SLOT);
  uint8_t* sValid = reinterpret_cast<uint8_t*>(vstage + AT * A_STAGE_LD);       // [2][128]
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (warp == 8) tmem_alloc(&tmem_slot, A_TMEM_COLS);
  if (tid == 0) {
    for (int i = 0; i < 2; ++i) { mbar_init(&kv_full[i], 4); mbar_init(&kv_empty[i], 1); }
    mbar_init(&s_full, 1); mbar_init(&q_full, 4); mbar_init(&p_ready, 4); mbar_init(&pv_done, 1);
    fence_mbar_init();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tbase = tmem_slot;      // two CTAs per SM: the allocation does not start at column 0

  const int n_qt = (g.T + AT - 1) / AT;
  const int n_items = g.B * g.n_heads * n_qt;
  const int ld = 3 * g.d_model;
  // key tiles of query tile qt: all of them, or (causal) up to and including qt
  auto n_ktiles = [&](int qt) { return MASK ? qt + 1 : n_qt; };

  if (warp < 4) {
    // ================= softmax / Q / O : thread = query row =================
    const int row = warp * 32 + lane;
    const uint32_t lane_addr = tbase + ((uint32_t)(warp * 32) << 16);
    const float scale = rsqrtf((float)AHD) * 1.44269504f;      // q / sqrt(depth), times log2(e): the softmax below works in base 2
    int kt = 0, it = 0;
    for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++it) {
      const int qt = item % n_qt, bh = item / n_qt, h = bh % g.n_heads, b = bh / g.n_heads;
      const int qi = qt * AT + row;
      const bool qin = qi < g.T;
      const bool qvalid = MASK ? (qin && g.valid[(size_t)b * g.T + qi]) : true;
      {   // Q row -> fp16 hi | lo in TMEM (the previous item's S products are complete: its last s_full was awaited)
        const float* qp = g.qkv + ((size_t)b * g.T + (qin ? qi : 0)) * ld + h * AHD;
        uint32_t hi[16], lo[16];
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          float4 v = qin ? __ldg(reinterpret_cast<const float4*>(qp + 4 * q)) : make_float4(0.f, 0.f, 0.f, 0.f);
          split_h2(v.x * scale, v.y * scale, hi[2 * q], lo[2 * q]);
          split_h2(v.z * scale, v.w * scale, hi[2 * q + 1], lo[2 * q + 1]);
        }
        tmem_st16(lane_addr + AC_QHI, hi);
        tmem_st16(lane_addr + AC_QLO, lo);
        tmem_wait_st(); tc_fence_before(); __syncwarp();
        if (lane == 0) mbar_arrive(&q_full);
      }
      float m_run = -INFINITY, l_run = 0.0f;
      const int nk = n_ktiles(qt);
      for (int j = 0; j < nk; ++j, ++kt) {
        const int k0 = j * AT;
        mbar_wait(&s_full, kt & 1); tc_fence_after();
        const uint32_t s_addr = lane_addr + AC_S;
        const uint8_t* vk = sValid + (kt & 1) * AT;
        auto logit = [&](float s, int c) -> float {          // c: key index inside the tile
          const int kj = k0 + c;
          if (kj >= g.T) return -INFINITY;                    // outside the batch: not a key at all
          if (MASK && !(kj <= qi && qvalid && vk[c])) return -1e10f;      // logits += -1e10 (1 - mask) absorbs s in fp32
          return s;
        };
        // interior tiles of the unmasked mode need no per-element tests (every key exists, nothing is masked)
        const bool plain = !MASK && k0 + AT <= g.T;
        float mx = -INFINITY;
#pragma unroll 1
        for (int c4 = 0; c4 < 4; ++c4) {
          float s[32];
          tmem_ld32(s_addr + 32 * c4, s); tmem_wait_ld();
          if (plain) {
#pragma unroll
            for (int e = 0; e < 32; e += 2) mx = fmaxf(mx, fmaxf(s[e], s[e + 1]));
          } else {
#pragma unroll
            for (int e = 0; e < 32; ++e) mx = fmaxf(mx, logit(s[e], 32 * c4 + e));
          }
        }
        const float m_new = fmaxf(m_run, mx);
        const float alpha = (m_run == -INFINITY) ? 0.0f : fast_ex2(m_run - m_new);
        if (kt > 0) { mbar_wait(&pv_done, (kt - 1) & 1); tc_fence_after(); }      // P and O are free again
        float psum = 0.0f;
#pragma unroll 1
        for (int c4 = 0; c4 < 4; ++c4) {
          float s[32];
          tmem_ld32(s_addr + 32 * c4, s); tmem_wait_ld();
          uint32_t hi[16], lo[16];
          if (plain) {
#pragma unroll
            for (int e = 0; e < 16; ++e) {
              const float2 pp = make_float2(fast_ex2(s[2 * e] - m_new), fast_ex2(s[2 * e + 1] - m_new));
              psum += pp.x + pp.y;
              split_h2x(pp, hi[e], lo[e]);
            }
          } else {
#pragma unroll
            for (int e = 0; e < 16; ++e) {
              const float l0 = logit(s[2 * e], 32 * c4 + 2 * e), l1 = logit(s[2 * e + 1], 32 * c4 + 2 * e + 1);
              const float2 pp = make_float2((l0 == -INFINITY) ? 0.0f : fast_ex2(l0 - m_new), (l1 == -INFINITY) ? 0.0f : fast_ex2(l1 - m_new));
              psum += pp.x + pp.y;
              split_h2x(pp, hi[e], lo[e]);
            }
          }
          tmem_st16(s_addr + 32 * c4, hi);             // in place: this chunk's logits are in registers
          tmem_st16(s_addr + 32 * c4 + 16, lo);
        }
        l_run = l_run * alpha + psum;
        m_run = m_new;
        if (j > 0) {      // rescale the running output
          float o[32];
          tmem_ld32(lane_addr + AC_O, o); tmem_wait_ld();
#pragma unroll
          for (int e = 0; e < 32; ++e) o[e] *= alpha;
          tmem_st32(lane_addr + AC_O, reinterpret_cast<const uint32_t(&)[32]>(o));
        }
        tmem_wait_st(); tc_fence_before(); __syncwarp();
        if (lane == 0) mbar_arrive(&p_ready);
      }
      // ---- O / l -> att
      mbar_wait(&pv_done, (kt - 1) & 1); tc_fence_after();
      float o[32];
      tmem_ld32(lane_addr + AC_O, o); tmem_wait_ld();
      tc_fence_before();
      if (qin) {
        const float inv = 1.0f / l_run;
        float* dst = g.att + ((size_t)b * g.T + qi) * g.d_model + h * AHD;
#pragma unroll
        for (int q = 0; q < 8; ++q)
          *reinterpret_cast<float4*>(dst + 4 * q) = make_float4(o[4 * q] * inv, o[4 * q + 1] * inv, o[4 * q + 2] * inv, o[4 * q + 3] * inv);
      }
    }
  } else if (warp < 8) {
    // ================= K / V tile loaders =================
    const int lt = tid - 128;
    auto ld_sync = [] { asm volatile("bar.sync 1, 128;" ::: "memory"); };
    int kt = 0;
    for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
      const int qt = item % n_qt, bh = item / n_qt, h = bh % g.n_heads, b = bh / g.n_heads;
      const float* kb = g.qkv + (size_t)b * g.T * ld + g.d_model + h * AHD;
      const float* vb = kb + g.d_model;
      const int nk = n_ktiles(qt);
      for (int j = 0; j < nk; ++j, ++kt) {
        const int slot = kt & 1, use = kt >> 1, k0 = j * AT;
        if (use >= 1) mbar_wait(&kv_empty[slot], (use - 1) & 1);
        unsigned char* sK = ring + slot * A_SLOT;
        unsigned char* sV = sK + 2 * AK_PART;
        // K: (key, 8 channels) -> one 16-byte row of a core matrix: off = (key>>3) 512 + u 128 + (key&7) 16
#pragma unroll
        for (int r = 0; r < 4; ++r) {
          const int i = lt + 128 * r, key = i >> 2, u = i & 3;
          float4 a = make_float4(0.f, 0.f, 0.f, 0.f), c = a;
          if (k0 + key < g.T) {
            const float* p = kb + (size_t)(k0 + key) * ld + 8 * u;
            a = __ldg(reinterpret_cast<const float4*>(p)); c = __ldg(reinterpret_cast<const float4*>(p + 4));
          }
          uint4 hi, lo;
          split_h2(a.x, a.y, hi.x, lo.x); split_h2(a.z, a.w, hi.y, lo.y);
          split_h2(c.x, c.y, hi.z, lo.z); split_h2(c.z, c.w, hi.w, lo.w);
          const uint32_t off = (uint32_t)(key >> 3) * 512 + u * 128 + (key & 7) * 16;
          *reinterpret_cast<uint4*>(sK + off) = hi;
          *reinterpret_cast<uint4*>(sK + AK_PART + off) = lo;
        }
        // V: coalesced rows into the fp32 stage, then (channel, 8 keys) -> one 16-byte row: off = (d>>3) 2048 + kg 128 + (d&7) 16
#pragma unroll
        for (int r = 0; r < 8; ++r) {
          const int i = lt + 128 * r, key = i >> 3, d4 = i & 7;
          const float4 a = (k0 + key < g.T) ? __ldg(reinterpret_cast<const float4*>(vb + (size_t)(k0 + key) * ld + 4 * d4)) : make_float4(0.f, 0.f, 0.f, 0.f);
          float* s = vstage + key * A_STAGE_LD + 4 * d4;
          s[0] = a.x; s[1] = a.y; s[2] = a.z; s[3] = a.w;
        }
        if (MASK) sValid[slot * AT + lt] = (k0 + lt < g.T) ? g.valid[(size_t)b * g.T + k0 + lt] : (uint8_t)0;
        ld_sync();
#pragma unroll
        for (int r = 0; r < 4; ++r) {
          const int i = lt + 128 * r, d = i & 31, kg = i >> 5;
          float x[8];
#pragma unroll
          for (int e = 0; e < 8; ++e) x[e] = vstage[(8 * kg + e) * A_STAGE_LD + d];
          uint4 hi, lo;
          split_h2(x[0], x[1], hi.x, lo.x); split_h2(x[2], x[3], hi.y, lo.y);
          split_h2(x[4], x[5], hi.z, lo.z); split_h2(x[6], x[7], hi.w, lo.w);
          const uint32_t off = (uint32_t)(d >> 3) * 2048 + kg * 128 + (d & 7) * 16;
          *reinterpret_cast<uint4*>(sV + off) = hi;
          *reinterpret_cast<uint4*>(sV + AK_PART + off) = lo;
        }
        fence_proxy_async();
        ld_sync();                              // the stage may be overwritten by the next tile
        if (lane == 0) mbar_arrive(&kv_full[slot]);
      }
    }
  } else {
    // ================= MMA issue =================
    constexpr uint32_t id_s = make_idesc_f16(AT, AT), id_o = make_idesc_f16(AT, AHD);
    auto issue_s = [&](int ktile) {      // S = Q K^T (issued behind the previous tile's P V: the tensor pipe runs in order)
      const int slot = ktile & 1;
      mbar_wait(&kv_full[slot], (ktile >> 1) & 1); tc_fence_after();
      const uint32_t k_hi = smem_u32(ring + slot * A_SLOT), k_lo = k_hi + AK_PART;
#pragma unroll
      for (int part = 0; part < 3; ++part) {
        const uint32_t a0 = part == 1 ? AC_QLO : AC_QHI, b0 = part == 2 ? k_lo : k_hi;
#pragma unroll
        for (int ks = 0; ks < 2; ++ks)
          mma_ts_elect(tbase + AC_S, tbase + a0 + 8 * ks, make_smem_desc_noswz(b0 + ks * 256, 128, 512), id_s, (part > 0 || ks > 0) ? 1u : 0u);
      }
      mma_commit_elect(&s_full);
    };
    int kt = 0, it = 0;
    for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++it) {
      const int qt = item % n_qt;
      const int nk = n_ktiles(qt);
      mbar_wait(&q_full, it & 1); tc_fence_after();
      for (int j = 0; j < nk; ++j, ++kt) {
        issue_s(kt);
        mbar_wait(&p_ready, kt & 1); tc_fence_after();
        const int slot = kt & 1;
        const uint32_t v_hi = smem_u32(ring + slot * A_SLOT) + 2 * AK_PART, v_lo = v_hi + AK_PART;
#pragma unroll
        for (int part = 0; part < 3; ++part) {
          const uint32_t b0 = part == 2 ? v_lo : v_hi;
#pragma unroll
          for (int ks = 0; ks < 8; ++ks)      // keys 16 ks ..: chunk ks >> 1, hi pairs at 32 c + 8 (ks & 1), lo pairs 16 columns further
            mma_ts_elect(tbase + AC_O, tbase + AC_S + 32 * (ks >> 1) + 8 * (ks & 1) + (part == 1 ? 16 : 0),
                         make_smem_desc_noswz(b0 + ks * 256, 128, 2048), id_o, (j > 0 || part > 0 || ks > 0) ? 1u : 0u);
        }
        mma_commit_elect(&kv_empty[slot]);
        mma_commit_elect(&pv_done);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 8) tmem_dealloc(tbase, A_TMEM_COLS);
}

int mhanet_umma_attention(const dxi_net& net, const float* qkv, const uint8_t* valid, int B, int T, float* att, cudaStream_t st) {
  const dxi_net_cfg& c = net.cfg;
  if (c.d_model / c.n_heads != AHD || (c.d_model & 3)) { set_error("tcgen05 attention is built for head size 32"); return DXI_E_INVALID; }
  AttnArgs a{qkv, valid, att, B, T, c.d_model, c.n_heads};
  // (per call: the attribute is per device, and a process may drive several)
  DXI_CUDA(cudaFuncSetAttribute(attn_umma_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, A_SMEM));
  DXI_CUDA(cudaFuncSetAttribute(attn_umma_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, A_SMEM));
  int n_sm = 148;
  { int dev = 0; cudaGetDevice(&dev); cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev); }
  const int items = B * c.n_heads * ((T + AT - 1) / AT);
  const int grid = items < 2 * n_sm ? items : 2 * n_sm;      // two CTAs per SM
  ProfScope prof("mha_attn", st, 1);
  if (c.mask_mode == DXI_MASK_CAUSAL_PAD) attn_umma_kernel<1><<<grid, A_THREADS, A_SMEM, st>>>(a);
  else attn_umma_kernel<0><<<grid, A_THREADS, A_SMEM, st>>>(a);
  DXI_LAUNCHED("attn_umma_kernel");
  return DXI_OK;
}

}  // namespace dxi
