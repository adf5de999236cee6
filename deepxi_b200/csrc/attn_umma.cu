// MHANetV3 multi-head attention on the 5th-generation tensor cores (tcgen05 + TMEM), precision DXI_PREC_F16X3.
// softmax((q / sqrt(32)) k^T [+ mask]) v per (utterance, head), head size 32 (tfa MultiHeadAttention einsum formulation,
// SURVEY F5; mask of attention.py:355-385 in mode CAUSAL_PAD), streaming over key tiles: the T x T logits never exist.
//
// Two launches per layer:
//   attn_pack_kv_kernel   K and V of every (utterance, head, tile of 128 keys) are split ONCE into fp16 hi | lo and written in the
//                         exact shared-memory image of the UMMA operands (K-major, 8 x 16-byte core matrices, no swizzle; V transposed):
//                         32 KB per tile, [K hi | K lo | V^T hi | V^T lo].  (The first version converted them inside the attention
//                         kernel, once per QUERY tile - 15 times at T = 1875 - and that conversion, not the softmax, bounded it.)
//   attn_umma_kernel      persistent CTAs walk (utterance, head, tile of 128 queries).  Per key tile of 128 keys:
//     S = Q K^T   : M 128 x N 128 x K 32, Q (fp16 hi | lo) in tensor memory (written once per query tile by the thread that owns the
//                   row), K tile in shared memory (one 32 KB bulk copy per tile into a ring of three), three MMAs per product, fp32 S
//                   in TMEM;
//     softmax     : thread = (query row, key half); running max / sum, P = exp(S - max) split into fp16 hi | lo and written back to
//                   TMEM in place as the A operand of the next product; the O accumulator (fp32 in TMEM) is rescaled in place, only
//                   when some row of the warp has a new maximum;
//     O += P V    : M 128 x N 32 x K 128, V^T tile in shared memory.
// Warps 0 .. 4 NH - 1: softmax.  TMEM lane quarter = warp & 3; with NH = 2, warps 0-3 take keys 0..63 of every key tile and warps
// 4-7 keys 64..127, each half with its OWN running max / sum and its own O accumulator (two independent streams over disjoint key
// sets, as in split-KV decoding): nothing is exchanged per key tile, and the two partial results of a row are merged once per
// query tile (m = max(m_a, m_b), O = O_a 2^(m_a - m) + O_b 2^(m_b - m), likewise the sums).  Then one warp that issues the bulk
// copies (and stages the tile's validity bytes in mask mode) and one warp that issues the MMAs.  A CTA needs 224 TMEM columns (it
// allocates 256) and 100 KB of shared memory, so TWO CTAs share an SM: while one waits for its softmax warps the tensor core
// serves the other.
// TMEM columns (relative to the allocation): S [0,128), overwritten IN PLACE by P (16 fp32 columns -> 8 columns of fp16 hi
// pairs + 8 of lo pairs); O_a [128,160), O_b [160,192); Q hi [192,208) lo [208,224).
#include <math.h>
#include "net.cuh"
#include "umma.cuh"
#include "gain_math.cuh"

namespace dxi {
using namespace umma;

constexpr int AT = 128, AHD = 32;
constexpr uint32_t AC_S = 0, AC_O = 128, AC_QHI = 192, AC_QLO = 208, A_TMEM_COLS = 256;
constexpr int AK_PART = AT * AHD * 2;                  // 8 KB: one precision part of a K tile / of a V^T tile
constexpr int A_SLOT = 4 * AK_PART;                    // K hi | K lo | V^T hi | V^T lo
constexpr int A_SLOTS = 3;                             // ring of key tiles in shared memory
constexpr int A_STAGE_LD = AHD + 1;
constexpr int A_SMEM = 1024 + A_SLOTS * A_SLOT + A_SLOTS * AT + 2 * AT * 8;
constexpr int A_NH = 2;                                // key halves per tile = softmax warps / 4

struct AttnArgs {
  const float* qkv;          // [B * T][3 * d_model]
  const unsigned char* kv;   // [B][n_heads][key tiles][A_SLOT]: the packed operands (attn_pack_kv_kernel)
  const uint8_t* valid;      // [B * T] (mask mode only)
  float* att;                // [B * T][d_model]
  int B, T, d_model, n_heads;
};

// ---- K / V of one (utterance, head, key tile) -> the 32 KB operand image.  Keys at or beyond T: zeros.
__global__ void __launch_bounds__(128) attn_pack_kv_kernel(const float* __restrict__ qkv, unsigned char* __restrict__ kv, int T, int d_model, int n_heads) {
  __shared__ float vstage[AT * A_STAGE_LD];
  const int n_kt = (T + AT - 1) / AT;
  const int j = blockIdx.x % n_kt, bh = blockIdx.x / n_kt, h = bh % n_heads, b = bh / n_heads;
  const int ld = 3 * d_model, k0 = j * AT, lt = threadIdx.x;
  const float* kb = qkv + (size_t)b * T * ld + d_model + h * AHD;
  const float* vb = kb + d_model;
  unsigned char* sK = kv + (size_t)blockIdx.x * A_SLOT;
  unsigned char* sV = sK + 2 * AK_PART;
  // V: coalesced rows into the fp32 stage first (their latency runs under the K conversion)
  float4 va[8];
#pragma unroll
  for (int r = 0; r < 8; ++r) {
    const int i = lt + 128 * r, key = i >> 3, d4 = i & 7;
    va[r] = (k0 + key < T) ? __ldg(reinterpret_cast<const float4*>(vb + (size_t)(k0 + key) * ld + 4 * d4)) : make_float4(0.f, 0.f, 0.f, 0.f);
  }
  // K: (key, 8 channels) -> one 16-byte row of a core matrix: off = (key>>3) 512 + u 128 + (key&7) 16
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    const int i = lt + 128 * r, key = i >> 2, u = i & 3;
    float4 a = make_float4(0.f, 0.f, 0.f, 0.f), c = a;
    if (k0 + key < T) {
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
#pragma unroll
  for (int r = 0; r < 8; ++r) {
    const int i = lt + 128 * r, key = i >> 3, d4 = i & 7;
    float* sp = vstage + key * A_STAGE_LD + 4 * d4;
    sp[0] = va[r].x; sp[1] = va[r].y; sp[2] = va[r].z; sp[3] = va[r].w;
  }
  __syncthreads();
  // V^T: (channel, 8 keys) -> one 16-byte row: off = (d>>3) 2048 + kg 128 + (d&7) 16
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
}

// PS: P = exp(S - max) as fp16 hi | lo (three products P V); !PS: P rounded to fp16 once (two products: P V_hi + P V_lo), as every
// fp16 / bf16 flash attention does - the weights of a row then carry a relative error of 2^-12 each, random in sign.
template <int MASK, int NH, bool PS>
__global__ void __launch_bounds__((4 * NH + 2) * 32, 2) attn_umma_kernel(const AttnArgs g) {
  constexpr int SW = 4 * NH, KH = AT / NH;      // softmax warps; keys of a tile per softmax thread
  extern __shared__ unsigned char smem_raw[];
  __shared__ __align__(8) uint64_t kv_full[A_SLOTS], kv_empty[A_SLOTS], s_full[NH], q_full, p_ready[NH], pv_done[NH];
  __shared__ uint32_t tmem_slot;
  unsigned char* ring = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* sValid = ring + A_SLOTS * A_SLOT;                                    // [A_SLOTS][128]
  float2* sML = reinterpret_cast<float2*>(sValid + A_SLOTS * AT);               // [2][128]: (max, sum) of the second key half of a row
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (warp == SW + 1) tmem_alloc(&tmem_slot, A_TMEM_COLS);
  if (tid == 0) {
    for (int i = 0; i < A_SLOTS; ++i) { mbar_init(&kv_full[i], 1); mbar_init(&kv_empty[i], 1); }
    for (int i = 0; i < NH; ++i) { mbar_init(&s_full[i], 1); mbar_init(&p_ready[i], 4); mbar_init(&pv_done[i], 1); }
    mbar_init(&q_full, 4);
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

  if (warp < SW) {
    // ================= softmax / Q / O : thread = (query row, key half) =================
    const int half = warp >> 2, wq = warp & 3;
    const int row = wq * 32 + lane;
    const uint32_t lane_addr = tbase + ((uint32_t)(wq * 32) << 16);
    const uint32_t s_addr = lane_addr + AC_S + KH * half, o_addr = lane_addr + AC_O + 32 * half;
    int slot = 0;
    const float scale = rsqrtf((float)AHD) * 1.44269504f;      // q / sqrt(depth), times log2(e): the softmax below works in base 2
    int kt = 0, it = 0;
    for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++it) {
      const int qt = item % n_qt, bh = item / n_qt, h = bh % g.n_heads, b = bh / g.n_heads;
      const int qi = qt * AT + row;
      const bool qin = qi < g.T;
      const bool qvalid = MASK ? (qin && g.valid[(size_t)b * g.T + qi]) : true;
      if (half == 0) {   // Q row -> fp16 hi | lo in TMEM (the previous item's S products are complete: its last s_full was awaited)
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
        const int k0 = j * AT + KH * half;                     // first key of this thread's half of the tile
        mbar_wait_bounded(&s_full[half], kt & 1); tc_fence_after();
        const uint8_t* vk = sValid + slot * AT + KH * half;
        slot = slot + 1 == A_SLOTS ? 0 : slot + 1;
        auto logit = [&](float s, int c) -> float {          // c: key index inside the half tile
          const int kj = k0 + c;
          if (kj >= g.T) return -INFINITY;                    // outside the batch: not a key at all
          if (MASK && !(kj <= qi && qvalid && vk[c])) return -1e10f;      // logits += -1e10 (1 - mask) absorbs s in fp32
          return s;
        };
        // interior tiles of the unmasked mode need no per-element tests (every key exists, nothing is masked)
        const bool plain = !MASK && k0 + KH <= g.T;
        float mx = -INFINITY;
        {   // the chunk loads are issued one ahead of their use: their latency is what this pass costs
          float sa[16], sb[16];
          tmem_ld16(s_addr, sa);
#pragma unroll
          for (int c = 0; c < KH / 16; ++c) {
            float (&cur)[16] = (c & 1) ? sb : sa;
            float (&nxt)[16] = (c & 1) ? sa : sb;
            tmem_wait_ld(); tmem_ld_fence16(cur);
            if (c + 1 < KH / 16) tmem_ld16(s_addr + 16 * (c + 1), nxt);
            if (plain) {
#pragma unroll
              for (int e = 0; e < 16; e += 2) mx = fmaxf(mx, fmaxf(cur[e], cur[e + 1]));
            } else {
#pragma unroll
              for (int e = 0; e < 16; ++e) mx = fmaxf(mx, logit(cur[e], 16 * c + e));
            }
          }
        }
        const float m_new = fmaxf(m_run, mx);
        const float alpha = (m_run == -INFINITY) ? 0.0f : fast_ex2(m_run - m_new);
        if (kt > 0) { mbar_wait_bounded(&pv_done[half], (kt - 1) & 1); tc_fence_after(); }      // P and O of this half are free again
        float2 psum = make_float2(0.0f, 0.0f);
        const float2 nm = make_float2(-m_new, -m_new);
        {
          float sa[16], sb[16];
          tmem_ld16(s_addr, sa);
#pragma unroll
          for (int c = 0; c < KH / 16; ++c) {
            float (&s)[16] = (c & 1) ? sb : sa;
            float (&nxt)[16] = (c & 1) ? sa : sb;
            tmem_wait_ld(); tmem_ld_fence16(s);
            if (c + 1 < KH / 16) tmem_ld16(s_addr + 16 * (c + 1), nxt);      // lands while this chunk is exponentiated
            uint32_t hi[8], lo[8];
            if (plain) {
#pragma unroll
              for (int e = 0; e < 8; ++e) {
                const float2 d = __fadd2_rn(make_float2(s[2 * e], s[2 * e + 1]), nm);
                const float2 pp = make_float2(fast_ex2(d.x), fast_ex2(d.y));
                psum = __fadd2_rn(psum, pp);
                if (PS) split_h2x(pp, hi[e], lo[e]); else hi[e] = pack_h2(pp.x, pp.y);
              }
            } else {
#pragma unroll
              for (int e = 0; e < 8; ++e) {
                const float l0 = logit(s[2 * e], 16 * c + 2 * e), l1 = logit(s[2 * e + 1], 16 * c + 2 * e + 1);
                const float2 pp = make_float2((l0 == -INFINITY) ? 0.0f : fast_ex2(l0 - m_new), (l1 == -INFINITY) ? 0.0f : fast_ex2(l1 - m_new));
                psum = __fadd2_rn(psum, pp);
                if (PS) split_h2x(pp, hi[e], lo[e]); else hi[e] = pack_h2(pp.x, pp.y);
              }
            }
            tmem_st8(s_addr + 16 * c, hi);               // in place: these 16 logits are in registers
            if (PS) tmem_st8(s_addr + 16 * c + 8, lo);
          }
        }
        l_run = fmaf(l_run, alpha, psum.x + psum.y);
        m_run = m_new;
        if (j > 0 && __any_sync(0xffffffffu, alpha != 1.0f)) {      // rescale the running output (rare once the maximum has settled)
          float o[32];
          tmem_ld32(o_addr, o); tmem_wait_ld();
#pragma unroll
          for (int e = 0; e < 32; ++e) o[e] *= alpha;
          tmem_st32(o_addr, reinterpret_cast<const uint32_t(&)[32]>(o));
        }
        tmem_wait_st(); tc_fence_before(); __syncwarp();
        if (lane == 0) mbar_arrive(&p_ready[half]);
      }
      // ---- merge of the key halves of a row, O / l -> att
      if (NH == 2) {
        if (half == 1) sML[(it & 1) * AT + row] = make_float2(m_run, l_run);
        asm volatile("bar.sync 2, 256;" ::: "memory");
      }
      if (half == 0) {
        float ia, ib = 0.0f;
        if (NH == 2) {
          const float2 ml = sML[(it & 1) * AT + row];
          const float m = fmaxf(m_run, ml.x);
          const float wa = (m_run == -INFINITY) ? 0.0f : fast_ex2(m_run - m), wb = (ml.x == -INFINITY) ? 0.0f : fast_ex2(ml.x - m);
          const float inv = 1.0f / fmaf(l_run, wa, ml.y * wb);
          ia = wa * inv; ib = wb * inv;
        } else {
          ia = 1.0f / l_run;
        }
        mbar_wait_bounded(&pv_done[0], (kt - 1) & 1);
        if (NH == 2) mbar_wait_bounded(&pv_done[NH - 1], (kt - 1) & 1);
        tc_fence_after();
        float* dst = g.att + ((size_t)b * g.T + qi) * g.d_model + h * AHD;
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          float oa[16], ob[16];
          tmem_ld16(lane_addr + AC_O + 16 * c, oa);
          if (NH == 2) tmem_ld16(lane_addr + AC_O + 32 + 16 * c, ob);
          tmem_wait_ld();
          if (qin) {
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              float4 o = make_float4(oa[4 * q] * ia, oa[4 * q + 1] * ia, oa[4 * q + 2] * ia, oa[4 * q + 3] * ia);
              if (NH == 2) { o.x = fmaf(ob[4 * q], ib, o.x); o.y = fmaf(ob[4 * q + 1], ib, o.y); o.z = fmaf(ob[4 * q + 2], ib, o.z); o.w = fmaf(ob[4 * q + 3], ib, o.w); }
              *reinterpret_cast<float4*>(dst + 16 * c + 4 * q) = o;
            }
          }
        }
        tc_fence_before();
      }
    }
  } else if (warp == SW) {
    // ================= key-tile loader: one 32 KB bulk copy per tile =================
    int slot = 0, use = 0;
    const int n_kt = (g.T + AT - 1) / AT;
    for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
      const int qt = item % n_qt, bh = item / n_qt, b = bh / g.n_heads;
      const unsigned char* src = g.kv + (size_t)bh * n_kt * A_SLOT;
      const int nk = n_ktiles(qt);
      for (int j = 0; j < nk; ++j) {
        if (use >= 1) while (!mbar_try_wait(&kv_empty[slot], (use - 1) & 1)) __nanosleep(200);      // not on the critical path: leave the issue slots to the softmax warps
        if (MASK) {
          const int k0 = j * AT + 4 * lane;
#pragma unroll
          for (int e = 0; e < 4; ++e) sValid[slot * AT + 4 * lane + e] = (k0 + e < g.T) ? g.valid[(size_t)b * g.T + k0 + e] : (uint8_t)0;
          __syncwarp();
        }
        if (lane == 0) {
          mbar_arrive_expect_tx(&kv_full[slot], A_SLOT);
          bulk_g2s(ring + slot * A_SLOT, src + (size_t)j * A_SLOT, A_SLOT / 2, &kv_full[slot]);
          bulk_g2s(ring + slot * A_SLOT + A_SLOT / 2, src + (size_t)j * A_SLOT + A_SLOT / 2, A_SLOT / 2, &kv_full[slot]);
        }
        if (++slot == A_SLOTS) { slot = 0; ++use; }
      }
    }
  } else {
    // ================= MMA issue =================
    // The key halves are independent streams (own S columns, own O, own barriers): per tile and half, P V of the tile is followed at
    // once by Q K^T of the NEXT tile for the same half, so that while the softmax warps of one half work, the tensor pipe serves the other.
    constexpr uint32_t id_s = make_idesc_f16(AT, KH), id_o = make_idesc_f16(AT, AHD);
    const uint32_t e = elect_leader();      // one election; descriptors as 32-bit words (the issue path of tcn_chain.cu)
    int slot = 0, use = 0;
    auto issue_s = [&](int hf, int sl) {      // S_hf = Q K_hf^T of the tile in ring slot sl (the caller has awaited kv_full)
      const uint32_t k_hi = smem_u32(ring + sl * A_SLOT) + hf * (KH / 8) * 512, k_lo = k_hi + AK_PART;
#pragma unroll
      for (int part = 0; part < 3; ++part) {
        const uint32_t a0 = part == 1 ? AC_QLO : AC_QHI, b0 = part == 2 ? k_lo : k_hi;
#pragma unroll
        for (int ks = 0; ks < 2; ++ks)
          mma_ts_lo<desc_hi_noswz(512)>(tbase + AC_S + KH * hf, tbase + a0 + 8 * ks, desc_lo_noswz(b0 + ks * 256, 128), id_s, (part > 0 || ks > 0) ? 1u : 0u, e);
      }
      mma_commit_lo(&s_full[hf], e);
    };
    int kt = 0, it = 0;
    for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++it) {
      const int qt = item % n_qt;
      const int nk = n_ktiles(qt);
      mbar_wait_bounded(&q_full, it & 1);
      mbar_wait_bounded(&kv_full[slot], use & 1); tc_fence_after();
#pragma unroll
      for (int hf = 0; hf < NH; ++hf) issue_s(hf, slot);
      for (int j = 0; j < nk; ++j, ++kt) {
        int nslot = slot + 1, nuse = use;
        if (nslot == A_SLOTS) { nslot = 0; ++nuse; }
        const uint32_t v_hi = smem_u32(ring + slot * A_SLOT) + 2 * AK_PART, v_lo = v_hi + AK_PART;
#pragma unroll
        for (int hf = 0; hf < NH; ++hf) {
          mbar_wait_bounded(&p_ready[hf], kt & 1); tc_fence_after();
#pragma unroll
          for (int part = 0; part < 3; ++part) {
            if (!PS && part == 1) continue;      // no P_lo
            const uint32_t b0 = part == 2 ? v_lo : v_hi;
#pragma unroll
            for (int k2 = 0; k2 < 8 / NH; ++k2) {      // keys 16 ks ..: hi pairs of P at column 16 ks, lo pairs 8 columns further
              const int ks = hf * (8 / NH) + k2;
              mma_ts_lo<desc_hi_noswz(2048)>(tbase + AC_O + 32 * hf, tbase + AC_S + 16 * ks + (part == 1 ? 8 : 0),
                                             desc_lo_noswz(b0 + ks * 256, 128), id_o, (j > 0 || part > 0 || k2 > 0) ? 1u : 0u, e);
            }
          }
          mma_commit_lo(&pv_done[hf], e);
          if (hf == NH - 1) mma_commit_lo(&kv_empty[slot], e);      // every product that reads the slot has been issued
          if (j + 1 < nk) {
            if (hf == 0) { mbar_wait_bounded(&kv_full[nslot], nuse & 1); tc_fence_after(); }
            issue_s(hf, nslot);
          }
        }
        slot = nslot; use = nuse;
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == SW + 1) tmem_dealloc(tbase, A_TMEM_COLS);
}

size_t mhanet_umma_attention_workspace(const dxi_net& net, int B, int T) {
  return (size_t)B * net.cfg.n_heads * ((T + AT - 1) / AT) * A_SLOT;
}

// kv: mhanet_umma_attention_workspace(net, B, T) bytes (16-byte aligned): the packed key tiles, already written by the fused QKV
// projection (kv_packed, mha_umma.cu) or packed here from the K / V columns of qkv
int mhanet_umma_attention(const dxi_net& net, const float* qkv, const uint8_t* valid, int B, int T, float* att, void* kv, bool kv_packed, cudaStream_t st) {
  const dxi_net_cfg& c = net.cfg;
  if (c.d_model / c.n_heads != AHD || (c.d_model & 3)) { set_error("tcgen05 attention is built for head size 32"); return DXI_E_INVALID; }
  AttnArgs a{qkv, reinterpret_cast<const unsigned char*>(kv), valid, att, B, T, c.d_model, c.n_heads};
  constexpr int threads = (4 * A_NH + 2) * 32;
  // The probabilities are rounded to fp16 once (Q, K and V keep the hi | lo split): measured against the float64 oracle at T = 1875 the
  // maximum |d xi_hat| is 1.5e-4 dB unmasked either way and 1.1e-3 dB (split: 1.3e-4) with the causal mask, whose first rows average over
  // few keys - 100 times inside the 0.1 dB tolerance, for 10 % of the kernel's time.  DXI_ATTN_P_SPLIT=1 restores P as fp16 hi | lo.
  const bool ps = getenv("DXI_ATTN_P_SPLIT") && atoi(getenv("DXI_ATTN_P_SPLIT"));
  const bool mk = c.mask_mode == DXI_MASK_CAUSAL_PAD;
  auto kern = mk ? (ps ? attn_umma_kernel<1, A_NH, true> : attn_umma_kernel<1, A_NH, false>) : (ps ? attn_umma_kernel<0, A_NH, true> : attn_umma_kernel<0, A_NH, false>);
  DXI_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, A_SMEM));      // per call: the attribute is per device
  int n_sm = 148;
  { int dev = 0; cudaGetDevice(&dev); cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev); }
  const int n_kt = (T + AT - 1) / AT;
  const int items = B * c.n_heads * n_kt;
  const int grid = items < 2 * n_sm ? items : 2 * n_sm;      // two CTAs per SM
  ProfScope prof("mha_attn", st, kv_packed ? 1 : 2);
  if (!kv_packed) {
    attn_pack_kv_kernel<<<items, 128, 0, st>>>(qkv, reinterpret_cast<unsigned char*>(kv), T, c.d_model, c.n_heads);
    DXI_LAUNCHED("attn_pack_kv_kernel");
  }
  kern<<<grid, threads, A_SMEM, st>>>(a);
  DXI_LAUNCHED("attn_umma_kernel");
  return DXI_OK;
}

}  // namespace dxi
