// ResNetV2 forward on the 5th-generation tensor cores (tcgen05 + TMEM), precision modes
// DXI_PREC_F16X3 (fp16 hi/lo split operands, 3 MMAs per product: fp32-level accuracy) and
// DXI_PREC_F16 (one fp16 MMA per product).  Restates deepxi/network/tcn.py:116-225.
//
// "Shifted" fusion.  A residual block is  h += W3 u(W2 (*)d u(W1 u(h)))  with u = ReLU -> LayerNorm.
// Only the dilated k=3 conv looks at other frames, and it looks at c1 = LN(ReLU(W1 LN(ReLU(h)) + b1)),
// a 64-channel tensor.  Stage s of the pipeline therefore computes, for a tile of 128 frames,
//     back half of block s-1 :  GEMM1  [128 x 192] x [192 x 64]   (three time-shifted copies of c1_{s-1})
//                               GEMM2  [128 x  64] x [ 64 x 256]  -> h += . + b3      (fp32 residual stream)
//     front half of block s  :  GEMM3  [128 x 256] x [256 x 64]   -> c1_s
// so no halo is ever recomputed and per frame and block HBM/L2 sees: h read + write (2 x 1 KB, fp32) and
// c1 read (3 shifted, L2 hits) + write (2 x 128 B fp16 hi/lo).  Stage 0 has no back half (h comes from the
// stem), stage n_blocks has no front half.
//
// Per CTA (persistent over tiles): the packed fp16 weights of the stage (176 KB hi+lo) stay in shared
// memory as tcgen05 B operands (128-byte swizzle); every A operand is produced by the epilogue threads
// straight into tensor memory (thread r owns frame r: TMEM lane r), so LayerNorm is a purely
// thread-local reduction over the accumulator row and nothing is staged through shared memory.
// TMEM columns: [0,256) accumulators / fp32 residual row, [256,512) A operands (hi | lo).
// Warps 0-7: epilogue (two threads per frame, half the columns each; TMEM lane quarter = warp & 3);
// warp 8: weight load (bulk async copy), L2 prefetch of the next tile, MMA issue.  GEMM2 is committed in four
// column groups and GEMM3 is issued in eight K-chunks so that tensor-core time hides behind the epilogue.
#include <vector>
#include "net.cuh"
#include "umma.cuh"

namespace dxi {
using namespace umma;

constexpr int TILE = 128;                 // frames per tile (UMMA M)
constexpr int C1_PAD = 32;                // zero rows before / after every utterance in the c1 planes (2 * max dilation)
constexpr int IMG_W2 = 0;                 // [192 -> 64] : 3 K-chunks x [64 rows x 128 B]
constexpr int IMG_W3 = 24576;             // [ 64 -> 256]: 1 K-chunk  x [256 rows x 128 B]
constexpr int IMG_W1 = 57344;             // [256 -> 64] : 4 K-chunks x [64 rows x 128 B]
constexpr int IMG_PART = 90112;           // bytes of one precision part (hi, then lo)
constexpr int IMG_BIAS = 2 * IMG_PART;    // b2[64], b3[256], b1[64] fp32
constexpr int IMG_BYTES = IMG_BIAS + 384 * 4;

// TMEM column map
constexpr uint32_t COL_ACC = 0;
constexpr uint32_t COL_A1_HI = 256, COL_A1_LO = 352;     // 3 taps x 32 columns each
constexpr uint32_t COL_A2_HI = 448, COL_A2_LO = 480;     // 32 columns each
constexpr uint32_t COL_A3_HI = 256, COL_A3_LO = 384;     // 128 columns each

struct StageArgs {
  const unsigned char* img;      // packed weights of this stage
  float* h;                      // residual stream, tiled: [tile][c/4][row][4] fp32
  const __half* c1_in;           // [B][2 planes][8 units][Ts][8] fp16 (hi plane, lo plane)
  __half* c1_out;
  int T, tiles_per_utt, n_tiles, Ts;
  int shift0, shift1, shift2;    // frame offsets of the three taps: tap j reads frame t - shift_j
  int has_back, has_front;
  long long* dbg;               // optional: clock64 stamps of the epilogue phases (16 per tile), see dxi_debug_tcn_clocks
};

__device__ __forceinline__ float relu(float x) { return fmaxf(x, 0.0f); }

constexpr int EPI_WARPS = 8;                      // 2 warps per TMEM lane quarter: each owns half the columns of a row
constexpr int EPI_THREADS = EPI_WARPS * 32;
constexpr int STAGE_THREADS = EPI_THREADS + 32;   // + 1 warp: weight load, L2 prefetch, MMA issue

__device__ __forceinline__ void epi_barrier() { asm volatile("bar.sync 1, %0;" ::"n"(EPI_THREADS) : "memory"); }
__device__ __forceinline__ void prefetch_l2(const void* p, uint32_t bytes) {
  asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p), "r"(bytes) : "memory");
}

// LayerNorm statistics of one row that is split between two threads (n values each): every thread
// contributes (mean, M2 = sum of squared deviations from its own mean); Chan's formula merges them.
__device__ __forceinline__ void ln_merge(float2* red, int row, int hb, float n, float mean_h, float m2_h, float& inv, float& off) {
  red[hb * TILE + row] = make_float2(mean_h, m2_h);
  epi_barrier();
  const float2 o = red[(hb ^ 1) * TILE + row];
  const float mean = 0.5f * (mean_h + o.x);
  const float dm = mean_h - o.x;
  const float var = (m2_h + o.y + dm * dm * (0.5f * n)) / (2.0f * n);    // biased variance over 2n values
  inv = rsqrtf(var + 1e-6f);
  off = -mean * inv;
}

// u(x) = ReLU -> LayerNorm over a 64-wide row of which this thread holds 32 values (two-pass locally)
__device__ __forceinline__ void relu_ln_half(float (&a)[32], const float* bias, float2* red, int row, int hb) {
  float s = 0.0f;
#pragma unroll
  for (int j = 0; j < 32; ++j) { a[j] = relu(a[j] + bias[j]); s += a[j]; }
  const float mean_h = s * (1.0f / 32.0f);
  float q = 0.0f;
#pragma unroll
  for (int j = 0; j < 32; ++j) { const float d = a[j] - mean_h; q = fmaf(d, d, q); }
  float inv, off;
  ln_merge(red, row, hb, 32.0f, mean_h, q, inv, off);
#pragma unroll
  for (int j = 0; j < 32; ++j) a[j] = fmaf(a[j], inv, off);
}

template <bool SPLIT>
__global__ void __launch_bounds__(STAGE_THREADS, 1) tcn_stage_kernel(const StageArgs p) {
  extern __shared__ unsigned char smem_raw[];
  __shared__ __align__(8) uint64_t bar_w, bar_a1, bar_a2, bar_a3[8], bar_d1, bar_d2[4], bar_d3;
  __shared__ uint32_t tmem_slot;
  __shared__ float2 red[3][2 * TILE];
  unsigned char* sW = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const float* sBias = reinterpret_cast<const float*>(sW + IMG_BIAS);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  if (warp == EPI_WARPS) tmem_alloc(&tmem_slot, 512);
  if (tid == 0) {
    mbar_init(&bar_w, 1);
    mbar_init(&bar_a1, EPI_THREADS);
    mbar_init(&bar_a2, EPI_THREADS);
    for (int i = 0; i < 8; ++i) mbar_init(&bar_a3[i], EPI_THREADS / 2);
    mbar_init(&bar_d1, 1);
    for (int i = 0; i < 4; ++i) mbar_init(&bar_d2[i], 1);
    mbar_init(&bar_d3, 1);
    fence_mbar_init();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  // The CTA owns the SM's whole tensor memory (512 columns, one CTA per SM), so the allocation starts at
  // lane 0 / column 0: TMEM addresses below are compile-time constants.
  if (tmem_slot != 0) __trap();
  constexpr uint32_t tb = 0;

  if (warp == EPI_WARPS) {
    // ================= weight load, L2 prefetch of the next tile, MMA issue =================
    // The whole warp runs this code convergently; single-thread operations elect a lane inside the asm.
    if (elect_one()) {
      mbar_arrive_expect_tx(&bar_w, IMG_BYTES);
      for (int off = 0; off < IMG_BYTES; off += 16384) {
        const int n = IMG_BYTES - off < 16384 ? IMG_BYTES - off : 16384;
        bulk_g2s(sW + off, p.img + off, n, &bar_w);
      }
    }
    __syncwarp();
    mbar_wait(&bar_w, 0);
    const uint32_t w_hi = smem_u32(sW), w_lo = w_hi + IMG_PART;
    constexpr uint32_t id64 = make_idesc_f16(TILE, 64);
    constexpr int NPART = SPLIT ? 3 : 1;          // (a_hi, w_hi) [+ (a_lo, w_hi) + (a_hi, w_lo)]
    uint32_t ph = 0;
    for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x) {
      {   // pull the next tile's residual rows and c1 rows into L2 while this tile is being computed
        const int nt = tile + gridDim.x;
        if (nt < p.n_tiles) {
          const char* hn = reinterpret_cast<const char*>(p.h + (size_t)nt * (TILE * 256));
          if (lane < 8) prefetch_l2(hn + lane * 16384, 16384);
          if (p.has_back && lane >= 16 && lane < (SPLIT ? 32 : 24)) {
            const int b = nt / p.tiles_per_utt, t0 = (nt - b * p.tiles_per_utt) * TILE;
            const int lo_row = t0 - (p.shift0 > 0 ? p.shift0 : 0) + C1_PAD;
            const int n_rows = TILE + (p.shift0 > 0 ? p.shift0 : 0) - (p.shift2 < 0 ? p.shift2 : 0);
            const __half* cb = p.c1_in + (size_t)b * 2 * 8 * p.Ts * 8;
            prefetch_l2(cb + ((size_t)(lane - 16) * p.Ts + lo_row) * 8, n_rows * 16);
          }
        }
        __syncwarp();
      }
      if (p.has_back) {
        mbar_wait(&bar_a1, ph); tc_fence_after();
        {   // c2 = W2 (*) [c1(t-s0) | c1(t-s1) | c1(t-s2)] : K = 192, N = 64
          uint32_t acc = 0;
#pragma unroll
          for (int part = 0; part < NPART; ++part) {
            const uint32_t a0 = part == 1 ? COL_A1_LO : COL_A1_HI, w0 = (part == 2 ? w_lo : w_hi) + IMG_W2;
#pragma unroll
            for (int ks = 0; ks < 12; ++ks) {
              mma_ts_elect(COL_ACC, a0 + 8 * ks, make_smem_desc_sw128(w0 + (ks >> 2) * 64 * 128 + (ks & 3) * 32), id64, acc);
              acc = 1;
            }
          }
        }
        mma_commit_elect(&bar_d1);
        mbar_wait(&bar_a2, ph); tc_fence_after();
#pragma unroll
        for (int g = 0; g < 4; ++g) {   // W3 u(c2): K = 64, N = 256 in four column groups, each committed on its own
          uint32_t acc = 0;
#pragma unroll
          for (int part = 0; part < NPART; ++part) {
            const uint32_t a0 = part == 1 ? COL_A2_LO : COL_A2_HI, w0 = (part == 2 ? w_lo : w_hi) + IMG_W3 + g * 64 * 128;
#pragma unroll
            for (int ks = 0; ks < 4; ++ks) {
              mma_ts_elect(COL_ACC + 64 * g, a0 + 8 * ks, make_smem_desc_sw128(w0 + ks * 32), id64, acc);
              acc = 1;
            }
          }
          mma_commit_elect(&bar_d2[g]);
        }
      }
      if (p.has_front) {
        uint32_t acc = 0;
#pragma unroll
        for (int cc = 0; cc < 8; ++cc) {   // W1' u(h): K = 256 in eight 32-channel chunks, issued as the epilogue produces them
          mbar_wait(&bar_a3[cc], ph);
          if (cc == 0) continue;           // D overwrites residual columns [0,64): wait until chunks 0 AND 1 are consumed
          tc_fence_after();
#pragma unroll
          for (int c2 = (cc == 1 ? 0 : cc); c2 <= cc; ++c2)
#pragma unroll
            for (int part = 0; part < NPART; ++part) {
              const uint32_t a0 = part == 1 ? COL_A3_LO : COL_A3_HI, w0 = (part == 2 ? w_lo : w_hi) + IMG_W1;
#pragma unroll
              for (int ks = 2 * c2; ks < 2 * c2 + 2; ++ks) {
                mma_ts_elect(COL_ACC, a0 + 8 * ks, make_smem_desc_sw128(w0 + (ks >> 2) * 64 * 128 + (ks & 3) * 32), id64, acc);
                acc = 1;
              }
            }
        }
        mma_commit_elect(&bar_d3);
      }
      ph ^= 1;
    }
  } else {
    // ================= epilogue: threads (q, lane) and (q+4, lane) share frame r = 32 q + lane =================
    mbar_wait(&bar_w, 0);        // biases live in the weight image
    const int hb = warp >> 2, row = (warp & 3) * 32 + lane;
    const uint32_t lane_addr = (uint32_t)((warp & 3) * 32) << 16;
    const uint32_t t_acc = tb + lane_addr + COL_ACC;
    uint32_t ph = 0;
    const bool stamp = p.dbg != nullptr && tid == 0;
#define DXI_STAMP(k) do { if (stamp) p.dbg[(size_t)tile * 16 + (k)] = clock64(); } while (0)
    for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x) {
      const int b = tile / p.tiles_per_utt, t0 = (tile - b * p.tiles_per_utt) * TILE;
      const int t = t0 + row;
      DXI_STAMP(0);
      const bool valid = t < p.T;
      float* hrow = p.h + (size_t)tile * (TILE * 256) + row * 4;          // + c4 * 512
      // residual values of the first chunk this thread owns: issue the loads before anything else
      float4 hv[8];
#pragma unroll
      for (int q = 0; q < 8; ++q) hv[q] = *reinterpret_cast<const float4*>(hrow + (size_t)(hb * 8 + q) * (TILE * 4));
      if (p.has_back) {
        // ---- A1: three shifted copies of c1 (hi | lo planes) -> TMEM; this thread moves units 4hb..4hb+3
        const __half* cb = p.c1_in + (size_t)b * 2 * 8 * p.Ts * 8;
        uint4 qv[3][SPLIT ? 2 : 1][4];
#pragma unroll
        for (int j = 0; j < 3; ++j) {
          const int shift = j == 0 ? p.shift0 : (j == 1 ? p.shift1 : p.shift2);
          const size_t r_in = (size_t)(t - shift + C1_PAD);
#pragma unroll
          for (int plane = 0; plane < (SPLIT ? 2 : 1); ++plane)
#pragma unroll
            for (int u = 0; u < 4; ++u)
              qv[j][plane][u] = __ldg(reinterpret_cast<const uint4*>(cb + ((size_t)(plane * 8 + 4 * hb + u) * p.Ts + r_in) * 8));
        }
#pragma unroll
        for (int j = 0; j < 3; ++j)
#pragma unroll
          for (int plane = 0; plane < (SPLIT ? 2 : 1); ++plane) {
            uint32_t r[16];
#pragma unroll
            for (int u = 0; u < 4; ++u) { r[4 * u] = qv[j][plane][u].x; r[4 * u + 1] = qv[j][plane][u].y; r[4 * u + 2] = qv[j][plane][u].z; r[4 * u + 3] = qv[j][plane][u].w; }
            tmem_st16(tb + lane_addr + (plane ? COL_A1_LO : COL_A1_HI) + 32 * j + 16 * hb, r);
          }
        tmem_wait_st(); tc_fence_before(); mbar_arrive(&bar_a1);
        DXI_STAMP(1);
        // ---- E1: c2 = acc + b2 -> u(.) -> A2 (this thread: channels 32hb .. 32hb+31)
        mbar_wait(&bar_d1, ph); tc_fence_after();
        DXI_STAMP(2);
        {
          float a[32];
          tmem_ld32(t_acc + 32 * hb, a); tmem_wait_ld();
          relu_ln_half(a, sBias + 32 * hb, red[0], row, hb);
          uint32_t hi[16], lo[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) split_h2(a[2 * j], a[2 * j + 1], hi[j], lo[j]);
          tmem_st16(tb + lane_addr + COL_A2_HI + 16 * hb, hi);
          if (SPLIT) tmem_st16(tb + lane_addr + COL_A2_LO + 16 * hb, lo);
        }
        tmem_wait_st(); tc_fence_before(); mbar_arrive(&bar_a2);
        DXI_STAMP(3);
      }
      // ---- E2: h_new = h + acc + b3 (fp32: back to HBM and kept in TMEM); this thread owns the 32-column
      //      chunks cc = 2 i + hb.  Statistics of ReLU(h_new) as shifted sums.
      float s1 = 0.0f, s2 = 0.0f, kshift = 0.0f;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int cc = 2 * i + hb;
        float4 hn[8];
        if (i < 3) {
#pragma unroll
          for (int q = 0; q < 8; ++q) hn[q] = *reinterpret_cast<const float4*>(hrow + (size_t)((cc + 2) * 8 + q) * (TILE * 4));
        }
        float v[32];
        if (p.has_back) {
          mbar_wait(&bar_d2[cc >> 1], ph); tc_fence_after();
          if (i == 0) DXI_STAMP(4);
          tmem_ld32(t_acc + 32 * cc, v); tmem_wait_ld();
          const float* bb = sBias + 64 + 32 * cc;
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            v[4 * q] += hv[q].x + bb[4 * q];
            v[4 * q + 1] += hv[q].y + bb[4 * q + 1];
            v[4 * q + 2] += hv[q].z + bb[4 * q + 2];
            v[4 * q + 3] += hv[q].w + bb[4 * q + 3];
          }
        } else {
#pragma unroll
          for (int q = 0; q < 8; ++q) { v[4 * q] = hv[q].x; v[4 * q + 1] = hv[q].y; v[4 * q + 2] = hv[q].z; v[4 * q + 3] = hv[q].w; }
        }
        if (!valid) {
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = 0.0f;
        }
        if (p.has_back) {
#pragma unroll
          for (int q = 0; q < 8; ++q)
            *reinterpret_cast<float4*>(hrow + (size_t)(cc * 8 + q) * (TILE * 4)) = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
        }
        if (p.has_front) {
          if (i == 0) kshift = relu(v[0]);
#pragma unroll
          for (int j = 0; j < 32; ++j) { const float r = relu(v[j]) - kshift; s1 += r; s2 = fmaf(r, r, s2); }
          tmem_st32(t_acc + 32 * cc, reinterpret_cast<const uint32_t(&)[32]>(v));
        }
        if (i < 3) {
#pragma unroll
          for (int q = 0; q < 8; ++q) hv[q] = hn[q];
        }
      }
      DXI_STAMP(5);
      if (p.has_front) {
        tmem_wait_st();
        float inv, off;
        {
          const float m1 = s1 * (1.0f / 128.0f);
          ln_merge(red[1], row, hb, 128.0f, kshift + m1, fmaxf(s2 - s1 * m1, 0.0f), inv, off);
        }
        DXI_STAMP(6);
        // ---- A3 = u(h_new) as fp16 hi | lo, chunk by chunk (the MMA warp starts on a chunk as soon as it lands)
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int cc = 2 * i + hb;
          float v[32];
          tmem_ld32(t_acc + 32 * cc, v); tmem_wait_ld();
          uint32_t hi[16], lo[16];
#pragma unroll
          for (int j = 0; j < 16; ++j)
            split_h2(fmaf(relu(v[2 * j]), inv, off), fmaf(relu(v[2 * j + 1]), inv, off), hi[j], lo[j]);
          tmem_st16(tb + lane_addr + COL_A3_HI + 16 * cc, hi);
          if (SPLIT) tmem_st16(tb + lane_addr + COL_A3_LO + 16 * cc, lo);
          tmem_wait_st(); tc_fence_before(); mbar_arrive(&bar_a3[cc]);
        }
        // ---- E3: c1' = u(acc + b1') -> fp16 hi | lo planes in HBM (zeros for frames beyond T)
        DXI_STAMP(7);
        mbar_wait(&bar_d3, ph); tc_fence_after();
        DXI_STAMP(8);
        float a[32];
        tmem_ld32(t_acc + 32 * hb, a); tmem_wait_ld();
        relu_ln_half(a, sBias + 320 + 32 * hb, red[2], row, hb);
        __half* ob = p.c1_out + (size_t)b * 2 * 8 * p.Ts * 8;
        const size_t r_out = (size_t)(t + C1_PAD);
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          uint32_t hi[4], lo[4];
#pragma unroll
          for (int j = 0; j < 4; ++j) split_h2(valid ? a[8 * u + 2 * j] : 0.0f, valid ? a[8 * u + 2 * j + 1] : 0.0f, hi[j], lo[j]);
          *reinterpret_cast<uint4*>(ob + ((size_t)(4 * hb + u) * p.Ts + r_out) * 8) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
          if (SPLIT) *reinterpret_cast<uint4*>(ob + ((size_t)(8 + 4 * hb + u) * p.Ts + r_out) * 8) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
        }
        tc_fence_before();
      }
      DXI_STAMP(9);
      ph ^= 1;
    }
#undef DXI_STAMP
  }
  tc_fence_before();
  __syncthreads();
  if (warp == EPI_WARPS) tmem_dealloc(tb, 512);
}

// ---------------------------------------------------------------------------------------------------
// Stem and head on the CUDA cores in fp32 (the stem is the layer the 0.1 dB budget is most sensitive to,
// SURVEY F8): Conv1D(256,1)+b -> LN(gamma) -> ReLU into the tiled residual layout, and
// Conv1D(257,1)+b -> sigmoid out of it.  One CTA = 32 frames.
// ---------------------------------------------------------------------------------------------------
constexpr int SM_ROWS = 32;

__device__ __forceinline__ float warp_sum_f(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__device__ __forceinline__ size_t h_tiled_index(int tile, int row, int c) {
  return (size_t)tile * (TILE * 256) + ((size_t)(c >> 2) * TILE + row) * 4 + (c & 3);
}

// out[r][n] = sum_k A[r][k] W[k][n]; A tile in shared memory [SM_ROWS][LDA]; thread computes 4 rows x (NOUT/32) cols
template <int KIN, int NOUT, bool STEM>
__global__ void __launch_bounds__(256) stem_head_kernel(const float* __restrict__ in, const float* __restrict__ W,
                                                        const float* __restrict__ bias, const float* __restrict__ gamma,
                                                        float* __restrict__ out, int T, int tiles_per_utt, int groups_per_utt) {
  constexpr int LDA = KIN + 1;
  constexpr int CPT = (NOUT + 31) / 32;       // columns per thread, interleaved by 32
  constexpr int LDO = NOUT + 1;
  extern __shared__ __align__(16) float sm[];
  float* A = sm;                                // [SM_ROWS][LDA], reused for the output tile [SM_ROWS][LDO]
  float* Wc = sm + SM_ROWS * (LDA > LDO ? LDA : LDO);   // [16][NOUT]
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int b = blockIdx.x / groups_per_utt, t0 = (blockIdx.x - b * groups_per_utt) * SM_ROWS;
  // ---- load the input rows (coalesced in either layout)
  if (STEM) {
    for (int r = warp; r < SM_ROWS; r += 8) {
      const int t = t0 + r;
      for (int c = lane; c < KIN; c += 32) A[r * LDA + c] = t < T ? __ldcs(in + ((size_t)b * T + t) * KIN + c) : 0.0f;
    }
  } else {
    for (int i = tid; i < SM_ROWS * (KIN / 4); i += 256) {
      const int c4 = i / SM_ROWS, r = i - c4 * SM_ROWS;
      const int t = t0 + r;
      float4 x = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
      if (t < T) x = *reinterpret_cast<const float4*>(in + h_tiled_index(b * tiles_per_utt + t / TILE, t % TILE, 4 * c4));
      float* dst = A + r * LDA + 4 * c4;
      dst[0] = x.x; dst[1] = x.y; dst[2] = x.z; dst[3] = x.w;
    }
  }
  const int ty = tid >> 5, tx = tid & 31;       // 8 row groups x 32 column lanes
  float acc[4][CPT];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int c = 0; c < CPT; ++c) acc[i][c] = 0.0f;
  for (int k0 = 0; k0 < KIN; k0 += 16) {
    __syncthreads();
    for (int i = tid; i < 16 * NOUT; i += 256) {
      const int kk = i / NOUT, c = i - kk * NOUT;
      Wc[i] = (k0 + kk < KIN) ? __ldg(W + (size_t)(k0 + kk) * NOUT + c) : 0.0f;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < 16; ++kk) {
      if (k0 + kk >= KIN) break;
      float a[4], w[CPT];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = A[(ty * 4 + i) * LDA + k0 + kk];
#pragma unroll
      for (int c = 0; c < CPT; ++c) w[c] = (tx + 32 * c < NOUT) ? Wc[kk * NOUT + tx + 32 * c] : 0.0f;
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int c = 0; c < CPT; ++c) acc[i][c] = fmaf(a[i], w[c], acc[i][c]);
    }
  }
  __syncthreads();
  float* O = A;
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int c = 0; c < CPT; ++c) {
      const int col = tx + 32 * c;
      if (col < NOUT) O[(ty * 4 + i) * LDO + col] = acc[i][c] + __ldg(bias + col);
    }
  __syncthreads();
  if (STEM) {
    // LayerNorm(scale gamma, no centre, eps 1e-6) -> ReLU (tcn.py:176-179); tiled store: for a fixed group of
    // 4 channels, consecutive frames are consecutive float4 -> each warp writes 32 frames x 16 B = 512 B runs
    for (int r = warp; r < SM_ROWS; r += 8) {
      const float* row = O + r * LDO;
      float s = 0.0f;
      for (int c = lane; c < NOUT; c += 32) s += row[c];
      const float mean = warp_sum_f(s) * (1.0f / NOUT);
      float q = 0.0f;
      for (int c = lane; c < NOUT; c += 32) { const float d = row[c] - mean; q = fmaf(d, d, q); }
      const float rs = rsqrtf(warp_sum_f(q) * (1.0f / NOUT) + 1e-6f);
      if (lane == 0) { Wc[r] = mean; Wc[SM_ROWS + r] = rs; }
    }
    __syncthreads();
    for (int i = tid; i < SM_ROWS * (NOUT / 4); i += 256) {
      const int c4 = i / SM_ROWS, r = i - c4 * SM_ROWS;
      const int t = t0 + r;
      if (t >= T) continue;
      const float mean = Wc[r], rs = Wc[SM_ROWS + r];
      float4 o;
      float* po = &o.x;
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int c = 4 * c4 + e;
        const float inv = rs * __ldg(gamma + c);
        po[e] = fmaxf(fmaf(O[r * LDO + c], inv, -mean * inv), 0.0f);
      }
      *reinterpret_cast<float4*>(out + h_tiled_index(b * tiles_per_utt + t / TILE, t % TILE, 4 * c4)) = o;
    }
  } else {
    for (int r = warp; r < SM_ROWS; r += 8) {
      const int t = t0 + r;
      if (t >= T) continue;
      for (int c = lane; c < NOUT; c += 32)
        __stcs(out + ((size_t)b * T + t) * NOUT + c, 1.0f / (1.0f + expf(-O[r * LDO + c])));
    }
  }
}

// ---------------------------------------------------------------------------------------------------
// Host side
// ---------------------------------------------------------------------------------------------------
static void pack_b_sw128(unsigned char* hi, unsigned char* lo, int N, int K, const float* W /* [K][N] */) {
  for (int n = 0; n < N; ++n)
    for (int k = 0; k < K; ++k) {
      const int c = k >> 6, kk = k & 63, u = kk >> 3, e = kk & 7;
      const size_t off = (size_t)c * N * 128 + (size_t)(n >> 3) * 1024 + (n & 7) * 128 + ((u ^ (n & 7)) * 16) + e * 2;
      const float w = W[(size_t)k * N + n];
      const __half h = __float2half_rn(w);
      const __half l = __float2half_rn(w - __half2float(h));
      memcpy(hi + off, &h, 2);
      memcpy(lo + off, &l, 2);
    }
}

static int n_dilations(int max_d_rate) { int n = 0; for (int m = max_d_rate; m > 0; m >>= 1) ++n; return n; }

int resnet_umma_prepare(dxi_net& net, cudaStream_t st) {
  const dxi_net_cfg& c = net.cfg;
  if (!(c.n_feat == 257 && c.n_outp == 257 && c.d_model == 256 && c.d_f == 64 && c.k == 3 && c.max_d_rate <= 16)) {
    set_error("tcgen05 ResNetV2 path is built for n_feat=n_outp=257, d_model=256, d_f=64, k=3, max_d_rate<=16");
    return DXI_E_INVALID;
  }
  const int n_stages = c.n_blocks + 1;
  std::vector<unsigned char> img((size_t)n_stages * IMG_BYTES, 0);
  for (int s = 0; s < n_stages; ++s) {
    unsigned char* base = img.data() + (size_t)s * IMG_BYTES;
    float* bias = reinterpret_cast<float*>(base + IMG_BIAS);
    if (s >= 1) {
      const int li = 2 + 3 * (s - 1);
      pack_b_sw128(base + IMG_W2, base + IMG_PART + IMG_W2, 64, 192, net.host_tensor(li + 1, "kernel")->data());   // [3*64][64]
      pack_b_sw128(base + IMG_W3, base + IMG_PART + IMG_W3, 256, 64, net.host_tensor(li + 2, "kernel")->data());   // [64][256]
      memcpy(bias, net.host_tensor(li + 1, "bias")->data(), 64 * 4);
      memcpy(bias + 64, net.host_tensor(li + 2, "bias")->data(), 256 * 4);
    }
    if (s < c.n_blocks) {
      const int li = 2 + 3 * s;
      pack_b_sw128(base + IMG_W1, base + IMG_PART + IMG_W1, 64, 256, net.host_tensor(li, "kernel")->data());       // [256][64]
      memcpy(bias + 320, net.host_tensor(li, "bias")->data(), 64 * 4);
    }
  }
  if (net.d_umma) { cudaFree(net.d_umma); net.d_umma = nullptr; }
  DXI_CUDA(cudaMalloc(&net.d_umma, img.size()));
  DXI_CUDA(cudaMemcpyAsync(net.d_umma, img.data(), img.size(), cudaMemcpyHostToDevice, st));
  DXI_CUDA(cudaStreamSynchronize(st));   // img is a local
  net.umma_bytes = img.size();
  return DXI_OK;
}

static thread_local long long* g_dbg_clocks = nullptr;
static thread_local int g_dbg_stage = -1;

static inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

int64_t resnet_umma_workspace_bytes(const dxi_net& net, int B, int T) {
  const int tiles = (T + TILE - 1) / TILE;
  const size_t Ts = (size_t)tiles * TILE + 2 * C1_PAD;
  const size_t h_bytes = (size_t)B * tiles * TILE * 256 * 4;
  const size_t c1_bytes = align_up((size_t)B * 2 * 8 * Ts * 16, 256);
  return (int64_t)(256 + h_bytes + 2 * c1_bytes);
}

int resnet_umma_forward(const dxi_net& net, const float* mag, int B, int T, float* xbar, void* ws, size_t ws_bytes,
                        cudaStream_t st) {
  const dxi_net_cfg& c = net.cfg;
  if ((int64_t)ws_bytes < resnet_umma_workspace_bytes(net, B, T)) { set_error("workspace too small"); return DXI_E_NOMEM; }
  const int tiles = (T + TILE - 1) / TILE;
  const int Ts = tiles * TILE + 2 * C1_PAD;
  const size_t h_bytes = (size_t)B * tiles * TILE * 256 * 4;
  const size_t c1_bytes = align_up((size_t)B * 2 * 8 * Ts * 16, 256);
  unsigned char* base = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(ws) + 255) & ~(uintptr_t)255);
  float* h = reinterpret_cast<float*>(base);
  __half* c1[2] = {reinterpret_cast<__half*>(base + h_bytes), reinterpret_cast<__half*>(base + h_bytes + c1_bytes)};
  // zero padding rows of the c1 planes (and everything else in them)
  DXI_CUDA(cudaMemsetAsync(c1[0], 0, 2 * c1_bytes, st));

  // ---- stem (fp32 CUDA cores) -> tiled h
  {
    constexpr int KIN = 257, NOUT = 256, LD = 258;
    const size_t smem = sizeof(float) * (SM_ROWS * LD + 16 * NOUT);
    auto kern = stem_head_kernel<KIN, NOUT, true>;
    DXI_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int groups = (T + SM_ROWS - 1) / SM_ROWS;
    ProfScope prof("tcn_stem", st, 1);
    kern<<<B * groups, 256, smem, st>>>(mag, net.dev_tensor(0, "kernel"), net.dev_tensor(0, "bias"), net.dev_tensor(1, "gamma"),
                                        h, T, tiles, groups);
    DXI_LAUNCHED("stem_head_kernel<stem>");
  }
  // ---- 41 tensor-core stages
  const bool split = c.precision == DXI_PREC_F16X3;
  const size_t smem = IMG_BYTES + 1024;
  DXI_CUDA(cudaFuncSetAttribute(tcn_stage_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  DXI_CUDA(cudaFuncSetAttribute(tcn_stage_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int n_tiles = B * tiles;
  int n_sm = 148;
  { int dev = 0; cudaGetDevice(&dev); cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev); }
  const int grid = n_tiles < n_sm ? n_tiles : n_sm;
  const int nd = n_dilations(c.max_d_rate);
  {
  ProfScope prof_stages("tcn_stage", st, c.n_blocks + 1);
  for (int s = 0; s <= c.n_blocks; ++s) {
    StageArgs a{};
    a.img = reinterpret_cast<const unsigned char*>(net.d_umma) + (size_t)s * IMG_BYTES;
    a.h = h;
    a.c1_in = c1[(s + 1) & 1];
    a.c1_out = c1[s & 1];
    a.T = T; a.tiles_per_utt = tiles; a.n_tiles = n_tiles; a.Ts = Ts;
    a.has_back = s >= 1; a.has_front = s < c.n_blocks;
    a.dbg = (s == g_dbg_stage) ? g_dbg_clocks : nullptr;
    const int d = s >= 1 ? 1 << ((s - 1) % nd) : 1;
    if (c.padding == DXI_PAD_CAUSAL) { a.shift0 = 2 * d; a.shift1 = d; a.shift2 = 0; }       // tap j reads t-(2-j)d
    else                             { a.shift0 = d;     a.shift1 = 0; a.shift2 = -d; }      // tap j reads t+(j-1)d
    if (split) tcn_stage_kernel<true><<<grid, STAGE_THREADS, smem, st>>>(a);
    else       tcn_stage_kernel<false><<<grid, STAGE_THREADS, smem, st>>>(a);
    DXI_LAUNCHED("tcn_stage_kernel");
  }
  }
  // ---- head (fp32 CUDA cores): tiled h -> sigmoid(W h + b)
  {
    constexpr int KIN = 256, NOUT = 257, LD = 258;
    const size_t smem2 = sizeof(float) * (SM_ROWS * LD + 16 * NOUT);
    auto kern = stem_head_kernel<KIN, NOUT, false>;
    DXI_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem2));
    const int groups = (T + SM_ROWS - 1) / SM_ROWS;
    const int li = 2 + 3 * c.n_blocks;
    ProfScope prof("tcn_head", st, 1);
    kern<<<B * groups, 256, smem2, st>>>(h, net.dev_tensor(li, "kernel"), net.dev_tensor(li, "bias"), nullptr, xbar, T, tiles, groups);
    DXI_LAUNCHED("stem_head_kernel<head>");
  }
  return DXI_OK;
}

}  // namespace dxi

// Debug aid: the epilogue of stage `stage` of the following forward calls (this thread) writes 16 clock64
// stamps per tile into dev_buf[n_tiles * 16]; pass nullptr to switch it off.
extern "C" DXI_API void dxi_debug_tcn_clocks(long long* dev_buf, int stage) {
  dxi::g_dbg_clocks = dev_buf;
  dxi::g_dbg_stage = dev_buf ? stage : -1;
}
