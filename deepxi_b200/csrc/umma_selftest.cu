// Self test of the tcgen05 building blocks in umma.cuh: one CTA computes D[128,N] = A[128,K] B[N,K]^T
// (fp16 operands, fp32 accumulate) and writes D to global memory; tests/test_umma_selftest.py compares
// with numpy.  `variant` selects the encodings under test:
//   bit0  A operand: 0 = tensor memory (the production path), 1 = shared memory (128B swizzle)
//   bit1  B operand layout: 0 = 128-byte swizzle (production), 1 = no swizzle (8x16B core matrices)
//   bit2  A-in-TMEM half order: 0 = even k in the low half (production), 1 = swapped
//   bit3  A operand in shared memory WITHOUT swizzle as [unit of 8 k][32 zero rows + 128 rows][16 B] and a descriptor whose start
//         address is shifted SHIFT = 5 rows back: D[m] = A[m - 5] B^T with zero rows before the first (the c1 taps of tcn_chain.cu)
//   bit4  B operand written by tensor-map TMA (cp.async.bulk.tensor.2d, 128-byte swizzle applied by the engine) instead of by threads
#include <cuda.h>
#include "umma.cuh"

namespace dxi {
using namespace umma;

constexpr int ST_SHIFT = 5, ST_ROWS = 160;

__global__ void __launch_bounds__(128) umma_selftest_kernel(const __grid_constant__ CUtensorMap tm_b, const __half* __restrict__ A,
                                                            const __half* __restrict__ Bm, int N, int K, int variant, float* __restrict__ D) {
  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  __shared__ __align__(8) uint64_t bar, bar_tma;
  __shared__ uint32_t tmem_base_slot;
  const int tid = threadIdx.x, warp = tid >> 5;
  const bool a_smem = variant & 1, b_noswz = variant & 2, a_swap = variant & 4, a_taps = variant & 8, b_tma = variant & 16;
  unsigned char* sB = smem;                               // K/64 chunks of [N x 128 B]  (or no-swizzle image)
  unsigned char* sA = smem + (size_t)N * K * 2;           // K/64 chunks of [128 x 128 B]

  if (warp == 0) tmem_alloc(&tmem_base_slot, 512);
  if (tid == 0) { mbar_init(&bar, 1); mbar_init(&bar_tma, 1); fence_mbar_init(); }
  __syncthreads();
  // ---- B image
  if (b_tma) {
    if (tid == 0) {
      mbar_arrive_expect_tx(&bar_tma, (uint32_t)(N * K * 2));
      for (int c = 0; c < K / 64; ++c) tma_load_2d(sB + (size_t)c * N * 128, &tm_b, c * 64, 0, &bar_tma);
    }
    mbar_wait_bounded(&bar_tma, 0);
  } else
  for (int i = tid; i < N * (K / 8); i += 128) {          // one 16-byte unit (8 halves) per step
    const int n = i / (K / 8), u = i - n * (K / 8);       // unit u covers k = 8u .. 8u+7
    uint4 val = *reinterpret_cast<const uint4*>(Bm + (size_t)n * K + 8 * u);
    size_t off;
    if (!b_noswz) {
      const int c = u >> 3, uu = u & 7;
      off = (size_t)c * N * 128 + (size_t)(n >> 3) * 1024 + (n & 7) * 128 + ((uu ^ (n & 7)) * 16);
    } else {
      off = (size_t)(n >> 3) * (K / 8) * 128 + (size_t)u * 128 + (n & 7) * 16;
    }
    *reinterpret_cast<uint4*>(sB + off) = val;
  }
  if (a_taps) {
    for (int i = tid; i < ST_ROWS * (K / 8); i += 128) {
      const int u = i / ST_ROWS, r = i - u * ST_ROWS;
      uint4 val = make_uint4(0, 0, 0, 0);
      if (r >= 32) val = *reinterpret_cast<const uint4*>(A + (size_t)(r - 32) * K + 8 * u);
      *reinterpret_cast<uint4*>(sA + ((size_t)u * ST_ROWS + r) * 16) = val;
    }
  } else if (a_smem) {
    for (int i = tid; i < 128 * (K / 8); i += 128) {
      const int m = i / (K / 8), u = i - m * (K / 8);
      uint4 val = *reinterpret_cast<const uint4*>(A + (size_t)m * K + 8 * u);
      const int c = u >> 3, uu = u & 7;
      size_t off = (size_t)c * 128 * 128 + (size_t)(m >> 3) * 1024 + (m & 7) * 128 + ((uu ^ (m & 7)) * 16);
      *reinterpret_cast<uint4*>(sA + off) = val;
    }
  }
  fence_proxy_async();            // generic-proxy smem writes -> visible to the tensor core (async proxy)
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tbase = tmem_base_slot;
  const uint32_t lane_base = (uint32_t)(warp * 32);
  const uint32_t a_col = 256;
  if (!a_smem && !a_taps) {
    // thread tid owns row tid: pack its K halves into K/2 columns starting at a_col
    for (int c0 = 0; c0 < K / 2; c0 += 16) {
      uint32_t r[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        const __half lo = A[(size_t)tid * K + 2 * (c0 + j)], hi = A[(size_t)tid * K + 2 * (c0 + j) + 1];
        const uint32_t l = __half_as_ushort(lo), h = __half_as_ushort(hi);
        r[j] = a_swap ? (h | (l << 16)) : (l | (h << 16));
      }
      tmem_st16(tmem_addr(tbase, lane_base, a_col + c0), r);
    }
    tmem_wait_st();
  }
  tc_fence_before();
  __syncthreads();
  if (tid == 0) {
    tc_fence_after();
    const uint32_t idesc = make_idesc_f16(128, N);
    for (int k16 = 0; k16 < K / 16; ++k16) {
      const int c = k16 >> 2, kk = k16 & 3;
      uint64_t bdesc;
      if (!b_noswz) bdesc = make_smem_desc_sw128(smem_u32(sB + (size_t)c * N * 128) + kk * 32);
      else          bdesc = make_smem_desc_noswz(smem_u32(sB) + k16 * 256, 128, (K / 8) * 128);
      if (a_taps) {
        uint64_t adesc = make_smem_desc_noswz(smem_u32(sA) + (2 * k16) * ST_ROWS * 16 + (32 - ST_SHIFT) * 16, ST_ROWS * 16, 128);
        mma_ss(tbase, adesc, bdesc, idesc, k16 > 0);
      } else if (!a_smem) {
        mma_ts(tbase, tmem_addr(tbase, 0, a_col + 8 * k16), bdesc, idesc, k16 > 0);
      } else {
        uint64_t adesc = make_smem_desc_sw128(smem_u32(sA + (size_t)c * 128 * 128) + kk * 32);
        mma_ss(tbase, adesc, bdesc, idesc, k16 > 0);
      }
    }
    mma_commit(&bar);
  }
  mbar_wait(&bar, 0);
  tc_fence_after();
  for (int n0 = 0; n0 < N; n0 += 32) {
    float v[32];
    tmem_ld32(tmem_addr(tbase, lane_base, n0), v);
    tmem_wait_ld();
#pragma unroll
    for (int j = 0; j < 32; ++j) D[(size_t)tid * N + n0 + j] = v[j];
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tbase, 512);
}


// CTA-pair self test: D[256, 256] = A[256, K] B[256, K]^T by ONE tcgen05.mma.cta_group::2 sequence.  CTA r of the cluster holds rows
// 128 r .. of A and rows (n) 128 r .. of B in its own shared memory (128-byte swizzle), the leader issues, each CTA reads its 128 rows.
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128) umma_pair_selftest_kernel(const __half* __restrict__ A, const __half* __restrict__ Bm,
                                                                                          int K, float* __restrict__ D) {
  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tmem_base_slot;
  const int tid = threadIdx.x, warp = tid >> 5;
  const uint32_t rank = cluster_ctarank();
  unsigned char* sB = smem;                               // K/64 chunks of [128 x 128 B]
  unsigned char* sA = smem + (size_t)128 * K * 2;         // K/64 chunks of [128 x 128 B]
  if (warp == 0) tmem_alloc_pair(&tmem_base_slot, 256);
  if (tid == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
  for (int i = tid; i < 128 * (K / 8); i += 128) {
    const int m = i / (K / 8), u = i - m * (K / 8), c = u >> 3, uu = u & 7;
    const size_t off = (size_t)c * 128 * 128 + (size_t)(m >> 3) * 1024 + (m & 7) * 128 + ((uu ^ (m & 7)) * 16);
    *reinterpret_cast<uint4*>(sA + off) = *reinterpret_cast<const uint4*>(A + (size_t)(128 * rank + m) * K + 8 * u);
    *reinterpret_cast<uint4*>(sB + off) = *reinterpret_cast<const uint4*>(Bm + (size_t)(128 * rank + m) * K + 8 * u);
  }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();            // both CTAs' operands and barriers are in place
  tc_fence_after();
  const uint32_t tbase = tmem_base_slot;
  if (rank == 0 && tid == 0) {
    const uint32_t idesc = make_idesc_f16(256, 256);
    for (int k16 = 0; k16 < K / 16; ++k16) {
      const int c = k16 >> 2, kk = k16 & 3;
      mma_ss_pair(tbase, make_smem_desc_sw128(smem_u32(sA + (size_t)c * 128 * 128) + kk * 32),
                  make_smem_desc_sw128(smem_u32(sB + (size_t)c * 128 * 128) + kk * 32), idesc, k16 > 0);
    }
    mma_commit_pair(&bar, 3);
  }
  mbar_wait_bounded(&bar, 0);
  tc_fence_after();
  for (int n0 = 0; n0 < 256; n0 += 32) {
    float v[32];
    tmem_ld32(tmem_addr(tbase, (uint32_t)(warp * 32), n0), v);
    tmem_wait_ld();
#pragma unroll
    for (int j = 0; j < 32; ++j) D[(size_t)(128 * rank + tid) * 256 + n0 + j] = v[j];
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (warp == 0) tmem_dealloc_pair(tbase, 256);
}

#ifdef DXI_ENABLE_DEBUG
// TMEM port micro-benchmark: `warps` warps (multiple of 4) each run `rounds` x (tcgen05.ld|st .32x32b.x32) on their
// lane quarter; out[0] = cycles for the whole CTA (max over warps).  mode 0 = loads, 1 = stores, 2 = load + store.
__global__ void __launch_bounds__(1024) tmem_bw_kernel(int mode, int rounds, long long* out) {
  __shared__ uint32_t slot;
  __shared__ long long t_start, t_end;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) tmem_alloc(&slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t lane_addr = (uint32_t)((warp & 3) * 32) << 16;
  const uint32_t col0 = (uint32_t)((warp >> 2) * 32) & 511;
  float v[32];
#pragma unroll
  for (int j = 0; j < 32; ++j) v[j] = (float)j;
  __syncthreads();
  const long long t0 = clock64();
  for (int r = 0; r < rounds; ++r) {
    const uint32_t col = (col0 + 128 * (r & 3)) & 511;
    if (mode == 0 || mode == 2) { tmem_ld32(lane_addr + col, v); }
    if (mode == 1 || mode == 2) { tmem_st32(lane_addr + ((col + 64) & 511), reinterpret_cast<const uint32_t(&)[32]>(v)); }
  }
  tmem_wait_ld(); tmem_wait_st();
  const long long t1 = clock64();
  if (threadIdx.x == 0) { t_start = t0; t_end = t1; }
  __syncthreads();
  if ((threadIdx.x & 31) == 0) { atomicMin(&t_start, t0); atomicMax(&t_end, t1); }
  __syncthreads();
  float acc = 0.0f;
#pragma unroll
  for (int j = 0; j < 32; ++j) acc += v[j];
  if (threadIdx.x == 0) { out[0] = t_end - t_start; out[1] = (long long)acc; }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(slot, 512);
}

#endif

}  // namespace dxi

using namespace dxi;

namespace dxi { int make_weight_map(void* out, const void* dev, int K, size_t rows, int box_rows); }
static int selftest_weight_map(CUtensorMap* tm, const void* b, int K, int N) { return dxi::make_weight_map(tm, b, K, (size_t)N, N); }

#ifdef DXI_ENABLE_DEBUG
// Layout discovery for the 16-lane TMEM access shapes (tuning build).  Phase 1: TMEM[lane][col] = lane * 256 + col for 64 columns is
// written with the 32x32b shape (thread = lane); every warp then reads its two 16-lane halves with 16x256b.x8 and dumps its 32 + 32
// registers: out[tid][0..63].  Phase 2: every thread writes (tid << 8 | register index) with 16x128b.x8 (16 registers per half) into
// columns 64..95; the 32x32b read-back of those columns is dumped to out[128 * 64 + lane * 32 + col].
namespace dxi {
__global__ void __launch_bounds__(128) tmem_layout_kernel(uint32_t* out) {
  __shared__ uint32_t slot;
  const int tid = threadIdx.x, warp = tid >> 5;
  if (warp == 0) tmem_alloc(&slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t base = slot + ((uint32_t)(warp * 32) << 16);
  {
    uint32_t v[32];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
#pragma unroll
      for (int c = 0; c < 32; ++c) v[c] = (uint32_t)tid * 256u + 32 * h + c;
      tmem_st32(base + 32 * h, v);
    }
    tmem_wait_st();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
#pragma unroll
  for (int half = 0; half < 2; ++half) {
    uint32_t r[32];
    asm volatile(
        "tcgen05.ld.sync.aligned.16x256b.x8.b32 "
        "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]),
          "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]),
          "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(base + ((uint32_t)(16 * half) << 16))
        : "memory");
    tmem_wait_ld();
#pragma unroll
    for (int i = 0; i < 32; ++i) out[(size_t)tid * 64 + 32 * half + i] = r[i];
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
#pragma unroll
  for (int half = 0; half < 2; ++half) {
    uint32_t r[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) r[i] = ((uint32_t)tid << 8) | (uint32_t)(16 * half + i);
    asm volatile(
        "tcgen05.st.sync.aligned.16x128b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};" ::"r"(
            base + ((uint32_t)(16 * half) << 16) + 64),
        "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]),
        "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
        : "memory");
  }
  tmem_wait_st();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  {
    float v[32];
    tmem_ld32(base + 64, v); tmem_wait_ld();
#pragma unroll
    for (int c = 0; c < 32; ++c) out[(size_t)128 * 64 + (size_t)tid * 32 + c] = __float_as_uint(v[c]);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(slot, 512);
}
}  // namespace dxi

extern "C" DXI_API int dxi_debug_tmem_layout(uint32_t* dev_out, void* stream) {
  if (int rc = check_device()) return rc;
  dxi::tmem_layout_kernel<<<1, 128, 0, as_stream(stream)>>>(dev_out);
  DXI_LAUNCHED("tmem_layout_kernel");
  return DXI_OK;
}

extern "C" DXI_API int dxi_debug_tmem_bw(int mode, int warps, int rounds, long long* dev_out, void* stream) {
  if (int rc = check_device()) return rc;
  DXI_REQUIRE(warps >= 4 && warps <= 32 && warps % 4 == 0 && rounds > 0 && dev_out, "dxi_debug_tmem_bw: bad argument");
  tmem_bw_kernel<<<1, warps * 32, 0, as_stream(stream)>>>(mode, rounds, dev_out);
  DXI_LAUNCHED("tmem_bw_kernel");
  return DXI_OK;
}
#endif

extern "C" DXI_API int dxi_selftest_umma_pair(const void* a_f16, const void* b_f16, int K, float* d_out, void* stream) {
  if (int rc = check_device()) return rc;
  DXI_REQUIRE(a_f16 && b_f16 && d_out, "dxi_selftest_umma_pair: null argument");
  DXI_REQUIRE(K >= 64 && K <= 256 && K % 64 == 0, "dxi_selftest_umma_pair: K must be a multiple of 64 in [64,256]");
  const size_t smem = (size_t)2 * 128 * K * 2 + 2048;
  DXI_CUDA(cudaFuncSetAttribute(umma_pair_selftest_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  umma_pair_selftest_kernel<<<2, 128, smem, as_stream(stream)>>>(reinterpret_cast<const __half*>(a_f16), reinterpret_cast<const __half*>(b_f16), K, d_out);
  DXI_LAUNCHED("umma_pair_selftest_kernel");
  return DXI_OK;
}

extern "C" DXI_API int dxi_selftest_umma(const void* a_f16, const void* b_f16, int N, int K, int variant, float* d_out,
                                 void* stream) {
  if (int rc = check_device()) return rc;
  DXI_REQUIRE(a_f16 && b_f16 && d_out, "dxi_selftest_umma: null argument");
  DXI_REQUIRE(N >= 32 && N <= 256 && N % 32 == 0, "dxi_selftest_umma: N must be a multiple of 32 in [32,256]");
  DXI_REQUIRE(K >= 64 && K <= 256 && K % 64 == 0, "dxi_selftest_umma: K must be a multiple of 64 in [64,256]");
  const size_t smem = (size_t)N * K * 2 + ST_ROWS * (size_t)K * 2 + 2048;
  DXI_CUDA(cudaFuncSetAttribute(umma_selftest_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  CUtensorMap tm{};
  if (variant & 16) {
    DXI_REQUIRE(!(variant & 2), "dxi_selftest_umma: the TMA-loaded B operand is the 128-byte swizzled one");
    if (int rc = selftest_weight_map(&tm, b_f16, K, N)) return rc;
  }
  umma_selftest_kernel<<<1, 128, smem, as_stream(stream)>>>(tm, reinterpret_cast<const __half*>(a_f16),
                                                            reinterpret_cast<const __half*>(b_f16), N, K, variant, d_out);
  DXI_LAUNCHED("umma_selftest_kernel");
  return DXI_OK;
}
