// Blackwell (sm_100a) primitives used by the tensor-core kernels: mbarrier, bulk async copy (TMA
// engine, 1-D), tensor memory (TMEM) allocation / load / store, tcgen05.mma with the A operand in
// TMEM or shared memory and the B operand in shared memory, tcgen05.commit.
//
// Operand conventions used throughout (all K-major, 16-bit elements, fp32 accumulate):
//   B in shared memory, 128-byte swizzle: a [N x 64] K-chunk is N rows of 128 bytes; rows are grouped
//     by 8 (1024-byte atoms, SBO = 1024); inside an atom the 16-byte unit u of row r sits at
//     r*128 + ((u ^ (r & 7)) * 16).  One tcgen05.mma consumes K = 16 elements = 32 bytes; stepping K
//     inside the 128-byte row adds 32 bytes to the descriptor start address.
//   A in tensor memory: element (m, k) lives in lane m, 32-bit column k/2, half k&1 (low half = even k),
//     so a [128 x 16] MMA operand spans 8 columns and a thread that owns row m writes it with
//     tcgen05.st.32x32b (thread i of warp w <-> lane 32*(w&3)+i).
//   D in tensor memory: element (m, n) at lane m, column n (fp32).
#pragma once
#include <cstdio>
#include "common.cuh"

namespace dxi {
namespace umma {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// ---- mbarrier ---------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  while (!mbar_try_wait(bar, parity)) {}
}
// Same, but a wait that lasts longer than ~2 s of SM clocks is a protocol error (a missed arrival, a producer that never
// ran): trap, so that the launch fails with an error instead of hanging the GPU.
constexpr long long SPIN_LIMIT_CYCLES = 4000000000ll;
constexpr uint32_t MBAR_SUSPEND_NS = 20000u;      // upper bound of one hardware-assisted sleep; the barrier event ends it earlier
// The retry loop is out of line: a kernel with ~25 waits per iteration would otherwise carry ~300 instructions of it in its hot loop,
// and these kernels are sensitive to the instruction-cache footprint of that loop.
#ifdef DXI_ENABLE_DEBUG
static __device__ int g_mbar_abort = 0;
#endif
// SLEEP = false: re-poll at once (a waiter on the critical path: the poll sees the completed phase a few cycles after it happens, but a
// polling warp takes issue slots - in tcn_chain_kernel 32 % of all executed warp instructions were such polls).
// SLEEP = true: try_wait WITH a suspend-time hint: ptxas emits TRYWAIT, NANOSLEEP.SYNCS (a sleep that the barrier's phase completion
// ends) and a re-check, so the waiting warp issues nothing; waking up costs more than a poll (measured: +11 % on the chain kernel when
// EVERY wait slept), so this is for the warps whose waits are off the critical path (weight loaders, flag agents, operand producers).
template <bool SLEEP>
static __device__ __noinline__ void mbar_wait_slow(uint32_t bar_addr, uint32_t parity) {
  const long long t0 = clock64();
  uint32_t n = 0;
  for (;;) {
    uint32_t ok;
    if (SLEEP) {
      asm volatile(
          "{\n\t.reg .pred p;\n\t"
          "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
          "selp.u32 %0, 1, 0, p;\n\t}"
          : "=r"(ok)
          : "r"(bar_addr), "r"(parity), "r"(MBAR_SUSPEND_NS)
          : "memory");
    } else {
      asm volatile(
          "{\n\t.reg .pred p;\n\t"
          "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
          "selp.u32 %0, 1, 0, p;\n\t}"
          : "=r"(ok)
          : "r"(bar_addr), "r"(parity)
          : "memory");
    }
    if (ok) return;
    if (!SLEEP && (++n & 1023u) != 0) continue;
#ifdef DXI_ENABLE_DEBUG
    // tuning build: name the wait that timed out (barrier address inside the CTA's shared window) and give up on it, so that the kernel
    // ends and the message is printed; the launch's results are then meaningless
    if (g_mbar_abort) return;
    if (clock64() - t0 > 200000000ll) {
      if ((threadIdx.x & 31) == 0) printf("mbarrier wait timed out: block %d warp %d barrier 0x%x parity %u\n", (int)blockIdx.x, (int)(threadIdx.x >> 5), bar_addr, parity);
      if (clock64() - t0 > 260000000ll) { g_mbar_abort = 1; return; }      // a little later, so that every wait stuck at the same time reports
    }
#else
    if (clock64() - t0 > SPIN_LIMIT_CYCLES) __trap();
#endif
  }
}
__device__ __forceinline__ void mbar_wait_bounded(uint64_t* bar, uint32_t parity) {
  if (!mbar_try_wait(bar, parity)) mbar_wait_slow<false>(smem_u32(bar), parity);
}
// ... for waits off the critical path: the warp sleeps until the barrier's phase completes
__device__ __forceinline__ void mbar_wait_relaxed(uint64_t* bar, uint32_t parity) {
  if (!mbar_try_wait(bar, parity)) mbar_wait_slow<true>(smem_u32(bar), parity);
}

// ---- bulk async copy global -> shared (TMA engine, no tensor map); bytes % 16 == 0 ---------------
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(smem_dst)),
               "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
// The same copy delivered to the same shared-memory offset of every CTA of the cluster named in cta_mask; each destination CTA's
// mbarrier (same offset) receives the byte count.
__device__ __forceinline__ void bulk_g2s_multicast(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar, uint16_t cta_mask) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1], %2, [%3], %4;" ::"r"(
                   smem_u32(smem_dst)),
               "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar)), "h"(cta_mask)
               : "memory");
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// ---- tensor-map TMA (cp.async.bulk.tensor): a 2-D box of a global tensor -> shared memory, swizzled by the engine as the
// tensor map says; `tmap` points at a CUtensorMap in kernel-parameter (__grid_constant__) or global memory.
__device__ __forceinline__ void tma_prefetch_desc(const void* tmap) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(tmap) : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const void* tmap, int c0, int c1, uint64_t* bar) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(
                   smem_u32(smem_dst)),
               "l"(tmap), "r"(c0), "r"(c1), "r"(smem_u32(bar))
               : "memory");
}

// ---- tensor memory ----------------------------------------------------------------------------------
// Whole-warp calls.  ncols: power of two >= 32.  The allocated base address is written to *smem_slot.
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_slot, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_slot)), "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// ---- CTA pair (cta_group::2): two CTAs of a cluster on the two SMs of a TPC run ONE MMA of M = 256: each supplies its 128 rows of A and
// its half of B's N rows from its own shared memory (same offsets in both), the leader CTA issues, each keeps its 128 rows of D.
__device__ __forceinline__ void tmem_alloc_pair(uint32_t* smem_slot, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_slot)), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_pair(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void mma_ss_pair(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrives on the mbarrier at this offset in every CTA of cta_mask once all MMAs issued so far by this thread have completed
__device__ __forceinline__ void mma_commit_pair(uint64_t* bar, uint16_t cta_mask) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(smem_u32(bar)),
               "h"(cta_mask)
               : "memory");
}
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
// arrive on the mbarrier at the same offset in CTA `rank` of the cluster
__device__ __forceinline__ void mbar_arrive_remote(uint64_t* bar, uint32_t rank) {
  asm volatile(
      "{\n\t.reg .b32 ra;\n\tmapa.shared::cluster.u32 ra, %0, %1;\n\t"
      "mbarrier.arrive.release.cluster.shared::cluster.b64 _, [ra];\n\t}" ::"r"(smem_u32(bar)),
      "r"(rank)
      : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// address of (lane, column) relative to an allocation base
__device__ __forceinline__ uint32_t tmem_addr(uint32_t base, uint32_t lane, uint32_t col) { return base + (lane << 16) + col; }

// Whole-warp load of 32 consecutive fp32 columns of this thread's lane (lane field of taddr must be
// the warp's first lane, 32*(warp&3)).
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&v)[32]) {
  uint32_t* r = reinterpret_cast<uint32_t*>(v);
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
// Whole-warp store of 32 / 16 consecutive 32-bit columns of this thread's lane.
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,%32};" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
      "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]),
      "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]),
      "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
      : "memory");
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
      "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}

__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t (&r)[8]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(taddr), "r"(r[0]), "r"(r[1]),
               "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
               : "memory");
}
// 16-lane shapes (layouts verified on B200 with scripts/tmem_layout.py; l = lane of the warp, taddr's lane field = first of the 16 lanes):
//   tcgen05.ld.16x256b.x8: 64 consecutive columns; register 4 j + e  <- (lane l / 4 + 8 (e >> 1), column 8 j + 2 (l & 3) + (e & 1)):
//     a thread holds two rows x 16 columns, and the 64 columns of a row sit in the 4 adjacent lanes l & ~3 .. | 3 (shuffle reductions);
//   tcgen05.st.16x128b.x8: 32 consecutive columns; register 2 j + g  -> (lane l / 4 + 8 g, column 4 j + (l & 3)):
//     exactly where the fp16 pair of columns (8 j + 2 (l & 3), + 1) of the load belongs in a K-major A operand.
__device__ __forceinline__ void tmem_ld_16x256b_x8(uint32_t taddr, float (&v)[32]) {
  uint32_t* r = reinterpret_cast<uint32_t*>(v);
  asm volatile(
      "tcgen05.ld.sync.aligned.16x256b.x8.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_st_16x128b_x8(uint32_t taddr, const uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.16x128b.x8.b32 [%0], "
      "{%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
      "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}

// Whole-warp load of 8 consecutive fp32 columns of this thread's lane.
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, float (&v)[8]) {
  uint32_t* r = reinterpret_cast<uint32_t*>(v);
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr)
               : "memory");
}
// Whole-warp load of 16 consecutive fp32 columns of this thread's lane.
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16]) {
  uint32_t* r = reinterpret_cast<uint32_t*>(v);
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}

// After a tcgen05.wait::ld that completes a load issued EARLIER than the statement before it (software-pipelined loads): ties the
// destination registers to this point of the program, so that the compiler cannot move their uses above the wait.
__device__ __forceinline__ void tmem_ld_fence16(float (&v)[16]) {
  uint32_t* r = reinterpret_cast<uint32_t*>(v);
  asm volatile("" : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]), "+r"(r[8]),
               "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15])::"memory");
}

// ---- descriptors ------------------------------------------------------------------------------------
// Shared-memory matrix descriptor, K-major, SWIZZLE_128B, 8-row atoms 1024 bytes apart.
__device__ __forceinline__ uint64_t make_smem_desc_sw128(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);        // start address, bits [0,14)
  d |= (uint64_t)1 << 16;                              // leading byte offset (unused with swizzle), bits [16,30)
  d |= (uint64_t)(1024 >> 4) << 32;                    // stride byte offset, bits [32,46)
  d |= (uint64_t)1 << 46;                              // descriptor version (Blackwell)
  d |= (uint64_t)2 << 61;                              // layout type: SWIZZLE_128B
  return d;
}
// Same, no swizzle ("interleave"): 8x16-byte core matrices; lbo = byte step between core matrices
// along K, sbo = byte step between 8-row groups.
__device__ __forceinline__ uint64_t make_smem_desc_noswz(uint32_t smem_addr, uint32_t lbo, uint32_t sbo) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
  d |= (uint64_t)(lbo >> 4) << 16;
  d |= (uint64_t)(sbo >> 4) << 32;
  d |= (uint64_t)1 << 46;
  return d;
}

// Instruction descriptor for kind::f16: fp16 A/B (K-major both), fp32 accumulate, shape M x N.
__host__ __device__ constexpr uint32_t make_idesc_f16(int M, int N) {
  return (1u << 4)                       // D format: F32
         | (0u << 7) | (0u << 10)        // A, B format: F16
         | (0u << 15) | (0u << 16)       // A, B K-major
         | ((uint32_t)(N >> 3) << 17)    // N / 8
         | ((uint32_t)(M >> 4) << 24);   // M / 16
}

// D[tmem_d] (+)= A[tmem_a] * B[smem desc]; issued by ONE thread.
__device__ __forceinline__ void mma_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(tmem_d),
      "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// D[tmem_d] (+)= A[smem desc] * B[smem desc]
__device__ __forceinline__ void mma_ss(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// Convergent-warp variants: executed by ALL 32 lanes of the issuing warp; one elected lane issues.  Keeping the
// election inside the asm statement keeps control flow warp-uniform, so descriptors and TMEM addresses stay in
// uniform registers (a divergent `if (lane == 0)` makes ptxas wrap every MMA in an R2UR + ELECT + BRA.U.ANY loop).
__device__ __forceinline__ void mma_ts_elect(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p, e;\n\tsetp.ne.b32 p, %4, 0;\n\telect.sync _|e, 0xffffffff;\n\t"
      "@e tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(tmem_d),
      "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// NK consecutive K = 16 steps of one 64-wide K chunk under ONE election: step i reads A at tmem_a + 8 i columns (16 fp16) and B at
// bdesc + 2 i (32 bytes further along the 128-byte swizzled row); the first step accumulates iff `accumulate`, the others always.
// One ELECT / predicate set-up per group instead of per MMA, and the derived operands are formed next to the instruction that
// uses them: the issuing warp spends ~19 SASS instructions per MMA with mma_ts_elect, and its issue time is on the critical
// path of every tile.
template <int NK>
__device__ __forceinline__ void mma_ts_elect_k(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  static_assert(NK == 2 || NK == 4, "NK");
  if (NK == 4) {
    asm volatile(
        "{\n\t.reg .pred p, e, t;\n\t.reg .b64 b1, b2, b3;\n\t.reg .b32 a1, a2, a3;\n\t"
        "setp.ne.b32 p, %4, 0;\n\tsetp.eq.b32 t, 0, 0;\n\telect.sync _|e, 0xffffffff;\n\t"
        "add.u32 a1, %1, 8;\n\tadd.u32 a2, %1, 16;\n\tadd.u32 a3, %1, 24;\n\t"
        "add.u64 b1, %2, 2;\n\tadd.u64 b2, %2, 4;\n\tadd.u64 b3, %2, 6;\n\t"
        "@e tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t"
        "@e tcgen05.mma.cta_group::1.kind::f16 [%0], [a1], b1, %3, t;\n\t"
        "@e tcgen05.mma.cta_group::1.kind::f16 [%0], [a2], b2, %3, t;\n\t"
        "@e tcgen05.mma.cta_group::1.kind::f16 [%0], [a3], b3, %3, t;\n\t}" ::"r"(tmem_d),
        "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
  } else {
    asm volatile(
        "{\n\t.reg .pred p, e, t;\n\t.reg .b64 b1;\n\t.reg .b32 a1;\n\t"
        "setp.ne.b32 p, %4, 0;\n\tsetp.eq.b32 t, 0, 0;\n\telect.sync _|e, 0xffffffff;\n\t"
        "add.u32 a1, %1, 8;\n\tadd.u64 b1, %2, 2;\n\t"
        "@e tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t"
        "@e tcgen05.mma.cta_group::1.kind::f16 [%0], [a1], b1, %3, t;\n\t}" ::"r"(tmem_d),
        "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
  }
}
// Four consecutive K = 16 steps with BOTH operands in shared memory under one election: step i reads A at adesc + i * A_STEP
// (descriptor start-address units of 16 bytes: the distance between two K16 slices of the A image) and B at bdesc + 2 i (32 bytes
// along the 128-byte swizzled row).  The first step accumulates iff `accumulate`, the others always.
template <int A_STEP>
__device__ __forceinline__ void mma_ss_elect_k4(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p, e, t;\n\t.reg .b64 a1, a2, a3, b1, b2, b3;\n\t"
      "setp.ne.b32 p, %4, 0;\n\tsetp.eq.b32 t, 0, 0;\n\telect.sync _|e, 0xffffffff;\n\t"
      "add.u64 a1, %1, %5;\n\tadd.u64 a2, %1, %6;\n\tadd.u64 a3, %1, %7;\n\t"
      "add.u64 b1, %2, 2;\n\tadd.u64 b2, %2, 4;\n\tadd.u64 b3, %2, 6;\n\t"
      "@e tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "@e tcgen05.mma.cta_group::1.kind::f16 [%0], a1, b1, %3, t;\n\t"
      "@e tcgen05.mma.cta_group::1.kind::f16 [%0], a2, b2, %3, t;\n\t"
      "@e tcgen05.mma.cta_group::1.kind::f16 [%0], a3, b3, %3, t;\n\t}" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate), "n"(A_STEP), "n"(2 * A_STEP), "n"(3 * A_STEP)
      : "memory");
}
// ---- issue path with the election hoisted and 32-bit descriptor words ---------------------------------------------------
// A kernel whose MMA warp issues many small (N = 64) MMAs is bound by that warp's instruction stream, not by the tensor pipe:
// with 64-bit descriptors formed per MMA ptxas emits ~13 SASS instructions per UTCHMMA (64-bit adds in vector registers, R2UR
// moves, a VOTEU per election).  Here the leader is elected ONCE (`e` = elect_leader(), non-zero in one lane), the high descriptor
// word is a compile-time constant and the low word is base + immediate, which ptxas keeps in uniform registers (UIADD3): 3-5
// instructions per UTCHMMA.  All lanes execute the calls convergently; only the leader issues.
__device__ __forceinline__ uint32_t elect_leader() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred e;\n\telect.sync _|e, 0xffffffff;\n\tselp.u32 %0, 1, 0, e;\n\t}" : "=r"(pred));
  return pred;
}
constexpr uint32_t DESC_HI_SW128 = (1024u >> 4) | (1u << 14) | (2u << 29);      // SBO 1024, version 1, SWIZZLE_128B
__host__ __device__ constexpr uint32_t desc_hi_noswz(uint32_t sbo) { return (sbo >> 4) | (1u << 14); }
// low descriptor word of a 128-byte-swizzled operand at shared address `a` (LBO field 1) / of a no-swizzle operand
__device__ __forceinline__ uint32_t desc_lo_sw128(uint32_t a) { return ((a & 0x3FFFFu) >> 4) | (1u << 16); }
__device__ __forceinline__ uint32_t desc_lo_noswz(uint32_t a, uint32_t lbo) { return ((a & 0x3FFFFu) >> 4) | ((lbo >> 4) << 16); }
template <uint32_t B_HI>
__device__ __forceinline__ void mma_ts_lo(uint32_t tmem_d, uint32_t tmem_a, uint32_t b_lo, uint32_t idesc, uint32_t accumulate, uint32_t e) {
  asm volatile(
      "{\n\t.reg .pred p, q;\n\t.reg .b64 bd;\n\tsetp.ne.b32 p, %4, 0;\n\tsetp.ne.b32 q, %5, 0;\n\tmov.b64 bd, {%2, %6};\n\t"
      "@q tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], bd, %3, p;\n\t}" ::"r"(tmem_d),
      "r"(tmem_a), "r"(b_lo), "r"(idesc), "r"(accumulate), "r"(e), "n"(B_HI)
      : "memory");
}
// COLL: use of the A collector buffer.  0 = default (A is fetched and discarded), 1 = fill (fetched and kept for the next MMA),
// 2 = lastuse (the A operand the previous MMA kept is reused, not fetched again): two adjacent MMAs with the same A operand read
// it from shared memory once, which is what bounds an SS-mode N = 64 MMA (4 KB of A + 2 KB of B per 32 tensor cycles).
template <uint32_t A_HI, uint32_t B_HI, int COLL = 0>
__device__ __forceinline__ void mma_ss_lo(uint32_t tmem_d, uint32_t a_lo, uint32_t b_lo, uint32_t idesc, uint32_t accumulate, uint32_t e) {
#define DXI_MMA_SS_LO(MOD)                                                                                                              \
  asm volatile(                                                                                                                         \
      "{\n\t.reg .pred p, q;\n\t.reg .b64 ad, bd;\n\tsetp.ne.b32 p, %4, 0;\n\tsetp.ne.b32 q, %5, 0;\n\t"                                 \
      "mov.b64 ad, {%1, %6};\n\tmov.b64 bd, {%2, %7};\n\t"                                                                              \
      "@q tcgen05.mma.cta_group::1.kind::f16" MOD " [%0], ad, bd, %3, p;\n\t}" ::"r"(tmem_d),                                             \
      "r"(a_lo), "r"(b_lo), "r"(idesc), "r"(accumulate), "r"(e), "n"(A_HI), "n"(B_HI)                                                   \
      : "memory")
  if (COLL == 1) DXI_MMA_SS_LO(".collector::a::fill");
  else if (COLL == 2) DXI_MMA_SS_LO(".collector::a::lastuse");
  else DXI_MMA_SS_LO("");
#undef DXI_MMA_SS_LO
}
// The CTA-pair forms of the two above (issued by the leader CTA's elected lane only)
template <uint32_t A_HI, uint32_t B_HI>
__device__ __forceinline__ void mma_ss_pair_lo(uint32_t tmem_d, uint32_t a_lo, uint32_t b_lo, uint32_t idesc, uint32_t accumulate, uint32_t e) {
  asm volatile(
      "{\n\t.reg .pred p, q;\n\t.reg .b64 ad, bd;\n\tsetp.ne.b32 p, %4, 0;\n\tsetp.ne.b32 q, %5, 0;\n\t"
      "mov.b64 ad, {%1, %6};\n\tmov.b64 bd, {%2, %7};\n\t"
      "@q tcgen05.mma.cta_group::2.kind::f16 [%0], ad, bd, %3, p;\n\t}" ::"r"(tmem_d),
      "r"(a_lo), "r"(b_lo), "r"(idesc), "r"(accumulate), "r"(e), "n"(A_HI), "n"(B_HI)
      : "memory");
}
__device__ __forceinline__ void mma_commit_pair_lo(uint64_t* bar, uint16_t cta_mask, uint32_t e) {
  asm volatile(
      "{\n\t.reg .pred q;\n\tsetp.ne.b32 q, %2, 0;\n\t"
      "@q tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;\n\t}" ::"r"(smem_u32(bar)),
      "h"(cta_mask), "r"(e)
      : "memory");
}
__device__ __forceinline__ void mma_commit_lo(uint64_t* bar, uint32_t e) {
  asm volatile(
      "{\n\t.reg .pred q;\n\tsetp.ne.b32 q, %1, 0;\n\t"
      "@q tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t}" ::"r"(smem_u32(bar)), "r"(e)
      : "memory");
}
// ... and arrives on the mbarrier at the same offset in every CTA of cta_mask (operands delivered by multicast are released cluster-wide)
__device__ __forceinline__ void mma_commit_multicast_lo(uint64_t* bar, uint16_t cta_mask, uint32_t e) {
  asm volatile(
      "{\n\t.reg .pred q;\n\tsetp.ne.b32 q, %2, 0;\n\t"
      "@q tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;\n\t}" ::"r"(smem_u32(bar)),
      "h"(cta_mask), "r"(e)
      : "memory");
}
__device__ __forceinline__ void mma_ss_elect(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p, e;\n\tsetp.ne.b32 p, %4, 0;\n\telect.sync _|e, 0xffffffff;\n\t"
      "@e tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void mma_commit_elect(uint64_t* bar) {
  asm volatile(
      "{\n\t.reg .pred e;\n\telect.sync _|e, 0xffffffff;\n\t"
      "@e tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t}" ::"r"(smem_u32(bar))
      : "memory");
}
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred e;\n\telect.sync _|e, 0xffffffff;\n\tselp.u32 %0, 1, 0, e;\n\t}" : "=r"(pred));
  return pred != 0;
}

// Makes the mbarrier track completion of all tcgen05 operations issued so far by this thread.
__device__ __forceinline__ void mma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// fp32 -> packed fp16 pair (lo = a, hi = b), and the residuals of that rounding.
__device__ __forceinline__ uint32_t pack_h2(float a, float b) {
  __half2 h = __floats2half2_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}
__device__ __forceinline__ void split_h2(float a, float b, uint32_t& hi, uint32_t& lo) {
  __half2 h = __floats2half2_rn(a, b);
  float2 f = __half22float2(h);
  __half2 l = __floats2half2_rn(a - f.x, b - f.y);
  hi = *reinterpret_cast<uint32_t*>(&h);
  lo = *reinterpret_cast<uint32_t*>(&l);
}

// hi/lo split of a float2 with few conversion-pipe (XU) instructions: hi = a truncated to an 11-bit significand by a
// bit mask (ALU), so its fp16 conversion is exact, lo = a - hi (one packed FADD2, exact) rounded to fp16: hi + lo
// carries >= 21 significant bits.  Two F2FP per pair instead of two F2FP + two fp16->fp32 conversions.
__device__ __forceinline__ void split_h2x(float2 a, uint32_t& hi, uint32_t& lo) {
  const float2 h = make_float2(__uint_as_float(__float_as_uint(a.x) & 0xFFFFE000u), __uint_as_float(__float_as_uint(a.y) & 0xFFFFE000u));
  const float2 l = __fadd2_rn(a, make_float2(-h.x, -h.y));
  __half2 hh = __float22half2_rn(h), ll = __float22half2_rn(l);
  hi = *reinterpret_cast<uint32_t*>(&hh);
  lo = *reinterpret_cast<uint32_t*>(&ll);
}
template <bool SPLIT>
__device__ __forceinline__ void to_h2(float a, float b, uint32_t& hi, uint32_t& lo) {
  if (SPLIT) split_h2x(make_float2(a, b), hi, lo);
  else { hi = pack_h2(a, b); lo = 0; }       // single-product mode: round to nearest
}

}  // namespace umma
}  // namespace dxi
