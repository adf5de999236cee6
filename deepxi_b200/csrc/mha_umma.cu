// MHANetV3 linear layers on the 5th-generation tensor cores (tcgen05 + TMEM): the four GEMMs of every block
// (fused QKV projection, output projection + residual + LayerNorm, feed-forward in / out; attention.py:327-353, :88-101)
// in DXI_PREC_F16X3: both operands split into fp16 hi + lo, three MMAs per product, fp32 accumulate in tensor memory.
//
//   out[M, N] = epilogue( A[M, K] (fp32, row-major) x W[K, N] )        K in {256, 1024}, N in {256, 768, 1024}
//
// One persistent CTA per SM walks (row tile of 128) x (column tile of 256) output tiles, n fastest, so that the column
// tiles of one row tile run on neighbouring SMs at the same time and share A through L2.  Warp roles:
//   warps 0-7  producers, two groups of four: group p owns ring slot p and produces the chunks of parity p, so the global
//              loads of one chunk are in flight while the other chunk is converted and multiplied.  A [128 x 64] chunk of A per
//              step: coalesced float4 loads, hi/lo split, stores into the 128-byte-swizzled K-major shared-memory operand (the
//              layout umma_selftest checks); the group's first lane also starts the bulk copy of the matching pre-packed
//              weight chunk (hi | lo, 64 KB);
//   warp 12    issues the 12 MMAs of a chunk (M 128 x N 256 x K 16, (a_hi, w_hi) + (a_lo, w_hi) + (a_hi, w_lo)) as soon as
//              both operands have landed, releases the ring slot with tcgen05.commit;
//   warps 8-11 epilogue: thread = output row, 256 fp32 columns from TMEM in chunks of 32: bias / ReLU, or
//              residual + LayerNorm (row statistics by Chan's merge over the 8 chunks, normalised in a second pass
//              over the tile kept in TMEM); two accumulator buffers, so the epilogue of a tile overlaps the next tile's MMAs.
// Shared memory: 2 ring slots x (A hi|lo 32 KB + W hi|lo 64 KB) = 192 KB + bias / gamma / beta of the column tile.
// The kernel is bound by L2 -> SM bytes (96 KB per 1536 tensor cycles and SM, two thirds of it the weight chunk every row tile re-reads),
// so CS CTAs form a thread-block cluster that walks CS row tiles of the SAME column tile in lock step: each CTA fetches 1 / CS of the
// weight chunk and multicasts it to all of them (a ring slot is re-filled when the MMAs of ALL CS CTAs on it have committed: the commit
// arrives on every CTA's barrier).
// PAIR (CS = 2): the two CTAs run ONE M = 256 product per step (tcgen05.mma.cta_group::2): each converts its own 128 rows of A and fetches
// only ITS half of the weight chunk (128 of the 256 output columns), the leader CTA issues the MMAs, each keeps its 128 rows of the
// accumulator.  64 KB instead of 96 KB enter an SM per 1536 tensor cycles - the kernel is bound by exactly that - and the ring holds
// three slots.  The peer's MMA warp only forwards "my operands of slot s have landed" to the leader's barrier.
#include <vector>
#include "net.cuh"
#include "umma.cuh"

namespace dxi {
using namespace umma;

constexpr int LM = 128, LNT = 256, LK = 64;
constexpr int LA_PART = LM * 128;                     // one precision part of an A chunk: [128 rows x 128 B]
constexpr int LW_PART = LNT * 128;                    // one precision part of a W chunk: [256 rows x 128 B]
constexpr int L_SLOT = 2 * LA_PART + 2 * LW_PART;     // 96 KB
constexpr int L_WCHUNK = 2 * LW_PART;                 // packed weight chunk in global memory: hi | lo
constexpr int L_PROD_WARPS = 8;                       // two groups of 4: group p fills ring slot p (chunks of parity p)
constexpr int L_THREADS = (L_PROD_WARPS + 5) * 32;    // + 4 epilogue warps + 1 MMA warp
constexpr int L_STAGE_LD = 36;                        // per-warp [32 rows][32 + 4] fp32 transposition stage: float4 accesses by row and by column group are conflict free
constexpr int L_SMEM = 1024 + 2 * L_SLOT + 3 * LNT * 4 + LM * L_STAGE_LD * 4;
enum { LEPI_PLAIN = 0, LEPI_BIAS_RELU = 1, LEPI_RES_LN = 2, LEPI_SIGMOID = 3, LEPI_QKV = 4 };
constexpr int L_KV_PART = 128 * 32 * 2, L_KV_TILE = 4 * L_KV_PART;      // the attention kernel's key-tile image (attn_umma.cu): K hi | K lo | V^T hi | V^T lo

struct LinArgs {
  const float* A; int lda;
  const unsigned char* W;      // packed: [column tile][K chunk][hi 32 KB | lo 32 KB]
  const float* bias;           // [N] or null
  const float* res;            // [M][N] (LEPI_RES_LN; N == 256)
  const float* gamma; const float* beta;
  const float* pos; int T;     // LEPI_RES_LN without residual: ReLU, then + pos[m % T][:] (input layer, attention.py:428-433)
  float* out; int ldo;
  int M, N, K, epi;            // N, K: padded to multiples of 256 / 64 (the packed weights carry the zeros)
  int Nr, Kr;                  // real sizes (Kr = 257 for the input layer, Nr = 257 for the output layer)
  // LEPI_QKV (fused QKV projection, N = 768 = Q | K | V): the row space is padded per utterance to whole 128-row tiles, row tile mt =
  // (utterance mt / n_kt, key tile mt % n_kt) reads rows b T + 128 j .. of A (rows at or beyond T: zeros).  Column tile 0 (Q) goes to
  // out as fp32; tiles 1, 2 (K, V) leave as the attention kernel's operand images, one head per 32-column chunk:
  // kv[((b n_heads + h) n_kt + j)][L_KV_TILE] - fp16 hi | lo, K-major core matrices, V transposed - zero padded for free.
  unsigned char* kv; int n_kt, n_heads;
};

template <int CS, bool PAIR = false>
__global__ void __launch_bounds__(L_THREADS, 1) lin_umma_kernel(const LinArgs g) {      // 13 warps are allocated as 16: 128 registers is the cap
  static_assert(!PAIR || CS == 2, "a CTA pair is a cluster of two");
  constexpr int NSLOT = PAIR ? 3 : 2;                              // ring slots
  constexpr int SLOT = PAIR ? 2 * LA_PART + LW_PART : L_SLOT;      // A hi | A lo | W hi | W lo (PAIR: this CTA's 128 of the 256 weight rows)
  constexpr int W_LO = PAIR ? LW_PART / 2 : LW_PART;               // bytes from the hi part of the slot's weights to the lo part
  extern __shared__ unsigned char smem_raw[];
  __shared__ __align__(8) uint64_t full_a[3], full_w[3], empty[3], peer_full[3], acc_full[2], acc_empty[2];
  __shared__ uint32_t tmem_slot;
  unsigned char* ring = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  float* sVec = reinterpret_cast<float*>(ring + 2 * L_SLOT);      // bias[256], gamma[256], beta[256] of the column tile
  float* sStage = sVec + 3 * LNT;                                 // [4 warps][32][36]: thread-owns-row <-> warp-writes-row re-ordering of a 32-column chunk
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (warp == L_PROD_WARPS + 4) { if (PAIR) tmem_alloc_pair(&tmem_slot, 512); else tmem_alloc(&tmem_slot, 512); }
  if (tid == 0) {
    for (int i = 0; i < 3; ++i) { mbar_init(&full_a[i], 4); mbar_init(&full_w[i], 1); mbar_init(&empty[i], PAIR ? 1 : CS); mbar_init(&peer_full[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&acc_full[i], 1); mbar_init(&acc_empty[i], PAIR ? 8 : 4); }
    fence_mbar_init();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (tmem_slot != 0) __trap();
  if (CS > 1) cluster_sync_all();      // every CTA's barriers exist before a peer's copy or commit can reach them

  // The cluster walks (group of CS row tiles) x (column tile), n fastest; CTA `rank` takes row tile CS * group + rank (beyond the
  // matrix: a tile of zeros that only keeps the multicast protocol in step).
  const int n_nt = g.N / LNT, n_mt = (g.M + LM - 1) / LM, n_kc = g.K / LK;
  const int rank = blockIdx.x % CS, cid = blockIdx.x / CS, n_cl = gridDim.x / CS;
  const int n_tiles = (n_mt + CS - 1) / CS * n_nt;      // per cluster: "tile" t = (row group t / n_nt, column tile t % n_nt)
  constexpr uint16_t cl_mask = (uint16_t)((1u << CS) - 1);
  // first source / destination row of row tile mt and how many of its 128 rows exist
  auto tile_rows = [&](int mt, int& src0, int& nvalid) {
    if (g.epi == LEPI_QKV) {
      const int b = mt / g.n_kt, j = mt - b * g.n_kt;
      src0 = b * g.T + j * LM;
      nvalid = mt < n_mt ? min(LM, g.T - j * LM) : 0;
    } else {
      src0 = mt * LM;
      nvalid = min(LM, g.M - mt * LM);
    }
  };

  if (warp < L_PROD_WARPS) {
    // ================= producers =================
    // Group grp owns ring slot grp and the chunks of parity grp in this CTA's chunk sequence.  The fp32 rows of a group's NEXT chunk are
    // loaded into registers as soon as the current one has been stored: those loads need no ring slot, so their latency (the
    // kernel's bound before) runs under the MMAs of the current chunk and the other group's work.
    const int grp = warp >> 2, ptid = tid & 127;                  // producer group = the ring slot it owns
    const int c4 = ptid & 15, r0 = ptid >> 4;                     // float4 column, first row; rows r0 + 8 j
    int t = cid, kc = 0, gch = 0;                                 // this group's current chunk: (tile, K chunk), position in the CTA's sequence
    auto advance = [&]() { if (++kc == n_kc) { kc = 0; t += n_cl; } ++gch; };
    auto seek = [&]() { while (t < n_tiles && (gch & 1) != grp) advance(); };
    float4 v[16];
    auto load = [&]() {
      if (t >= n_tiles) return;
      const int mt = (t / n_nt) * CS + rank;
      int m0, nvalid;
      tile_rows(mt, m0, nvalid);
      const float* ap = g.A + (size_t)m0 * g.lda + kc * LK + 4 * c4;
      const bool vec = (g.lda & 3) == 0 && (kc + 1) * LK <= g.Kr;      // else: rows not 16-byte aligned / ragged K (input layer)
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        const int r = r0 + 8 * j;
        if (vec) {
          v[j] = (r < nvalid) ? __ldg(reinterpret_cast<const float4*>(ap + (size_t)r * g.lda)) : make_float4(0.f, 0.f, 0.f, 0.f);
        } else {
          const float* q = ap + (size_t)r * g.lda;
          const int k = kc * LK + 4 * c4;
          const bool in = r < nvalid;
          v[j].x = (in && k < g.Kr) ? __ldg(q) : 0.f;         v[j].y = (in && k + 1 < g.Kr) ? __ldg(q + 1) : 0.f;
          v[j].z = (in && k + 2 < g.Kr) ? __ldg(q + 2) : 0.f; v[j].w = (in && k + 3 < g.Kr) ? __ldg(q + 3) : 0.f;
        }
      }
    };
    seek();
    load();
    while (t < n_tiles) {
      const int slot = PAIR ? gch % NSLOT : grp, use = PAIR ? gch / NSLOT : gch >> 1;
      const int nt = t % n_nt;
      if (use >= 1) mbar_wait_relaxed(&empty[slot], (use - 1) & 1);       // the MMAs of every CTA of the cluster that read this slot have completed
      unsigned char* sA = ring + slot * SLOT;
      if (ptid == 0) {
        mbar_arrive_expect_tx(&full_w[slot], PAIR ? LW_PART : L_WCHUNK);
        const unsigned char* src = g.W + ((size_t)nt * n_kc + kc) * L_WCHUNK;
        if (PAIR) {      // this CTA's 128 of the chunk's 256 weight rows: the second / first half of the hi part and of the lo part
          bulk_g2s(sA + 2 * LA_PART, src + rank * (LW_PART / 2), LW_PART / 2, &full_w[slot]);
          bulk_g2s(sA + 2 * LA_PART + LW_PART / 2, src + LW_PART + rank * (LW_PART / 2), LW_PART / 2, &full_w[slot]);
        } else if (CS == 1) {
          for (int off = 0; off < L_WCHUNK; off += 16384) bulk_g2s(sA + 2 * LA_PART + off, src + off, 16384, &full_w[slot]);
        } else {      // this CTA's share of the chunk, to every CTA of the cluster
          constexpr int share = L_WCHUNK / CS;
          for (int off = rank * share; off < (rank + 1) * share; off += 16384)
            bulk_g2s_multicast(sA + 2 * LA_PART + off, src + off, 16384, &full_w[slot], cl_mask);
        }
      }
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        const int r = r0 + 8 * j;
        uint32_t h0, l0, h1, l1;
        split_h2x(make_float2(v[j].x, v[j].y), h0, l0);      // hi by truncation (a mask), lo = the exact remainder: 5 instructions per pair, not 7 -
        split_h2x(make_float2(v[j].z, v[j].w), h1, l1);      // this conversion, repeated for every column tile, is what bounds the kernel
        const uint32_t off = (uint32_t)(r >> 3) * 1024 + (r & 7) * 128 + ((((uint32_t)c4 >> 1) ^ (r & 7)) << 4) + (c4 & 1) * 8;
        *reinterpret_cast<uint2*>(sA + off) = make_uint2(h0, h1);
        *reinterpret_cast<uint2*>(sA + LA_PART + off) = make_uint2(l0, l1);
      }
      fence_proxy_async();                                       // generic-proxy stores -> visible to the tensor core
      __syncwarp();
      if (lane == 0) mbar_arrive(&full_a[slot]);
      advance(); seek();
      load();
    }
  } else if (warp == L_PROD_WARPS + 4) {
    // ================= MMA issue =================
    constexpr uint32_t idesc = make_idesc_f16(PAIR ? 2 * LM : LM, LNT);
    const uint32_t e = elect_leader();      // issue path as in tcn_chain.cu: one election, 32-bit descriptor words in uniform registers
    const uint32_t ring_a = smem_u32(ring);
    int gch = 0, lt = 0;
    if (PAIR && rank != 0) {
      // the peer of a CTA pair issues nothing: it tells the leader when its own operands of a slot have landed
      for (int t = cid; t < n_tiles; t += n_cl)
        for (int kc = 0; kc < n_kc; ++kc, ++gch) {
          const int slot = gch % NSLOT, use = gch / NSLOT;
          mbar_wait_relaxed(&full_a[slot], use & 1);
          mbar_wait_relaxed(&full_w[slot], use & 1);
          __syncwarp();
          if (lane == 0) mbar_arrive_remote(&peer_full[slot], 0);
        }
    } else
    for (int t = cid; t < n_tiles; t += n_cl, ++lt) {
      const int buf = lt & 1, u = lt >> 1;
      if (u >= 1) { mbar_wait_relaxed(&acc_empty[buf], (u - 1) & 1); tc_fence_after(); }      // the epilogue has drained this accumulator
      const uint32_t d_col = 256u * buf;
      for (int kc = 0; kc < n_kc; ++kc, ++gch) {
        const int slot = gch % NSLOT, use = gch / NSLOT;
        mbar_wait_relaxed(&full_a[slot], use & 1);      // the kernel is bound by operand delivery, not by this warp's reaction time:
        mbar_wait_relaxed(&full_w[slot], use & 1);      // sleeping leaves the issue slots to the producers and the epilogue
        if (PAIR) mbar_wait_relaxed(&peer_full[slot], use & 1);
        tc_fence_after();
        const uint32_t a32 = desc_lo_sw128(ring_a + slot * SLOT), w32 = desc_lo_sw128(ring_a + slot * SLOT + 2 * LA_PART);
#pragma unroll
        for (int part = 0; part < 3; ++part)
#pragma unroll
          for (int ks = 0; ks < 4; ++ks) {
            const uint32_t ad = a32 + (((part == 1 ? LA_PART : 0) + ks * 32) >> 4), wd = w32 + (((part == 2 ? W_LO : 0) + ks * 32) >> 4);
            const uint32_t acc = (kc > 0 || part > 0 || ks > 0) ? 1u : 0u;
            if (PAIR) mma_ss_pair_lo<DESC_HI_SW128, DESC_HI_SW128>(d_col, ad, wd, idesc, acc, e);
            else mma_ss_lo<DESC_HI_SW128, DESC_HI_SW128>(d_col, ad, wd, idesc, acc, e);
          }
        if (PAIR) mma_commit_pair_lo(&empty[slot], cl_mask, e);
        else if (CS == 1) mma_commit_lo(&empty[slot], e);
        else mma_commit_multicast_lo(&empty[slot], cl_mask, e);
      }
      if (PAIR) mma_commit_pair_lo(&acc_full[buf], cl_mask, e); else mma_commit_lo(&acc_full[buf], e);
    }
  } else {
    // ================= epilogue (warps 8-11: TMEM lane quarter = warp & 3) =================
    // A thread owns one output row in TMEM, but a warp must touch global memory one row SEGMENT at a time (the first version stored
    // 16 bytes per thread into 32 different rows per instruction: 32 line requests and half-written sectors, and that - not the MMAs -
    // set the pace of the K = 256 layers).  Every 32-column chunk therefore passes through a per-warp stage: rows in, 4 rows x 128
    // bytes out per instruction (and the reverse for the residual rows).
    const int ew = warp & 3, row = ew * 32 + lane, et = tid - L_PROD_WARPS * 32;      // et: 0..127
    const uint32_t lane_addr = (uint32_t)(ew * 32) << 16;
    float* wst = sStage + ew * (32 * L_STAGE_LD);
    const int sub_r = lane >> 3, sub_c = 4 * (lane & 7);      // coalesced side: 4 rows per instruction, 8 lanes x float4 per row
    auto epi_sync = [] { asm volatile("bar.sync 1, 128;" ::: "memory"); };
    int lt = 0;
    for (int t = cid; t < n_tiles; t += n_cl, ++lt) {
      const int mt = (t / n_nt) * CS + rank, nt = t % n_nt, n0 = nt * LNT;
      int src0, nvalid;
      tile_rows(mt, src0, nvalid);
      const int mw = src0 + ew * 32, nw = nvalid - ew * 32;      // first row of this warp; rows r < nw of the warp exist
      const int buf = lt & 1, u = lt >> 1;
      // registers (thread = row) -> global rows of the warp, columns n0 + c0 .. + 31
      auto store_rows = [&](const float (&v)[32], int c0) {
#pragma unroll
        for (int q = 0; q < 8; ++q) *reinterpret_cast<float4*>(wst + lane * L_STAGE_LD + 4 * q) = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
        __syncwarp();
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const int r = 4 * j + sub_r;
          if (r < nw)
            *reinterpret_cast<float4*>(g.out + (size_t)(mw + r) * g.ldo + n0 + c0 + sub_c) = *reinterpret_cast<const float4*>(wst + r * L_STAGE_LD + sub_c);
        }
        __syncwarp();
      };
      // global rows (row r of the warp at rowptr(r), 32 floats from column c0) -> registers (thread = row); rows beyond M: zeros.
      // Two halves, so that the loads of the NEXT chunk are in flight while this one is processed (their latency, eight times per
      // tile, was what the LayerNorm epilogue of the K = 256 projection cost).
      auto fetch_rows = [&](auto rowptr, int c0, float4 (&x)[8]) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const int r = 4 * j + sub_r;
          x[j] = (r < nw) ? __ldg(reinterpret_cast<const float4*>(rowptr(mw + r) + c0 + sub_c)) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
      };
      // ... stage_rows puts them into the warp's stage; thread = row then reads its 32 values as staged(q), q = 0..7, and the warp
      // synchronises before the stage is written again
      auto stage_rows = [&](const float4 (&x)[8]) {
#pragma unroll
        for (int j = 0; j < 8; ++j) *reinterpret_cast<float4*>(wst + (4 * j + sub_r) * L_STAGE_LD + sub_c) = x[j];
        __syncwarp();
      };
      auto staged = [&](int q) { return *reinterpret_cast<const float4*>(wst + lane * L_STAGE_LD + 4 * q); };
      epi_sync();                                                  // previous tile's readers of sVec are done
      for (int i = et; i < LNT; i += 128) {
        sVec[i] = (g.bias && n0 + i < g.Nr) ? __ldg(g.bias + n0 + i) : 0.0f;
        if (g.epi == LEPI_RES_LN) { sVec[LNT + i] = __ldg(g.gamma + i); sVec[2 * LNT + i] = __ldg(g.beta + i); }
      }
      epi_sync();
      mbar_wait_relaxed(&acc_full[buf], u & 1); tc_fence_after();
      const uint32_t d_addr = lane_addr + 256u * buf;
      if (g.epi == LEPI_SIGMOID) {
        // sigmoid(acc + bias) -> [M][Nr] with an unaligned row pitch: a warp writes 32 consecutive floats of one row per instruction
#pragma unroll 1
        for (int c8 = 0; c8 < 8 && n0 + 32 * c8 < g.Nr; ++c8) {
          float v[32];
          tmem_ld32(d_addr + 32 * c8, v); tmem_wait_ld();
#pragma unroll
          for (int e = 0; e < 32; ++e) wst[lane * L_STAGE_LD + e] = 1.0f / (1.0f + __expf(-(v[e] + sVec[32 * c8 + e])));
          __syncwarp();
          const int col = n0 + 32 * c8 + lane;
          for (int r = 0; r < 32; ++r)
            if (r < nw && col < g.Nr) g.out[(size_t)(mw + r) * g.ldo + col] = wst[r * L_STAGE_LD + lane];
          __syncwarp();
        }
      } else if (g.epi == LEPI_QKV && nt > 0) {
        // K (nt = 1) / V (nt = 2) of key tile (b, j), head c8 -> the operand image; the tile's rows are its keys (rows beyond T: zeros)
        if (mt < n_mt) {
          const int b = mt / g.n_kt, j = mt - b * g.n_kt;
#pragma unroll 1
          for (int c8 = 0; c8 < 8; ++c8) {
            float v[32];
            tmem_ld32(d_addr + 32 * c8, v); tmem_wait_ld();
            unsigned char* img = g.kv + ((size_t)(b * g.n_heads + c8) * g.n_kt + j) * L_KV_TILE;
            if (nt == 1) {      // (key, 8 channels) -> one 16-byte row of a core matrix: off = (key >> 3) 512 + u 128 + (key & 7) 16
              const uint32_t off0 = (uint32_t)(row >> 3) * 512 + (row & 7) * 16;
#pragma unroll
              for (int u = 0; u < 4; ++u) {
                uint4 hi, lo;
                split_h2(v[8 * u], v[8 * u + 1], hi.x, lo.x); split_h2(v[8 * u + 2], v[8 * u + 3], hi.y, lo.y);
                split_h2(v[8 * u + 4], v[8 * u + 5], hi.z, lo.z); split_h2(v[8 * u + 6], v[8 * u + 7], hi.w, lo.w);
                *reinterpret_cast<uint4*>(img + off0 + u * 128) = hi;
                *reinterpret_cast<uint4*>(img + L_KV_PART + off0 + u * 128) = lo;
              }
            } else {            // V^T: (channel, 8 keys) -> one 16-byte row: off = (d >> 3) 2048 + kg 128 + (d & 7) 16; lane = channel
#pragma unroll
              for (int q = 0; q < 8; ++q) *reinterpret_cast<float4*>(wst + lane * L_STAGE_LD + 4 * q) = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
              __syncwarp();
              unsigned char* vimg = img + 2 * L_KV_PART + (uint32_t)(lane >> 3) * 2048 + (lane & 7) * 16;
#pragma unroll
              for (int kg = 0; kg < 4; ++kg) {
                float x[8];
#pragma unroll
                for (int e = 0; e < 8; ++e) x[e] = wst[(8 * kg + e) * L_STAGE_LD + lane];
                uint4 hi, lo;
                split_h2(x[0], x[1], hi.x, lo.x); split_h2(x[2], x[3], hi.y, lo.y);
                split_h2(x[4], x[5], hi.z, lo.z); split_h2(x[6], x[7], hi.w, lo.w);
                *reinterpret_cast<uint4*>(vimg + (4 * ew + kg) * 128) = hi;
                *reinterpret_cast<uint4*>(vimg + L_KV_PART + (4 * ew + kg) * 128) = lo;
              }
              __syncwarp();
            }
          }
        }
      } else if (g.epi != LEPI_RES_LN) {
#pragma unroll 1
        for (int c8 = 0; c8 < 8; ++c8) {
          float v[32];
          tmem_ld32(d_addr + 32 * c8, v); tmem_wait_ld();
#pragma unroll
          for (int e = 0; e < 32; ++e) {
            v[e] += sVec[32 * c8 + e];
            if (g.epi == LEPI_BIAS_RELU) v[e] = fmaxf(v[e], 0.f);
          }
          store_rows(v, 32 * c8);
        }
      } else {
        // y = acc + bias + residual, kept in TMEM; LayerNorm(y) * gamma + beta (Keras non-fused order, eps 1e-6)
        float mean = 0.0f, m2 = 0.0f;
        auto res_row = [&](int mm) { return g.res + (size_t)mm * g.N; };
        auto pos_row = [&](int mm) { return g.pos + (size_t)(mm % g.T) * LNT; };
        float4 nx[8];
        if (g.res) fetch_rows(res_row, 0, nx);
#pragma unroll 1
        for (int c8 = 0; c8 < 8; ++c8) {
          float v[32];
          tmem_ld32(d_addr + 32 * c8, v);
          if (g.res) {
            stage_rows(nx);
            if (c8 + 1 < 8) fetch_rows(res_row, 32 * (c8 + 1), nx);
          }
          tmem_wait_ld();
          float s = 0.0f;
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            const float4 r4 = g.res ? staged(q) : make_float4(0.f, 0.f, 0.f, 0.f);
            v[4 * q] += sVec[32 * c8 + 4 * q] + r4.x;         v[4 * q + 1] += sVec[32 * c8 + 4 * q + 1] + r4.y;
            v[4 * q + 2] += sVec[32 * c8 + 4 * q + 2] + r4.z; v[4 * q + 3] += sVec[32 * c8 + 4 * q + 3] + r4.w;
          }
          __syncwarp();      // the stage is free again
#pragma unroll
          for (int q = 0; q < 8; ++q) s += (v[4 * q] + v[4 * q + 1]) + (v[4 * q + 2] + v[4 * q + 3]);
          const float mc = s * (1.0f / 32.0f);
          float qc = 0.0f;
#pragma unroll
          for (int j = 0; j < 32; ++j) { const float d = v[j] - mc; qc = fmaf(d, d, qc); }
          // Chan: merge (32 c8 values, mean, m2) with (32 values, mc, qc)
          const float na = 32.0f * c8, nb = 32.0f, dlt = mc - mean, nn = na + nb;
          mean += dlt * (nb / nn);
          m2 += qc + dlt * dlt * (na * nb / nn);
          tmem_st32(d_addr + 32 * c8, reinterpret_cast<const uint32_t(&)[32]>(v));
        }
        tmem_wait_st();
        const float rstd = rsqrtf(m2 * (1.0f / LNT) + 1e-6f);
        if (g.pos) fetch_rows(pos_row, 0, nx);
#pragma unroll 1
        for (int c8 = 0; c8 < 8; ++c8) {
          float v[32];
          tmem_ld32(d_addr + 32 * c8, v);
          if (g.pos) {
            stage_rows(nx);
            if (c8 + 1 < 8) fetch_rows(pos_row, 32 * (c8 + 1), nx);
          }
          tmem_wait_ld();
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            const float4 p4 = g.pos ? staged(q) : make_float4(0.f, 0.f, 0.f, 0.f);
            const float pe[4] = {p4.x, p4.y, p4.z, p4.w};
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const int c = 32 * c8 + 4 * q + e;
              const float inv = rstd * sVec[LNT + c];
              float y = fmaf(v[4 * q + e], inv, sVec[2 * LNT + c] - mean * inv);
              if (g.pos) y = fmaxf(y, 0.0f) + pe[e];
              v[4 * q + e] = y;
            }
          }
          __syncwarp();      // the stage is free again
          store_rows(v, 32 * c8);
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) { if (PAIR && rank != 0) mbar_arrive_remote(&acc_empty[buf], 0); else mbar_arrive(&acc_empty[buf]); }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (CS > 1) cluster_sync_all();      // no CTA leaves while a peer's commit may still arrive on its barriers
  if (warp == L_PROD_WARPS + 4) { if (PAIR) tmem_dealloc_pair(0, 512); else tmem_dealloc(0, 512); }
}

// ---- host side ------------------------------------------------------------------------------------------------
// W [K][N] fp32 row-major -> [N / 256][K / 64][hi | lo] chunks, each part [256 rows (n) x 128 B (64 k as fp16)], 128-byte swizzle
static void pack_lin(std::vector<unsigned char>& img, size_t base, const float* W, int Kr, int Nr) {
  const int K = (Kr + LK - 1) / LK * LK, N = (Nr + LNT - 1) / LNT * LNT;      // padded with zero weights
  const int n_nt = N / LNT, n_kc = K / LK;
  for (int nt = 0; nt < n_nt; ++nt)
    for (int kc = 0; kc < n_kc; ++kc) {
      unsigned char* hi = img.data() + base + ((size_t)nt * n_kc + kc) * L_WCHUNK;
      unsigned char* lo = hi + LW_PART;
      for (int n = 0; n < LNT; ++n)
        for (int kk = 0; kk < LK; ++kk) {
          const int k = kc * LK + kk, nn = nt * LNT + n;
          const float w = (k < Kr && nn < Nr) ? W[(size_t)k * Nr + nn] : 0.0f;
          const __half h = __float2half_rn(w);
          const __half l = __float2half_rn(w - __half2float(h));
          const size_t off = (size_t)(n >> 3) * 1024 + (n & 7) * 128 + (((kk >> 3) ^ (n & 7)) * 16) + (kk & 7) * 2;
          memcpy(hi + off, &h, 2);
          memcpy(lo + off, &l, 2);
        }
    }
}

// Packs the 4 x n_blocks weight matrices; net.umma_stage_offset[4 blk + {0: qkv, 1: projection, 2: ffn in, 3: ffn out}]
int mhanet_umma_prepare(dxi_net& net, cudaStream_t st) {
  const dxi_net_cfg& c = net.cfg;
  const int d = c.d_model;
  if (d != 256) { set_error("tcgen05 MHANetV3 path is built for d_model = 256"); return DXI_E_INVALID; }
  const size_t sz[4] = {(size_t)d * 3 * d * 4, (size_t)d * d * 4, (size_t)d * 4 * d * 4, (size_t)4 * d * d * 4};   // hi + lo = 4 B per weight
  std::vector<size_t> offs;
  size_t total = 0;
  for (int blk = 0; blk < c.n_blocks; ++blk)
    for (int i = 0; i < 4; ++i) { offs.push_back(total); total += sz[i]; }
  // input layer [n_feat -> d] (K padded to 320) and output layer [d -> n_outp] (N padded to 512)
  const int Kin = (c.n_feat + LK - 1) / LK * LK, Nout = (c.n_outp + LNT - 1) / LNT * LNT;
  offs.push_back(total); total += (size_t)Kin * d * 4;
  offs.push_back(total); total += (size_t)d * Nout * 4;
  std::vector<unsigned char> img(total, 0);
  pack_lin(img, offs[4 * c.n_blocks], net.host_tensor(0, "kernel")->data(), c.n_feat, d);
  pack_lin(img, offs[4 * c.n_blocks + 1], net.host_tensor(3 + 5 * c.n_blocks, "kernel")->data(), d, c.n_outp);
  for (int blk = 0; blk < c.n_blocks; ++blk) {
    const int li = 3 + 5 * blk;
    char nm[64];
    snprintf(nm, sizeof(nm), "packed-%d/qkv", li);
    auto it = net.host.find(nm);
    if (it == net.host.end()) { set_error("packed QKV weights missing"); return DXI_E_STATE; }
    pack_lin(img, offs[4 * blk + 0], it->second.data(), d, 3 * d);
    pack_lin(img, offs[4 * blk + 1], net.host_tensor(li, "projection_kernel")->data(), d, d);
    pack_lin(img, offs[4 * blk + 2], net.host_tensor(li + 2, "kernel")->data(), d, 4 * d);
    pack_lin(img, offs[4 * blk + 3], net.host_tensor(li + 3, "kernel")->data(), 4 * d, d);
  }
  net.umma_stage_offset = offs;
  if (net.d_umma) { cudaFree(net.d_umma); net.d_umma = nullptr; }
  DXI_CUDA(cudaMalloc(&net.d_umma, img.size()));
  DXI_CUDA(cudaMemcpyAsync(net.d_umma, img.data(), img.size(), cudaMemcpyHostToDevice, st));
  DXI_CUDA(cudaStreamSynchronize(st));      // img is a local
  net.umma_bytes = img.size();
  return DXI_OK;
}

static int lin_launch(const LinArgs& g, cudaStream_t st) {
  int n_sm = 148;
  { int dev = 0; cudaGetDevice(&dev); cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev); }
  const int n_mt = (g.M + LM - 1) / LM, n_nt = g.N / LNT;
  // DXI_LIN_CLUSTER = 2 / 4: clusters of that many row tiles share every weight chunk by multicast.  Measured on B200 at 64 x 1875 frames:
  // 4.85 ms (1), 4.92 ms (2), slower at 4 (fewer SMs fit whole clusters) - the L2 already merges the concurrent reads of a chunk.
  int cs = 1;
  if (const char* e = getenv("DXI_LIN_CLUSTER")) { const int v = atoi(e); if (v == 1 || v == 2 || v == 4) cs = v; }
  // DXI_LIN_PAIR=1: CTA pairs (tcgen05.mma.cta_group::2, M = 256 per weight fetch)
  const bool pair = getenv("DXI_LIN_PAIR") && atoi(getenv("DXI_LIN_PAIR")) && n_mt >= 2;
  if (pair) cs = 2;
  auto kern = pair ? lin_umma_kernel<2, true> : cs == 4 ? lin_umma_kernel<4> : cs == 2 ? lin_umma_kernel<2> : lin_umma_kernel<1>;
  DXI_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, L_SMEM));      // per device, so per call
  cudaLaunchConfig_t cfg{};
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = cs; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  cfg.blockDim = dim3(L_THREADS); cfg.dynamicSmemBytes = L_SMEM; cfg.stream = st; cfg.attrs = at; cfg.numAttrs = 1;
  int n_cl = n_sm / cs;
  if (cs > 1) {
    cfg.gridDim = dim3(n_sm / cs * cs);
    int max_cl = 0;
    DXI_CUDA(cudaOccupancyMaxActiveClusters(&max_cl, kern, &cfg));      // GPCs whose SM count is not a multiple of cs hold fewer clusters
    if (max_cl < 1) { set_error("lin_umma: no cluster of %d CTAs fits the device", cs); return DXI_E_STATE; }
    n_cl = max_cl < n_cl ? max_cl : n_cl;
  }
  const int tiles = (n_mt + cs - 1) / cs * n_nt;
  cfg.gridDim = dim3((tiles < n_cl ? tiles : n_cl) * cs);
  ProfScope prof("mha_gemm", st, 1);
  DXI_CUDA(cudaLaunchKernelEx(&cfg, kern, g));
  DXI_LAUNCHED("lin_umma_kernel");
  return DXI_OK;
}

// image: index into net.umma_stage_offset (4 blk + {0 qkv, 1 projection, 2 ffn in, 3 ffn out}; 4 n_blocks: input layer, + 1: output)
int mhanet_umma_linear(const dxi_net& net, int image, int epi, const float* A, int lda, const float* bias, const float* res,
                       const float* gamma, const float* beta, const float* pos, int T, float* out, int ldo, int M, int Nr, int Kr,
                       cudaStream_t st) {
  if (!net.d_umma || (size_t)image >= net.umma_stage_offset.size()) { set_error("tcgen05 weight images missing"); return DXI_E_STATE; }
  const int N = (Nr + LNT - 1) / LNT * LNT, K = (Kr + LK - 1) / LK * LK;
  LinArgs g{A, lda, reinterpret_cast<const unsigned char*>(net.d_umma) + net.umma_stage_offset[image], bias, res, gamma, beta, pos, T,
            out, ldo, M, N, K, epi, Nr, Kr};
  if (epi == LEPI_QKV) { set_error("lin_umma: the fused QKV epilogue is launched through mhanet_umma_qkv"); return DXI_E_INVALID; }
  if (epi == LEPI_RES_LN && Nr != LNT) { set_error("lin_umma: LayerNorm epilogue needs N = 256"); return DXI_E_INVALID; }
  if (epi != LEPI_SIGMOID && ((ldo & 3) || Nr != N)) { set_error("lin_umma: unaligned outputs only through the sigmoid epilogue"); return DXI_E_INVALID; }
  return lin_launch(g, st);
}

// Fused QKV projection of block blk: x [B T][d] -> Q (fp32, columns 0 .. d - 1 of qkv [B T][3 d]) and the packed K / V key tiles of the
// tensor-core attention (kv: mhanet_umma_attention_workspace bytes).  The K / V columns of qkv are NOT written.
int mhanet_umma_qkv(const dxi_net& net, int blk, const float* x, int B, int T, float* qkv, void* kv, cudaStream_t st) {
  const dxi_net_cfg& c = net.cfg;
  const int d = c.d_model, image = 4 * blk;
  if (!net.d_umma || (size_t)image >= net.umma_stage_offset.size()) { set_error("tcgen05 weight images missing"); return DXI_E_STATE; }
  if (d != LNT || d / c.n_heads != 32) { set_error("lin_umma: the fused QKV epilogue is built for d_model 256, head size 32"); return DXI_E_INVALID; }
  const int n_kt = (T + LM - 1) / LM;
  LinArgs g{x, d, reinterpret_cast<const unsigned char*>(net.d_umma) + net.umma_stage_offset[image], nullptr, nullptr, nullptr, nullptr, nullptr, T,
            qkv, 3 * d, B * n_kt * LM, 3 * d, d, LEPI_QKV, 3 * d, d};
  g.kv = reinterpret_cast<unsigned char*>(kv); g.n_kt = n_kt; g.n_heads = c.n_heads;
  return lin_launch(g, st);
}

}  // namespace dxi
