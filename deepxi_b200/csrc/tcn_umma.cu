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
// straight into tensor memory (TMEM lane r = frame r), so nothing is staged through shared memory.
//
// Deferred normalisation.  u(x) = (r - mu) * inv with r = ReLU(x), so  W^T u = inv * (W^T r) - inv * mu * colsum(W).
// The MMAs therefore consume the UN-normalised ReLU output (written to TMEM the moment it exists) and the
// row statistics are applied to the small GEMM outputs afterwards; no second pass over the 256-wide row, and the
// statistics merge runs while the tensor core is already busy.  colsum(W3), colsum(W1') ride in the weight image
// (summed over the effective fp16 hi[+lo] weights so the mean component cancels to rounding).
//
// TMEM columns: [0,256) GEMM2 accumulators, overwritten IN PLACE by the A operand of GEMM3 (32 fp32 columns
// -> 16 hi + 16 lo); [256,320) A operand of GEMM2, later GEMM3 accumulators; [320,512) the three c1 taps of
// the NEXT tile (loaded while this tile's GEMM3 drains); GEMM1 accumulators reuse [0,64).
// Warps 0-15: epilogue, four threads per frame (a quarter of the columns each; TMEM lane quarter = warp & 3;
// LayerNorm statistics merged with Chan's formula through shared memory).  Warp 16: weight load (bulk async
// copy), L2 prefetch of the next tile, warp-convergent MMA issue.  GEMM2 is committed in four column groups and
// GEMM3 is issued in eight K-chunks as the epilogue produces them.
//
// Launch structure: stem x 2, 41 stage launches, output layer, all with programmatic dependent launch.  The stage launches depend on
// each other TILE BY TILE through counters in the workspace (flags[tile] = number of stages that have published the tile, see
// the kernel), not launch by launch: a stage's CTAs start on the SMs its predecessor has already left and the 41 launches
// form one uninterrupted stream of tiles.  Registers: the MMA warp's warpgroup is launched whole and setmaxnreg moves
// registers from it (64) to the epilogue (104); the MMA warp re-forms its shared-memory descriptors per tile instead of
// keeping them as (spilled) loop invariants.  DESIGN.md section 4 has the measurements behind both.
#include <cstdlib>
#include <vector>
#include "launch.cuh"
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
constexpr int IMG_BIAS = 2 * IMG_PART;    // fp32: b2[64], b3[256], b1'[64], colsum(W3)[256], colsum(W1')[64], 1/s1, 1/s2, 1/s3 (the matrices are stored as W * s, see weight_pow2_scale)
constexpr int OFF_B2 = 0, OFF_B3 = 64, OFF_B1 = 320, OFF_CS3 = 384, OFF_CS1 = 640, OFF_IS = 704, N_AUX = 708;      // OFF_IS: 1/s1, 1/s2, 1/s3 (weight scales)
constexpr int IMG_BYTES = IMG_BIAS + N_AUX * 4;

// TMEM column map
constexpr uint32_t COL_D2 = 0;                            // 256 fp32; chunk cc becomes A3: hi at 32cc, lo at 32cc+16
constexpr uint32_t COL_D1 = 0;                            // 64 fp32 (dead before GEMM2 writes)
constexpr uint32_t COL_A2_HI = 256, COL_A2_LO = 288;      // 32 columns each
constexpr uint32_t COL_D3 = 256;                          // 64 fp32 (A2 is dead by then)
constexpr uint32_t COL_A1_HI = 320, COL_A1_LO = 416;      // 3 taps x 32 columns each

struct StageArgs {
  const unsigned char* img;      // packed weights of this stage
  float* h;                      // residual stream, tiled: [tile][c/4][row][4] fp32
  const __half* c1_in;           // [B][2 planes][8 units][Ts][8] fp16 (hi plane, lo plane)
  __half* c1_out;
  int T, tiles_per_utt, n_tiles, Ts;
  int shift0, shift1, shift2;    // frame offsets of the three taps: tap j reads frame t - shift_j
  int has_back, has_front;
  int reverse;                   // walk this CTA's tiles last-to-first: consecutive stages alternate, so each starts on what the previous one left in L2
  const float2* stem_stats;     // stage 0 only: per-row partial statistics of the stem pre-activation (8 parts of 32 channels);
                                // the row is then LayerNorm(gamma) + ReLU'd on load (tcn.py:176-179), gamma in the b3 slot
  const float* sc_in;           // per-row power-of-two operand scale left by the previous stage (nullptr in stage 0: 1.0)
  float sc0;                    // stage 0: power of two next to 1 / rms(gamma) of the first layer
  float* sc_out;                // ... and the one this stage leaves for the next: the exponent of the row's 1 / std (fp16 range guard, see P2)
  int zero;                     // always 0 (see the MMA warp's descriptor arithmetic)
  int wait_val;                 // with flags: a tile of this stage may start once flags[.] >= wait_val for it and its neighbours
  int* flags;                   // flags[tile] = number of stages that have published the tile (zeroed by the host before the stem); nullptr:
                                // every stage waits for the whole previous launch (griddepcontrol.wait) instead
#ifdef DXI_ENABLE_DEBUG         // tuning build only (DXI_DEBUG_BUILD=1 python -m deepxi_b200.build): none of this exists in the product library
  int dbg_flags;                // tuning experiments: 1 = skip residual loads, 2 = skip residual stores, 4 = skip c1 tap loads
  long long* dbg_cta;           // per-CTA timeline of this launch in globaltimer ns (8 slots per CTA), see scripts/tcn_timeline.py
  long long* dbg;               // clock64 stamps of the epilogue phases (16 per tile), see dxi_debug_tcn_clocks
#endif
};
#ifdef DXI_ENABLE_DEBUG
#define DBG_CTA (p.dbg_cta)
#define DBG_FLAGS (p.dbg_flags)
#define DBG_STAMPS (p.dbg)
#else
#define DBG_CTA (static_cast<long long*>(nullptr))
#define DBG_FLAGS 0
#define DBG_STAMPS (static_cast<long long*>(nullptr))
#endif

__device__ __forceinline__ float relu(float x) { return fmaxf(x, 0.0f); }

constexpr int NSPLIT = 4;                         // threads per frame
constexpr int EPI_WARPS = 4 * NSPLIT;
constexpr int EPI_THREADS = EPI_WARPS * 32;
constexpr int STAGE_THREADS = EPI_THREADS + 32;   // + 1 warp: weight load, L2 prefetch, MMA issue
// tcn_stage_kernel: the MMA warp's warpgroup is launched whole (3 idle warps) so that setmaxnreg can move registers from it to the
// epilogue.  A CTA is allocated in units of four warps anyway (17 warps cost 20; measured: 21 warps x 96 registers do not launch),
// the kernel is compiled for 96 registers per thread = a pool of 640 x 96 = 61 440, redistributed as 512 x 104 + 128 x 64.  The
// epilogue's eight extra registers remove most of its spills (-4.5 % stage time); below 64 the MMA warp spills its descriptors and
// the kernel gets slower (40 / 104: +2.5 %).
constexpr int TCN_THREADS = EPI_THREADS + 128;
constexpr int TCN_REGS_MMA = 64, TCN_REGS_EPI = 104;

__device__ __forceinline__ void epi_barrier() { asm volatile("bar.sync 1, %0;" ::"n"(EPI_THREADS) : "memory"); }
// The NSPLIT threads that share a frame sit in the NSPLIT warps with the same lane quarter q = warp & 3: statistics
// merges only need those 128 threads (named barrier 2 + q), not the whole epilogue.
__device__ __forceinline__ void quarter_barrier(int q) { asm volatile("bar.sync %0, %1;" ::"r"(2 + q), "n"(NSPLIT * 32) : "memory"); }
__device__ __forceinline__ void prefetch_l2(const void* p, uint32_t bytes) {
  asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p), "r"(bytes) : "memory");
}

__device__ __forceinline__ long long globaltimer_ns() {
  long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
__device__ __forceinline__ int ld_acquire_gpu(const int* p) {
  int v;
  asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void red_release_gpu_add(int* p, int v) {
  asm volatile("red.release.gpu.global.add.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// LayerNorm statistics of one row split between NSPLIT threads (n values each): every thread contributes
// (mean_i, M2_i = sum of squared deviations from mean_i); Chan's formula merges them.  Returns mean and
// 1/sqrt(biased variance + 1e-6) of the whole row.
__device__ __forceinline__ void ln_merge(float2* red, int row, int qd, float n, float mean_i, float m2_i, float& mean, float& inv, float eps = 1e-6f) {
  red[qd * TILE + row] = make_float2(mean_i, m2_i);
  quarter_barrier(row >> 5);
  float2 pt[NSPLIT];
#pragma unroll
  for (int i = 0; i < NSPLIT; ++i) pt[i] = red[i * TILE + row];
  float m = 0.0f;
#pragma unroll
  for (int i = 0; i < NSPLIT; ++i) m += pt[i].x;
  m *= (1.0f / NSPLIT);
  float m2 = 0.0f;
#pragma unroll
  for (int i = 0; i < NSPLIT; ++i) { const float d = pt[i].x - m; m2 += pt[i].y + n * d * d; }
  mean = m;
  inv = rsqrtf(m2 / (n * NSPLIT) + eps);
}

template <bool SPLIT>
__global__ void __launch_bounds__(TCN_THREADS, 1) tcn_stage_kernel(const StageArgs p) {
  extern __shared__ unsigned char smem_raw[];
  __shared__ __align__(8) uint64_t bar_w, bar_a1, bar_a2, bar_a3[8], bar_d1, bar_d2[4], bar_d3, bar_dep, bar_pub;
  __shared__ uint32_t tmem_slot;
  __shared__ float2 red[3][NSPLIT * TILE];
  unsigned char* sW = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const float* sAux = reinterpret_cast<const float*>(sW + IMG_BIAS);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  // (debug) per-CTA timeline slots: the pointer is re-formed from the kernel parameters at every use instead of living in registers
#define tl (DBG_CTA + (size_t)blockIdx.x * 8)
  if (DBG_CTA && tid == 0) {
    uint32_t smid;
    asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
    tl[0] = globaltimer_ns(); tl[6] = smid;
  }

  if (warp == EPI_WARPS) tmem_alloc(&tmem_slot, 512);
  if (tid == 0) {
    mbar_init(&bar_w, 1);
    // epilogue -> MMA barriers count WARPS: after the warp-wide tcgen05.wait one lane arrives for its warp
    mbar_init(&bar_a1, EPI_WARPS);
    mbar_init(&bar_a2, EPI_WARPS);
    for (int i = 0; i < 8; ++i) mbar_init(&bar_a3[i], 4);        // a 32-channel chunk belongs to the 4 warps of one column quarter
    mbar_init(&bar_d1, 1);
    for (int i = 0; i < 4; ++i) mbar_init(&bar_d2[i], 1);
    mbar_init(&bar_d3, 1);
    mbar_init(&bar_dep, 1);
    mbar_init(&bar_pub, EPI_WARPS);
    fence_mbar_init();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  // The CTA owns the SM's whole tensor memory (512 columns, one CTA per SM), so the allocation starts at
  // lane 0 / column 0: TMEM addresses below are compile-time constants.
  if (tmem_slot != 0) __trap();
  // Programmatic dependent launch: the next stage's CTAs may be scheduled as SMs drain (their prologue -- barrier
  // init, TMEM allocation, weight load -- overlaps this launch's tail); nothing written by the PREVIOUS launch is
  // touched before griddepcontrol.wait below.
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  // Tile-level dependencies (p.flags): a stage with a back half does NOT wait for the whole previous launch.  Its CTAs start on
  // the SMs the previous stage has already left (ragged last round) and take every tile as soon as the previous stage has
  // published it and its two neighbours in the utterance (left / self: producers of the c1 rows and of h; right: done READING the
  // c1 plane this stage overwrites, and a producer when the padding is 'same').  So the stages form one continuous stream of
  // tiles: no idle tail, no cold first round per launch.  Stage 0 still waits for the stem through griddepcontrol.wait.  The two
  // launches of a pair can be co-resident only because every CTA of the older one is already running when the younger starts
  // (launch_dependents above is executed by all of them first), so a waiting consumer never keeps its producer off the machine.
  const bool dep_flags = p.flags != nullptr && p.has_back;      // consume flags instead of griddepcontrol.wait
  const bool pub_flags = p.flags != nullptr && p.has_front;     // publish this stage's tiles for the next launch

  // this CTA's tiles: blockIdx.x, blockIdx.x + grid, ... in ascending or (p.reverse) descending order
  const int n_r = (int)blockIdx.x < p.n_tiles ? (p.n_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;
  // Order: the full rounds (every CTA has a tile) ascending or, in a `reverse` stage, descending; consecutive stages alternate so
  // that each starts on the ~2.5 rounds its predecessor left in L2.  The tile of the ragged last round always comes LAST: the CTAs
  // of the next stage that start on the SMs this stage has already left then begin with tiles that are complete (and L2-hot)
  // instead of waiting for the very tiles the ragged round is still producing.
  const int n_full = p.n_tiles / (int)gridDim.x;
  auto tile_at = [&](int r) { return (int)blockIdx.x + (r < n_full ? (p.reverse ? n_full - 1 - r : r) : n_full) * (int)gridDim.x; };
  const int tile_first = tile_at(0);

  if (warp >= EPI_WARPS) asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(TCN_REGS_MMA));
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
    const uint32_t w_base = smem_u32(sW);
    constexpr uint32_t id64 = make_idesc_f16(TILE, 64);
    constexpr int NPART = SPLIT ? 3 : 1;          // (a_hi, w_hi) [+ (a_lo, w_hi) + (a_hi, w_lo)]
    uint32_t ph = 0, ppub = 0;
    // This warp is also the CTA's synchronisation agent, so that the epilogue never pays a gpu-scope fence or a flag round trip.
    // publish(t): once every epilogue warp has written tile t's h and c1 rows (bar_pub), release them at gpu scope by bumping
    // flags[t].  await(t): lanes 0..2 poll tiles t-1, t, t+1 of the utterance until the previous stage has published them, then
    // bar_dep tells the epilogue (one phase per awaited tile; never more than one phase ahead of its consumer, see below).
    auto publish = [&](int t) {
      mbar_wait(&bar_pub, ppub); ppub ^= 1;
      if (elect_one()) { __threadfence(); red_release_gpu_add(p.flags + t, 1); }
      __syncwarp();
    };
    auto await = [&](int t) {
      const int j = t % p.tiles_per_utt, dj = lane - 1;
      if (lane < 3 && j + dj >= 0 && j + dj < p.tiles_per_utt) {
        const int* f = p.flags + t + dj;
        if (ld_acquire_gpu(f) < p.wait_val) {      // bounded: a broken co-residency argument must fail the launch, not hang the GPU
          const long long t_spin = clock64();
          uint32_t spins = 0;
          while (ld_acquire_gpu(f) < p.wait_val)
            if ((++spins & 255u) == 0 && clock64() - t_spin > SPIN_LIMIT_CYCLES) __trap();
        }
      }
      __syncwarp();
      if (elect_one()) mbar_arrive(&bar_dep);
      __syncwarp();
    };
    if (DBG_CTA && lane == 0) tl[1] = globaltimer_ns();
    if (dep_flags) { if (n_r > 0) await(tile_first); }
    else asm volatile("griddepcontrol.wait;" ::: "memory");
    if (DBG_CTA && lane == 0) tl[2] = globaltimer_ns();
    for (int r = 0; r < n_r; ++r) {
      const int tile = tile_at(r);
      // The shared-memory descriptors are recomputed per tile (a handful of uniform-datapath adds) instead of being kept as ~100
      // spilled loop invariants: p.zero (= 0, but only the host knows) makes the base depend on the iteration without taking it
      // out of the uniform registers the MMA instruction reads its operands from.
      const uint32_t w_hi = w_base + (uint32_t)(p.zero * r), w_lo = w_hi + IMG_PART;
      if (pub_flags && !p.has_back && r > 0) publish(tile_at(r - 1));
      {   // pull the next tile's residual rows / c1 rows into L2 ahead of their use
        const int nt = tile_at(r + 1);
        if (r + 1 < n_r) {
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
            for (int kc = 0; kc < 3; ++kc) {      // one tap = one 64-wide K chunk = four K16 steps
              mma_ts_elect_k<4>(COL_D1, a0 + 32 * kc, make_smem_desc_sw128(w0 + kc * 64 * 128), id64, acc);
              acc = 1;
            }
          }
        }
        mma_commit_elect(&bar_d1);
        mbar_wait(&bar_a2, ph); tc_fence_after();
#pragma unroll
        for (int g = 0; g < 4; ++g) {   // W3 ReLU(c2): K = 64, N = 256 in four column groups, each committed on its own
          uint32_t acc = 0;
#pragma unroll
          for (int part = 0; part < NPART; ++part) {
            const uint32_t a0 = part == 1 ? COL_A2_LO : COL_A2_HI, w0 = (part == 2 ? w_lo : w_hi) + IMG_W3 + g * 64 * 128;
            mma_ts_elect_k<4>(COL_D2 + 64 * g, a0, make_smem_desc_sw128(w0), id64, acc);
            acc = 1;
          }
          mma_commit_elect(&bar_d2[g]);
        }
        // bar_a1 of this tile has been seen, so the epilogue has consumed this tile's bar_dep phase and is past the stores of the
        // previous tile: publish that one, then look at the next tile's dependencies (the epilogue loads its taps after P2 of
        // this tile).  Both round trips hide behind GEMM2 and the first chunks of P2.
        if (pub_flags && r > 0) publish(tile_at(r - 1));
        if (dep_flags && r + 1 < n_r) await(tile_at(r + 1));
      }
      if (p.has_front) {
        uint32_t acc = 0;
#pragma unroll
        for (int cc = 0; cc < 8; ++cc) {   // W1' ReLU(h): K = 256 in eight 32-channel chunks, issued as the epilogue produces them
          mbar_wait(&bar_a3[cc], ph); tc_fence_after();
#pragma unroll
          for (int part = 0; part < NPART; ++part) {
            const uint32_t a0 = COL_D2 + 32 * cc + (part == 1 ? 16 : 0), w0 = (part == 2 ? w_lo : w_hi) + IMG_W1;
            // chunk cc holds K16 steps 2cc, 2cc + 1: half of the 64-wide K chunk cc >> 1
            mma_ts_elect_k<2>(COL_D3, a0, make_smem_desc_sw128(w0 + (cc >> 1) * 64 * 128 + (cc & 1) * 64), id64, acc);
            acc = 1;
          }
        }
        mma_commit_elect(&bar_d3);
      }
      ph ^= 1;
    }
    if (pub_flags && n_r > 0) publish(tile_at(n_r - 1));
  } else if (warp < EPI_WARPS) {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(TCN_REGS_EPI));
    // ================= epilogue: warps (q, qd): TMEM lane quarter q = warp & 3, column quarter qd = warp >> 2 =================
    mbar_wait(&bar_w, 0);        // biases / column sums live in the weight image
    const int qd = warp >> 2, row = (warp & 3) * 32 + lane;
    const uint32_t lane_addr = (uint32_t)((warp & 3) * 32) << 16;
    uint32_t ph = 0, pdep = 0;
    // one arrival per warp: tcgen05.wait is warp-wide, so once it returns every lane's TMEM traffic is done
    auto warp_arrive = [&](uint64_t* bar) { tc_fence_before(); __syncwarp(); if (lane == 0) mbar_arrive(bar); };
    const bool stamp = DBG_STAMPS != nullptr && tid == 0;
#define DXI_STAMP(k) do { if (stamp) DBG_STAMPS[(size_t)tile * 16 + (k)] = clock64(); } while (0)

    // c1 taps of a tile -> TMEM [320,512): this thread moves units 2qd, 2qd+1 of every (tap, plane)
    auto load_a1 = [&](int tile_) {
      const int b = tile_ / p.tiles_per_utt, t = (tile_ - b * p.tiles_per_utt) * TILE + row;
      const __half* cb = p.c1_in + (size_t)b * 2 * 8 * p.Ts * 8;
      uint4 qv[3][SPLIT ? 2 : 1][2];
#pragma unroll
      for (int j = 0; j < 3; ++j) {
        const int shift = j == 0 ? p.shift0 : (j == 1 ? p.shift1 : p.shift2);
        const size_t r_in = (size_t)(t - shift + C1_PAD);
#pragma unroll
        for (int plane = 0; plane < (SPLIT ? 2 : 1); ++plane)
#pragma unroll
          for (int u = 0; u < 2; ++u)
            qv[j][plane][u] = (DBG_FLAGS & 4) ? make_uint4(0, 0, 0, 0) : __ldcg(reinterpret_cast<const uint4*>(cb + ((size_t)(plane * 8 + 2 * qd + u) * p.Ts + r_in) * 8));
      }
#pragma unroll
      for (int j = 0; j < 3; ++j)
#pragma unroll
        for (int plane = 0; plane < (SPLIT ? 2 : 1); ++plane) {
          const uint32_t r[8] = {qv[j][plane][0].x, qv[j][plane][0].y, qv[j][plane][0].z, qv[j][plane][0].w,
                                 qv[j][plane][1].x, qv[j][plane][1].y, qv[j][plane][1].z, qv[j][plane][1].w};
          tmem_st8(lane_addr + (plane ? COL_A1_LO : COL_A1_HI) + 32 * j + 8 * qd, r);
        }
      tmem_wait_st(); tc_fence_before();
    };

    // h and the c1 planes are read with ld.global.cg (L2, the point of coherence): with tile-level dependencies the producer may be a
    // CTA of the previous launch that is still running on another SM
    auto dep_wait = [&]() { if (dep_flags) { mbar_wait(&bar_dep, pdep); pdep ^= 1; } };
    if (!dep_flags) asm volatile("griddepcontrol.wait;" ::: "memory");      // the previous launch's h / c1 are complete and visible
    if (p.has_back && n_r > 0) { dep_wait(); load_a1(tile_first); warp_arrive(&bar_a1); }
    if (DBG_CTA && tid == 0) { tl[3] = globaltimer_ns(); tl[7] = n_r; }
    for (int r = 0; r < n_r; ++r) {
      const int tile = tile_at(r);
      const int b = tile / p.tiles_per_utt, t0 = (tile - b * p.tiles_per_utt) * TILE;
      const int t = t0 + row;
      const bool valid = t < p.T;
      float* hrow = p.h + (size_t)tile * (TILE * 256) + row * 4;          // + c4 * 512
      DXI_STAMP(0);
      // The first residual chunk of the tile is requested here, ahead of P1: since the MMA warp issues GEMM2 promptly its wait no
      // longer covers the round trip of these loads, P1 and the statistics merge do.
      float4 hv[8];
      auto load_h = [&](int cc_) {
#pragma unroll
        for (int q = 0; q < 8; ++q) hv[q] = (DBG_FLAGS & 1) ? make_float4(0.f, 0.f, 0.f, 0.f) : __ldcg(reinterpret_cast<const float4*>(hrow + (size_t)(cc_ * 8 + q) * (TILE * 4)));
      };
      load_h(qd);
      float mu2 = 0.0f, inv2 = 0.0f;
      if (p.has_back) {
        // ---- P1: r2 = ReLU(acc1 + b2) -> A2 at once (un-normalised); statistics merged while GEMM2 runs
        mbar_wait(&bar_d1, ph); tc_fence_after();
        DXI_STAMP(1);
        float a[16];
        tmem_ld16(lane_addr + COL_D1 + 16 * qd, a); tmem_wait_ld();
        float s = 0.0f;
#pragma unroll
        for (int j = 0; j < 16; ++j) { a[j] = relu(fmaf(a[j], sAux[OFF_IS + 1], sAux[OFF_B2 + 16 * qd + j])); s += a[j]; }
        {
          uint32_t hi[8], lo[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) to_h2<SPLIT>(a[2 * j], a[2 * j + 1], hi[j], lo[j]);
          tmem_st8(lane_addr + COL_A2_HI + 8 * qd, hi);
          if (SPLIT) tmem_st8(lane_addr + COL_A2_LO + 8 * qd, lo);
        }
        tmem_wait_st(); warp_arrive(&bar_a2);
        DXI_STAMP(2);
        const float mean_i = s * (1.0f / 16.0f);
        float q2 = 0.0f;
#pragma unroll
        for (int j = 0; j < 16; ++j) { const float d = a[j] - mean_i; q2 = fmaf(d, d, q2); }
        ln_merge(red[0], row, qd, 16.0f, mean_i, q2, mu2, inv2);
        inv2 *= sAux[OFF_IS + 2];      // W3 is stored as W3 * s3
      }
      DXI_STAMP(3);
      // ---- P2: h_new = h + b3 + inv2 (acc2 - mu2 colsum(W3)); r3 = ReLU(h_new) -> A3 in place, chunk by chunk
      float mean0 = 0.0f, inv0 = 0.0f;
      if (p.stem_stats) {     // merge the 8 partial statistics (32 channels each) the stem kernels left for this row
        const float2* sp = p.stem_stats + ((size_t)tile * TILE + row) * 8;
        float2 pt[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) pt[i] = sp[i];
        float m = 0.0f;
#pragma unroll
        for (int i = 0; i < 8; ++i) m += pt[i].x;
        m *= 0.125f;
        float m2 = 0.0f;
#pragma unroll
        for (int i = 0; i < 8; ++i) { const float d = pt[i].x - m; m2 += pt[i].y + 32.0f * d * d; }
        mean0 = m;
        inv0 = rsqrtf(m2 * (1.0f / 256.0f) + 1e-6f);
      }
      float2 s1v = make_float2(0.0f, 0.0f), s2v = make_float2(0.0f, 0.0f);     // sums of r and r^2 over this thread's 64 channels
      const float nim = -inv2 * mu2;
      // fp16 range guard: the MMAs consume the UN-normalised ReLU(h) as fp16 hi | lo, so the row is pre-scaled by an exact power of
      // two taken from its 1 / std one block earlier (the residual stream changes slowly); LayerNorm is scale invariant, the epsilon
      // is scaled by sc^2, so the result does not change -- but |h| of 1e-4 or 1e5 no longer leaves fp16's normal range.
      const float sc = !p.has_front ? 1.0f : (p.sc_in ? __ldcg(p.sc_in + (size_t)tile * TILE + row) : p.sc0);
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        const int cc = qd + 4 * i;
        if (i == 1) load_h(cc);
        float v[32];
        if (p.has_back) {
          mbar_wait(&bar_d2[cc >> 1], ph); tc_fence_after();
          if (i == 0) DXI_STAMP(4); else DXI_STAMP(13);
          tmem_ld32(lane_addr + COL_D2 + 32 * cc, v); tmem_wait_ld();
          if (i == 0) DXI_STAMP(10); else DXI_STAMP(14);
          // packed fp32x2 arithmetic (FADD2 / FFMA2): v = inv2 * acc + (nim * colsum + (h + b3))
          const float4* b3 = reinterpret_cast<const float4*>(sAux + OFF_B3 + 32 * cc);
          const float4* cs = reinterpret_cast<const float4*>(sAux + OFF_CS3 + 32 * cc);
          const float2 inv2v = make_float2(inv2, inv2), nimv = make_float2(nim, nim);
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            const float4 bq = b3[q], cq = cs[q];
            float2 t0 = __fadd2_rn(make_float2(hv[q].x, hv[q].y), make_float2(bq.x, bq.y));
            float2 t1 = __fadd2_rn(make_float2(hv[q].z, hv[q].w), make_float2(bq.z, bq.w));
            t0 = __ffma2_rn(nimv, make_float2(cq.x, cq.y), t0);
            t1 = __ffma2_rn(nimv, make_float2(cq.z, cq.w), t1);
            t0 = __ffma2_rn(inv2v, make_float2(v[4 * q], v[4 * q + 1]), t0);
            t1 = __ffma2_rn(inv2v, make_float2(v[4 * q + 2], v[4 * q + 3]), t1);
            v[4 * q] = t0.x; v[4 * q + 1] = t0.y; v[4 * q + 2] = t1.x; v[4 * q + 3] = t1.y;
          }
        } else if (p.stem_stats) {
          // stage 0: the loaded row is the stem pre-activation z; h0 = ReLU(z * inv0 * gamma - mean0 * inv0 * gamma)
          const float* gm = sAux + OFF_B3 + 32 * cc;
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            const float g0 = inv0 * gm[4 * q], g1 = inv0 * gm[4 * q + 1], g2 = inv0 * gm[4 * q + 2], g3 = inv0 * gm[4 * q + 3];
            v[4 * q]     = relu(fmaf(hv[q].x, g0, -mean0 * g0));
            v[4 * q + 1] = relu(fmaf(hv[q].y, g1, -mean0 * g1));
            v[4 * q + 2] = relu(fmaf(hv[q].z, g2, -mean0 * g2));
            v[4 * q + 3] = relu(fmaf(hv[q].w, g3, -mean0 * g3));
          }
        } else {
#pragma unroll
          for (int q = 0; q < 8; ++q) { v[4 * q] = hv[q].x; v[4 * q + 1] = hv[q].y; v[4 * q + 2] = hv[q].z; v[4 * q + 3] = hv[q].w; }
        }
        if (!valid) {
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = 0.0f;
        }
        if ((p.has_back || p.stem_stats) && !(DBG_FLAGS & 2)) {
#pragma unroll
          for (int q = 0; q < 8; ++q)
            *reinterpret_cast<float4*>(hrow + (size_t)(cc * 8 + q) * (TILE * 4)) = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
        }
        if (i == 0) DXI_STAMP(11);
        if (p.has_front) {
          uint32_t hi[16], lo[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const float2 r = __fmul2_rn(make_float2(relu(v[2 * j]), relu(v[2 * j + 1])), make_float2(sc, sc));
            s1v = __fadd2_rn(s1v, r);
            s2v = __ffma2_rn(r, r, s2v);
            to_h2<SPLIT>(r.x, r.y, hi[j], lo[j]);
          }
          tmem_st16(lane_addr + COL_D2 + 32 * cc, hi);
          if (SPLIT) tmem_st16(lane_addr + COL_D2 + 32 * cc + 16, lo);
          tmem_wait_st(); warp_arrive(&bar_a3[cc]);
        }
        if (i == 0) DXI_STAMP(12);
      }
      DXI_STAMP(5);
      float mu3 = 0.0f, inv3 = 0.0f;
      if (p.has_front) {
        const float s1 = s1v.x + s1v.y, s2 = s2v.x + s2v.y, m1 = s1 * (1.0f / 64.0f);
        ln_merge(red[1], row, qd, 64.0f, m1, fmaxf(s2 - s1 * m1, 0.0f), mu3, inv3, 1e-6f * sc * sc);
        if (qd == 0 && p.sc_out)
          p.sc_out[(size_t)tile * TILE + row] = __uint_as_float(__float_as_uint(fminf(fmaxf(inv3 * sc, 1e-30f), 1e30f)) & 0x7F800000u);
      }
      DXI_STAMP(6);
      // ---- the next tile's c1 taps travel to TMEM while this tile's GEMM3 drains
      const int next = tile_at(r + 1);
      const bool has_next = p.has_back && r + 1 < n_r;
      if (has_next) { dep_wait(); load_a1(next); }
      DXI_STAMP(7);
      if (p.has_front) {
        // ---- P3: c1' = LN(ReLU(inv3 (acc3 - mu3 colsum(W1')) + b1')) -> fp16 hi | lo planes in HBM (zeros beyond T)
        mbar_wait(&bar_d3, ph); tc_fence_after();
        DXI_STAMP(8);
        float a[16];
        tmem_ld16(lane_addr + COL_D3 + 16 * qd, a); tmem_wait_ld();
        if (has_next) warp_arrive(&bar_a1);      // every D3 read is done: GEMM1 / A2 of the next tile may reuse the columns
        const float inv3w = inv3 * sAux[OFF_IS], nim3 = -inv3w * mu3;      // W1 is stored as W1 * s1
        float s = 0.0f;
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          a[j] = relu(fmaf(inv3w, a[j], fmaf(nim3, sAux[OFF_CS1 + 16 * qd + j], sAux[OFF_B1 + 16 * qd + j])));
          s += a[j];
        }
        const float mean_i = s * (1.0f / 16.0f);
        float q2 = 0.0f;
#pragma unroll
        for (int j = 0; j < 16; ++j) { const float d = a[j] - mean_i; q2 = fmaf(d, d, q2); }
        float mean, inv;
        ln_merge(red[2], row, qd, 16.0f, mean_i, q2, mean, inv);
        const float off = -mean * inv;
        __half* ob = p.c1_out + (size_t)b * 2 * 8 * p.Ts * 8;
        const size_t r_out = (size_t)(t + C1_PAD);
#pragma unroll
        for (int u = 0; u < 2; ++u) {
          uint32_t hi[4], lo[4];
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const float x0 = valid ? fmaf(a[8 * u + 2 * j], inv, off) : 0.0f, x1 = valid ? fmaf(a[8 * u + 2 * j + 1], inv, off) : 0.0f;
            to_h2<SPLIT>(x0, x1, hi[j], lo[j]);
          }
          *reinterpret_cast<uint4*>(ob + ((size_t)(2 * qd + u) * p.Ts + r_out) * 8) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
          if (SPLIT) *reinterpret_cast<uint4*>(ob + ((size_t)(8 + 2 * qd + u) * p.Ts + r_out) * 8) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
        }
        if (pub_flags) { __syncwarp(); if (lane == 0) mbar_arrive(&bar_pub); }      // this warp's share of the tile's h and c1 rows is written
      } else if (has_next) {
        warp_arrive(&bar_a1);
      }
      DXI_STAMP(9);
      ph ^= 1;
    }
#undef DXI_STAMP
    if (DBG_CTA && tid == 0) tl[4] = globaltimer_ns();
  }
  tc_fence_before();
  __syncthreads();
  if (DBG_CTA && tid == 0) tl[5] = globaltimer_ns();
#undef tl
  if (warp == EPI_WARPS) tmem_dealloc(0, 512);
}

// ---------------------------------------------------------------------------------------------------
// Stem (tcn.py:166-180) and output layer (tcn.py:158-161) on the tensor cores.
//
// stem_umma_kernel<half>: z[:, 128 half .. +128) = mag[:, 0..255] W0[0..255, ...] (fp16 hi/lo split, 3 MMAs: the
//   stem is the layer the 0.1 dB budget is most sensitive to) + mag[:, 256] W0[256, ...] + b0 (CUDA cores, exact).
//   Two launches (the hi+lo weights of one half are 128 KB of shared memory); z goes to the tiled residual
//   buffer, per-row partial statistics to a side buffer; stage 0 applies LayerNorm(gamma) + ReLU on load.
// head_umma_kernel: x_bar = sigmoid(h Wo + bo): columns 0..255 by tcgen05 (activation hi + lo times fp16
//   weights, 2 MMAs: measured max 0.012 dB), column 256 as an fp32 dot product on the CUDA cores.
// Both: 16 epilogue warps (4 threads per frame) + 1 MMA / loader warp, A operand written straight to TMEM.
// ---------------------------------------------------------------------------------------------------
constexpr int STEM_PART = 4 * 128 * 128;             // one precision part: 4 K-chunks x [128 rows x 128 B]
constexpr int STEM_AUX = 2 * STEM_PART;              // fp32: b0[128], W0[256][128 half]
constexpr int STEM_IMG_BYTES = STEM_AUX + 260 * 4;      // b0[128], W0[256][half 128], 1 / s0 (+ pad)
constexpr int STEM_STAGE_LD = TILE + 1;              // transposed input staging [64 cols][129]
constexpr int HEAD_W = 4 * 256 * 128;                // hi only: 4 K-chunks x [256 rows x 128 B]
constexpr int HEAD_IMG_BYTES = HEAD_W + (260 + 256) * 4;   // fp32: bo[257 (+3 pad)], Wo[:, 256]
constexpr int HEAD_W_PAD = 134144;                   // staging buffer starts here (past the image)
static_assert(HEAD_W_PAD >= HEAD_IMG_BYTES, "head staging buffer overlaps the weight image");
constexpr int HEAD_STAGE_LD = 33;

struct StemArgs {
  const unsigned char* img;
  const float* mag;        // [B, T, 257]
  float* h;                // tiled residual buffer (receives the pre-activation z)
  float2* stats;           // [n_tiles * 128][8] (mean_i, M2_i) of 32-channel parts
  int T, tiles_per_utt, n_tiles, half, n_feat;
};

__global__ void __launch_bounds__(STAGE_THREADS, 1) stem_umma_kernel(const StemArgs p) {
  extern __shared__ unsigned char smem_raw[];
  __shared__ __align__(8) uint64_t bar_w, bar_a, bar_d;
  __shared__ uint32_t tmem_slot;
  unsigned char* sW = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const float* sAux = reinterpret_cast<const float*>(sW + STEM_AUX);
  float* stage = reinterpret_cast<float*>(sW + STEM_IMG_BYTES);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  constexpr uint32_t COL_A_HI = 0, COL_A_LO = 128, COL_D = 256;
  if (warp == EPI_WARPS) tmem_alloc(&tmem_slot, 512);
  if (tid == 0) { mbar_init(&bar_w, 1); mbar_init(&bar_a, EPI_THREADS); mbar_init(&bar_d, 1); fence_mbar_init(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (tmem_slot != 0) __trap();
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  if (warp == EPI_WARPS) {
    if (elect_one()) {
      mbar_arrive_expect_tx(&bar_w, STEM_IMG_BYTES);
      for (int off = 0; off < STEM_IMG_BYTES; off += 16384) {
        const int n = STEM_IMG_BYTES - off < 16384 ? STEM_IMG_BYTES - off : 16384;
        bulk_g2s(sW + off, p.img + off, n, &bar_w);
      }
    }
    __syncwarp();
    mbar_wait(&bar_w, 0);
    asm volatile("griddepcontrol.wait;" ::: "memory");      // everything the previous launch wrote is visible
    const uint32_t w_hi = smem_u32(sW), w_lo = w_hi + STEM_PART;
    constexpr uint32_t id128 = make_idesc_f16(TILE, 128);
    uint32_t ph = 0;
    for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x) {
      mbar_wait(&bar_a, ph); tc_fence_after();
      uint32_t acc = 0;
#pragma unroll
      for (int part = 0; part < 3; ++part) {
        const uint32_t a0 = part == 1 ? COL_A_LO : COL_A_HI, w0 = part == 2 ? w_lo : w_hi;
#pragma unroll
        for (int ks = 0; ks < 16; ++ks) {
          mma_ts_elect(COL_D, a0 + 8 * ks, make_smem_desc_sw128(w0 + (ks >> 2) * 128 * 128 + (ks & 3) * 32), id128, acc);
          acc = 1;
        }
      }
      mma_commit_elect(&bar_d);
      ph ^= 1;
    }
  } else {
    mbar_wait(&bar_w, 0);
    asm volatile("griddepcontrol.wait;" ::: "memory");      // everything the previous launch wrote is visible
    const int qd = warp >> 2, row = (warp & 3) * 32 + lane;
    const uint32_t lane_addr = (uint32_t)((warp & 3) * 32) << 16;
    uint32_t ph = 0;
    for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x) {
      const int b = tile / p.tiles_per_utt, t0 = (tile - b * p.tiles_per_utt) * TILE;
      const int t = t0 + row;
      const bool valid = t < p.T;
      const float* mb = p.mag + (size_t)b * p.T * p.n_feat;
      // ---- A operand: magnitudes of bins 0..255 as fp16 hi | lo, through a transposed shared-memory stage
      for (int kq = 0; kq < 4; ++kq) {
        {   // 16 independent loads per thread in flight, then the transposed stores
          float x[16];
#pragma unroll
          for (int u = 0; u < 16; ++u) {
            const int i = tid + u * EPI_THREADS, r = i >> 6, c = i & 63;
            x[u] = (t0 + r < p.T) ? __ldcs(mb + (size_t)(t0 + r) * p.n_feat + 64 * kq + c) : 0.0f;
          }
#pragma unroll
          for (int u = 0; u < 16; ++u) {
            const int i = tid + u * EPI_THREADS, r = i >> 6, c = i & 63;
            stage[c * STEM_STAGE_LD + r] = x[u];
          }
        }
        epi_barrier();
        uint32_t hi[8], lo[8];
#pragma unroll
        for (int j = 0; j < 8; ++j)
          split_h2(stage[(16 * qd + 2 * j) * STEM_STAGE_LD + row], stage[(16 * qd + 2 * j + 1) * STEM_STAGE_LD + row], hi[j], lo[j]);
        tmem_st8(lane_addr + COL_A_HI + 32 * kq + 8 * qd, hi);
        tmem_st8(lane_addr + COL_A_LO + 32 * kq + 8 * qd, lo);
        epi_barrier();
      }
      const float x256 = valid ? __ldg(mb + (size_t)t * p.n_feat + 256) : 0.0f;
      tmem_wait_st(); tc_fence_before(); mbar_arrive(&bar_a);
      // ---- z = acc + b0 + mag[256] W0[256, :]; partial statistics; tiled store
      mbar_wait(&bar_d, ph); tc_fence_after();
      float z[32];
      tmem_ld32(lane_addr + COL_D + 32 * qd, z); tmem_wait_ld();
      tc_fence_before();
      float s = 0.0f;
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        z[j] = valid ? fmaf(z[j], sAux[256], fmaf(x256, sAux[128 + 32 * qd + j], sAux[32 * qd + j])) : 0.0f;      // sAux[256] = 1 / s0
        s += z[j];
      }
      const float mean_i = s * (1.0f / 32.0f);
      float m2 = 0.0f;
#pragma unroll
      for (int j = 0; j < 32; ++j) { const float d = z[j] - mean_i; m2 = fmaf(d, d, m2); }
      p.stats[((size_t)tile * TILE + row) * 8 + p.half * 4 + qd] = make_float2(mean_i, m2);
      float* hrow = p.h + (size_t)tile * (TILE * 256) + row * 4 + (size_t)(32 * p.half + 8 * qd) * (TILE * 4);
#pragma unroll
      for (int q = 0; q < 8; ++q)
        *reinterpret_cast<float4*>(hrow + (size_t)q * (TILE * 4)) = make_float4(z[4 * q], z[4 * q + 1], z[4 * q + 2], z[4 * q + 3]);
      ph ^= 1;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == EPI_WARPS) tmem_dealloc(0, 512);
}

struct HeadArgs {
  const unsigned char* img;
  const float* h;          // tiled residual buffer
  float* xbar;             // [B, T, 257]
  int T, tiles_per_utt, n_tiles, n_outp;
};

__global__ void __launch_bounds__(STAGE_THREADS, 1) head_umma_kernel(const HeadArgs p) {
  extern __shared__ unsigned char smem_raw[];
  __shared__ __align__(8) uint64_t bar_w, bar_a, bar_d[2];
  __shared__ uint32_t tmem_slot;
  __shared__ float dot[NSPLIT * TILE], rmax[NSPLIT * TILE];
  unsigned char* sW = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const float* sBias = reinterpret_cast<const float*>(sW + HEAD_W);          // [260]
  const float* sLast = sBias + 260;                                           // Wo[:, 256]
  float* stage = reinterpret_cast<float*>(sW + HEAD_W_PAD);                   // [4][128][33]
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  constexpr uint32_t COL_A_HI = 0, COL_A_LO = 128, COL_D = 256;
  if (warp == EPI_WARPS) tmem_alloc(&tmem_slot, 512);
  if (tid == 0) { mbar_init(&bar_w, 1); mbar_init(&bar_a, EPI_THREADS); mbar_init(&bar_d[0], 1); mbar_init(&bar_d[1], 1); fence_mbar_init(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (tmem_slot != 0) __trap();
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  if (warp == EPI_WARPS) {
    if (elect_one()) {
      mbar_arrive_expect_tx(&bar_w, HEAD_IMG_BYTES);
      for (int off = 0; off < HEAD_IMG_BYTES; off += 16384) {
        const int n = HEAD_IMG_BYTES - off < 16384 ? HEAD_IMG_BYTES - off : 16384;
        bulk_g2s(sW + off, p.img + off, n, &bar_w);
      }
    }
    __syncwarp();
    mbar_wait(&bar_w, 0);
    asm volatile("griddepcontrol.wait;" ::: "memory");      // everything the previous launch wrote is visible
    const uint32_t w_hi = smem_u32(sW);
    constexpr uint32_t id128 = make_idesc_f16(TILE, 128);
    uint32_t ph = 0;
    for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x) {
      {
        const int nt = tile + gridDim.x;
        if (nt < p.n_tiles && lane < 8) prefetch_l2(reinterpret_cast<const char*>(p.h + (size_t)nt * (TILE * 256)) + lane * 16384, 16384);
        __syncwarp();
      }
      mbar_wait(&bar_a, ph); tc_fence_after();
#pragma unroll
      for (int nh = 0; nh < 2; ++nh) {      // output columns 128 nh .. +128, committed separately
        uint32_t acc = 0;
#pragma unroll
        for (int part = 0; part < 2; ++part) {
          const uint32_t a0 = part == 1 ? COL_A_LO : COL_A_HI;
#pragma unroll
          for (int ks = 0; ks < 16; ++ks) {
            mma_ts_elect(COL_D + 128 * nh, a0 + 8 * ks, make_smem_desc_sw128(w_hi + (ks >> 2) * 256 * 128 + nh * 128 * 128 + (ks & 3) * 32), id128, acc);
            acc = 1;
          }
        }
        mma_commit_elect(&bar_d[nh]);
      }
      ph ^= 1;
    }
  } else {
    mbar_wait(&bar_w, 0);
    asm volatile("griddepcontrol.wait;" ::: "memory");      // everything the previous launch wrote is visible
    const int qd = warp >> 2, row = (warp & 3) * 32 + lane;
    const uint32_t lane_addr = (uint32_t)((warp & 3) * 32) << 16;
    uint32_t ph = 0;
    for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x) {
      const int b = tile / p.tiles_per_utt, t0 = (tile - b * p.tiles_per_utt) * TILE;
      const int t = t0 + row;
      const bool valid = t < p.T;
      const float* hrow = p.h + (size_t)tile * (TILE * 256) + row * 4 + (size_t)(16 * qd) * (TILE * 4);
      // ---- A operand: h (raw residual sum, tcn.py:158-159) as fp16 hi | lo; column 256 as an fp32 dot product.
      // fp16 range guard: the row is scaled by an exact power of two that brings its largest magnitude into [1, 2) before the
      // split and the accumulator is scaled back, so a residual stream of 1e-4 or 1e5 neither underflows nor overflows fp16.
      float4 hv[16];
#pragma unroll
      for (int q = 0; q < 16; ++q) hv[q] = *reinterpret_cast<const float4*>(hrow + (size_t)q * (TILE * 4));
      float mx = 0.0f;
#pragma unroll
      for (int q = 0; q < 16; ++q) mx = fmaxf(fmaxf(mx, fmaxf(fabsf(hv[q].x), fabsf(hv[q].y))), fmaxf(fabsf(hv[q].z), fabsf(hv[q].w)));
      rmax[qd * TILE + row] = mx;
      epi_barrier();
      mx = fmaxf(fmaxf(rmax[row], rmax[TILE + row]), fmaxf(rmax[2 * TILE + row], rmax[3 * TILE + row]));
      const uint32_t ex = min(max((__float_as_uint(mx) >> 23) & 0xFFu, 1u), 253u);      // biased exponent of the row maximum (1 if the row is zero)
      const float sc = __uint_as_float((254u - ex) << 23), isc = __uint_as_float(ex << 23) * sBias[257];      // 2^-(ex-127); its inverse times 1 / s_o (Wo is stored as Wo * s_o)
      float d256 = 0.0f;
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        uint32_t hi[8], lo[8];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const float4 h4 = hv[4 * g + q];
          const float* wl = sLast + 64 * qd + 16 * g + 4 * q;
          d256 = fmaf(h4.x, wl[0], d256); d256 = fmaf(h4.y, wl[1], d256);
          d256 = fmaf(h4.z, wl[2], d256); d256 = fmaf(h4.w, wl[3], d256);
          split_h2(h4.x * sc, h4.y * sc, hi[2 * q], lo[2 * q]);
          split_h2(h4.z * sc, h4.w * sc, hi[2 * q + 1], lo[2 * q + 1]);
        }
        tmem_st8(lane_addr + COL_A_HI + 32 * qd + 8 * g, hi);
        tmem_st8(lane_addr + COL_A_LO + 32 * qd + 8 * g, lo);
      }
      dot[qd * TILE + row] = d256;
      tmem_wait_st(); tc_fence_before(); mbar_arrive(&bar_a);
      epi_barrier();
      float* xb = p.xbar + (size_t)b * p.T * p.n_outp;
      if (qd == 0 && valid) {
        const float z = dot[row] + dot[TILE + row] + dot[2 * TILE + row] + dot[3 * TILE + row] + sBias[256];
        xb[(size_t)t * p.n_outp + 256] = 1.0f / (1.0f + expf(-z));
      }
      // ---- x_bar = sigmoid(acc + bo), staged through shared memory for row-major coalesced stores
#pragma unroll
      for (int nh = 0; nh < 2; ++nh) {
        mbar_wait(&bar_d[nh], ph); tc_fence_after();
        float z[32];
        tmem_ld32(lane_addr + COL_D + 128 * nh + 32 * qd, z); tmem_wait_ld();
        tc_fence_before();
        float* st = stage + ((size_t)qd * TILE + row) * HEAD_STAGE_LD;
#pragma unroll
        for (int j = 0; j < 32; ++j) st[j] = 1.0f / (1.0f + expf(-fmaf(z[j], isc, sBias[128 * nh + 32 * qd + j])));
        epi_barrier();
        for (int i = tid; i < TILE * 128; i += EPI_THREADS) {
          const int r = i >> 7, c = i & 127;
          if (t0 + r < p.T) __stcs(xb + (size_t)(t0 + r) * p.n_outp + 128 * nh + c, stage[((size_t)(c >> 5) * TILE + r) * HEAD_STAGE_LD + (c & 31)]);
        }
        epi_barrier();
      }
      ph ^= 1;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == EPI_WARPS) tmem_dealloc(0, 512);
}

// ---------------------------------------------------------------------------------------------------
// Host side
// ---------------------------------------------------------------------------------------------------
// Packs W [K][N] (fp32) as fp16 hi / lo UMMA B images and returns the column sums of the EFFECTIVE weights
// (hi, or hi + lo) in colsum[N]: that is what the deferred normalisation has to subtract.
static void pack_b_sw128(unsigned char* hi, unsigned char* lo, int N, int K, const float* W /* [K][N] */, bool split,
                         float* colsum, float scale = 1.0f) {
  std::vector<double> cs(N, 0.0);
  for (int n = 0; n < N; ++n)
    for (int k = 0; k < K; ++k) {
      const int c = k >> 6, kk = k & 63, u = kk >> 3, e = kk & 7;
      const size_t off = (size_t)c * N * 128 + (size_t)(n >> 3) * 1024 + (n & 7) * 128 + ((u ^ (n & 7)) * 16) + e * 2;
      const float w = W[(size_t)k * N + n] * scale;
      const __half h = __float2half_rn(w);
      const __half l = __float2half_rn(w - __half2float(h));
      memcpy(hi + off, &h, 2);
      memcpy(lo + off, &l, 2);
      cs[n] += (double)__half2float(h) + (split ? (double)__half2float(l) : 0.0);
    }
  if (colsum)
    for (int n = 0; n < N; ++n) colsum[n] = (float)cs[n];
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
    float* aux = reinterpret_cast<float*>(base + IMG_BIAS);
    const bool split = c.precision == DXI_PREC_F16X3;
    if (s >= 1) {
      const int li = 2 + 3 * (s - 1);
      const float* k2 = net.host_tensor(li + 1, "kernel")->data();
      const float* k3 = net.host_tensor(li + 2, "kernel")->data();
      const float s2 = weight_pow2_scale(k2, 192 * 64, false), s3 = weight_pow2_scale(k3, 64 * 256, false);
      pack_b_sw128(base + IMG_W2, base + IMG_PART + IMG_W2, 64, 192, k2, split, nullptr, s2);        // [3*64][64]
      pack_b_sw128(base + IMG_W3, base + IMG_PART + IMG_W3, 256, 64, k3, split, aux + OFF_CS3, s3);  // [64][256]
      aux[OFF_IS + 1] = 1.0f / s2; aux[OFF_IS + 2] = 1.0f / s3;
      memcpy(aux + OFF_B2, net.host_tensor(li + 1, "bias")->data(), 64 * 4);
      memcpy(aux + OFF_B3, net.host_tensor(li + 2, "bias")->data(), 256 * 4);
    }
    if (s < c.n_blocks) {
      const int li = 2 + 3 * s;
      const float* k1 = net.host_tensor(li, "kernel")->data();
      const float s1 = weight_pow2_scale(k1, 256 * 64, false);
      pack_b_sw128(base + IMG_W1, base + IMG_PART + IMG_W1, 64, 256, k1, split, aux + OFF_CS1, s1);      // [256][64]
      aux[OFF_IS] = 1.0f / s1;
      memcpy(aux + OFF_B1, net.host_tensor(li, "bias")->data(), 64 * 4);
    }
  }
  // ---- stem (two 128-column halves) and output layer images
  const size_t off_stem = align_up(img.size(), 1024), stem_stride = align_up(STEM_IMG_BYTES, 1024);
  const size_t off_head = off_stem + 2 * stem_stride;
  img.resize(off_head + align_up(HEAD_IMG_BYTES, 1024), 0);
  {
    const float* W0 = net.host_tensor(0, "kernel")->data();      // [1][257][256]
    const float* b0 = net.host_tensor(0, "bias")->data();
    std::vector<float> sub((size_t)256 * 128);
    for (int half = 0; half < 2; ++half) {
      unsigned char* base = img.data() + off_stem + half * stem_stride;
      for (int k = 0; k < 256; ++k)
        for (int n = 0; n < 128; ++n) sub[(size_t)k * 128 + n] = W0[(size_t)k * 256 + 128 * half + n];
      const float s0 = weight_pow2_scale(W0, (size_t)256 * 256, false);
      pack_b_sw128(base, base + STEM_PART, 128, 256, sub.data(), true, nullptr, s0);
      float* aux = reinterpret_cast<float*>(base + STEM_AUX);
      for (int n = 0; n < 128; ++n) { aux[n] = b0[128 * half + n]; aux[128 + n] = W0[(size_t)256 * 256 + 128 * half + n]; }
      aux[256] = 1.0f / s0;
    }
    // stage 0 applies the stem's LayerNorm scale: gamma rides in its (otherwise unused) b3 slot
    memcpy(reinterpret_cast<float*>(img.data() + IMG_BIAS) + OFF_B3, net.host_tensor(1, "gamma")->data(), 256 * 4);
    const int lo_ = 2 + 3 * c.n_blocks;
    const float* Wo = net.host_tensor(lo_, "kernel")->data();    // [1][256][257]
    const float* bo = net.host_tensor(lo_, "bias")->data();
    std::vector<float> sq((size_t)256 * 256);
    std::vector<unsigned char> scratch(HEAD_W);
    for (int k = 0; k < 256; ++k)
      for (int n = 0; n < 256; ++n) sq[(size_t)k * 256 + n] = Wo[(size_t)k * 257 + n];
    unsigned char* hb = img.data() + off_head;
    const float so = weight_pow2_scale(sq.data(), sq.size(), false);
    pack_b_sw128(hb, scratch.data(), 256, 256, sq.data(), false, nullptr, so);
    float* aux = reinterpret_cast<float*>(hb + HEAD_W);
    for (int n = 0; n < 257; ++n) aux[n] = bo[n];
    aux[257] = 1.0f / so;
    for (int k = 0; k < 256; ++k) aux[260 + k] = Wo[(size_t)k * 257 + 256];
  }
  net.umma_stage_offset = {off_stem, off_stem + stem_stride, off_head};
  if (net.d_umma) { cudaFree(net.d_umma); net.d_umma = nullptr; }
  DXI_CUDA(cudaMalloc(&net.d_umma, img.size()));
  DXI_CUDA(cudaMemcpyAsync(net.d_umma, img.data(), img.size(), cudaMemcpyHostToDevice, st));
  DXI_CUDA(cudaStreamSynchronize(st));   // img is a local
  net.umma_bytes = img.size();
  if (resnet_chain_supported(net)) return resnet_chain_prepare(net, st);
  return DXI_OK;
}

static bool env_flag(const char* name) { const char* v = getenv(name); return v && *v && *v != '0'; }
#ifdef DXI_ENABLE_DEBUG
static thread_local long long* g_dbg_clocks = nullptr;
static thread_local int g_dbg_stage = -1, g_dbg_flags = 0, g_dbg_stop_after = -1;
#else
constexpr int g_dbg_stop_after = -1;
#endif

int64_t resnet_umma_workspace_bytes(const dxi_net& net, int B, int T) {
  const int tiles = (T + TILE - 1) / TILE;
  if (resnet_chain_fused(net))      // one launch, residual tile in tensor memory: only the tile flags, the work counter and the halo records
    return (int64_t)(512 + resnet_chain_extra_workspace(net, B, tiles));
  const size_t Ts = (size_t)tiles * TILE + 2 * C1_PAD;
  const size_t h_bytes = (size_t)B * tiles * TILE * 256 * 4;
  const size_t c1_bytes = align_up((size_t)B * 2 * 8 * Ts * 16, 256);
  const size_t stats_bytes = (size_t)B * tiles * TILE * 8 * sizeof(float2);
  const size_t flag_bytes = align_up((size_t)B * tiles * sizeof(int), 256);
  // behind the stem statistics: the depth-first path's flags / halo records, or the stage-per-launch path's per-row operand scales
  const size_t chain_bytes = 256 + (resnet_chain_supported(net) ? resnet_chain_extra_workspace(net, B, tiles) : 2 * (size_t)B * tiles * TILE * sizeof(float));
  return (int64_t)(256 + h_bytes + 2 * c1_bytes + flag_bytes + stats_bytes + chain_bytes);
}

// One group of whole utterances through stem -> 41 stages -> output layer; every group uses the same workspace region.
static int resnet_umma_group(const dxi_net& net, const float* mag, int B, int T, float* xbar, void* ws, cudaStream_t st, bool dbg) {
  const dxi_net_cfg& c = net.cfg;
  const int tiles = (T + TILE - 1) / TILE;
  const int Ts = tiles * TILE + 2 * C1_PAD;
  const size_t h_bytes = (size_t)B * tiles * TILE * 256 * 4;
  const size_t c1_bytes = align_up((size_t)B * 2 * 8 * Ts * 16, 256);
  unsigned char* base = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(ws) + 255) & ~(uintptr_t)255);
  float* h = reinterpret_cast<float*>(base);
  __half* c1[2] = {reinterpret_cast<__half*>(base + h_bytes), reinterpret_cast<__half*>(base + h_bytes + c1_bytes)};
  const size_t flag_bytes = align_up((size_t)B * tiles * sizeof(int), 256);
  int* flags = reinterpret_cast<int*>(base + h_bytes + 2 * c1_bytes);
  float2* stem_stats = reinterpret_cast<float2*>(base + h_bytes + 2 * c1_bytes + flag_bytes);
  if (resnet_chain_fused(net)) {      // first layer, blocks and output layer in ONE launch (tcn_chain.cu)
    int n_sm_ = 148;
    { int dev = 0; cudaGetDevice(&dev); cudaDeviceGetAttribute(&n_sm_, cudaDevAttrMultiProcessorCount, dev); }
    DXI_CUDA(cudaMemsetAsync(base, 0, align_up((size_t)B * tiles * sizeof(int) + 64, 256), st));      // tile flags + work counter
    return resnet_chain_network(net, mag, xbar, B, T, base, n_sm_, st);
  }
  const bool chain = resnet_chain_supported(net);      // causal padding: depth-first residual blocks (tcn_chain.cu)
  const size_t stats_bytes = (size_t)B * tiles * TILE * 8 * sizeof(float2);
  unsigned char* chain_ws = reinterpret_cast<unsigned char*>(align_up(reinterpret_cast<uintptr_t>(base + h_bytes + 2 * c1_bytes + flag_bytes + stats_bytes), 256));
  if (chain) {
    // per-tile published-block counters + the work-item counter (the halo records behind them need no initialisation)
    DXI_CUDA(cudaMemsetAsync(chain_ws, 0, align_up((size_t)B * tiles * sizeof(int) + 64, 256), st));
  } else {
    // zero padding rows of the c1 planes (and everything else in them) and the per-tile stage counters
    DXI_CUDA(cudaMemsetAsync(c1[0], 0, 2 * c1_bytes + flag_bytes, st));
  }
  static const bool no_flags = env_flag("DXI_TCN_NO_FLAGS");      // A/B switch: every stage waits for the whole previous launch

  const int n_tiles = B * tiles;
  int n_sm = 148;
  { int dev = 0; cudaGetDevice(&dev); cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev); }
  const int grid = n_tiles < n_sm ? n_tiles : n_sm;
  const unsigned char* images = reinterpret_cast<const unsigned char*>(net.d_umma);
  // ---- stem (tcgen05, two 128-column halves) -> pre-activation z in the tiled buffer + partial row statistics
  {
    const size_t smem_stem = STEM_IMG_BYTES + 64 * STEM_STAGE_LD * sizeof(float) + 1024;
    DXI_CUDA(cudaFuncSetAttribute(stem_umma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_stem));
    ProfScope prof("tcn_stem", st, 2);
    for (int half = 0; half < 2; ++half) {
      StemArgs a{images + net.umma_stage_offset[half], mag, h, stem_stats, T, tiles, n_tiles, half, c.n_feat};
      DXI_CUDA(launch_pdl(stem_umma_kernel, grid, STAGE_THREADS, smem_stem, st, a));
      DXI_LAUNCHED("stem_umma_kernel");
    }
  }
  if (chain) {
    if (int rc = resnet_chain_blocks(net, h, stem_stats, B, T, chain_ws, n_sm, st)) return rc;
  } else {
  // ---- 41 tensor-core stages
  const bool split = c.precision == DXI_PREC_F16X3;
  const size_t smem = IMG_BYTES + 1024;
  DXI_CUDA(cudaFuncSetAttribute(tcn_stage_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  DXI_CUDA(cudaFuncSetAttribute(tcn_stage_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int nd = n_dilations(c.max_d_rate);
  {
  ProfScope prof_stages("tcn_stage", st, c.n_blocks + 1);
  for (int s = 0; s <= c.n_blocks; ++s) {
    if (g_dbg_stop_after >= 0 && s > g_dbg_stop_after) break;
    StageArgs a{};
    a.img = images + (size_t)s * IMG_BYTES;
    a.stem_stats = s == 0 ? stem_stats : nullptr;
    a.h = h;
    a.c1_in = c1[(s + 1) & 1];
    a.c1_out = c1[s & 1];
    a.T = T; a.tiles_per_utt = tiles; a.n_tiles = n_tiles; a.Ts = Ts;
    a.has_back = s >= 1; a.has_front = s < c.n_blocks;
    a.reverse = (s & 1) == 0 && !env_flag("DXI_TCN_NO_REVERSE");      // the stem leaves the last tiles in L2, the output layer starts on the first
#ifdef DXI_ENABLE_DEBUG
    a.dbg = (dbg && s == g_dbg_stage) ? g_dbg_clocks : nullptr;
    a.dbg_cta = (dbg && g_dbg_stage == 255 && g_dbg_clocks) ? g_dbg_clocks + (size_t)s * grid * 8 : nullptr;      // stage 255: CTA timelines of all stages
    a.dbg_flags = g_dbg_flags;
#endif
    a.flags = (no_flags || (dbg && g_dbg_stop_after >= 0)) ? nullptr : flags;
    float* row_scale = reinterpret_cast<float*>(chain_ws);      // [2][n_tiles * 128], alternating between stages
    a.sc_in = s >= 1 ? row_scale + (size_t)((s + 1) & 1) * n_tiles * TILE : nullptr;
    a.sc_out = row_scale + (size_t)(s & 1) * n_tiles * TILE;
    a.sc0 = resnet_first_operand_scale(net);
    a.wait_val = s;               // stage s-1 has published a tile when its counter reaches s
    const int d = s >= 1 ? 1 << ((s - 1) % nd) : 1;
    if (c.padding == DXI_PAD_CAUSAL) { a.shift0 = 2 * d; a.shift1 = d; a.shift2 = 0; }       // tap j reads t-(2-j)d
    else                             { a.shift0 = d;     a.shift1 = 0; a.shift2 = -d; }      // tap j reads t+(j-1)d
    if (split) DXI_CUDA(launch_pdl(tcn_stage_kernel<true>, grid, TCN_THREADS, smem, st, a));
    else       DXI_CUDA(launch_pdl(tcn_stage_kernel<false>, grid, TCN_THREADS, smem, st, a));
    DXI_LAUNCHED("tcn_stage_kernel");
  }
  }
  }
  // ---- output layer (tcgen05 + one fp32 column): tiled h -> sigmoid(W h + b), row-major x_bar
  {
    const size_t smem_head = HEAD_W_PAD + (size_t)NSPLIT * TILE * HEAD_STAGE_LD * sizeof(float) + 1024;
    DXI_CUDA(cudaFuncSetAttribute(head_umma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_head));
    HeadArgs a{images + net.umma_stage_offset[2], h, xbar, T, tiles, n_tiles, c.n_outp};
    ProfScope prof("tcn_head", st, 1);
    DXI_CUDA(launch_pdl(head_umma_kernel, grid, STAGE_THREADS, smem_head, st, a));
    DXI_LAUNCHED("head_umma_kernel");
  }
  return DXI_OK;
}

int resnet_umma_forward(const dxi_net& net, const float* mag, int B, int T, float* xbar, void* ws, size_t ws_bytes,
                        cudaStream_t st) {
  const dxi_net_cfg& c = net.cfg;
  if ((int64_t)ws_bytes < resnet_umma_workspace_bytes(net, B, T)) { set_error("workspace too small"); return DXI_E_NOMEM; }
  // DXI_TCN_GROUP_UTTS=n (tuning experiment): run the whole network over groups of n utterances, one group after the other, so
  // that a group's residual stream and c1 planes stay closer to L2; 0 / unset: the whole batch at once.
  static const int group = [] { const char* v = getenv("DXI_TCN_GROUP_UTTS"); return v && *v ? atoi(v) : 0; }();
  if (group <= 0 || group >= B) return resnet_umma_group(net, mag, B, T, xbar, ws, st, true);
  for (int b0 = 0; b0 < B; b0 += group) {
    const int Bg = B - b0 < group ? B - b0 : group;
    if (int rc = resnet_umma_group(net, mag + (size_t)b0 * T * c.n_feat, Bg, T, xbar + (size_t)b0 * T * c.n_outp, ws, st, b0 == 0)) return rc;
  }
  return DXI_OK;
}

}  // namespace dxi

#ifdef DXI_ENABLE_DEBUG
// Debug aid: the epilogue of stage `stage` of the following forward calls (this thread) writes 16 clock64
// stamps per tile into dev_buf[n_tiles * 16]; pass nullptr to switch it off.
extern "C" DXI_API void dxi_debug_tcn_clocks(long long* dev_buf, int stage) {
  dxi::g_dbg_clocks = dev_buf;
  dxi::g_dbg_stage = dev_buf ? (stage & 0xff) : -1;
  dxi::g_dbg_flags = dev_buf ? (stage >> 8) : 0;     // bits 8.. : tuning-experiment flags (results are then wrong)
}

// Debug aid: run only the stem and stages 0..stage of subsequent forwards of this thread (-1 = everything), so
// that the residual / c1 buffers in the caller's workspace can be inspected between stages.
extern "C" DXI_API void dxi_debug_tcn_stop_after(int stage) { dxi::g_dbg_stop_after = stage; }
#endif
