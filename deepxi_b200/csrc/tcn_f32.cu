// ResNetV2 forward on the fp32 CUDA cores (precision mode DXI_PREC_F32, "exact mode"), and the two sibling architectures of
// deepxi/network/tcn.py that only exist in this mode:
//   ResNet   (tcn.py:17-114)   first layer Conv1D(no bias) -> LN(gamma, beta) -> ReLU; unit LN(gamma, beta) -> ReLU -> Conv1D
//                              (bias only in the third unit of a block)
//   ResNetV3 (tcn.py:227-245)  ResNetV2 with the first layer Conv1D+b -> ReLU -> LN(no affine)
//
// One kernel per convolutional unit of deepxi/network/tcn.py:116-225:
//   stem  (tcn.py:166-180)  Conv1D(256,1)+b -> LayerNorm(scale gamma, no centre) -> ReLU
//   unit  (tcn.py:218-223)  ReLU -> LayerNorm(no affine, eps 1e-6) -> Conv1D(k, dilation d, causal|same)+b
//   block (tcn.py:182-197)  three units (1x1 256->64, k=3 64->64 dilated, 1x1 64->256) + residual add
//   head  (tcn.py:158-161)  Conv1D(257,1)+b on the raw residual sum -> sigmoid
// Each CTA owns TM consecutive frames of one utterance: it normalises the input rows it needs (one
// row per tap) into shared memory, multiplies by the weights streamed through shared memory in
// K-chunks, then runs a row-wise epilogue with coalesced stores.  This path exists for bit-level
// fidelity (fp32 everywhere, two-pass LayerNorm statistics as Keras does); the tcgen05 path in
// tcn_umma.cu is the fast one.
#include "common.cuh"
#include "net.cuh"

namespace dxi {

constexpr int TM = 64;       // frames per CTA
constexpr int KC = 16;       // K-chunk of the weight stream

enum { PRE_NONE = 0, PRE_RELU_LN = 1, PRE_LN_AFF_RELU = 2 };
enum { POST_BIAS = 0, POST_RESIDUAL = 1, POST_LN_SCALE_RELU = 2, POST_SIGMOID = 3, POST_RELU_LN = 4, POST_LN_AFF_RELU = 5 };

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

template <int CIN, int COUT, int TAPS, int PRE, int POST>
__global__ void __launch_bounds__(256) unit_f32_kernel(const float* __restrict__ in, const float* __restrict__ W,
                                                       const float* __restrict__ bias, const float* __restrict__ gamma,
                                                       const float* __restrict__ beta, const float* __restrict__ pre_gamma,
                                                       const float* __restrict__ pre_beta, const float* res, float* out, int T,
                                                       int tiles_per_utt, int d_rate, int causal) {
  constexpr int KT = TAPS * CIN;
  constexpr int LDA = KT + 1;
  constexpr int CPT = (COUT + 15) / 16;      // output columns per thread (interleaved by 16)
  constexpr int RM = TM / 16;                // rows per thread
  constexpr int LDO = COUT + 1;
  extern __shared__ __align__(16) float smem[];
  float* A = smem;                            // [TM][LDA]; reused as the output tile [TM][LDO]
  float* Wc = smem + TM * (LDA > LDO ? LDA : LDO);   // [KC][COUT]
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int b = blockIdx.x / tiles_per_utt, t0 = (blockIdx.x - b * tiles_per_utt) * TM;
  const float* in_b = in + (int64_t)b * T * CIN;

  // ---- phase 1: gather + (ReLU -> LayerNorm) the rows each tap needs
  for (int rj = warp; rj < TM * TAPS; rj += 8) {
    const int r = rj / TAPS, j = rj - r * TAPS;
    // causal: out[t] = sum_j W[j] x[t-(TAPS-1-j)d];  same: out[t] = sum_j W[j] x[t+(j-(TAPS-1)/2)d]
    const int shift = causal ? (TAPS - 1 - j) * d_rate : ((TAPS - 1) / 2 - j) * d_rate;
    const int ts = t0 + r - shift;
    float* dst = A + r * LDA + j * CIN;
    if (t0 + r >= T || ts < 0 || ts >= T) {   // zero padding is applied to the conv INPUT (after ReLU->LN)
      for (int c = lane; c < CIN; c += 32) dst[c] = 0.0f;
      continue;
    }
    const float* src = in_b + (int64_t)ts * CIN;
    if (PRE == PRE_NONE) {
      for (int c = lane; c < CIN; c += 32) dst[c] = src[c];
    } else {
      constexpr int PER = (CIN + 31) / 32;
      float x[PER];
      float s = 0.0f;
#pragma unroll
      for (int i = 0; i < PER; ++i) {
        const int c = lane + 32 * i;
        x[i] = c < CIN ? (PRE == PRE_RELU_LN ? fmaxf(src[c], 0.0f) : src[c]) : 0.0f;
        s += x[i];
      }
      const float mean = warp_sum(s) * (1.0f / CIN);
      float q = 0.0f;
#pragma unroll
      for (int i = 0; i < PER; ++i) {
        const int c = lane + 32 * i;
        const float dlt = c < CIN ? x[i] - mean : 0.0f;
        q += dlt * dlt;
      }
      const float inv = rsqrtf(warp_sum(q) * (1.0f / CIN) + 1e-6f);
      const float off = -mean * inv;            // tf.nn.batch_normalization: x*inv + (-mean*inv)
#pragma unroll
      for (int i = 0; i < PER; ++i) {
        const int c = lane + 32 * i;
        if (c >= CIN) continue;
        if (PRE == PRE_RELU_LN) dst[c] = fmaf(x[i], inv, off);
        else {      // LayerNormalization(gamma, beta) -> ReLU   (tcn.py:108-110)
          const float g = inv * __ldg(pre_gamma + c);
          dst[c] = fmaxf(fmaf(x[i], g, __ldg(pre_beta + c) - mean * g), 0.0f);
        }
      }
    }
  }
  // ---- phase 2: [TM x KT] x [KT x COUT]
  const int ty = tid >> 4, tx = tid & 15;
  float acc[RM][CPT];
#pragma unroll
  for (int i = 0; i < RM; ++i)
#pragma unroll
    for (int c = 0; c < CPT; ++c) acc[i][c] = 0.0f;
  for (int k0 = 0; k0 < KT; k0 += KC) {
    __syncthreads();
    for (int i = tid; i < KC * COUT; i += 256) {
      const int kk = i / COUT, c = i - kk * COUT;
      Wc[i] = (k0 + kk < KT) ? __ldg(W + (int64_t)(k0 + kk) * COUT + c) : 0.0f;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < KC; ++kk) {
      if (k0 + kk >= KT) break;
      float a[RM], w[CPT];
#pragma unroll
      for (int i = 0; i < RM; ++i) a[i] = A[(ty * RM + i) * LDA + k0 + kk];
#pragma unroll
      for (int c = 0; c < CPT; ++c) w[c] = (tx + 16 * c < COUT) ? Wc[kk * COUT + tx + 16 * c] : 0.0f;
#pragma unroll
      for (int i = 0; i < RM; ++i)
#pragma unroll
        for (int c = 0; c < CPT; ++c) acc[i][c] = fmaf(a[i], w[c], acc[i][c]);
    }
  }
  __syncthreads();
  float* O = A;
#pragma unroll
  for (int i = 0; i < RM; ++i)
#pragma unroll
    for (int c = 0; c < CPT; ++c) {
      const int col = tx + 16 * c;
      if (col < COUT) O[(ty * RM + i) * LDO + col] = acc[i][c] + (bias ? __ldg(bias + col) : 0.0f);
    }
  __syncthreads();
  // ---- phase 3: row-wise epilogue, one warp per row, coalesced stores
  for (int r = warp; r < TM; r += 8) {
    const int t = t0 + r;
    if (t >= T) continue;
    const int64_t o = ((int64_t)b * T + t) * COUT;
    const float* row = O + r * LDO;
    if (POST == POST_BIAS) {
      for (int c = lane; c < COUT; c += 32) out[o + c] = row[c];
    } else if (POST == POST_RESIDUAL) {
      for (int c = lane; c < COUT; c += 32) out[o + c] = res[o + c] + row[c];
    } else if (POST == POST_SIGMOID) {
      for (int c = lane; c < COUT; c += 32) out[o + c] = 1.0f / (1.0f + expf(-row[c]));
    } else {
      // POST_LN_SCALE_RELU: LayerNorm(scale=gamma, centre=False, eps 1e-6) -> ReLU   (ResNetV2, tcn.py:176-179)
      // POST_LN_AFF_RELU:   LayerNorm(gamma, beta) -> ReLU                           (ResNet,   tcn.py:73-76)
      // POST_RELU_LN:       ReLU -> LayerNorm(no affine)                             (ResNetV3, tcn.py:241-244)
      float s = 0.0f;
      for (int c = lane; c < COUT; c += 32) s += POST == POST_RELU_LN ? fmaxf(row[c], 0.0f) : row[c];
      const float mean = warp_sum(s) * (1.0f / COUT);
      float q = 0.0f;
      for (int c = lane; c < COUT; c += 32) { const float dlt = (POST == POST_RELU_LN ? fmaxf(row[c], 0.0f) : row[c]) - mean; q += dlt * dlt; }
      const float rs = rsqrtf(warp_sum(q) * (1.0f / COUT) + 1e-6f);
      for (int c = lane; c < COUT; c += 32) {
        if (POST == POST_RELU_LN) { out[o + c] = fmaf(fmaxf(row[c], 0.0f), rs, -mean * rs); continue; }
        const float inv = rs * __ldg(gamma + c);
        const float off = (POST == POST_LN_AFF_RELU ? __ldg(beta + c) : 0.0f) - mean * inv;
        out[o + c] = fmaxf(fmaf(row[c], inv, off), 0.0f);
      }
    }
  }
}

template <int CIN, int COUT, int TAPS, int PRE, int POST>
static int launch_unit(const float* in, const float* W, const float* bias, const float* gamma, const float* res,
                       float* out, int B, int T, int d_rate, int causal, cudaStream_t st, const float* beta = nullptr,
                       const float* pre_gamma = nullptr, const float* pre_beta = nullptr) {
  constexpr int KT = TAPS * CIN;
  constexpr int LD = (KT + 1) > (COUT + 1) ? (KT + 1) : (COUT + 1);
  const size_t smem = sizeof(float) * (TM * LD + KC * COUT);
  auto kern = unit_f32_kernel<CIN, COUT, TAPS, PRE, POST>;
  DXI_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int tiles = (T + TM - 1) / TM;
  kern<<<B * tiles, 256, smem, st>>>(in, W, bias, gamma, beta, pre_gamma, pre_beta, res, out, T, tiles, d_rate, causal);
  DXI_LAUNCHED("unit_f32_kernel");
  return DXI_OK;
}

int64_t resnet_f32_workspace_bytes(const dxi_net& net, int B, int T) {
  const int64_t rows = (int64_t)B * T;
  return sizeof(float) * rows * (net.cfg.d_model + 2 * net.cfg.d_f) + 256;
}

// Forward of the whole network; weights are the fp32 device copies made by dxi_net_finalize.
int resnet_f32_forward(const dxi_net& net, const float* mag, int B, int T, float* xbar, void* ws, size_t ws_bytes,
                       cudaStream_t st) {
  const dxi_net_cfg& c = net.cfg;
  if (!(c.n_feat == 257 && c.n_outp == 257 && c.d_model == 256 && c.d_f == 64 && c.k == 3)) {
    set_error("fp32 ResNetV2 path is instantiated for n_feat=n_outp=257, d_model=256, d_f=64, k=3");
    return DXI_E_INVALID;
  }
  if ((int64_t)ws_bytes < resnet_f32_workspace_bytes(net, B, T)) { set_error("workspace too small"); return DXI_E_NOMEM; }
  const int64_t rows = (int64_t)B * T;
  float* h = reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(ws) + 255) & ~(uintptr_t)255);
  float* u1 = h + rows * 256;
  float* u2 = u1 + rows * 64;
  const int causal = c.padding == DXI_PAD_CAUSAL;
  auto Wt = [&](int li, const char* var) { return net.dev_tensor(li, var); };
  int n_rates = 0;
  for (int m = c.max_d_rate; m > 0; m >>= 1) ++n_rates;          // log2(max_d_rate) + 1   (tcn.py:55-56, :156-157)
  int rc;
  if (net.kind == DXI_NET_RESNET) {
    // layer_with_weights: 0 conv (no bias), 1 LN; per block LN, conv, LN, conv, LN, conv(+bias); last conv (+bias)
    rc = launch_unit<257, 256, 1, PRE_NONE, POST_LN_AFF_RELU>(mag, Wt(0, "kernel"), nullptr, Wt(1, "gamma"), nullptr, h, B, T, 1,
                                                              causal, st, Wt(1, "beta"));
    if (rc) return rc;
    int li = 2;
    for (int i = 0; i < c.n_blocks; ++i) {
      const int d = 1 << (i % n_rates);
      rc = launch_unit<256, 64, 1, PRE_LN_AFF_RELU, POST_BIAS>(h, Wt(li + 1, "kernel"), nullptr, nullptr, nullptr, u1, B, T, 1, causal, st,
                                                               nullptr, Wt(li, "gamma"), Wt(li, "beta"));
      if (rc) return rc;
      rc = launch_unit<64, 64, 3, PRE_LN_AFF_RELU, POST_BIAS>(u1, Wt(li + 3, "kernel"), nullptr, nullptr, nullptr, u2, B, T, d, causal, st,
                                                              nullptr, Wt(li + 2, "gamma"), Wt(li + 2, "beta"));
      if (rc) return rc;
      rc = launch_unit<64, 256, 1, PRE_LN_AFF_RELU, POST_RESIDUAL>(u2, Wt(li + 5, "kernel"), Wt(li + 5, "bias"), nullptr, h, h, B, T, 1,
                                                                   causal, st, nullptr, Wt(li + 4, "gamma"), Wt(li + 4, "beta"));
      if (rc) return rc;
      li += 6;
    }
    return launch_unit<256, 257, 1, PRE_NONE, POST_SIGMOID>(h, Wt(li, "kernel"), Wt(li, "bias"), nullptr, nullptr, xbar, B, T, 1, causal, st);
  }
  int li;
  if (net.kind == DXI_NET_RESNETV3) {      // Conv1D+b -> ReLU -> LN(no affine): no LayerNorm weights, the blocks start at layer 1
    rc = launch_unit<257, 256, 1, PRE_NONE, POST_RELU_LN>(mag, Wt(0, "kernel"), Wt(0, "bias"), nullptr, nullptr, h, B, T, 1, causal, st);
    li = 1;
  } else {
    rc = launch_unit<257, 256, 1, PRE_NONE, POST_LN_SCALE_RELU>(mag, Wt(0, "kernel"), Wt(0, "bias"), Wt(1, "gamma"),
                                                               nullptr, h, B, T, 1, causal, st);
    li = 2;
  }
  if (rc) return rc;
  for (int i = 0; i < c.n_blocks; ++i) {
    const int d = 1 << (i % n_rates);
    rc = launch_unit<256, 64, 1, PRE_RELU_LN, POST_BIAS>(h, Wt(li, "kernel"), Wt(li, "bias"), nullptr, nullptr, u1, B, T, 1, causal, st);
    if (rc) return rc;
    rc = launch_unit<64, 64, 3, PRE_RELU_LN, POST_BIAS>(u1, Wt(li + 1, "kernel"), Wt(li + 1, "bias"), nullptr, nullptr, u2, B, T, d, causal, st);
    if (rc) return rc;
    rc = launch_unit<64, 256, 1, PRE_RELU_LN, POST_RESIDUAL>(u2, Wt(li + 2, "kernel"), Wt(li + 2, "bias"), nullptr, h, h, B, T, 1, causal, st);
    if (rc) return rc;
    li += 3;
  }
  return launch_unit<256, 257, 1, PRE_NONE, POST_SIGMOID>(h, Wt(li, "kernel"), Wt(li, "bias"), nullptr, nullptr, xbar, B, T, 1, causal, st);
}

}  // namespace dxi
