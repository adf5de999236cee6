/*
 * deepxi_b200.h -- C ABI of libdeepxi_b200.so: the Deep Xi inference hot path on B200 (sm_100a).
 *
 * The reference (golfbears/DeepXi) has no native / FFI interface: its boundary is the Python API of
 * deepxi/sig.py, deepxi/map.py, deepxi/gain.py, deepxi/inp_tgt.py, deepxi/network/*.py and
 * DeepXi.infer (deepxi/model.py).  Each entry point below replaces the arithmetic that one of those
 * Python functions delegates to TensorFlow / scipy; the comment above it cites the reference
 * file:line.  The Python package deepxi_b200/ binds these with ctypes (deepxi_b200/_lib.py);
 * INTEGRATION.md shows the stub a maintainer of the reference would add.
 *
 * Conventions
 *   - every function returns 0 on success or a negative DXI_E_* code; dxi_last_error() returns a
 *     thread-local message for the last failure;
 *   - all tensor pointers are DEVICE pointers owned by the caller, row-major, dense; the library
 *     never frees them and never retains them beyond stream order of the call;
 *   - `stream` is a cudaStream_t passed as void*; all work is enqueued asynchronously on it, there
 *     is no hidden synchronisation and no host fallback: on a device that is not sm_100 the calls
 *     return DXI_E_ARCH;
 *   - opaque handles (dxi_net_t) are created / destroyed explicitly; a handle belongs to the device
 *     that was current at creation.  Stateless functions are re-entrant.
 */
#ifndef DEEPXI_B200_H_
#define DEEPXI_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define DXI_API __attribute__((visibility("default")))
#else
#define DXI_API
#endif

#define DXI_OK            0
#define DXI_E_INVALID   (-1)   /* bad argument (shape, null pointer, enum) */
#define DXI_E_CUDA      (-2)   /* CUDA runtime error (message in dxi_last_error) */
#define DXI_E_ARCH      (-3)   /* current device is not sm_100 (no fallback path exists) */
#define DXI_E_STATE     (-4)   /* handle not finalised / missing tensors */
#define DXI_E_NOMEM     (-5)   /* workspace too small */

/* gain function types of deepxi/gain.py:168-191 (gfunc) */
#define DXI_G_MMSE_LSA   0
#define DXI_G_MMSE_STSA  1
#define DXI_G_WF         2
#define DXI_G_SRWF       3
#define DXI_G_CWF        4
#define DXI_G_IRM        5
#define DXI_G_IBM        6
#define DXI_G_DEEPMMSE   7

/* network kinds of deepxi/network/selector.py:8-132 that have committed checkpoints */
#define DXI_NET_RESNETV2  0
#define DXI_NET_MHANETV3  1
#define DXI_NET_RESNET    2   /* tcn.py:17-114  (resnet-1.0c architecture; precision f32 only) */
#define DXI_NET_RESNETV3  3   /* tcn.py:227-245 (precision f32 only) */

/* Conv1D padding of the TCN (main.py:24-25: causal -> "causal", else "same") */
#define DXI_PAD_CAUSAL 0
#define DXI_PAD_SAME   1

/* arithmetic of the network GEMMs */
#define DXI_PREC_F32     0   /* fp32 CUDA-core path (exact mode) */
#define DXI_PREC_F16X3   1   /* tcgen05, fp16 hi/lo split operands, 3 MMAs per product, fp32 accumulate */
#define DXI_PREC_F16     2   /* tcgen05, single fp16 MMA per product, fp32 accumulate */

/* attention mask mode of MHANetV3 */
#define DXI_MASK_NONE        0   /* what the shipped model computes (tfa ignores the mask input) */
#define DXI_MASK_CAUSAL_PAD  1   /* the mask attention.py:355-385 builds */

DXI_API const char* dxi_last_error(void);
/* Pinned host memory for the host <-> device copies of a serving loop (deepxi_b200.model.HostPipeline; the reference's counterpart
 * is the numpy batch of deepxi/se_batch.py:12-55).  write_combined: for buffers the host only fills and the device reads. */
DXI_API int dxi_host_alloc(void** ptr, size_t bytes, int write_combined);
DXI_API int dxi_host_free(void* ptr);
DXI_API int dxi_version(void);
/* 0 when the current CUDA device can run the library (compute capability 10.x). */
DXI_API int dxi_device_check(void);

/*
 * STFT analysis.  Replaces InputTarget.normalise + AnalysisSynthesis.polar_analysis
 * (deepxi/sig.py:189-199, :43-55) and the per-utterance loop of DeepXi.observation_batch
 * (deepxi/model.py:2232-2254) for N_d=512, N_s=256, K=512, Hamming(periodic=False).
 *   wav        [B, wav_stride] int16 (wav_is_i16=1; normalised by 1/32768) or float32
 *   lens       [B] int32 valid samples per utterance (device); NULL = wav_stride for all
 *   mag, phase [B, Tmax, 257] float32; frames at or beyond ceil(len/256) are written as zeros
 */
DXI_API int dxi_stft(const void* wav, int wav_is_i16, const int32_t* lens, int B, int64_t wav_stride,
             int Tmax, float* mag, float* phase, void* stream);

/*
 * iSTFT synthesis.  Replaces AnalysisSynthesis.polar_synthesis (deepxi/sig.py:57-69):
 * Y = (mag*gain) e^{j phase} -> irfft512 -> synthesis window -> overlap-add (hop 256), and the
 * int16 rule of utils.save_wav (deepxi/utils.py:28: (y*32768) truncated toward zero).
 *   gain      nullable [B, Tmax, 257]
 *   n_frames  nullable [B] int32 (device): frames at or beyond n_frames[b] contribute nothing
 *   wav_f32 / wav_i16  nullable outputs [B, out_stride], out_stride >= (Tmax+1)*256
 */
DXI_API int dxi_istft(const float* mag, const float* gain, const float* phase, const int32_t* n_frames,
              int B, int Tmax, float* wav_f32, int16_t* wav_i16, int64_t out_stride, void* stream);

/*
 * Inverse CDF map + gain.  Replaces NormalCDF.inverse with map_type 'DBNormalCDF'
 * (deepxi/map.py:373-390, :75-85), MagXi.xi_hat / gamma_hat (deepxi/inp_tgt.py:216-240) and
 * gfunc(xi_hat, xi_hat+1, gtype) (deepxi/gain.py:168-191) in one pass.
 *   xbar [n_rows, n_bins]; mu, sigma [n_bins]; outputs nullable:
 *   xi_hat f32, gain f32, ibm uint8 (xi_hat > 1, bit-exact with the float32 chain of the oracle)
 */
DXI_API int dxi_map_gain(const float* xbar, const float* mu, const float* sigma, int64_t n_rows, int n_bins,
                 int gtype, float* xi_hat, float* gain, uint8_t* ibm, void* stream);

/* out_type 'deepmmse' of DeepXi.infer (deepxi/model.py:314-318): d_psd = |X|^2 * gfunc(xi_hat, xi_hat + 1, 'deepmmse') with
 * xi_hat the inverse map of xbar, in one pass.  mag, xbar, d_psd: [n_rows, n_bins]. */
DXI_API int dxi_deepmmse(const float* mag, const float* xbar, const float* mu, const float* sigma, int64_t n_rows, int n_bins,
                 float* d_psd, void* stream);

/* gfunc(xi, gamma, gtype) element-wise (deepxi/gain.py:168-191); gamma may be NULL for the
 * xi-only gains (wf, srwf, cwf, irm, ibm). */
DXI_API int dxi_gfunc(const float* xi, const float* gamma, int64_t n, int gtype, float* G, void* stream);

/* NormalCDF.map with 'DBNormalCDF' (deepxi/map.py:356-371, :62-73): xi -> xbar. */
DXI_API int dxi_cdf_map(const float* xi, const float* mu, const float* sigma, int64_t n_rows, int n_bins,
                float* xbar, void* stream);

/* Subband a priori SNR and its ideal binary mask (deepxi/model.py:323-328 with the mel filter bank of deepxi/sig.py:301-346):
 * xi_sub[r][m] = sum_k xi[r][k] H[m][k], ibm = xi_sub > 1.  H: float32 [M][n_bins] on the device; xi_sub or ibm may be NULL. */
DXI_API int dxi_subband_ibm(const float* xi, const float* H, int64_t n_rows, int n_bins, int M, float* xi_sub,
                    uint8_t* ibm, void* stream);

/* ---- training-target pipeline ------------------------------------------------------------ */

/*
 * InputTarget.mix for a padded int16 batch (deepxi/sig.py:162-187, add_noise_pad :231-254, add_noise :256-284):
 * s / 32768; the noise section d[offset : offset + s_len] / 32768 scaled so that the SNR over the s_len samples is
 * snr_db; x = s + d.  Outputs are float32 [B, out_stride], zero from s_len on; any of them may be NULL.  The
 * reference draws `offset` with tf.random.uniform([1], 0, 1 + d_len - s_len) (sig.py:277): here the caller draws it.
 * workspace: dxi_mix_workspace_bytes(B) bytes of device memory.
 */
DXI_API int64_t dxi_mix_workspace_bytes(int B);
DXI_API int dxi_mix(const int16_t* s, const int16_t* d, const int32_t* s_len, const int32_t* d_len, const float* snr_db,
            const int32_t* offsets, int B, int64_t s_stride, int64_t d_stride, float* s_out, float* d_out,
            float* x_out, int64_t out_stride, void* workspace, void* stream);

/* InputTarget.xi (deepxi/sig.py:110-121) = S^2 / max(D^2, 1e-12), optionally fused with NormalCDF.map
 * (deepxi/map.py:356-371): the tail of MagXi.example (deepxi/inp_tgt.py:192-195).  xi or xi_bar may be NULL. */
DXI_API int dxi_xi_map(const float* S, const float* D, const float* mu, const float* sigma, int64_t n_rows, int n_bins,
               float* xi, float* xi_bar, void* stream);

/*
 * MagXi.stats / NormalCDF.stats (deepxi/inp_tgt.py:160-171, deepxi/map.py:392-402) as mergeable moments:
 * acc[0][k] += number of frames, acc[1][k] += sum, acc[2][k] += sum of squares of 10 log10 max(xi, 1e-12) over the
 * first n_frames[b] frames of every utterance (n_frames NULL: all Tmax).  acc: float64 [3][n_bins] on the device,
 * zeroed by the caller; mu = acc[1] / acc[0], sigma = sqrt(acc[2] / acc[0] - mu^2) (population std).  Shards held by
 * different GPUs are combined by summing acc (one all-reduce of 3 x n_bins doubles).
 */
DXI_API int dxi_xi_db_moments(const float* S, const float* D, const int32_t* n_frames, int B, int Tmax, int n_bins,
                      double* acc, void* stream);

/*
 * Fused enhancement back end: MagXi.enhanced_speech (deepxi/inp_tgt.py:198-214) =
 * inverse map -> gamma_hat = xi_hat + 1 -> gfunc -> |Y| = |X| G -> polar_synthesis.
 * Same outputs as dxi_istft; never materialises xi_hat or G in HBM.
 */
DXI_API int dxi_enhance(const float* mag, const float* phase, const float* xbar, const float* mu,
                const float* sigma, int gtype, const int32_t* n_frames, int B, int Tmax,
                float* wav_f32, int16_t* wav_i16, int64_t out_stride, void* stream);

/*
 * A priori SNR estimator networks.  Replace the Keras model built by network_selector
 * (deepxi/network/selector.py:86-99 -> deepxi/network/tcn.py:116-225 ResNetV2;
 * selector.py:9-21 -> deepxi/network/attention.py:387-442 MHANetV3) and model.predict
 * (deepxi/model.py:286).
 */
typedef struct dxi_net dxi_net_t;

typedef struct dxi_net_cfg {
  int32_t n_feat;      /* 257 */
  int32_t n_outp;      /* 257 */
  int32_t d_model;     /* 256 */
  int32_t n_blocks;    /* 40 (ResNetV2) / 5 (MHANetV3) */
  int32_t d_f;         /* 64   bottleneck size (ResNetV2) */
  int32_t k;           /* 3    kernel size (ResNetV2) */
  int32_t max_d_rate;  /* 16   (ResNetV2) */
  int32_t padding;     /* DXI_PAD_* (ResNetV2) */
  int32_t n_heads;     /* 8    (MHANetV3) */
  int32_t max_len;     /* 2048 (MHANetV3) */
  int32_t mask_mode;   /* DXI_MASK_* (MHANetV3) */
  int32_t precision;   /* DXI_PREC_* */
} dxi_net_cfg;

DXI_API int dxi_net_create(dxi_net_t** h, int kind, const dxi_net_cfg* cfg);
/* Supplies one checkpoint tensor by its Keras name, e.g. "layer_with_weights-3/kernel" (HOST fp32,
 * Keras layout [k, C_in, C_out]); shape is validated against the configuration. */
DXI_API int dxi_net_load(dxi_net_t* h, const char* tensor_name, const float* host_data, const int64_t* shape,
                 int rank);
/* Uploads / repacks the weights for the chosen precision; must follow the last dxi_net_load. */
DXI_API int dxi_net_finalize(dxi_net_t* h, void* stream);
/* Bytes of caller-provided device workspace needed by dxi_net_forward for this batch shape. */
DXI_API int64_t dxi_net_workspace_bytes(const dxi_net_t* h, int B, int Tmax);
/*
 * Forward: mag [B, Tmax, n_feat] -> xbar [B, Tmax, n_outp] in (0,1).  All Tmax frames of every
 * utterance are computed as the reference does (zero-padded frames are zero-input frames).
 * The workspace belongs to the call in stream order (activations and, for the tcgen05 ResNetV2 path, the per-tile
 * stage counters that chain its 41 stage launches; zeroed by the call): forward passes that may overlap on
 * different streams need different workspaces.
 */
DXI_API int dxi_net_forward(dxi_net_t* h, const float* mag, int B, int Tmax, float* xbar, void* workspace,
                    size_t workspace_bytes, void* stream);
DXI_API int dxi_net_destroy(dxi_net_t* h);

/* Number of kernel launches this library has enqueued from the calling thread since the last call to
 * dxi_launch_count_reset (used by bench.py for its "gpu_launches" claim). */
DXI_API int64_t dxi_launch_count(void);
DXI_API void dxi_launch_count_reset(void);

/* Per-kernel timing for bench.py: when enabled, the library brackets each named group of launches
 * ("stft", "istft", "map_gain", "tcn_stage", "tcn_stem", "tcn_head", ...) with CUDA events on the
 * launching stream.  dxi_profile_read synchronises on those events, returns the summed milliseconds and
 * the number of launches they covered, and clears the group.  Per calling thread. */
DXI_API void dxi_profile_enable(int on);
DXI_API int dxi_profile_read(const char* key, double* total_ms, int64_t* launches);

/* Self test of the tcgen05 / TMEM building blocks: D[128,N] = A[128,K] * B[K,N] with fp16 operands
 * (A from tensor memory, B from shared memory) written to `d_out` (float32 [128,N]).
 * a_host_layout / b: device fp16 row-major [128,K] and [N,K].  variant selects descriptor encodings
 * under test (0 = the encoding the production kernels use). */
DXI_API int dxi_selftest_umma(const void* a_f16, const void* b_f16, int N, int K, int variant, float* d_out,
                      void* stream);
/* The same for the CTA pair (tcgen05.mma.cta_group::2, a cluster of two CTAs on the SMs of one TPC): D[256,256] = A[256,K] * B[256,K]^T;
 * CTA r holds rows 128 r .. of A and of B, the leader issues, each CTA reads its 128 rows of D. */
DXI_API int dxi_selftest_umma_pair(const void* a_f16, const void* b_f16, int K, float* d_out, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* DEEPXI_B200_H_ */
