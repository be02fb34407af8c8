/*
 * fadb.h — C ABI of libfadb200.so, the B200 (sm_100a) Frechet-Audio-Distance hot path.
 *
 * Drop-in boundary for gibiansky/frechet-audio-distance-exported.  The reference has no FFI of its
 * own (it is pure Python); each entry point below names the reference interface it replaces
 * (paths relative to frechet_audio_distance_exported/).  The Python binding a maintainer adds is
 * the ctypes stub in INTEGRATION.md (shipped as frechet_audio_distance_exported_b200/_lib.py).
 *
 * Conventions
 *   - plain C types only; no torch / C++ types cross this boundary;
 *   - `*_dev` pointers are device pointers owned by the caller (e.g. tensor.data_ptr());
 *     `*_host` pointers are host pointers (pinned memory recommended);
 *   - `stream` is a cudaStream_t passed as void* (0 = legacy default stream); every call that
 *     takes a stream is stream-ordered and asynchronous unless stated otherwise;
 *   - return value: 0 = ok, negative = error (see FADB_E_*); fadb_last_error() returns a
 *     thread-local message; no exception crosses the boundary;
 *   - a handle is not thread-safe; distinct handles are independent.
 */
#ifndef FADB_H_
#define FADB_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif
#if defined(__GNUC__)
#pragma GCC visibility push(default) /* the library is built with -fvisibility=hidden */
#endif

#define FADB_ABI_VERSION 1

/* error codes */
#define FADB_OK 0
#define FADB_E_INVALID (-1)   /* bad argument / unsupported shape */
#define FADB_E_CUDA (-2)      /* CUDA runtime / driver error */
#define FADB_E_STATE (-3)     /* call sequence error (weights not loaded, ...) */
#define FADB_E_NOMEM (-4)
#define FADB_E_DEVICE (-5)    /* device-side failure flag (pipeline timeout, non-finite result) */

/* model = front-end variant + network; names follow fad.py:109-117 (VALID_MODELS) */
#define FADB_MODEL_VGGISH 0   /* "vggish"   16 kHz, 128-d per 0.96 s patch  */
#define FADB_MODEL_PANN8K 1   /* "pann-8k"   8 kHz, 2048-d per clip         */
#define FADB_MODEL_PANN16K 2  /* "pann-16k" 16 kHz                           */
#define FADB_MODEL_PANN32K 3  /* "pann-32k" 32 kHz                           */
#define FADB_MODEL_CLAP 4     /* "clap"     48 kHz CNN14 branch + head, 512-d L2-normalised */

/* arithmetic of the tensor-core layers (all: tcgen05 kind::f16, fp32 accumulation in tensor memory) */
#define FADB_PREC_BF16 0      /* bf16 activations x bf16 weights, 1 MMA per product: FAD within ~1e-4 .. 1e-3 of fp32 */
#define FADB_PREC_BF16X3 1    /* split-bf16 (hi+lo) activations and weights, 3 MMAs per product + exact segment sums:
                               * ~fp32 accuracy (embeddings 1e-5), slow; the strict parity mode */
#define FADB_PREC_FP16 2      /* IEEE fp16 activations x fp16 weights, 1 MMA per product (same speed as BF16, 8x smaller
                               * operand rounding; activations saturate at +-65504) */
#define FADB_PREC_FP16X2 3    /* fp16 activations x split-fp16 (hi+lo) weights, 2 MMAs per product: the weights are exact to
                               * ~2^-22, FAD within 1e-4 of the fp32 reference.  DEFAULT of a new handle. */

typedef struct fadb_handle fadb_handle;

/* ---------------------------------------------------------------- lifetime */
int fadb_abi_version(void);
const char* fadb_last_error(void);
/* Replaces device selection at fad.py:228-233.  Fails (FADB_E_CUDA) when no sm_100 GPU is present:
 * there is no CPU fallback. */
int fadb_create(fadb_handle** out, int device);
void fadb_destroy(fadb_handle* h);
/* Packed weights are stored in the 16-bit format of the precision that was current at fadb_weights_commit (bf16 or
 * fp16): after switching between the two families the weights must be committed again (FADB_E_STATE otherwise). */
int fadb_set_precision(fadb_handle* h, int prec);
/* Statistics of large embedding sets (>= 8192 rows per call, d >= 512 and a multiple of 128) on the tensor cores:
 * sum (x-K)(x-K)^T as a split-fp16 tcgen05 GEMM (3 MMAs per product) instead of the fp64 DFMA kernel.  About 8x
 * faster at d = 2048; the covariance is then accurate to ~1e-6 (truncation inside the MMA's fp32 accumulation)
 * instead of ~1e-14.  Off by default: the reference computes np.cov in fp64 (fad.py:495), and the Frechet distance
 * of two SIMILAR sets amplifies a covariance error by tr(S) / FAD. */
int fadb_set_tensor_syrk(fadb_handle* h, int on);
/* CLAP only: whether the front end applies the int16 truncation of clap.py:70-72 to the samples it reads (default 1).
 * The reference quantises BEFORE it resamples (clap.py:70-80), so a caller that resamples CLAP input itself quantises
 * at the source rate and switches this off for that call (fad.py does, for get_embeddings(x, sr != 48000)). */
int fadb_set_clap_quantize(fadb_handle* h, int on);
/* max patches (VGGish) / clips (CNN14) per internal batch; sizes the activation workspace */
int fadb_set_max_batch(fadb_handle* h, int max_items);

/* ---------------------------------------------------------------- weights
 * Replaces fad.py:249-300 (_load_model): instead of a torch.export artefact the caller hands over
 * the state_dict tensors of VGGishCore (models/vggish.py:69-78) or PANNCore (models/pann.py:223-234)
 * under their own key names, fp32, contiguous, PyTorch layout ([Cout,Cin,3,3], [out,in]).
 * CLAP head keys: "clap_head.0.weight/bias", "clap_head.2.weight/bias" (README.md:195-199).
 * fadb_weights_commit folds eval-mode BatchNorm (eps 1e-5), repacks to K-major bf16 (hi/lo planes)
 * on the device and frees the staging copies.  Synchronous. */
int fadb_weights_begin(fadb_handle* h, int model);
int fadb_weights_tensor(fadb_handle* h, const char* name, const float* data_host, const int64_t* shape, int ndim);
int fadb_weights_commit(fadb_handle* h);

/* ---------------------------------------------------------------- front end
 * Replaces models/vggish.py:230-279 (waveform_to_examples), models/pann.py:68-145
 * (waveform_to_logmel) + fad.py:41-66 (_pad_to_valid_pann_time), models/clap.py:41-80
 * (preprocess_for_clap) + fad.py:69-91,356-359.  Native-rate mono PCM only.
 *   pcm_dev   : [n_clips, n_samples] fp32, clip stride = pcm_stride elements
 *   feats_dev : VGGish  [n_clips * patches, 96, 64] fp32
 *               PANN    [n_clips, T', 64] fp32 (zero rows appended, T' = 32k-24)
 *               CLAP    [n_clips, 1001, 64] fp32 (n_samples <= 480000; zero-padded like fad.py:356)
 * fadb_frontend_rows() returns rows (patches / frames T') per clip for (model, n_samples). */
int64_t fadb_frontend_rows(int model, int64_t n_samples);
int fadb_frontend(fadb_handle* h, int model, const float* pcm_dev, int64_t n_clips, int64_t n_samples,
                  int64_t pcm_stride, float* feats_dev, void* stream);

/* ---------------------------------------------------------------- network
 * Replaces `self.model(x)` at fad.py:367,382,393 (VGGishCore.forward models/vggish.py:80-95,
 * PANNCore.forward models/pann.py:236-273).
 *   VGGish: feats [n_items, 96, 64]  -> emb [n_items, 128]
 *   PANN  : feats [n_items, T, 64]   -> emb [n_items, 2048]   (T = 32k-24)
 *   CLAP  : feats [n_items, 1001, 64]-> emb [n_items, 512] */
int fadb_embed_dim(int model);
int fadb_embed(fadb_handle* h, const float* feats_dev, int64_t n_items, int64_t t_frames, float* emb_dev,
               void* stream);

/* PCM -> embeddings in one call (front end + network, internally batched).  Replaces the per-clip
 * loop of get_embeddings, fad.py:302-408.  emb_dev holds n_clips * fadb_frontend_rows() rows for
 * VGGish, n_clips rows otherwise. */
int fadb_embed_pcm(fadb_handle* h, const float* pcm_dev, int64_t n_clips, int64_t n_samples,
                   int64_t pcm_stride, float* emb_dev, void* stream);
/* Same with raw 16-bit PCM (the WAV sample format; the reference's dtype="int16" path, fad.py:145-149): the
 * front end scales by 1/32768 while loading, which is exact, so the result equals fadb_embed_pcm on
 * float(pcm)/32768.  Halves the host->device and HBM bytes of the PCM.  pcm_stride in samples. */
int fadb_embed_pcm16(fadb_handle* h, const int16_t* pcm_dev, int64_t n_clips, int64_t n_samples,
                     int64_t pcm_stride, float* emb_dev, void* stream);

/* ---------------------------------------------------------------- resampling
 * resampy.resample(x, sr_orig, sr_new) semantics (filter kaiser_best) for clips already on the device; replaces
 * fad.py:159, models/vggish.py:250, models/pann.py:101.  ratio = sr_new / sr_orig, n_out = int(n_in * ratio).
 * win_dev: the interpolation filter's right wing, double[nwin], with num_table entries per zero crossing, already
 * multiplied by ratio when ratio < 1 (frechet_audio_distance_exported_b200/resample.py builds it); the kernel replays
 * resample.py's fp64 arithmetic operation by operation.  Parity against resampy itself is unpinned (not installable). */
int fadb_resample(fadb_handle* h, const float* in_dev, int64_t n_clips, int64_t n_in, int64_t in_stride, double ratio,
                  const double* win_dev, int nwin, int num_table, float* out_dev, int64_t n_out, int64_t out_stride,
                  void* stream);

/* ---------------------------------------------------------------- statistics
 * Replaces calculate_embd_statistics, fad.py:483-496, as a summable sufficient statistic.
 *   acc_dev : double[1 + d + d*d] = { n, sum(x-K), sum (x-K)(x-K)^T }   (caller zero-initialises)
 *   shift_dev: double[d] common shift K (may be NULL = 0); must be identical on all ranks.
 * Partials from different calls / GPUs add (allreduce acc between accumulate and finalize).
 * finalize: mu[d] (fp64), sigma[d*d] (fp64, ddof=1, np.cov semantics). */
int fadb_stats_accumulate(fadb_handle* h, const float* emb_dev, int64_t n_rows, int d, int64_t row_stride,
                          const double* shift_dev, double* acc_dev, void* stream);
/* same, for embeddings stored as fp64 (e.g. a float64 .npy cache, fad.py:619) */
int fadb_stats_accumulate_f64(fadb_handle* h, const double* emb_dev, int64_t n_rows, int d, int64_t row_stride,
                              const double* shift_dev, double* acc_dev, void* stream);
int fadb_stats_finalize(fadb_handle* h, const double* acc_dev, int d, const double* shift_dev, double* mu_dev,
                        double* sigma_dev, void* stream);

/* ---------------------------------------------------------------- Frechet distance
 * Replaces calculate_frechet_distance, fad.py:498-555:
 *   ||mu1-mu2||^2 + tr S1 + tr S2 - 2 tr sqrtm(S1 S2)
 * with tr sqrtm(S1 S2) = sum_i sqrt(lambda_i(L^T S2 L)), S1 = L L^T (semi-definite Cholesky),
 * eigenvalues by Householder tridiagonalisation + Sturm bisection, all fp64 on the device.
 * out_dev: double[4] = { fad, tr sqrtm(S1 S2), tr S1 + tr S2, ||mu1-mu2||^2 }.  sigma inputs are
 * not modified.  Stream-ordered; uses workspace inside the handle. */
int fadb_frechet(fadb_handle* h, const double* mu1_dev, const double* sigma1_dev, const double* mu2_dev,
                 const double* sigma2_dev, int d, double* out_dev, void* stream);

/* ---------------------------------------------------------------- whole path, host buffers
 * score() minus file decoding (fad.py:621-656): PCM of both sets in HOST memory -> FAD.
 * Copies host->device in chunks overlapped with compute, returns the scalar to the host.
 * Synchronous.  emb_*_host may be NULL; otherwise receive the embeddings (for the .npy cache,
 * fad.py:623-637). */
int fadb_fad_from_pcm_host(fadb_handle* h, const float* pcm_bg_host, int64_t n_bg, const float* pcm_ev_host,
                           int64_t n_ev, int64_t n_samples, float* emb_bg_host, float* emb_ev_host,
                           double* fad_out);
int fadb_fad_from_pcm16_host(fadb_handle* h, const int16_t* pcm_bg_host, int64_t n_bg, const int16_t* pcm_ev_host,
                             int64_t n_ev, int64_t n_samples, float* emb_bg_host, float* emb_ev_host,
                             double* fad_out);

/* ---------------------------------------------------------------- introspection / tests */
/* Per-launch CUDA-event timing of the tensor-core layers (for the roofline figure in bench.py):
 * enable(1) resets the counters; read() synchronises the recorded events and returns
 * out4 = { sum of layer launch durations in ms, algorithmic FLOPs of those launches, launches,
 *          sum of front-end (+ fused conv1) launch durations in ms }. */
int fadb_profile_enable(fadb_handle* h, int on);
int fadb_profile_read(fadb_handle* h, double* out4);
/* number of kernels this library has launched on this handle since creation */
int64_t fadb_launch_count(const fadb_handle* h);
/* device-side error flag (0 = none); non-zero after a pipeline timeout */
int fadb_device_status(fadb_handle* h);
/* Single tensor-core layer, exposed for parity tests of the implicit-GEMM kernel:
 *   act NHWC bf16 hi (+lo or NULL) [B,H,W,Cin]; w fp32 PyTorch layout [Cout,Cin,kh,kw] (kh=kw=3, pad 1, or 1)
 *   out fp32 NHWC [B,H',W',Cout] after bias, optional ReLU, optional 2x2 pool (0 none,1 max,2 avg). */
int fadb_debug_conv_layer(fadb_handle* h, const float* x_nhwc_f32_dev, int B, int H, int W, int Cin,
                          const float* w_dev, const float* bias_dev, int Cout, int ksize, int relu, int pool,
                          float* out_nhwc_f32_dev, void* stream);

#if defined(__GNUC__)
#pragma GCC visibility pop
#endif
#ifdef __cplusplus
}
#endif
#endif /* FADB_H_ */
