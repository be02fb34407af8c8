// common.cuh — shared declarations of libfadb200 (B200 / sm_100a FAD hot path).
#pragma once

#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include <map>
#include <string>
#include <vector>

#include "../../include/fadb.h"

namespace fadb {

void set_error(const char* fmt, ...);

#define FADB_CUDA_CHECK(expr)                                                                         \
    do {                                                                                              \
        cudaError_t _e = (expr);                                                                      \
        if (_e != cudaSuccess) {                                                                      \
            ::fadb::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
            return FADB_E_CUDA;                                                                       \
        }                                                                                             \
    } while (0)

#define FADB_CHECK(expr)            \
    do {                            \
        int _r = (expr);            \
        if (_r != FADB_OK) return _r; \
    } while (0)

#define FADB_REQUIRE(cond, ...)                 \
    do {                                        \
        if (!(cond)) {                          \
            ::fadb::set_error(__VA_ARGS__);     \
            return FADB_E_INVALID;              \
        }                                       \
    } while (0)

// device error codes written to handle->err_flag
enum : int { DEVERR_NONE = 0, DEVERR_PIPE_TIMEOUT = 1, DEVERR_NONFINITE = 2 };

// ------------------------------------------------------------------------------------------------
// A growable device buffer owned by the handle.
struct DevBuf {
    void* p = nullptr;
    size_t bytes = 0;
    int reserve(size_t need) {
        if (need <= bytes) return FADB_OK;
        if (p) cudaFree(p);
        p = nullptr;
        bytes = 0;
        cudaError_t e = cudaMalloc(&p, need);
        if (e != cudaSuccess) {
            set_error("cudaMalloc(%zu) failed: %s", need, cudaGetErrorString(e));
            return FADB_E_NOMEM;
        }
        bytes = need;
        return FADB_OK;
    }
    void release() {
        if (p) cudaFree(p);
        p = nullptr;
        bytes = 0;
    }
    template <class T> T* as() const { return reinterpret_cast<T*>(p); }
};

// 16-bit operand format of a precision mode: IEEE fp16 (FP16 / FP16X2) or bf16 (BF16 / BF16X3)
inline bool prec_is_f16(int prec) { return prec == FADB_PREC_FP16 || prec == FADB_PREC_FP16X2; }

// A packed tensor-core layer: B operand [N][K] K-major, 16-bit (hi + lo planes, bf16 or fp16) and fp32 bias.
struct PackedLayer {
    int N = 0, K = 0;      // K = taps * Cin
    int Cin = 0, taps = 0;
    int f16 = 0;           // planes hold IEEE fp16 instead of bf16
    __nv_bfloat16* w_hi = nullptr;
    __nv_bfloat16* w_lo = nullptr;
    uint8_t* w8 = nullptr;   // fp16x2: e4m3((w - fp16(w)) / lo_scale), [N][K] K-major, when Cin % 128 == 0 (else nullptr)
    float lo_scale = 1.f;    // power of two; the e4m3 pass's sums are multiplied by it
    float* bias = nullptr;   // [N] (folded BN shift or conv/linear bias)
};

// Front-end constant tables of one model (twiddles, window, sparse mel bands), on the handle's device (frontend.cu)
struct FrontTables {
    double2* tw = nullptr;     // [NF]  exp(-2 pi i k / NF)
    double* win = nullptr;     // [WIN] periodic Hann
    int* band_start = nullptr; // [64]
    int* band_len = nullptr;   // [64]
    float* band_wt = nullptr;  // [wt_rows][32] fp32: row (h2 ? wt_off1 : 0) + i, column lane = weight i of band lane + 32*h2
    int wt_rows = 0, wt_off1 = 0;
    int nfft = 0, win_len = 0, hop = 0;
    bool ready = false;
};

struct HostTensor {
    std::vector<int64_t> shape;
    float* dev = nullptr;    // staging copy on device (fp32, PyTorch layout)
    size_t numel = 0;
};

}  // namespace fadb

// ------------------------------------------------------------------------------------------------
struct fadb_handle {
    int device = 0;
    int sm_count = 148;
    int precision = FADB_PREC_FP16X2;
    int max_batch = 16384;          // VGGish patches per internal batch (fewer, larger launches: ~5 us gap each)
    int max_batch_cnn14 = 128;      // CNN14 clips per internal batch
    int gemm_smem_budget = 231168;  // bytes of smem for resident weights + pipeline stages per GEMM CTA
                                    // (227 KB opt-in maximum minus barriers/alignment slack)
    int fused_front = 1;            // VGGish: PCM -> conv1 output in one kernel (features stay in shared memory)
    int halo = 1;                   // 3x3 layers on large maps: one halo tile per channel block feeds all 9 taps
    int resident_b = 1;             // keep short-K weight slabs resident in smem (see gemm_tc.cu)
    int gemm_cluster = 1;           // 1 = CTA clusters share weight tiles through TMA multicast (plain single-pass layers)
    int gemm_cluster_size = 2;      // CTAs per cluster: 2 or 4
    int gemm_pair_halo = 1;         // CTA pairs for the halo-mode layers too (CNN14 blocks 1-4: +10 %)
    int gemm_twocta = 1;            // clusters of 2: 1 = cta_group::2 MMAs (M = 256 across the pair) instead of weight multicast
    int lo_fp8 = 1;                 // fp16x2: the low-order weight pass runs in e4m3 at twice the fp16 rate where Cin % 128 == 0
                                    // (FADB_LO_FP8=0: every lo pass in fp16)
    unsigned x2_mask = 0xffffffffu; // fp16x2: bit i = tensor-core layer i multiplies the lo weight plane too (FADB_X2_MASK)
    int model = -1;                 // model whose weights are committed
    bool weights_ready = false;
    int64_t launches = 0;
    int* err_flag = nullptr;        // device int
    int* err_flag_host = nullptr;   // pinned mirror

    std::map<std::string, fadb::HostTensor> staged;    // between weights_begin and commit
    int staged_model = -1;

    // first conv layer (Cin = 1): fp32 weights [9][64] (+BN folded), bias [64]; bn0 affine for CNN14
    float* conv1_w = nullptr;
    float* conv1_b = nullptr;
    float* bn0_scale = nullptr;
    float* bn0_shift = nullptr;
    std::vector<fadb::PackedLayer> layers;             // tensor-core layers in execution order
    fadb::FrontTables front_tables[5];                 // per model, built on first use ON THIS HANDLE'S DEVICE
    fadb::DevBuf weight_pool;                          // backing store of all packed weights

    // activation workspace
    fadb::DevBuf ws_a1[1];      // conv1 output (input of the first tensor-core layer)
    fadb::DevBuf ws_a1_8;       // its e4m3 copy, W-padded (fp16x2 with the e4m3 low-order pass: LayerIO::in8_wpad)
    fadb::DevBuf ws_feats;      // fp32 features of one batch
    fadb::DevBuf ws_act[2];     // ping-pong bf16 activations (hi plane followed by lo plane)
    fadb::DevBuf ws_act8[2];    // fp16x2: e4m3 copies of the same activations
    fadb::DevBuf ws_misc;       // pooled vectors etc.
    fadb::DevBuf ws_frechet;    // fp64 workspace of fadb_frechet
    fadb::DevBuf ws_stats;      // fp64 workspace of fadb_fad_from_pcm_host
    fadb::DevBuf ws_syrk;       // tensor-core syrk: transposed split-fp16 planes of one row chunk + its fp32 product
    int lo_fp8_c64 = 1;         // ... and of the 64-channel 3x3 layers (two taps per 128-byte K block; FADB_LO_FP8_C64=0: fp16 pass)
    int clap_quantize = 1;      // CLAP front end applies clap.py:70-72's int16 truncation (fadb_set_clap_quantize)
    int tc_syrk = 0;            // 1: second moments of >= 8192 rows at d >= 512 on the tensor cores (fadb_set_tensor_syrk /
                                // FADB_TC_SYRK): ~8x faster, covariance accurate to ~1e-6 instead of 1e-14 — OFF by default
                                // because FAD between two SIMILAR sets amplifies that error by tr(S) / FAD
    fadb::DevBuf ws_pcm[2];     // double-buffered PCM chunks for the host path
    fadb::DevBuf ws_emb;        // embeddings of the host path
    // optional per-launch timing of the tensor-core layers (bench.py roofline leg)
    bool profile = false;
    std::vector<cudaEvent_t> prof_events;   // pairs (start, stop) around tensor-core layer launches
    std::vector<cudaEvent_t> prof_front;    // pairs (start, stop) around front-end (+ conv1) launches
    double prof_flops = 0.0;
    cudaStream_t copy_stream = nullptr;
    cudaEvent_t ev_copy[2] = {nullptr, nullptr};
    cudaEvent_t ev_compute[2] = {nullptr, nullptr};
};

namespace fadb {

// ---- kernels / launchers implemented across the .cu files ---------------------------------------
// PCM source on the device: fp32 samples, or raw int16 PCM that the front end normalises by 1/32768 while loading
// (the reference's dtype="int16" path, fad.py:145-149; exact in fp32, so both forms give identical features)
struct PcmSrc {
    const void* ptr;
    int i16;
    PcmSrc offset(int64_t samples) const { return PcmSrc{static_cast<const char*>(ptr) + samples * (i16 ? 2 : 4), i16}; }
};

// frontend.cu
int launch_frontend(fadb_handle* h, int model, PcmSrc pcm, int64_t n_clips, int64_t n_samples,
                    int64_t pcm_stride, float* feats, cudaStream_t st);
int frontend_init(fadb_handle* h);
void frontend_release(fadb_handle* h);
int launch_vggish_front_conv1(fadb_handle* h, PcmSrc pcm, int64_t n_clips, int64_t n_samples, int64_t pcm_stride,
                              __nv_bfloat16* out_hi, __nv_bfloat16* out_lo, uint8_t* out8p, cudaStream_t st);

int64_t frontend_rows(int model, int64_t n_samples);

// resample.cu
int launch_resample(fadb_handle* h, const float* in, int64_t n_clips, int64_t n_in, int64_t in_stride, double ratio,
                    const double* win, int nwin, int num_table, float* out, int64_t n_out, int64_t out_stride,
                    cudaStream_t st);

// conv1.cu — Cin = 1 direct 3x3 conv on CUDA cores
int launch_conv1_vggish(fadb_handle* h, const float* feats, int64_t n_patches, __nv_bfloat16* out_hi,
                        __nv_bfloat16* out_lo, cudaStream_t st);
int launch_conv1_cnn14(fadb_handle* h, const float* feats, int64_t n_clips, int T, __nv_bfloat16* out_hi,
                       __nv_bfloat16* out_lo, uint8_t* out8p, cudaStream_t st);

// gemm_tc.cu — tcgen05 implicit-GEMM layer
struct LayerIO {
    const __nv_bfloat16* in_hi = nullptr;
    const __nv_bfloat16* in_lo = nullptr;     // may be null (bf16 mode)
    const uint8_t* in8 = nullptr;              // fp16x2: e4m3 copy of the input (for the e4m3 low-order weight pass)
    uint8_t* out8 = nullptr;                   // fp16x2: where to write the e4m3 copy of the output (Cout % 128 == 0)
    int out8_wpad = 0;                         // write out8 in that W-padded layout (the consumer is a 64-channel layer)
    int in8_wpad = 0;                          // in8 is [B][H][W+2][Cin] with a zero column left and right (64-channel layers:
                                               // GemmParams::c64 reads it through overlapping two-pixel rows)
    int B = 0, H = 0, W = 0, Cin = 0;          // NHWC input; a linear layer is B=1,H=1,W=rows
    int taps = 9;                              // 9 = 3x3 pad 1, 1 = pointwise / linear
    int relu = 1;
    int pool = 0;                              // 0 none, 1 max 2x2, 2 avg 2x2
    __nv_bfloat16* out_hi = nullptr;
    __nv_bfloat16* out_lo = nullptr;           // may be null
    float* out_f32 = nullptr;                  // if set, fp32 output instead of bf16
    int use_lo_weights = 1;                    // fp16x2: 0 = this layer runs single-pass (FADB_X2_MASK sensitivity sweeps)
    // statistics (stats.cu): C = Y Y^T as a "linear layer" whose activations and weights are the same split-fp16 matrix —
    // fp16 hi/lo planes on both sides, 3 MMAs per product, N tiles of 128, tiles strictly below the diagonal skipped.
    // Overrides the handle's precision for this launch.
    int syrk = 0;
    int seg_blocks = 0;                        // K blocks per accumulation segment (0 = the default 16 = 64 MMAs)
    double* out_f64 = nullptr;                 // syrk: fp64 matrix the finished tiles are ADDED to
    long long row_stride = 0;                  // linear layers: elements between rows of the activation AND weight
                                               // matrices (0 = dense); lets a caller avoid power-of-two strides
};
int launch_gemm_layer(fadb_handle* h, const PackedLayer& L, const LayerIO& io, cudaStream_t st);
int gemm_init(fadb_handle* h);

// pack.cu — weight repacking, BN folding, fp32 -> bf16 hi/lo split
// (the planes are 16-bit words: bf16, or IEEE fp16 when f16 is set)
int pack_conv_weight(fadb_handle* h, const float* w_oihw, int Cout, int Cin, int ksize, const float* scale,
                     __nv_bfloat16* w_hi, __nv_bfloat16* w_lo, bool f16, cudaStream_t st);
// fp16x2: w8 = e4m3((w - fp16(w)) / lo_scale) in the packed [Cout][tap][Cin] order; *lo_scale (a power of two) is chosen so
// that the largest residual lands near the top of the e4m3 range.  Synchronises the stream.
int pack_conv_weight_lo8(fadb_handle* h, const float* w_oihw, int Cout, int Cin, int ksize, const float* scale,
                         uint8_t* w8, float* lo_scale, cudaStream_t st);
int quantize_e4m3(fadb_handle* h, const float* x, int64_t n, uint8_t* out, cudaStream_t st);
// [rows][W][64] fp32 or fp16 activations -> [rows][W+2][64] e4m3 with a zero column left and right (LayerIO::in8_wpad)
int quantize_e4m3_wpad(fadb_handle* h, const float* x32, const __nv_bfloat16* x16, int64_t rows, int W, uint8_t* out,
                       cudaStream_t st);
int split_f32_to_bf16(fadb_handle* h, const float* x, int64_t n, __nv_bfloat16* hi, __nv_bfloat16* lo, bool f16,
                      cudaStream_t st);
int fold_bn(fadb_handle* h, const float* gamma, const float* beta, const float* mean, const float* var, int C,
            float* scale, float* shift, cudaStream_t st);

// cnn14 tail (global pooling)  — cnn14.cu
int launch_cnn14_global_pool(fadb_handle* h, const __nv_bfloat16* x_hi, const __nv_bfloat16* x_lo, int64_t B, int Ht,
                             int Wf, int C, __nv_bfloat16* out_hi, __nv_bfloat16* out_lo, uint8_t* out8, cudaStream_t st);
int launch_l2_normalize(fadb_handle* h, float* x, int64_t rows, int d, cudaStream_t st);

// stats.cu
int launch_stats_accumulate(fadb_handle* h, const float* emb, int64_t n, int d, int64_t ld, const double* shift,
                            double* acc, cudaStream_t st);
int launch_stats_accumulate_f64(fadb_handle* h, const double* emb, int64_t n, int d, int64_t ld, const double* shift,
                                double* acc, cudaStream_t st);
int launch_stats_finalize(fadb_handle* h, const double* acc, int d, const double* shift, double* mu, double* sigma,
                          cudaStream_t st);

// frechet.cu
int frechet_init(fadb_handle* h);
int launch_frechet(fadb_handle* h, const double* mu1, const double* s1, const double* mu2, const double* s2, int d,
                   double* out, cudaStream_t st);

}  // namespace fadb
