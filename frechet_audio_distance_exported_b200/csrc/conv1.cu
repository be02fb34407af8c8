// conv1.cu — the Cin = 1 first convolution of both networks, on CUDA cores in fp32 (it is 0.2-0.4 %
// of the FLOPs and not GEMM-shaped: K = 9).  Reads the fp32 log-mel features, writes NHWC bf16
// (hi / optional lo plane) ready for the tcgen05 layers.
//
//   VGGish : conv3x3(1->64)+bias+ReLU+maxpool2x2   models/vggish.py:44-49 (first "64","M")
//            feats [P,96,64] -> [P,48,32,64]
//   CNN14  : bn0 (per-mel affine, applied to the zero time-pad rows too, models/pann.py:249-251)
//            -> conv3x3(1->64, no bias) -> BN1 (folded) -> ReLU      models/pann.py:190
//            feats [B,T,64] -> [B,T,64,64]
#include "common.cuh"
#include "conv1_dev.cuh"
#include "tc_ptx.cuh"

namespace fadb {

// ---------------------------------------------------------------- VGGish: conv + ReLU + maxpool
// grid = (6 bands of 8 pooled rows, P patches); block = 256 threads = the 8 x 32 pooled pixels of the band.
// A thread owns ONE pooled pixel (its 4x4 input window lives in registers as packed f32x2 pairs) and loops
// over the 64 output channels in 4 groups of 16.  The weight address is therefore warp-uniform: every
// LDS.128 of weights is a single broadcast wavefront (the first version indexed weights by lane and was
// shared-memory-bandwidth bound: 85 % LSU wavefront utilisation, ncu r01).
constexpr int kC1Band = 8;                       // pooled rows per CTA
__global__ void __launch_bounds__(256, 2) conv1_vggish_kernel(const float* __restrict__ feats, const float* __restrict__ w,
                                                              const float* __restrict__ bias,
                                                              __nv_bfloat16* __restrict__ out_hi,
                                                              __nv_bfloat16* __restrict__ out_lo, int f16) {
    constexpr int H = 96, W = 64, HP = 48, WP = 32;
    constexpr int ROWS = 2 * kC1Band + 2;
    __shared__ __align__(16) float s_in[ROWS][W + 4];      // pitch 68: rows stay 16-byte aligned
    __shared__ __align__(16) float s_w[9][64];
    __shared__ __align__(16) float s_b[64];
    const int patch = blockIdx.y;
    const int prow_base = blockIdx.x * kC1Band;     // first pooled row of this CTA
    const int y_in0 = prow_base * 2 - 1;            // first input row staged (with halo)
    const float* src = feats + size_t(patch) * H * W;
    for (int i = threadIdx.x; i < ROWS * (W + 2); i += 256) {
        const int r = i / (W + 2), c = i % (W + 2);
        const int y = y_in0 + r, x = c - 1;
        s_in[r][c] = (y >= 0 && y < H && x >= 0 && x < W) ? __ldg(src + y * W + x) : 0.f;
    }
    for (int i = threadIdx.x; i < 9 * 64; i += 256) s_w[i / 64][i % 64] = w[i];
    if (threadIdx.x < 64) s_b[threadIdx.x] = bias[threadIdx.x];
    __syncthreads();

    const int lr = threadIdx.x >> 5;            // pooled row within the band (= warp)
    const int pc = threadIdx.x & 31;            // pooled column (= lane)
    // 4x4 input window, each value duplicated into both halves of an f32x2 register
    unsigned long long in[4][4];
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        const float2 a = *reinterpret_cast<const float2*>(&s_in[lr * 2 + r][pc * 2]);
        const float2 b = *reinterpret_cast<const float2*>(&s_in[lr * 2 + r][pc * 2 + 2]);
        in[r][0] = pack_f32x2(a.x, a.x); in[r][1] = pack_f32x2(a.y, a.y);
        in[r][2] = pack_f32x2(b.x, b.x); in[r][3] = pack_f32x2(b.y, b.y);
    }
    const size_t obase = ((size_t(patch) * HP + prow_base + lr) * WP + pc) * 64;
    conv1_vggish_pixel(in, s_w, s_b, out_hi, out_lo, obase, f16);
}

// ---------------------------------------------------------------- CNN14: bn0 + conv + BN + ReLU
// grid = (ceil(T / 16) row bands, B clips); block = 256 threads = 64 mel columns x 4 row groups; a thread owns
// FOUR vertically adjacent pixels (rows 4q .. 4q+3 of the band) so every warp-uniform weight load feeds 4 x 8
// packed FMAs (one pixel per thread left the loop dominated by loads and epilogue: 23 % of the FMA rate).
constexpr int kC14Rows = 16;
__global__ void __launch_bounds__(256, 2) conv1_cnn14_kernel(const float* __restrict__ feats, int T,
                                                             const float* __restrict__ bn0_scale,
                                                             const float* __restrict__ bn0_shift,
                                                             const float* __restrict__ w, const float* __restrict__ bias,
                                                             __nv_bfloat16* __restrict__ out_hi,
                                                             __nv_bfloat16* __restrict__ out_lo,
                                                             uint8_t* __restrict__ out8p, int f16) {
    constexpr int W = 64;
    __shared__ float s_in[kC14Rows + 2][W + 2];
    __shared__ __align__(16) float s_w[9][64];
    __shared__ __align__(16) float s_b[64];
    const int clip = blockIdx.y;
    const int y0 = blockIdx.x * kC14Rows;
    const float* src = feats + size_t(clip) * T * W;
    for (int i = threadIdx.x; i < (kC14Rows + 2) * (W + 2); i += 256) {
        const int r = i / (W + 2), c = i % (W + 2);
        const int y = y0 - 1 + r, x = c - 1;
        float v = 0.f;      // conv zero padding happens AFTER bn0 (pann.py:249-255), so it stays a literal 0
        if (y >= 0 && y < T && x >= 0 && x < W) v = fmaf(__ldg(src + y * W + x), bn0_scale[x], bn0_shift[x]);
        s_in[r][c] = v;
    }
    for (int i = threadIdx.x; i < 9 * 64; i += 256) s_w[i / 64][i % 64] = w[i];
    if (threadIdx.x < 64) s_b[threadIdx.x] = bias[threadIdx.x];
    __syncthreads();

    const int q = threadIdx.x >> 6;             // row group: rows 4q .. 4q+3 of the band (warp-uniform)
    const int x = threadIdx.x & 63;
    unsigned long long in[6][3];                // 6 input rows x 3 columns, duplicated into f32x2
#pragma unroll
    for (int r = 0; r < 6; ++r)
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const float v = s_in[4 * q + r][x + c];
            in[r][c] = pack_f32x2(v, v);
        }
#pragma unroll 1
    for (int g = 0; g < 4; ++g) {
        unsigned long long acc[4][8];
#pragma unroll
        for (int r = 0; r < 4; ++r)
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[r][j] = 0ull;
#pragma unroll
        for (int ky = 0; ky < 3; ++ky)
#pragma unroll
            for (int kx = 0; kx < 3; ++kx) {
                const ulonglong2* wp = reinterpret_cast<const ulonglong2*>(&s_w[ky * 3 + kx][g * 16]);   // warp-uniform
                unsigned long long wv[8];
#pragma unroll
                for (int v = 0; v < 4; ++v) {
                    const ulonglong2 t = wp[v];
                    wv[2 * v] = t.x;
                    wv[2 * v + 1] = t.y;
                }
#pragma unroll
                for (int r = 0; r < 4; ++r)
#pragma unroll
                    for (int j = 0; j < 8; ++j) acc[r][j] = ffma2(in[r + ky][kx], wv[j], acc[r][j]);
            }
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            const int y = y0 + 4 * q + r;
            if (y < T) {
                float v[16];
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    float a0, a1;
                    unpack_f32x2(acc[r][j], a0, a1);
                    const float2 bb = *reinterpret_cast<const float2*>(&s_b[g * 16 + 2 * j]);
                    v[2 * j] = fmaxf(a0 + bb.x, 0.f);
                    v[2 * j + 1] = fmaxf(a1 + bb.y, 0.f);
                }
                store16(v, out_hi, out_lo, ((size_t(clip) * T + y) * W + x) * 64 + g * 16, f16);
                if (out8p) {    // e4m3 copy, W-padded [B][T][66][64] (gemm_tc.cu GemmParams::c64)
                    uint8_t* o8 = out8p + ((size_t(clip) * T + y) * (W + 2) + x + 1) * 64 + g * 16;
                    *reinterpret_cast<uint4*>(o8) = make_uint4(pack4_e4m3(v[0], v[1], v[2], v[3]), pack4_e4m3(v[4], v[5], v[6], v[7]),
                                                               pack4_e4m3(v[8], v[9], v[10], v[11]), pack4_e4m3(v[12], v[13], v[14], v[15]));
                    if (x == 0) *reinterpret_cast<uint4*>(o8 - 64) = make_uint4(0u, 0u, 0u, 0u);
                    if (x == W - 1) *reinterpret_cast<uint4*>(o8 + 64) = make_uint4(0u, 0u, 0u, 0u);
                }
            }
        }
    }
}

int launch_conv1_vggish(fadb_handle* h, const float* feats, int64_t n_patches, __nv_bfloat16* out_hi,
                        __nv_bfloat16* out_lo, cudaStream_t st) {
    if (n_patches <= 0) return FADB_OK;
    FADB_REQUIRE(n_patches <= 65535, "conv1: at most 65535 patches per batch");
    dim3 grid(48 / kC1Band, (unsigned)n_patches);
    conv1_vggish_kernel<<<grid, 256, 0, st>>>(feats, h->conv1_w, h->conv1_b, out_hi,
                                             h->precision == FADB_PREC_BF16X3 ? out_lo : nullptr, (int)prec_is_f16(h->precision));
    h->launches++;
    FADB_CUDA_CHECK(cudaGetLastError());
    return FADB_OK;
}

int launch_conv1_cnn14(fadb_handle* h, const float* feats, int64_t n_clips, int T, __nv_bfloat16* out_hi,
                       __nv_bfloat16* out_lo, uint8_t* out8p, cudaStream_t st) {
    if (n_clips <= 0) return FADB_OK;
    FADB_REQUIRE(n_clips <= 65535, "conv1: at most 65535 clips per batch");
    // (a tcgen05 version of this kernel, like the fused VGGish one, measured the same 0.22 ms per 64 clips: the
    // kernel is bound by writing its 8.45 MB of bf16 activations per clip to HBM, so the CUDA-core version stays)
    dim3 grid((unsigned)((T + kC14Rows - 1) / kC14Rows), (unsigned)n_clips);
    conv1_cnn14_kernel<<<grid, 256, 0, st>>>(feats, T, h->bn0_scale, h->bn0_shift, h->conv1_w, h->conv1_b, out_hi,
                                            h->precision == FADB_PREC_BF16X3 ? out_lo : nullptr, out8p,
                                            (int)prec_is_f16(h->precision));
    h->launches++;
    FADB_CUDA_CHECK(cudaGetLastError());
    return FADB_OK;
}

}  // namespace fadb
