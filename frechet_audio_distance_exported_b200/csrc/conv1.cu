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

namespace fadb {

__device__ __forceinline__ uint32_t pack2(float a, float b) {
    __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ float bfr(float x) { return __bfloat162float(__float2bfloat16_rn(x)); }

__device__ __forceinline__ void store16(const float (&v)[16], __nv_bfloat16* hi, __nv_bfloat16* lo, size_t o) {
    uint4 a, b;
    a.x = pack2(v[0], v[1]); a.y = pack2(v[2], v[3]); a.z = pack2(v[4], v[5]); a.w = pack2(v[6], v[7]);
    b.x = pack2(v[8], v[9]); b.y = pack2(v[10], v[11]); b.z = pack2(v[12], v[13]); b.w = pack2(v[14], v[15]);
    uint4* d = reinterpret_cast<uint4*>(hi + o);
    d[0] = a;
    d[1] = b;
    if (lo) {
        float r[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) r[j] = v[j] - bfr(v[j]);
        a.x = pack2(r[0], r[1]); a.y = pack2(r[2], r[3]); a.z = pack2(r[4], r[5]); a.w = pack2(r[6], r[7]);
        b.x = pack2(r[8], r[9]); b.y = pack2(r[10], r[11]); b.z = pack2(r[12], r[13]); b.w = pack2(r[14], r[15]);
        uint4* e = reinterpret_cast<uint4*>(lo + o);
        e[0] = a;
        e[1] = b;
    }
}

// ---------------------------------------------------------------- VGGish: conv + ReLU + maxpool
// grid = (6 bands of 8 pooled rows, P patches); block = 256 = 64 pooled pixels x 4 channel groups of 16.
// The 18 input rows of the band (16 + halo), the 9x64 weights and the bias are staged once per CTA and
// reused by 4 iterations of 2 pooled rows each.
constexpr int kC1Band = 8;                       // pooled rows per CTA
__global__ void __launch_bounds__(256, 2) conv1_vggish_kernel(const float* __restrict__ feats, const float* __restrict__ w,
                                                           const float* __restrict__ bias,
                                                           __nv_bfloat16* __restrict__ out_hi,
                                                           __nv_bfloat16* __restrict__ out_lo) {
    constexpr int H = 96, W = 64, HP = 48, WP = 32;
    constexpr int ROWS = 2 * kC1Band + 2;
    __shared__ float s_in[ROWS][W + 2];
    __shared__ __align__(16) float s_w[9][64];
    __shared__ float s_b[64];
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

    const int g = threadIdx.x & 3;              // channel group
    const int pp = threadIdx.x >> 2;            // pooled pixel within an iteration: 0..63
    const int pr = pp >> 5, pc = pp & 31;
#pragma unroll 1
    for (int itr = 0; itr < kC1Band / 2; ++itr) {
        const int lr = itr * 2 + pr;            // pooled row within the band
        float acc[4][16];                       // 4 positions of the pooling window x 16 channels
#pragma unroll
        for (int q = 0; q < 4; ++q)
#pragma unroll
            for (int j = 0; j < 16; ++j) acc[q][j] = 0.f;
#pragma unroll 1
        for (int ky = 0; ky < 3; ++ky) {
            // the two input rows this kernel row touches, 4 columns each (8-byte aligned pairs)
            const float2 r0a = *reinterpret_cast<const float2*>(&s_in[lr * 2 + ky][pc * 2]);
            const float2 r0b = *reinterpret_cast<const float2*>(&s_in[lr * 2 + ky][pc * 2 + 2]);
            const float2 r1a = *reinterpret_cast<const float2*>(&s_in[lr * 2 + ky + 1][pc * 2]);
            const float2 r1b = *reinterpret_cast<const float2*>(&s_in[lr * 2 + ky + 1][pc * 2 + 2]);
            const float in0[4] = {r0a.x, r0a.y, r0b.x, r0b.y};
            const float in1[4] = {r1a.x, r1a.y, r1b.x, r1b.y};
#pragma unroll
            for (int kx = 0; kx < 3; ++kx) {
                const float4* wp = reinterpret_cast<const float4*>(&s_w[ky * 3 + kx][g * 16]);
                float wv[16];
#pragma unroll
                for (int v = 0; v < 4; ++v) {
                    const float4 t = wp[v];
                    wv[v * 4 + 0] = t.x; wv[v * 4 + 1] = t.y; wv[v * 4 + 2] = t.z; wv[v * 4 + 3] = t.w;
                }
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    acc[0][j] = fmaf(in0[kx], wv[j], acc[0][j]);
                    acc[1][j] = fmaf(in0[kx + 1], wv[j], acc[1][j]);
                    acc[2][j] = fmaf(in1[kx], wv[j], acc[2][j]);
                    acc[3][j] = fmaf(in1[kx + 1], wv[j], acc[3][j]);
                }
            }
        }
        float v[16];
#pragma unroll
        for (int j = 0; j < 16; ++j)            // relu(max(.)+b) == max(relu(.+b))
            v[j] = fmaxf(fmaxf(fmaxf(acc[0][j], acc[1][j]), fmaxf(acc[2][j], acc[3][j])) + s_b[g * 16 + j], 0.f);
        const size_t o = ((size_t(patch) * HP + prow_base + lr) * WP + pc) * 64 + g * 16;
        store16(v, out_hi, out_lo, o);
    }
}

// ---------------------------------------------------------------- CNN14: bn0 + conv + BN + ReLU
// grid = (T rows, B clips); block = 256 = 64 mel columns x 4 channel groups of 16
__global__ void __launch_bounds__(256) conv1_cnn14_kernel(const float* __restrict__ feats, int T,
                                                          const float* __restrict__ bn0_scale,
                                                          const float* __restrict__ bn0_shift,
                                                          const float* __restrict__ w, const float* __restrict__ bias,
                                                          __nv_bfloat16* __restrict__ out_hi,
                                                          __nv_bfloat16* __restrict__ out_lo) {
    constexpr int W = 64;
    __shared__ float s_in[3][W + 2];
    __shared__ __align__(16) float s_w[9][64];
    __shared__ float s_b[64];
    const int clip = blockIdx.y;
    const int y0 = blockIdx.x;
    const float* src = feats + size_t(clip) * T * W;
    for (int i = threadIdx.x; i < 3 * (W + 2); i += 256) {
        const int r = i / (W + 2), c = i % (W + 2);
        const int y = y0 - 1 + r, x = c - 1;
        float v = 0.f;
        if (y >= 0 && y < T && x >= 0 && x < W) v = fmaf(__ldg(src + y * W + x), bn0_scale[x], bn0_shift[x]);
        s_in[r][c] = v;
    }
    for (int i = threadIdx.x; i < 9 * 64; i += 256) s_w[i / 64][i % 64] = w[i];
    if (threadIdx.x < 64) s_b[threadIdx.x] = bias[threadIdx.x];
    __syncthreads();

    const int g = threadIdx.x & 3;
    const int x = threadIdx.x >> 2;
    float acc[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) acc[j] = 0.f;
#pragma unroll
    for (int ky = 0; ky < 3; ++ky)
#pragma unroll
        for (int kx = 0; kx < 3; ++kx) {
            const float xv = s_in[ky][x + kx];
            const float4* wp = reinterpret_cast<const float4*>(&s_w[ky * 3 + kx][g * 16]);
#pragma unroll
            for (int v = 0; v < 4; ++v) {
                const float4 wv = wp[v];
                acc[v * 4 + 0] = fmaf(xv, wv.x, acc[v * 4 + 0]);
                acc[v * 4 + 1] = fmaf(xv, wv.y, acc[v * 4 + 1]);
                acc[v * 4 + 2] = fmaf(xv, wv.z, acc[v * 4 + 2]);
                acc[v * 4 + 3] = fmaf(xv, wv.w, acc[v * 4 + 3]);
            }
        }
    float v[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) v[j] = fmaxf(acc[j] + s_b[g * 16 + j], 0.f);
    const size_t o = ((size_t(clip) * T + y0) * W + x) * 64 + g * 16;
    store16(v, out_hi, out_lo, o);
}

int launch_conv1_vggish(fadb_handle* h, const float* feats, int64_t n_patches, __nv_bfloat16* out_hi,
                        __nv_bfloat16* out_lo, cudaStream_t st) {
    if (n_patches <= 0) return FADB_OK;
    FADB_REQUIRE(n_patches <= 65535, "conv1: at most 65535 patches per batch");
    dim3 grid(48 / kC1Band, (unsigned)n_patches);
    conv1_vggish_kernel<<<grid, 256, 0, st>>>(feats, h->conv1_w, h->conv1_b, out_hi,
                                             h->precision == FADB_PREC_BF16X3 ? out_lo : nullptr);
    h->launches++;
    FADB_CUDA_CHECK(cudaGetLastError());
    return FADB_OK;
}

int launch_conv1_cnn14(fadb_handle* h, const float* feats, int64_t n_clips, int T, __nv_bfloat16* out_hi,
                       __nv_bfloat16* out_lo, cudaStream_t st) {
    if (n_clips <= 0) return FADB_OK;
    FADB_REQUIRE(n_clips <= 65535, "conv1: at most 65535 clips per batch");
    dim3 grid((unsigned)T, (unsigned)n_clips);
    conv1_cnn14_kernel<<<grid, 256, 0, st>>>(feats, T, h->bn0_scale, h->bn0_shift, h->conv1_w, h->conv1_b, out_hi,
                                            h->precision == FADB_PREC_BF16X3 ? out_lo : nullptr);
    h->launches++;
    FADB_CUDA_CHECK(cudaGetLastError());
    return FADB_OK;
}

}  // namespace fadb
