// pack.cu — weight repacking on the device: PyTorch layouts -> K-major bf16 (hi/lo planes) for the
// tcgen05 B operand, eval-mode BatchNorm folding (models/pann.py:177-178,190-191; eps = 1e-5).
#include <cmath>
#include <cstring>

#include "common.cuh"

namespace fadb {

// x -> {hi, lo} 16-bit words with x ~ hi + lo: bf16, or IEEE fp16 (lo may be an fp16 subnormal: absolute precision 6e-8)
__device__ __forceinline__ void split16(float v, bool f16, __nv_bfloat16* hi, __nv_bfloat16* lo) {
    if (f16) {
        const __half h = __float2half_rn(fminf(fmaxf(v, -65504.f), 65504.f));
        *reinterpret_cast<__half*>(hi) = h;
        if (lo) *reinterpret_cast<__half*>(lo) = __float2half_rn(v - __half2float(h));
    } else {
        const __nv_bfloat16 h = __float2bfloat16_rn(v);
        *hi = h;
        if (lo) *lo = __float2bfloat16_rn(v - __bfloat162float(h));
    }
}

// [Cout][Cin][k][k] fp32  ->  [Cout][tap = ky*k+kx][Cin] bf16 hi (+lo), optionally scaled per Cout
__global__ void pack_conv_weight_kernel(const float* __restrict__ w, int Cout, int Cin, int kk,
                                        const float* __restrict__ scale, __nv_bfloat16* __restrict__ hi,
                                        __nv_bfloat16* __restrict__ lo, bool f16) {
    const size_t total = size_t(Cout) * Cin * kk;
    for (size_t i = blockIdx.x * size_t(blockDim.x) + threadIdx.x; i < total; i += size_t(gridDim.x) * blockDim.x) {
        const int ci = int(i % Cin);
        const size_t t = i / Cin;
        const int tap = int(t % kk);
        const int co = int(t / kk);
        float v = w[(size_t(co) * Cin + ci) * kk + tap];
        if (scale) v *= scale[co];
        split16(v, f16, hi + i, lo ? lo + i : nullptr);
    }
}

__global__ void split_kernel(const float* __restrict__ x, size_t n, __nv_bfloat16* __restrict__ hi,
                             __nv_bfloat16* __restrict__ lo, bool f16) {
    for (size_t i = blockIdx.x * size_t(blockDim.x) + threadIdx.x; i < n; i += size_t(gridDim.x) * blockDim.x)
        split16(x[i], f16, hi + i, lo ? lo + i : nullptr);
}

__global__ void fold_bn_kernel(const float* gamma, const float* beta, const float* mean, const float* var, int C,
                               float* scale, float* shift) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c < C) {
        const float s = gamma[c] / sqrtf(var[c] + 1e-5f);
        scale[c] = s;
        shift[c] = beta[c] - mean[c] * s;
    }
}

// |v| maximum as the bit pattern of a non-negative float (atomicMax on ints orders them like the floats)
__global__ void absmax_weight_kernel(const float* __restrict__ w, int Cout, int Cin, int kk, const float* __restrict__ scale,
                                     int* __restrict__ out_bits) {
    const size_t total = size_t(Cout) * Cin * kk;
    float m = 0.f;
    for (size_t i = blockIdx.x * size_t(blockDim.x) + threadIdx.x; i < total; i += size_t(gridDim.x) * blockDim.x) {
        float v = w[i];
        if (scale) v *= scale[i / (size_t(Cin) * kk)];
        m = fmaxf(m, fabsf(v));
    }
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((threadIdx.x & 31) == 0) atomicMax(out_bits, __float_as_int(m));
}

__device__ __forceinline__ uint8_t to_e4m3(float v) {
    uint16_t r;
    asm("cvt.rn.satfinite.e4m3x2.f32 %0, %1, %2;" : "=h"(r) : "f"(0.f), "f"(v));
    return (uint8_t)(r & 0xffu);
}

__global__ void pack_lo8_kernel(const float* __restrict__ w, int Cout, int Cin, int kk, const float* __restrict__ scale,
                                float inv_lo_scale, uint8_t* __restrict__ out) {
    const size_t total = size_t(Cout) * Cin * kk;
    for (size_t i = blockIdx.x * size_t(blockDim.x) + threadIdx.x; i < total; i += size_t(gridDim.x) * blockDim.x) {
        const int ci = int(i % Cin);
        const size_t t = i / Cin;
        const int tap = int(t % kk);
        const int co = int(t / kk);
        float v = w[(size_t(co) * Cin + ci) * kk + tap];
        if (scale) v *= scale[co];
        const float hi = __half2float(__float2half_rn(fminf(fmaxf(v, -65504.f), 65504.f)));
        out[i] = to_e4m3((v - hi) * inv_lo_scale);
    }
}

__global__ void quantize_e4m3_kernel(const float* __restrict__ x, size_t n, uint8_t* __restrict__ out) {
    for (size_t i = blockIdx.x * size_t(blockDim.x) + threadIdx.x; i < n; i += size_t(gridDim.x) * blockDim.x)
        out[i] = to_e4m3(x[i]);
}

// one thread per 16 output bytes; pixel column 0 and W+1 of every row are zero
__global__ void quantize_e4m3_wpad_kernel(const float* __restrict__ x32, const __half* __restrict__ x16, size_t rows, int W,
                                          uint8_t* __restrict__ out) {
    const size_t total = rows * (size_t)(W + 2) * 4;
    for (size_t i = blockIdx.x * size_t(blockDim.x) + threadIdx.x; i < total; i += size_t(gridDim.x) * blockDim.x) {
        const int q = int(i & 3);
        const size_t px = i >> 2;
        const int xp = int(px % (size_t)(W + 2));
        const size_t row = px / (size_t)(W + 2);
        uint4 o = make_uint4(0u, 0u, 0u, 0u);
        if (xp > 0 && xp <= W) {
            const size_t src = (row * W + (xp - 1)) * 64 + q * 16;
            uint32_t w[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                uint32_t v = 0;
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const float f = x32 ? x32[src + 4 * k + e] : __half2float(x16[src + 4 * k + e]);
                    v |= (uint32_t)to_e4m3(f) << (8 * e);
                }
                w[k] = v;
            }
            o = make_uint4(w[0], w[1], w[2], w[3]);
        }
        reinterpret_cast<uint4*>(out)[i] = o;
    }
}

static inline int grid_for(size_t n) {
    size_t g = (n + 255) / 256;
    return (int)(g > 4096 ? 4096 : (g ? g : 1));
}

int pack_conv_weight(fadb_handle* h, const float* w_oihw, int Cout, int Cin, int ksize, const float* scale,
                     __nv_bfloat16* w_hi, __nv_bfloat16* w_lo, bool f16, cudaStream_t st) {
    const size_t total = size_t(Cout) * Cin * ksize * ksize;
    pack_conv_weight_kernel<<<grid_for(total), 256, 0, st>>>(w_oihw, Cout, Cin, ksize * ksize, scale, w_hi, w_lo, f16);
    h->launches++;
    FADB_CUDA_CHECK(cudaGetLastError());
    return FADB_OK;
}

int pack_conv_weight_lo8(fadb_handle* h, const float* w_oihw, int Cout, int Cin, int ksize, const float* scale,
                         uint8_t* w8, float* lo_scale, cudaStream_t st) {
    const size_t total = size_t(Cout) * Cin * ksize * ksize;
    int* bits = nullptr;
    FADB_CUDA_CHECK(cudaMalloc(&bits, sizeof(int)));
    FADB_CUDA_CHECK(cudaMemsetAsync(bits, 0, sizeof(int), st));
    absmax_weight_kernel<<<grid_for(total), 256, 0, st>>>(w_oihw, Cout, Cin, ksize * ksize, scale, bits);
    int hb = 0;
    FADB_CUDA_CHECK(cudaMemcpyAsync(&hb, bits, sizeof(int), cudaMemcpyDeviceToHost, st));
    FADB_CUDA_CHECK(cudaStreamSynchronize(st));
    cudaFree(bits);
    float wmax;
    memcpy(&wmax, &hb, sizeof(float));
    // |w - fp16(w)| <= 2^-11 * 2^ceil(log2 wmax); put that bound at 256 (e4m3 tops out at 448)
    int e = 0;
    if (wmax > 0.f) frexpf(wmax, &e);                    // wmax = f * 2^e, f in [0.5, 1)
    const float s = ldexpf(1.f, e - 11 - 8);             // lo_scale: residual / s <= 256
    *lo_scale = s;
    pack_lo8_kernel<<<grid_for(total), 256, 0, st>>>(w_oihw, Cout, Cin, ksize * ksize, scale, 1.f / s, w8);
    h->launches += 2;
    FADB_CUDA_CHECK(cudaGetLastError());
    return FADB_OK;
}

int quantize_e4m3(fadb_handle* h, const float* x, int64_t n, uint8_t* out, cudaStream_t st) {
    if (n <= 0) return FADB_OK;
    quantize_e4m3_kernel<<<grid_for((size_t)n), 256, 0, st>>>(x, (size_t)n, out);
    h->launches++;
    FADB_CUDA_CHECK(cudaGetLastError());
    return FADB_OK;
}

int quantize_e4m3_wpad(fadb_handle* h, const float* x32, const __nv_bfloat16* x16, int64_t rows, int W, uint8_t* out,
                       cudaStream_t st) {
    if (rows <= 0) return FADB_OK;
    const size_t total = (size_t)rows * (size_t)(W + 2) * 4;
    quantize_e4m3_wpad_kernel<<<grid_for(total), 256, 0, st>>>(x32, reinterpret_cast<const __half*>(x16), (size_t)rows, W, out);
    h->launches++;
    FADB_CUDA_CHECK(cudaGetLastError());
    return FADB_OK;
}

int split_f32_to_bf16(fadb_handle* h, const float* x, int64_t n, __nv_bfloat16* hi, __nv_bfloat16* lo, bool f16,
                      cudaStream_t st) {
    if (n <= 0) return FADB_OK;
    split_kernel<<<grid_for((size_t)n), 256, 0, st>>>(x, (size_t)n, hi, lo, f16);
    h->launches++;
    FADB_CUDA_CHECK(cudaGetLastError());
    return FADB_OK;
}

int fold_bn(fadb_handle* h, const float* gamma, const float* beta, const float* mean, const float* var, int C,
            float* scale, float* shift, cudaStream_t st) {
    fold_bn_kernel<<<(C + 127) / 128, 128, 0, st>>>(gamma, beta, mean, var, C, scale, shift);
    h->launches++;
    FADB_CUDA_CHECK(cudaGetLastError());
    return FADB_OK;
}

}  // namespace fadb
