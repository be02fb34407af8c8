// conv1_dev.cuh — device helpers shared by conv1.cu and the fused front-end + conv1 kernel (frontend.cu).
#pragma once
#include <cuda_bf16.h>
#include <stdint.h>

namespace fadb {

__device__ __forceinline__ uint32_t pack2(float a, float b) {
    __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ float bfr(float x) { return __bfloat162float(__float2bfloat16_rn(x)); }

// Packed fp32x2 FMA (sm_100+): d.lo = a.lo*b.lo + c.lo, d.hi = a.hi*b.hi + c.hi in ONE issue slot.  A 3-register
// scalar FFMA issues every other cycle per scheduler on Blackwell, so the packed form is what reaches the
// 128-lane fp32 rate in an FMA-bound loop.
__device__ __forceinline__ unsigned long long pack_f32x2(float lo, float hi) {
    unsigned long long r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ void unpack_f32x2(unsigned long long v, float& lo, float& hi) {
    asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ unsigned long long ffma2(unsigned long long a, unsigned long long b, unsigned long long c) {
    unsigned long long d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}

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


// One pooled output pixel of VGGish conv1 (+bias +ReLU +2x2 maxpool), all 64 channels in 4 groups of 16 with
// warp-uniform weight loads.  `in` = the 4x4 input window, every value duplicated into an f32x2 register.
__device__ __forceinline__ void conv1_vggish_pixel(const unsigned long long (&in)[4][4], const float (*s_w)[64],
                                                   const float* s_b, __nv_bfloat16* out_hi, __nv_bfloat16* out_lo,
                                                   size_t obase) {
#pragma unroll 1
    for (int g = 0; g < 4; ++g) {               // 16 output channels per pass
        unsigned long long acc[4][8];           // 4 positions of the pooling window x 8 channel PAIRS
#pragma unroll
        for (int q = 0; q < 4; ++q)
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[q][j] = 0ull;
#pragma unroll
        for (int ky = 0; ky < 3; ++ky)
#pragma unroll
            for (int kx = 0; kx < 3; ++kx) {
                const ulonglong2* wp = reinterpret_cast<const ulonglong2*>(&s_w[ky * 3 + kx][g * 16]);   // warp-uniform
                unsigned long long wv[8];
#pragma unroll
                for (int v = 0; v < 4; ++v) {
                    const ulonglong2 t = wp[v];
                    wv[v * 2 + 0] = t.x;
                    wv[v * 2 + 1] = t.y;
                }
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    acc[0][j] = ffma2(in[ky][kx], wv[j], acc[0][j]);
                    acc[1][j] = ffma2(in[ky][kx + 1], wv[j], acc[1][j]);
                    acc[2][j] = ffma2(in[ky + 1][kx], wv[j], acc[2][j]);
                    acc[3][j] = ffma2(in[ky + 1][kx + 1], wv[j], acc[3][j]);
                }
            }
        float v[16];
#pragma unroll
        for (int j = 0; j < 8; ++j) {           // relu(max(.)+b) == max(relu(.+b))
            float a0, a1, b0, b1, c0, c1, d0, d1;
            unpack_f32x2(acc[0][j], a0, a1); unpack_f32x2(acc[1][j], b0, b1);
            unpack_f32x2(acc[2][j], c0, c1); unpack_f32x2(acc[3][j], d0, d1);
            const float2 bb = *reinterpret_cast<const float2*>(&s_b[g * 16 + 2 * j]);
            v[2 * j] = fmaxf(fmaxf(fmaxf(a0, b0), fmaxf(c0, d0)) + bb.x, 0.f);
            v[2 * j + 1] = fmaxf(fmaxf(fmaxf(a1, b1), fmaxf(c1, d1)) + bb.y, 0.f);
        }
        store16(v, out_hi, out_lo, obase + g * 16);
    }
}

}  // namespace fadb
