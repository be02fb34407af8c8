// conv1_dev.cuh — device helpers shared by conv1.cu and the fused front-end + conv1 kernel (frontend.cu).
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <stdint.h>

namespace fadb {

__device__ __forceinline__ uint32_t pack2(float a, float b) {
    __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ float bfr(float x) { return __bfloat162float(__float2bfloat16_rn(x)); }
// activation pair in the handle's 16-bit format: bf16, or IEEE fp16 with saturation (FADB_PREC_FP16 / FP16X2)
__device__ __forceinline__ uint32_t pack2a(float a, float b, int f16) {
    if (f16) {
        uint32_t r;
        asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(b), "f"(a));
        return r;
    }
    return pack2(a, b);
}

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

// one 32-byte global store (256-bit LSU access of sm_100): 16 bf16 channels = one full sector per instruction
__device__ __forceinline__ void st256(void* p, const uint4& a, const uint4& b) {
    asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
                 ::"l"(p), "r"(a.x), "r"(a.y), "r"(a.z), "r"(a.w), "r"(b.x), "r"(b.y), "r"(b.z), "r"(b.w)
                 : "memory");
}

__device__ __forceinline__ void store16(const float (&v)[16], __nv_bfloat16* hi, __nv_bfloat16* lo, size_t o, int f16) {
    uint4 a, b;
    a.x = pack2a(v[0], v[1], f16); a.y = pack2a(v[2], v[3], f16); a.z = pack2a(v[4], v[5], f16); a.w = pack2a(v[6], v[7], f16);
    b.x = pack2a(v[8], v[9], f16); b.y = pack2a(v[10], v[11], f16); b.z = pack2a(v[12], v[13], f16); b.w = pack2a(v[14], v[15], f16);
    st256(hi + o, a, b);
    if (lo) {
        float r[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) r[j] = v[j] - bfr(v[j]);
        a.x = pack2(r[0], r[1]); a.y = pack2(r[2], r[3]); a.z = pack2(r[4], r[5]); a.w = pack2(r[6], r[7]);
        b.x = pack2(r[8], r[9]); b.y = pack2(r[10], r[11]); b.z = pack2(r[12], r[13]); b.w = pack2(r[14], r[15]);
        st256(lo + o, a, b);
    }
}


// One pooled output pixel of VGGish conv1 (+bias +ReLU +2x2 maxpool), all 64 channels in 4 groups of 16 with
// warp-uniform weight loads.  `in` = the 4x4 input window, every value duplicated into an f32x2 register.
__device__ __forceinline__ void conv1_vggish_pixel(const unsigned long long (&in)[4][4], const float (*s_w)[64],
                                                   const float* s_b, __nv_bfloat16* out_hi, __nv_bfloat16* out_lo,
                                                   size_t obase, int f16) {
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
        store16(v, out_hi, out_lo, obase + g * 16, f16);
    }
}

// ---- conv1 as a tcgen05 GEMM (K = 32, see frontend.cu fadb_vggish_front_conv1_tc_kernel) --------------------------
// Operand tiles are un-swizzled core-matrix tiles: element (row r, 16-byte K chunk c) at
// (r >> 3) * kC1Sbo + c * kC1Lbo + (r & 7) * 16.
constexpr uint32_t kC1Lbo = 128;                 // core matrix -> next one along K
constexpr uint32_t kC1Sbo = 512;                 // 8-row group -> next one (4 K chunks each)
constexpr int kC1ATileBytes = 128 * 64;          // 128 pixels x K=32 bf16
constexpr int kC1BBytes = 64 * 64;               // 64 couts x K=32 bf16

// {bf16 hi, bf16 lo} of x packed in one word (hi in the low half), x = hi + lo to ~2^-17
__device__ __forceinline__ uint32_t split_hi_lo(float x) {
    const __nv_bfloat16 hi = __float2bfloat16_rn(x);
    const __nv_bfloat16 lo = __float2bfloat16_rn(x - __bfloat162float(hi));
    return (uint32_t)__bfloat16_as_ushort(hi) | ((uint32_t)__bfloat16_as_ushort(lo) << 16);
}

// B operand [64 couts][K=32] = [w_hi(tap 0..7) | w_hi(tap 0..7) | w_lo(tap 0..7) | w_hi8, w_hi8, w_lo8, 0...];
// conv_w is [tap][cout] fp32; call with tid = 0..255 (thread = cout n, chunk c)
__device__ __forceinline__ void c1tc_build_b(uint8_t* s_b, const float* __restrict__ conv_w, int tid) {
    const int n = tid >> 2, c = tid & 3;
    uint4 v = make_uint4(0u, 0u, 0u, 0u);
    if (c < 3) {
        float w[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const float f = __ldg(conv_w + j * 64 + n);
            w[j] = (c == 2) ? f - bfr(f) : f;
        }
        v.x = pack2(w[0], w[1]); v.y = pack2(w[2], w[3]); v.z = pack2(w[4], w[5]); v.w = pack2(w[6], w[7]);
    } else {
        const float f = __ldg(conv_w + 8 * 64 + n);
        v.x = pack2(f, f);
        v.y = pack2(f - bfr(f), 0.f);
    }
    *reinterpret_cast<uint4*>(s_b + (n >> 3) * kC1Sbo + c * kC1Lbo + (n & 7) * 16) = v;
}

// A row = [x_hi(tap 0..7) | x_lo(tap 0..7) | x_hi(tap 0..7) | x_hi8, x_lo8, x_hi8, 0...] from the 9 {hi, lo} words
// of a pixel's 3x3 window; arow = address of the row's chunk 0
__device__ __forceinline__ void c1tc_store_a_row(uint8_t* arow, const uint32_t (&w)[9]) {
    uint4 hi4, lo4;
    hi4.x = __byte_perm(w[0], w[1], 0x5410); hi4.y = __byte_perm(w[2], w[3], 0x5410);
    hi4.z = __byte_perm(w[4], w[5], 0x5410); hi4.w = __byte_perm(w[6], w[7], 0x5410);
    lo4.x = __byte_perm(w[0], w[1], 0x7632); lo4.y = __byte_perm(w[2], w[3], 0x7632);
    lo4.z = __byte_perm(w[4], w[5], 0x7632); lo4.w = __byte_perm(w[6], w[7], 0x7632);
    *reinterpret_cast<uint4*>(arow) = hi4;
    *reinterpret_cast<uint4*>(arow + kC1Lbo) = lo4;
    *reinterpret_cast<uint4*>(arow + 2 * kC1Lbo) = hi4;
    *reinterpret_cast<uint4*>(arow + 3 * kC1Lbo) = make_uint4(w[8], w[8] & 0xffffu, 0u, 0u);
}

}  // namespace fadb
