// api.cu — the C ABI of libfadb200.so (see include/fadb.h): handle lifetime, weight ingestion,
// the per-model layer schedules and the host-buffer end-to-end path.
#include <cstdarg>
#include <cstdlib>
#include <cstring>

#include "common.cuh"

namespace fadb {

static thread_local char g_err[1024] = "";

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

static int embed_dim(int model) {
    switch (model) {
        case FADB_MODEL_VGGISH: return 128;
        case FADB_MODEL_PANN8K:
        case FADB_MODEL_PANN16K:
        case FADB_MODEL_PANN32K: return 2048;
        case FADB_MODEL_CLAP: return 512;
        default: return -1;
    }
}

// ---------------------------------------------------------------- conv1 packing ([64][1][3][3] -> [9][64])
__global__ void pack_conv1_kernel(const float* __restrict__ w, const float* __restrict__ scale, float* __restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < 9 * 64) {
        const int tap = i / 64, co = i % 64;
        float v = w[co * 9 + tap];
        if (scale) v *= scale[co];
        out[i] = v;
    }
}

// ---------------------------------------------------------------- CNN14 tail
// x [B, Ht, Wf, C] bf16 (hi + optional lo) -> mean over Wf, then max over Ht + mean over Ht  (pann.py:263-268)
__global__ void cnn14_global_pool_kernel(const __nv_bfloat16* __restrict__ hi, const __nv_bfloat16* __restrict__ lo,
                                         int Ht, int Wf, int C, __nv_bfloat16* __restrict__ ohi,
                                         __nv_bfloat16* __restrict__ olo, int f16, uint8_t* __restrict__ o8) {
    const int b = blockIdx.y;
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    const size_t base = (size_t)b * Ht * Wf * C + c;
    float mx = -INFINITY, sum = 0.f;
    for (int t = 0; t < Ht; ++t) {
        float m = 0.f;
        for (int f = 0; f < Wf; ++f) {
            const size_t o = base + ((size_t)t * Wf + f) * C;
            float v = f16 ? __half2float(reinterpret_cast<const __half*>(hi)[o]) : __bfloat162float(hi[o]);
            if (lo) v += __bfloat162float(lo[o]);
            m += v;
        }
        m /= (float)Wf;
        mx = fmaxf(mx, m);
        sum += m;
    }
    const float r = mx + sum / (float)Ht;
    if (f16) {
        reinterpret_cast<__half*>(ohi)[(size_t)b * C + c] = __float2half_rn(fminf(r, 65504.f));
        if (o8) {
            uint16_t q;
            asm("cvt.rn.satfinite.e4m3x2.f32 %0, %1, %2;" : "=h"(q) : "f"(0.f), "f"(r));
            o8[(size_t)b * C + c] = (uint8_t)(q & 0xffu);
        }
        return;
    }
    const __nv_bfloat16 h = __float2bfloat16_rn(r);
    ohi[(size_t)b * C + c] = h;
    if (olo) olo[(size_t)b * C + c] = __float2bfloat16_rn(r - __bfloat162float(h));
}

__global__ void l2_normalize_kernel(float* __restrict__ x, int d) {   // F.normalize(dim=-1), eps 1e-12
    __shared__ float red[32];
    float* row = x + (size_t)blockIdx.x * d;
    float s = 0.f;
    for (int i = threadIdx.x; i < d; i += blockDim.x) s += row[i] * row[i];
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x < 32) {
        float t = (threadIdx.x < (blockDim.x >> 5)) ? red[threadIdx.x] : 0.f;
        for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
        if (threadIdx.x == 0) red[0] = t;
    }
    __syncthreads();
    const float inv = 1.f / fmaxf(sqrtf(red[0]), 1e-12f);
    for (int i = threadIdx.x; i < d; i += blockDim.x) row[i] *= inv;
}

int launch_cnn14_global_pool(fadb_handle* h, const __nv_bfloat16* x_hi, const __nv_bfloat16* x_lo, int64_t B, int Ht,
                             int Wf, int C, __nv_bfloat16* out_hi, __nv_bfloat16* out_lo, uint8_t* out8, cudaStream_t st) {
    dim3 grid((C + 255) / 256, (unsigned)B);
    const bool x3 = h->precision == FADB_PREC_BF16X3;
    cnn14_global_pool_kernel<<<grid, 256, 0, st>>>(x_hi, x3 ? x_lo : nullptr, Ht, Wf, C, out_hi, x3 ? out_lo : nullptr,
                                                   (int)prec_is_f16(h->precision),
                                                   (h->precision == FADB_PREC_FP16X2 && h->lo_fp8) ? out8 : nullptr);
    h->launches++;
    FADB_CUDA_CHECK(cudaGetLastError());
    return FADB_OK;
}

int launch_l2_normalize(fadb_handle* h, float* x, int64_t rows, int d, cudaStream_t st) {
    if (rows <= 0) return FADB_OK;
    l2_normalize_kernel<<<(unsigned)rows, 256, 0, st>>>(x, d);
    h->launches++;
    FADB_CUDA_CHECK(cudaGetLastError());
    return FADB_OK;
}

// ---------------------------------------------------------------- weights
static void free_layers(fadb_handle* h) {
    for (auto& L : h->layers) {
        if (L.w_hi) cudaFree(L.w_hi);
        if (L.w_lo) cudaFree(L.w_lo);
        if (L.w8) cudaFree(L.w8);
        if (L.bias) cudaFree(L.bias);
    }
    h->layers.clear();
    if (h->conv1_w) cudaFree(h->conv1_w);
    if (h->conv1_b) cudaFree(h->conv1_b);
    if (h->bn0_scale) cudaFree(h->bn0_scale);
    if (h->bn0_shift) cudaFree(h->bn0_shift);
    h->conv1_w = h->conv1_b = h->bn0_scale = h->bn0_shift = nullptr;
    h->weights_ready = false;
}
static void free_staged(fadb_handle* h) {
    for (auto& kv : h->staged)
        if (kv.second.dev) cudaFree(kv.second.dev);
    h->staged.clear();
}

static int get_staged(fadb_handle* h, const std::string& name, std::initializer_list<int64_t> shape, const float** out) {
    auto it = h->staged.find(name);
    if (it == h->staged.end()) {
        set_error("missing weight tensor '%s'", name.c_str());
        return FADB_E_STATE;
    }
    const HostTensor& t = it->second;
    bool ok = t.shape.size() == shape.size();
    size_t i = 0;
    for (int64_t s : shape) {
        if (ok && t.shape[i] != s) ok = false;
        ++i;
    }
    if (!ok) {
        set_error("weight tensor '%s' has the wrong shape", name.c_str());
        return FADB_E_INVALID;
    }
    *out = t.dev;
    return FADB_OK;
}

// pack one tensor-core layer from a staged weight (+ optional bias / folded BN)
static int add_layer(fadb_handle* h, const float* w, int Cout, int Cin, int ksize, const float* scale,
                     const float* bias_or_shift, cudaStream_t st) {
    PackedLayer L;
    L.N = Cout;
    L.Cin = Cin;
    L.taps = ksize * ksize;
    L.K = L.taps * Cin;
    const size_t n = (size_t)L.N * L.K;
    // 16-bit planes in the format of the handle's precision; the lo plane (w - hi) only for the split-weight modes
    L.f16 = prec_is_f16(h->precision) ? 1 : 0;
    const bool want_lo = h->precision == FADB_PREC_BF16X3 || h->precision == FADB_PREC_FP16X2;
    FADB_CUDA_CHECK(cudaMalloc(&L.w_hi, n * sizeof(__nv_bfloat16)));
    if (want_lo) FADB_CUDA_CHECK(cudaMalloc(&L.w_lo, n * sizeof(__nv_bfloat16)));
    FADB_CUDA_CHECK(cudaMalloc(&L.bias, (size_t)Cout * sizeof(float)));
    FADB_CHECK(pack_conv_weight(h, w, Cout, Cin, ksize, scale, L.w_hi, L.w_lo, L.f16 != 0, st));
    // e4m3 low-order plane (gemm_tc.cu lo8; 64-channel 3x3 layers: c64)
    if (h->precision == FADB_PREC_FP16X2 && h->lo_fp8 && (Cin % 128 == 0 || (Cin == 64 && ksize == 3 && h->lo_fp8_c64))) {
        FADB_CUDA_CHECK(cudaMalloc(&L.w8, n));
        FADB_CHECK(pack_conv_weight_lo8(h, w, Cout, Cin, ksize, scale, L.w8, &L.lo_scale, st));
    }
    if (bias_or_shift)
        FADB_CUDA_CHECK(cudaMemcpyAsync(L.bias, bias_or_shift, (size_t)Cout * sizeof(float), cudaMemcpyDeviceToDevice, st));
    else
        FADB_CUDA_CHECK(cudaMemsetAsync(L.bias, 0, (size_t)Cout * sizeof(float), st));
    h->layers.push_back(L);
    return FADB_OK;
}

static int commit_vggish(fadb_handle* h, cudaStream_t st) {
    static const int slots[6] = {0, 3, 6, 8, 11, 13};
    static const int chans[7] = {1, 64, 128, 256, 256, 512, 512};
    char name[64];
    const float *w, *b;
    snprintf(name, sizeof(name), "features.0.weight");
    FADB_CHECK(get_staged(h, name, {64, 1, 3, 3}, &w));
    FADB_CHECK(get_staged(h, "features.0.bias", {64}, &b));
    FADB_CUDA_CHECK(cudaMalloc(&h->conv1_w, 9 * 64 * sizeof(float)));
    FADB_CUDA_CHECK(cudaMalloc(&h->conv1_b, 64 * sizeof(float)));
    pack_conv1_kernel<<<3, 256, 0, st>>>(w, nullptr, h->conv1_w);
    h->launches++;
    FADB_CUDA_CHECK(cudaMemcpyAsync(h->conv1_b, b, 64 * sizeof(float), cudaMemcpyDeviceToDevice, st));
    for (int i = 1; i < 6; ++i) {
        snprintf(name, sizeof(name), "features.%d.weight", slots[i]);
        FADB_CHECK(get_staged(h, name, {chans[i + 1], chans[i], 3, 3}, &w));
        snprintf(name, sizeof(name), "features.%d.bias", slots[i]);
        FADB_CHECK(get_staged(h, name, {chans[i + 1]}, &b));
        FADB_CHECK(add_layer(h, w, chans[i + 1], chans[i], 3, nullptr, b, st));
    }
    static const int fslot[3] = {0, 2, 4};
    static const int fin[3] = {12288, 4096, 4096}, fout[3] = {4096, 4096, 128};
    for (int i = 0; i < 3; ++i) {
        snprintf(name, sizeof(name), "embeddings.%d.weight", fslot[i]);
        FADB_CHECK(get_staged(h, name, {fout[i], fin[i]}, &w));
        snprintf(name, sizeof(name), "embeddings.%d.bias", fslot[i]);
        FADB_CHECK(get_staged(h, name, {fout[i]}, &b));
        FADB_CHECK(add_layer(h, w, fout[i], fin[i], 1, nullptr, b, st));
    }
    return FADB_OK;
}

static int folded_bn(fadb_handle* h, const std::string& prefix, int C, float** scale, float** shift, cudaStream_t st) {
    const float *g, *b, *m, *v;
    FADB_CHECK(get_staged(h, prefix + ".weight", {C}, &g));
    FADB_CHECK(get_staged(h, prefix + ".bias", {C}, &b));
    FADB_CHECK(get_staged(h, prefix + ".running_mean", {C}, &m));
    FADB_CHECK(get_staged(h, prefix + ".running_var", {C}, &v));
    FADB_CUDA_CHECK(cudaMalloc(scale, C * sizeof(float)));
    FADB_CUDA_CHECK(cudaMalloc(shift, C * sizeof(float)));
    return fold_bn(h, g, b, m, v, C, *scale, *shift, st);
}

static int commit_cnn14(fadb_handle* h, bool clap, cudaStream_t st) {
    static const int ch[7] = {1, 64, 128, 256, 512, 1024, 2048};
    FADB_CHECK(folded_bn(h, "bn0", 64, &h->bn0_scale, &h->bn0_shift, st));
    std::vector<float*> tmp;
    int rc = FADB_OK;
    for (int blk = 1; blk <= 6 && rc == FADB_OK; ++blk) {
        for (int cv = 1; cv <= 2 && rc == FADB_OK; ++cv) {
            const int cin = (cv == 1) ? ch[blk - 1] : ch[blk], cout = ch[blk];
            char wname[64], bname[64];
            snprintf(wname, sizeof(wname), "conv_block%d.conv%d.weight", blk, cv);
            snprintf(bname, sizeof(bname), "conv_block%d.bn%d", blk, cv);
            const float* w;
            rc = get_staged(h, wname, {cout, cin, 3, 3}, &w);
            if (rc != FADB_OK) break;
            float *scale = nullptr, *shift = nullptr;
            rc = folded_bn(h, bname, cout, &scale, &shift, st);
            tmp.push_back(scale);
            tmp.push_back(shift);
            if (rc != FADB_OK) break;
            if (blk == 1 && cv == 1) {
                if (cudaMalloc(&h->conv1_w, 9 * 64 * sizeof(float)) != cudaSuccess ||
                    cudaMalloc(&h->conv1_b, 64 * sizeof(float)) != cudaSuccess) {
                    set_error("cudaMalloc conv1 failed");
                    rc = FADB_E_NOMEM;
                    break;
                }
                pack_conv1_kernel<<<3, 256, 0, st>>>(w, scale, h->conv1_w);
                h->launches++;
                cudaMemcpyAsync(h->conv1_b, shift, 64 * sizeof(float), cudaMemcpyDeviceToDevice, st);
            } else {
                rc = add_layer(h, w, cout, cin, 3, scale, shift, st);
            }
        }
    }
    if (rc == FADB_OK) {
        const float *w, *b;
        rc = get_staged(h, "fc1.weight", {2048, 2048}, &w);
        if (rc == FADB_OK) rc = get_staged(h, "fc1.bias", {2048}, &b);
        if (rc == FADB_OK) rc = add_layer(h, w, 2048, 2048, 1, nullptr, b, st);
        if (rc == FADB_OK && clap) {
            rc = get_staged(h, "clap_head.0.weight", {512, 2048}, &w);
            if (rc == FADB_OK) rc = get_staged(h, "clap_head.0.bias", {512}, &b);
            if (rc == FADB_OK) rc = add_layer(h, w, 512, 2048, 1, nullptr, b, st);
            if (rc == FADB_OK) rc = get_staged(h, "clap_head.2.weight", {512, 512}, &w);
            if (rc == FADB_OK) rc = get_staged(h, "clap_head.2.bias", {512}, &b);
            if (rc == FADB_OK) rc = add_layer(h, w, 512, 512, 1, nullptr, b, st);
        }
    }
    cudaStreamSynchronize(st);
    for (float* p : tmp)
        if (p) cudaFree(p);
    return rc;
}

// ---------------------------------------------------------------- layer schedules
static inline __nv_bfloat16* lo_plane(fadb_handle* h, int buf, size_t plane_elems) {
    return h->ws_act[buf].as<__nv_bfloat16>() + plane_elems;
}

constexpr size_t kVggishActElems = 98304;        // max activation elements per patch (48*32*64)

// tensor-core part of VGGishCore: a1 = conv1 output [P,48,32,64] (hi/lo) -> emb [P,128]
static int vggish_tc_layers(fadb_handle* h, const __nv_bfloat16* a1_hi, const __nv_bfloat16* a1_lo, const uint8_t* a1_8p,
                            int64_t P, float* emb, cudaStream_t st) {
    // sized for THIS batch (the buffers only ever grow); the lo plane exists in the bf16x3 mode only
    const size_t plane = (size_t)P * kVggishActElems;                       // largest later activation: 24*16*256
    const size_t nplanes = h->precision == FADB_PREC_BF16X3 ? 2 : 1;
    FADB_CHECK(h->ws_act[0].reserve(plane * nplanes * sizeof(__nv_bfloat16)));
    FADB_CHECK(h->ws_act[1].reserve(plane * nplanes * sizeof(__nv_bfloat16)));
    __nv_bfloat16* a[2] = {h->ws_act[0].as<__nv_bfloat16>(), h->ws_act[1].as<__nv_bfloat16>()};
    __nv_bfloat16* l[2] = {lo_plane(h, 0, plane), lo_plane(h, 1, plane)};
    uint8_t* a8[2] = {nullptr, nullptr};                                    // e4m3 copies (fp16x2 with the e4m3 lo pass)
    if (h->precision == FADB_PREC_FP16X2 && h->lo_fp8) {
        FADB_CHECK(h->ws_act8[0].reserve(plane));
        FADB_CHECK(h->ws_act8[1].reserve(plane));
        a8[0] = h->ws_act8[0].as<uint8_t>(); a8[1] = h->ws_act8[1].as<uint8_t>();
    }
    const uint8_t* in8 = a1_8p;                                             // conv1's 64-channel output: W-padded e4m3 copy, or none
    const int B = (int)P;
    struct Step { int H, W, Cin, pool; };
    static const Step steps[5] = {{48, 32, 64, 1}, {24, 16, 128, 0}, {24, 16, 256, 1}, {12, 8, 256, 0}, {12, 8, 512, 1}};
    const __nv_bfloat16* in_hi = a1_hi;
    const __nv_bfloat16* in_lo = a1_lo;
    int cur = 0;
    for (int i = 0; i < 5; ++i) {
        LayerIO io;
        io.in_hi = in_hi; io.in_lo = in_lo;
        io.B = B; io.H = steps[i].H; io.W = steps[i].W; io.Cin = steps[i].Cin;
        io.taps = 9; io.relu = 1; io.pool = steps[i].pool;
        io.out_hi = a[cur]; io.out_lo = l[cur];
        io.in8 = in8; io.out8 = a8[cur];
        io.in8_wpad = (i == 0);
        io.use_lo_weights = (h->x2_mask >> i) & 1u;
        FADB_CHECK(launch_gemm_layer(h, h->layers[i], io, st));
        in_hi = a[cur]; in_lo = l[cur]; in8 = a8[cur];
        cur ^= 1;
    }
    // NHWC flatten (vggish.py:91-94) is the memory order already: [P, 6*4*512]
    static const int fin[3] = {12288, 4096, 4096};
    for (int i = 0; i < 3; ++i) {
        LayerIO io;
        io.in_hi = in_hi; io.in_lo = in_lo;
        io.B = 1; io.H = 1; io.W = B; io.Cin = fin[i];
        io.taps = 1; io.relu = (i < 2); io.pool = 0;
        if (i < 2) { io.out_hi = a[cur]; io.out_lo = l[cur]; }
        else io.out_f32 = emb;                                                          // no final ReLU, vggish.py:76-77
        io.in8 = in8; io.out8 = a8[cur];
        io.use_lo_weights = (h->x2_mask >> (5 + i)) & 1u;
        FADB_CHECK(launch_gemm_layer(h, h->layers[5 + i], io, st));
        in_hi = a[cur]; in_lo = l[cur]; in8 = a8[cur];
        cur ^= 1;
    }
    return FADB_OK;
}

static int reserve_a1(fadb_handle* h, int buf, int64_t P) {
    const size_t nplanes = h->precision == FADB_PREC_BF16X3 ? 2 : 1;
    return h->ws_a1[buf].reserve((size_t)P * kVggishActElems * nplanes * sizeof(__nv_bfloat16));
}
static inline __nv_bfloat16* a1_lo_plane(fadb_handle* h, int buf, int64_t P) {
    return h->ws_a1[buf].as<__nv_bfloat16>() + (size_t)P * kVggishActElems;
}

// W-padded e4m3 copy of conv1's output [P][48][34][64], when the first tensor-core layer's low-order pass wants it
static int reserve_a1_8p(fadb_handle* h, int64_t P, uint8_t** out) {
    *out = nullptr;
    if (!(h->precision == FADB_PREC_FP16X2 && h->lo_fp8 && h->lo_fp8_c64 && !h->layers.empty() && h->layers[0].w8)) return FADB_OK;
    FADB_CHECK(h->ws_a1_8.reserve((size_t)P * 48 * 34 * 64));
    *out = h->ws_a1_8.as<uint8_t>();
    return FADB_OK;
}

static int vggish_forward(fadb_handle* h, const float* feats, int64_t P, float* emb, cudaStream_t st) {
    FADB_CHECK(reserve_a1(h, 0, P));
    __nv_bfloat16* a1 = h->ws_a1[0].as<__nv_bfloat16>();
    uint8_t* a1_8p = nullptr;
    FADB_CHECK(reserve_a1_8p(h, P, &a1_8p));
    FADB_CHECK(launch_conv1_vggish(h, feats, P, a1, a1_lo_plane(h, 0, P), st));        // [P,48,32,64]
    if (a1_8p) FADB_CHECK(quantize_e4m3_wpad(h, nullptr, a1, P * 48, 32, a1_8p, st));
    return vggish_tc_layers(h, a1, a1_lo_plane(h, 0, P), a1_8p, P, emb, st);
}

// PCM -> embeddings for VGGish, chunked.  (A side stream running front end + conv1 of chunk i+1 under the tcgen05
// layers of chunk i was measured and gave no gain at any smem budget / stream priority; removed.)
static int vggish_embed_pcm(fadb_handle* h, PcmSrc pcm, int64_t n_clips, int64_t n_samples, int64_t pcm_stride,
                            int64_t rows, float* emb, cudaStream_t st) {
    const int d = 128;
    int64_t cpc = h->max_batch / rows;                  // clips per chunk
    FADB_REQUIRE(cpc >= 1, "max_batch %d smaller than patches per clip %lld", h->max_batch, (long long)rows);
    for (int64_t c0 = 0; c0 < n_clips; c0 += cpc) {
        const int64_t nc = (n_clips - c0 < cpc) ? n_clips - c0 : cpc;
        if (h->fused_front) {
            // front end + conv1 in one kernel: the fp32 features never touch HBM
            // (chunks only shrink at the tail, so the lo-plane offset of the first chunk is the largest)
            const int64_t Pc = nc * rows;
            FADB_CHECK(reserve_a1(h, 0, Pc));
            __nv_bfloat16* a1 = h->ws_a1[0].as<__nv_bfloat16>();
            uint8_t* a1_8p = nullptr;
            FADB_CHECK(reserve_a1_8p(h, Pc, &a1_8p));
            FADB_CHECK(launch_vggish_front_conv1(h, pcm.offset(c0 * pcm_stride), nc, n_samples, pcm_stride, a1,
                                                 a1_lo_plane(h, 0, Pc), a1_8p, st));
            FADB_CHECK(vggish_tc_layers(h, a1, a1_lo_plane(h, 0, Pc), a1_8p, Pc, emb + c0 * rows * d, st));
        } else {
            FADB_CHECK(h->ws_feats.reserve((size_t)(n_clips < cpc ? n_clips : cpc) * rows * 96 * 64 * sizeof(float)));
            FADB_CHECK(launch_frontend(h, FADB_MODEL_VGGISH, pcm.offset(c0 * pcm_stride), nc, n_samples, pcm_stride,
                                       h->ws_feats.as<float>(), st));
            FADB_CHECK(vggish_forward(h, h->ws_feats.as<float>(), nc * rows, emb + c0 * rows * d, st));
        }
    }
    return FADB_OK;
}

static int cnn14_forward(fadb_handle* h, const float* feats, int64_t B64, int T, float* emb, cudaStream_t st) {
    const int B = (int)B64;
    const size_t per_clip = (size_t)T * 64 * 64;
    const size_t plane = (size_t)B * per_clip;                                // this batch; the buffers only ever grow
    const size_t nplanes = h->precision == FADB_PREC_BF16X3 ? 2 : 1;
    FADB_CHECK(h->ws_act[0].reserve(plane * nplanes * sizeof(__nv_bfloat16)));
    FADB_CHECK(h->ws_act[1].reserve(plane * nplanes * sizeof(__nv_bfloat16)));
    __nv_bfloat16* a[2] = {h->ws_act[0].as<__nv_bfloat16>(), h->ws_act[1].as<__nv_bfloat16>()};
    __nv_bfloat16* l[2] = {lo_plane(h, 0, plane), lo_plane(h, 1, plane)};
    uint8_t* a8[2] = {nullptr, nullptr};                                    // e4m3 copies (fp16x2 with the e4m3 lo pass)
    if (h->precision == FADB_PREC_FP16X2 && h->lo_fp8) {
        // the largest map with an unpadded copy is block 2's [B, T/2, 32, 128]; the 64-channel maps of block 1 get
        // W-padded copies (c64): conv1's [B, T, 66, 64] in a8[0], conv1_2's pooled [B, T/2, 34, 64] in a8[1]
        const bool c64 = h->lo_fp8_c64 && !h->layers.empty() && h->layers[0].w8;
        FADB_CHECK(h->ws_act8[0].reserve(c64 ? plane / 64 * 66 : plane / 2));
        FADB_CHECK(h->ws_act8[1].reserve(plane / 2));
        a8[0] = h->ws_act8[0].as<uint8_t>(); a8[1] = h->ws_act8[1].as<uint8_t>();
    }
    const bool c64 = a8[0] && h->lo_fp8_c64 && !h->layers.empty() && h->layers[0].w8;
    bool have8 = c64;                                                       // does a8[cur] hold the current activations?
    FADB_CHECK(launch_conv1_cnn14(h, feats, B, T, a[0], l[0], c64 ? a8[0] : nullptr, st));   // [B,T,64,64]
    int cur = 0, H = T, W = 64, C = 64, li = 0;
    for (int blk = 1; blk <= 6; ++blk) {
        for (int cv = 1; cv <= 2; ++cv) {
            if (blk == 1 && cv == 1) continue;
            LayerIO io;
            io.in_hi = a[cur]; io.in_lo = l[cur];
            io.B = B; io.H = H; io.W = W; io.Cin = C;
            io.taps = 9; io.relu = 1;
            io.pool = (cv == 2 && blk < 6) ? 2 : 0;                                     // avg_pool2d, pann.py:192,255-260
            io.out_hi = a[cur ^ 1]; io.out_lo = l[cur ^ 1];
            io.in8 = have8 ? a8[cur] : nullptr;
            io.in8_wpad = (have8 && C == 64) ? 1 : 0;                                   // block 1's maps and conv2_1's input
            io.out8_wpad = (c64 && h->layers[li].N == 64) ? 1 : 0;                      // conv1_2 -> conv2_1
            io.out8 = (h->layers[li].N % 128 == 0 || io.out8_wpad) ? a8[cur ^ 1] : nullptr;
            io.use_lo_weights = (h->x2_mask >> li) & 1u;
            FADB_CHECK(launch_gemm_layer(h, h->layers[li], io, st));
            have8 = io.out8 != nullptr;
            C = h->layers[li].N;
            ++li;
            cur ^= 1;
            if (io.pool) { H /= 2; W /= 2; }
        }
    }
    // global pooling -> [B, 2048]
    FADB_CHECK(launch_cnn14_global_pool(h, a[cur], l[cur], B, H, W, C, a[cur ^ 1], l[cur ^ 1], a8[cur ^ 1], st));
    cur ^= 1;
    const bool clap = (h->model == FADB_MODEL_CLAP);
    {
        LayerIO io;
        io.in_hi = a[cur]; io.in_lo = l[cur];
        io.B = 1; io.H = 1; io.W = B; io.Cin = 2048; io.taps = 1; io.relu = 1; io.pool = 0;     // pann.py:271
        io.in8 = a8[cur]; io.out8 = a8[cur ^ 1];
        if (clap) { io.out_hi = a[cur ^ 1]; io.out_lo = l[cur ^ 1]; }
        else io.out_f32 = emb;
        FADB_CHECK(launch_gemm_layer(h, h->layers[li++], io, st));
        cur ^= 1;
    }
    if (clap) {
        LayerIO io;
        io.in_hi = a[cur]; io.in_lo = l[cur];
        io.B = 1; io.H = 1; io.W = B; io.Cin = 2048; io.taps = 1; io.relu = 1; io.pool = 0;
        io.out_hi = a[cur ^ 1]; io.out_lo = l[cur ^ 1];
        io.in8 = a8[cur]; io.out8 = a8[cur ^ 1];
        FADB_CHECK(launch_gemm_layer(h, h->layers[li++], io, st));
        cur ^= 1;
        LayerIO io2;
        io2.in_hi = a[cur]; io2.in_lo = l[cur];
        io2.B = 1; io2.H = 1; io2.W = B; io2.Cin = 512; io2.taps = 1; io2.relu = 0; io2.pool = 0;
        io2.in8 = a8[cur];
        io2.out_f32 = emb;
        FADB_CHECK(launch_gemm_layer(h, h->layers[li++], io2, st));
        FADB_CHECK(launch_l2_normalize(h, emb, B, 512, st));
    }
    return FADB_OK;
}

static int check_device_flag(fadb_handle* h) {
    if (h->err_flag_host && *h->err_flag_host != 0) {
        set_error("device-side failure flag = %d (1 = tensor-core pipeline timeout)", *h->err_flag_host);
        return FADB_E_DEVICE;
    }
    return FADB_OK;
}

}  // namespace fadb

using namespace fadb;

// ================================================================================================
// extern "C"
// ================================================================================================
extern "C" {

int fadb_abi_version(void) { return FADB_ABI_VERSION; }
const char* fadb_last_error(void) { return g_err; }

int fadb_create(fadb_handle** out, int device) {
    if (!out) { set_error("fadb_create: out is NULL"); return FADB_E_INVALID; }
    *out = nullptr;
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0) {
        set_error("no CUDA device available (%s): libfadb200 has no CPU fallback", cudaGetErrorString(e));
        return FADB_E_CUDA;
    }
    if (device < 0 || device >= count) { set_error("device %d out of range (%d devices)", device, count); return FADB_E_INVALID; }
    FADB_CUDA_CHECK(cudaSetDevice(device));
    cudaDeviceProp prop;
    FADB_CUDA_CHECK(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10) {
        set_error("device %d is sm_%d%d; libfadb200 is built for sm_100a (B200) only", device, prop.major, prop.minor);
        return FADB_E_CUDA;
    }
    fadb_handle* h = new fadb_handle();
    h->device = device;
    h->sm_count = prop.multiProcessorCount;
    FADB_CUDA_CHECK(cudaHostAlloc(&h->err_flag_host, sizeof(int), cudaHostAllocMapped));
    *h->err_flag_host = 0;
    FADB_CUDA_CHECK(cudaHostGetDevicePointer(&h->err_flag, h->err_flag_host, 0));
    FADB_CUDA_CHECK(cudaStreamCreateWithFlags(&h->copy_stream, cudaStreamNonBlocking));
    for (int i = 0; i < 2; ++i) {
        FADB_CUDA_CHECK(cudaEventCreateWithFlags(&h->ev_copy[i], cudaEventDisableTiming));
        FADB_CUDA_CHECK(cudaEventCreateWithFlags(&h->ev_compute[i], cudaEventDisableTiming));
    }
    if (const char* e = getenv("FADB_GEMM_SMEM")) h->gemm_smem_budget = atoi(e);
    if (const char* e = getenv("FADB_RESIDENT_B")) h->resident_b = atoi(e);
    if (const char* e = getenv("FADB_CLUSTER")) h->gemm_cluster = atoi(e);
    if (const char* e = getenv("FADB_CLUSTER_SIZE")) h->gemm_cluster_size = atoi(e);
    if (const char* e = getenv("FADB_TWOCTA")) h->gemm_twocta = atoi(e);
    if (const char* e = getenv("FADB_PAIR_HALO")) h->gemm_pair_halo = atoi(e);
    if (const char* e = getenv("FADB_FUSED_FRONT")) h->fused_front = atoi(e);
    if (const char* e = getenv("FADB_HALO")) h->halo = atoi(e);
    if (const char* e = getenv("FADB_TC_SYRK")) h->tc_syrk = atoi(e);
    if (const char* e = getenv("FADB_LO_FP8")) h->lo_fp8 = atoi(e);
    if (const char* e = getenv("FADB_LO_FP8_C64")) h->lo_fp8_c64 = atoi(e);
    if (const char* e = getenv("FADB_X2_MASK")) h->x2_mask = (unsigned)strtoul(e, nullptr, 0);
    int rc = gemm_init(h);
    if (rc == FADB_OK) rc = frontend_init(h);
    if (rc == FADB_OK) rc = frechet_init(h);
    if (rc != FADB_OK) { fadb_destroy(h); return rc; }
    *out = h;
    return FADB_OK;
}

void fadb_destroy(fadb_handle* h) {
    if (!h) return;
    cudaSetDevice(h->device);
    free_layers(h);
    free_staged(h);
    frontend_release(h);
    h->weight_pool.release();
    h->ws_feats.release(); h->ws_act[0].release(); h->ws_act[1].release(); h->ws_misc.release();
    h->ws_act8[0].release(); h->ws_act8[1].release();
    h->ws_syrk.release();
    h->ws_frechet.release(); h->ws_stats.release(); h->ws_pcm[0].release(); h->ws_pcm[1].release(); h->ws_emb.release();
    for (cudaEvent_t e : h->prof_events) cudaEventDestroy(e);
    h->ws_a1[0].release();
    h->ws_a1_8.release();
    if (h->copy_stream) cudaStreamDestroy(h->copy_stream);
    for (int i = 0; i < 2; ++i) {
        if (h->ev_copy[i]) cudaEventDestroy(h->ev_copy[i]);
        if (h->ev_compute[i]) cudaEventDestroy(h->ev_compute[i]);
    }
    if (h->err_flag_host) cudaFreeHost(h->err_flag_host);
    delete h;
}

int fadb_set_precision(fadb_handle* h, int prec) {
    if (!h || prec < FADB_PREC_BF16 || prec > FADB_PREC_FP16X2) { set_error("bad precision"); return FADB_E_INVALID; }
    h->precision = prec;
    return FADB_OK;
}

int fadb_set_tensor_syrk(fadb_handle* h, int on) {
    if (!h) { set_error("fadb_set_tensor_syrk: NULL handle"); return FADB_E_INVALID; }
    h->tc_syrk = on ? 1 : 0;
    return FADB_OK;
}

int fadb_set_clap_quantize(fadb_handle* h, int on) {
    if (!h) { set_error("fadb_set_clap_quantize: NULL handle"); return FADB_E_INVALID; }
    h->clap_quantize = on ? 1 : 0;
    return FADB_OK;
}

int fadb_set_max_batch(fadb_handle* h, int max_items) {
    if (!h || max_items < 1 || max_items > 65535) { set_error("max_batch must be in [1, 65535]"); return FADB_E_INVALID; }
    h->max_batch = max_items;
    // CNN14 clips per batch: the call can lower it below its default (128 clips = 2.2 GB of activations), never raise it
    h->max_batch_cnn14 = max_items < 128 ? max_items : 128;
    return FADB_OK;
}

int fadb_weights_begin(fadb_handle* h, int model) {
    if (!h || embed_dim(model) < 0) { set_error("unknown model %d", model); return FADB_E_INVALID; }
    cudaSetDevice(h->device);
    free_staged(h);
    h->staged_model = model;
    return FADB_OK;
}

int fadb_weights_tensor(fadb_handle* h, const char* name, const float* data_host, const int64_t* shape, int ndim) {
    if (!h || !name || !data_host || ndim < 0 || ndim > 4) { set_error("bad weights_tensor arguments"); return FADB_E_INVALID; }
    if (h->staged_model < 0) { set_error("fadb_weights_tensor before fadb_weights_begin"); return FADB_E_STATE; }
    cudaSetDevice(h->device);
    HostTensor t;
    t.numel = 1;
    for (int i = 0; i < ndim; ++i) { t.shape.push_back(shape[i]); t.numel *= (size_t)shape[i]; }
    FADB_CUDA_CHECK(cudaMalloc(&t.dev, (t.numel ? t.numel : 1) * sizeof(float)));
    FADB_CUDA_CHECK(cudaMemcpy(t.dev, data_host, t.numel * sizeof(float), cudaMemcpyHostToDevice));
    auto it = h->staged.find(name);
    if (it != h->staged.end() && it->second.dev) cudaFree(it->second.dev);
    h->staged[name] = t;
    return FADB_OK;
}

int fadb_weights_commit(fadb_handle* h) {
    if (!h || h->staged_model < 0) { set_error("fadb_weights_commit before fadb_weights_begin"); return FADB_E_STATE; }
    cudaSetDevice(h->device);
    free_layers(h);
    cudaStream_t st = 0;
    int rc = (h->staged_model == FADB_MODEL_VGGISH) ? commit_vggish(h, st)
                                                    : commit_cnn14(h, h->staged_model == FADB_MODEL_CLAP, st);
    cudaError_t e = cudaStreamSynchronize(st);
    free_staged(h);
    if (rc != FADB_OK) { free_layers(h); return rc; }
    if (e != cudaSuccess) { set_error("weight packing failed: %s", cudaGetErrorString(e)); free_layers(h); return FADB_E_CUDA; }
    h->model = h->staged_model;
    h->staged_model = -1;
    h->weights_ready = true;
    return FADB_OK;
}

int64_t fadb_frontend_rows(int model, int64_t n_samples) { return frontend_rows(model, n_samples); }

int fadb_frontend(fadb_handle* h, int model, const float* pcm_dev, int64_t n_clips, int64_t n_samples,
                  int64_t pcm_stride, float* feats_dev, void* stream) {
    if (!h || !pcm_dev || !feats_dev) { set_error("fadb_frontend: NULL argument"); return FADB_E_INVALID; }
    cudaSetDevice(h->device);
    const int64_t kMax = 32768;
    const int64_t rows = frontend_rows(model, n_samples);
    if (rows < 0) { set_error("unknown model %d", model); return FADB_E_INVALID; }
    const int64_t row_elems = (model == FADB_MODEL_VGGISH) ? 96 * 64 : 64;
    for (int64_t c0 = 0; c0 < n_clips; c0 += kMax) {
        const int64_t nc = (n_clips - c0 < kMax) ? n_clips - c0 : kMax;
        FADB_CHECK(launch_frontend(h, model, PcmSrc{pcm_dev + c0 * pcm_stride, 0}, nc, n_samples, pcm_stride,
                                   feats_dev + c0 * rows * row_elems, (cudaStream_t)stream));
    }
    return FADB_OK;
}

int fadb_resample(fadb_handle* h, const float* in_dev, int64_t n_clips, int64_t n_in, int64_t in_stride, double ratio,
                  const double* win_dev, int nwin, int num_table, float* out_dev, int64_t n_out, int64_t out_stride,
                  void* stream) {
    if (!h || !in_dev || !win_dev || !out_dev) { set_error("fadb_resample: NULL argument"); return FADB_E_INVALID; }
    if (n_in <= 0 || n_out != (int64_t)((double)n_in * ratio)) {
        set_error("fadb_resample: n_out must be int(n_in * ratio) (resampy semantics), got %lld for %lld x %g",
                  (long long)n_out, (long long)n_in, ratio);
        return FADB_E_INVALID;
    }
    cudaSetDevice(h->device);
    return launch_resample(h, in_dev, n_clips, n_in, in_stride, ratio, win_dev, nwin, num_table, out_dev, n_out, out_stride,
                           (cudaStream_t)stream);
}

int fadb_embed_dim(int model) { return embed_dim(model); }

int fadb_embed(fadb_handle* h, const float* feats_dev, int64_t n_items, int64_t t_frames, float* emb_dev, void* stream) {
    if (!h || !feats_dev || !emb_dev) { set_error("fadb_embed: NULL argument"); return FADB_E_INVALID; }
    if (!h->weights_ready) { set_error("fadb_embed: weights not committed"); return FADB_E_STATE; }
    cudaSetDevice(h->device);
    cudaStream_t st = (cudaStream_t)stream;
    const int d = embed_dim(h->model);
    if (h->model == FADB_MODEL_VGGISH) {
        if (t_frames != 96) { set_error("VGGish patches have 96 frames, got %lld", (long long)t_frames); return FADB_E_INVALID; }
        const int64_t batch = h->max_batch;
        for (int64_t i0 = 0; i0 < n_items; i0 += batch) {
            const int64_t n = (n_items - i0 < batch) ? n_items - i0 : batch;
            FADB_CHECK(vggish_forward(h, feats_dev + i0 * 96 * 64, n, emb_dev + i0 * d, st));
        }
    } else {
        if (t_frames < 32 || t_frames > 16384) { set_error("CNN14 needs 32 <= T <= 16384 frames, got %lld", (long long)t_frames); return FADB_E_INVALID; }
        for (int64_t i0 = 0; i0 < n_items; i0 += h->max_batch_cnn14) {
            const int64_t n = (n_items - i0 < h->max_batch_cnn14) ? n_items - i0 : h->max_batch_cnn14;
            FADB_CHECK(cnn14_forward(h, feats_dev + i0 * t_frames * 64, n, (int)t_frames, emb_dev + i0 * d, st));
        }
    }
    return check_device_flag(h);
}

static int embed_pcm_any(fadb_handle* h, PcmSrc pcm, int64_t n_clips, int64_t n_samples, int64_t pcm_stride,
                         float* emb_dev, void* stream, const char* who) {
    if (!h || !pcm.ptr || !emb_dev) { set_error("%s: NULL argument", who); return FADB_E_INVALID; }
    if (!h->weights_ready) { set_error("%s: weights not committed", who); return FADB_E_STATE; }
    cudaSetDevice(h->device);
    cudaStream_t st = (cudaStream_t)stream;
    const int model = h->model;
    const int64_t rows = frontend_rows(model, n_samples);
    if (rows <= 0) return FADB_OK;       // clips too short: zero rows, like waveform_to_examples -> [0,1,96,64]
    const int d = embed_dim(model);
    if (model == FADB_MODEL_VGGISH) {
        FADB_CHECK(vggish_embed_pcm(h, pcm, n_clips, n_samples, pcm_stride, rows, emb_dev, st));
    } else {
        const int64_t cpc = h->max_batch_cnn14;
        FADB_CHECK(h->ws_feats.reserve((size_t)(n_clips < cpc ? n_clips : cpc) * rows * 64 * sizeof(float)));
        for (int64_t c0 = 0; c0 < n_clips; c0 += cpc) {
            const int64_t nc = (n_clips - c0 < cpc) ? n_clips - c0 : cpc;
            FADB_CHECK(launch_frontend(h, model, pcm.offset(c0 * pcm_stride), nc, n_samples, pcm_stride,
                                       h->ws_feats.as<float>(), st));
            FADB_CHECK(cnn14_forward(h, h->ws_feats.as<float>(), nc, (int)rows, emb_dev + c0 * d, st));
        }
    }
    return check_device_flag(h);
}

int fadb_embed_pcm(fadb_handle* h, const float* pcm_dev, int64_t n_clips, int64_t n_samples, int64_t pcm_stride,
                   float* emb_dev, void* stream) {
    return embed_pcm_any(h, PcmSrc{pcm_dev, 0}, n_clips, n_samples, pcm_stride, emb_dev, stream, "fadb_embed_pcm");
}

int fadb_embed_pcm16(fadb_handle* h, const int16_t* pcm_dev, int64_t n_clips, int64_t n_samples, int64_t pcm_stride,
                     float* emb_dev, void* stream) {
    return embed_pcm_any(h, PcmSrc{pcm_dev, 1}, n_clips, n_samples, pcm_stride, emb_dev, stream, "fadb_embed_pcm16");
}

int fadb_stats_accumulate(fadb_handle* h, const float* emb_dev, int64_t n_rows, int d, int64_t row_stride,
                          const double* shift_dev, double* acc_dev, void* stream) {
    if (!h || !emb_dev || !acc_dev) { set_error("fadb_stats_accumulate: NULL argument"); return FADB_E_INVALID; }
    cudaSetDevice(h->device);
    return launch_stats_accumulate(h, emb_dev, n_rows, d, row_stride, shift_dev, acc_dev, (cudaStream_t)stream);
}

int fadb_stats_accumulate_f64(fadb_handle* h, const double* emb_dev, int64_t n_rows, int d, int64_t row_stride,
                              const double* shift_dev, double* acc_dev, void* stream) {
    if (!h || !emb_dev || !acc_dev) { set_error("fadb_stats_accumulate_f64: NULL argument"); return FADB_E_INVALID; }
    cudaSetDevice(h->device);
    return launch_stats_accumulate_f64(h, emb_dev, n_rows, d, row_stride, shift_dev, acc_dev, (cudaStream_t)stream);
}

int fadb_stats_finalize(fadb_handle* h, const double* acc_dev, int d, const double* shift_dev, double* mu_dev,
                        double* sigma_dev, void* stream) {
    if (!h || !acc_dev || !sigma_dev) { set_error("fadb_stats_finalize: NULL argument"); return FADB_E_INVALID; }
    cudaSetDevice(h->device);
    return launch_stats_finalize(h, acc_dev, d, shift_dev, mu_dev, sigma_dev, (cudaStream_t)stream);
}

int fadb_frechet(fadb_handle* h, const double* mu1_dev, const double* sigma1_dev, const double* mu2_dev,
                 const double* sigma2_dev, int d, double* out_dev, void* stream) {
    if (!h || !mu1_dev || !sigma1_dev || !mu2_dev || !sigma2_dev || !out_dev) { set_error("fadb_frechet: NULL argument"); return FADB_E_INVALID; }
    cudaSetDevice(h->device);
    return launch_frechet(h, mu1_dev, sigma1_dev, mu2_dev, sigma2_dev, d, out_dev, (cudaStream_t)stream);
}

static int fad_from_pcm_host_any(fadb_handle* h, const void* pcm_bg_host, int64_t n_bg, const void* pcm_ev_host,
                                 int64_t n_ev, int64_t n_samples, int i16, float* emb_bg_host, float* emb_ev_host,
                                 double* fad_out) {
    if (!h || !pcm_bg_host || !pcm_ev_host || !fad_out) { set_error("fadb_fad_from_pcm_host: NULL argument"); return FADB_E_INVALID; }
    const size_t esz = i16 ? sizeof(int16_t) : sizeof(float);
    if (!h->weights_ready) { set_error("weights not committed"); return FADB_E_STATE; }
    cudaSetDevice(h->device);
    const int model = h->model;
    const int d = embed_dim(model);
    const int64_t rows = (model == FADB_MODEL_VGGISH) ? frontend_rows(model, n_samples) : 1;
    if (rows <= 0 || n_bg <= 0 || n_ev <= 0) { set_error("empty embedding set (fad.py:640-645)"); return FADB_E_INVALID; }
    int64_t cpc = (model == FADB_MODEL_VGGISH) ? h->max_batch / rows : h->max_batch_cnn14;
    if (cpc < 1) cpc = 1;
    const size_t acc_n = 1 + (size_t)d + (size_t)d * d;
    // stats workspace: acc[2] | mu[2] | sigma[2] | out[4]
    FADB_CHECK(h->ws_stats.reserve((2 * acc_n + 2 * (size_t)d + 2 * (size_t)d * d + 8) * sizeof(double)));
    double* acc[2] = {h->ws_stats.as<double>(), h->ws_stats.as<double>() + acc_n};
    double* mu[2] = {acc[1] + acc_n, acc[1] + acc_n + d};
    double* sg[2] = {mu[1] + d, mu[1] + d + (size_t)d * d};
    double* outd = sg[1] + (size_t)d * d;
    FADB_CHECK(h->ws_pcm[0].reserve((size_t)cpc * n_samples * esz));
    FADB_CHECK(h->ws_pcm[1].reserve((size_t)cpc * n_samples * esz));
    FADB_CHECK(h->ws_emb.reserve((size_t)cpc * rows * d * sizeof(float) * 2));
    cudaStream_t cs = h->copy_stream;
    cudaStream_t st = 0;   // legacy default stream orders against nothing else here; use a dedicated one
    static thread_local cudaStream_t s_compute = nullptr;
    if (!s_compute) FADB_CUDA_CHECK(cudaStreamCreateWithFlags(&s_compute, cudaStreamNonBlocking));
    st = s_compute;
    FADB_CUDA_CHECK(cudaMemsetAsync(acc[0], 0, 2 * acc_n * sizeof(double), st));
    int buf = 0;
    int64_t chunk_idx = 0;
    for (int set = 0; set < 2; ++set) {
        const char* src = static_cast<const char*>(set == 0 ? pcm_bg_host : pcm_ev_host);
        float* emb_host = set == 0 ? emb_bg_host : emb_ev_host;
        const int64_t n = set == 0 ? n_bg : n_ev;
        for (int64_t c0 = 0, step = 0; c0 < n; c0 += step, ++chunk_idx) {
            // the very first chunk is a quarter of the rest: the kernels start after a quarter of a chunk's copy time
            step = (chunk_idx == 0 && n > cpc && cpc >= 4) ? cpc / 4 : cpc;
            const int64_t nc = (n - c0 < step) ? n - c0 : step;
            void* dpcm = h->ws_pcm[buf].as<char>();
            float* demb = h->ws_emb.as<float>() + (size_t)buf * cpc * rows * d;
            if (chunk_idx >= 2) FADB_CUDA_CHECK(cudaStreamWaitEvent(cs, h->ev_compute[buf], 0));   // buffer free again
            FADB_CUDA_CHECK(cudaMemcpyAsync(dpcm, src + (size_t)c0 * n_samples * esz, (size_t)nc * n_samples * esz,
                                            cudaMemcpyHostToDevice, cs));
            FADB_CUDA_CHECK(cudaEventRecord(h->ev_copy[buf], cs));
            FADB_CUDA_CHECK(cudaStreamWaitEvent(st, h->ev_copy[buf], 0));
            FADB_CHECK(embed_pcm_any(h, PcmSrc{dpcm, i16}, nc, n_samples, n_samples, demb, st, "fadb_fad_from_pcm_host"));
            FADB_CHECK(launch_stats_accumulate(h, demb, nc * rows, d, d, nullptr, acc[set], st));
            if (emb_host)
                FADB_CUDA_CHECK(cudaMemcpyAsync(emb_host + c0 * rows * d, demb, (size_t)nc * rows * d * sizeof(float),
                                                cudaMemcpyDeviceToHost, st));
            FADB_CUDA_CHECK(cudaEventRecord(h->ev_compute[buf], st));
            buf ^= 1;
        }
    }
    for (int set = 0; set < 2; ++set) FADB_CHECK(launch_stats_finalize(h, acc[set], d, nullptr, mu[set], sg[set], st));
    FADB_CHECK(launch_frechet(h, mu[0], sg[0], mu[1], sg[1], d, outd, st));
    double res[4];
    FADB_CUDA_CHECK(cudaMemcpyAsync(res, outd, sizeof(res), cudaMemcpyDeviceToHost, st));
    FADB_CUDA_CHECK(cudaStreamSynchronize(st));
    *fad_out = res[0];
    return check_device_flag(h);
}

int fadb_fad_from_pcm_host(fadb_handle* h, const float* pcm_bg_host, int64_t n_bg, const float* pcm_ev_host,
                           int64_t n_ev, int64_t n_samples, float* emb_bg_host, float* emb_ev_host, double* fad_out) {
    return fad_from_pcm_host_any(h, pcm_bg_host, n_bg, pcm_ev_host, n_ev, n_samples, 0, emb_bg_host, emb_ev_host, fad_out);
}

int fadb_fad_from_pcm16_host(fadb_handle* h, const int16_t* pcm_bg_host, int64_t n_bg, const int16_t* pcm_ev_host,
                             int64_t n_ev, int64_t n_samples, float* emb_bg_host, float* emb_ev_host, double* fad_out) {
    return fad_from_pcm_host_any(h, pcm_bg_host, n_bg, pcm_ev_host, n_ev, n_samples, 1, emb_bg_host, emb_ev_host, fad_out);
}

int fadb_profile_enable(fadb_handle* h, int on) {
    if (!h) return FADB_E_INVALID;
    for (cudaEvent_t e : h->prof_events) cudaEventDestroy(e);
    for (cudaEvent_t e : h->prof_front) cudaEventDestroy(e);
    h->prof_events.clear();
    h->prof_front.clear();
    h->prof_flops = 0.0;
    h->profile = on != 0;
    return FADB_OK;
}

int fadb_profile_read(fadb_handle* h, double* out4) {
    if (!h || !out4) return FADB_E_INVALID;
    cudaSetDevice(h->device);
    double ms = 0.0;
    for (size_t i = 0; i + 1 < h->prof_events.size(); i += 2) {
        FADB_CUDA_CHECK(cudaEventSynchronize(h->prof_events[i + 1]));
        float t = 0.f;
        FADB_CUDA_CHECK(cudaEventElapsedTime(&t, h->prof_events[i], h->prof_events[i + 1]));
        ms += t;
    }
    out4[0] = ms;
    out4[1] = h->prof_flops;
    out4[2] = (double)(h->prof_events.size() / 2);
    double fms = 0.0;
    for (size_t i = 0; i + 1 < h->prof_front.size(); i += 2) {
        FADB_CUDA_CHECK(cudaEventSynchronize(h->prof_front[i + 1]));
        float t = 0.f;
        FADB_CUDA_CHECK(cudaEventElapsedTime(&t, h->prof_front[i], h->prof_front[i + 1]));
        fms += t;
    }
    out4[3] = fms;
    return FADB_OK;
}

int64_t fadb_launch_count(const fadb_handle* h) { return h ? h->launches : 0; }

int fadb_device_status(fadb_handle* h) { return (h && h->err_flag_host) ? *h->err_flag_host : 0; }

// single tensor-core layer for parity tests of the implicit-GEMM kernel
int fadb_debug_conv_layer(fadb_handle* h, const float* x, int B, int H, int W, int Cin, const float* w_dev,
                          const float* bias_dev, int Cout, int ksize, int relu, int pool, float* out, void* stream) {
    if (!h || !x || !w_dev || !out) { set_error("fadb_debug_conv_layer: NULL argument"); return FADB_E_INVALID; }
    if (ksize != 3 && ksize != 1) { set_error("ksize must be 1 or 3"); return FADB_E_INVALID; }
    cudaSetDevice(h->device);
    cudaStream_t st = (cudaStream_t)stream;
    const size_t n_in = (size_t)B * H * W * Cin;
    __nv_bfloat16 *xh = nullptr, *xl = nullptr;
    FADB_CUDA_CHECK(cudaMalloc(&xh, n_in * 2));
    FADB_CUDA_CHECK(cudaMalloc(&xl, n_in * 2));
    const bool f16 = prec_is_f16(h->precision);
    int rc = split_f32_to_bf16(h, x, (int64_t)n_in, xh, xl, f16, st);
    PackedLayer L;
    L.N = Cout; L.Cin = Cin; L.taps = ksize * ksize; L.K = L.taps * Cin;
    const size_t nw = (size_t)L.N * L.K;
    if (rc == FADB_OK && (cudaMalloc(&L.w_hi, nw * 2) != cudaSuccess || cudaMalloc(&L.w_lo, nw * 2) != cudaSuccess)) {
        set_error("cudaMalloc failed");
        rc = FADB_E_NOMEM;
    }
    L.f16 = f16 ? 1 : 0;
    if (rc == FADB_OK) rc = pack_conv_weight(h, w_dev, Cout, Cin, ksize, nullptr, L.w_hi, L.w_lo, f16, st);
    uint8_t* x8 = nullptr;
    const bool wpad = Cin == 64 && ksize == 3 && h->lo_fp8_c64;       // 64-channel layers: W-padded e4m3 input (GemmParams::c64)
    if (rc == FADB_OK && h->precision == FADB_PREC_FP16X2 && h->lo_fp8 && (Cin % 128 == 0 || wpad)) {
        const size_t n8 = wpad ? (size_t)B * H * (W + 2) * Cin : n_in;
        if (cudaMalloc(&x8, n8) != cudaSuccess || cudaMalloc(&L.w8, nw) != cudaSuccess) { set_error("cudaMalloc failed"); rc = FADB_E_NOMEM; }
        if (rc == FADB_OK) rc = wpad ? quantize_e4m3_wpad(h, x, nullptr, (int64_t)B * H, W, x8, st) : quantize_e4m3(h, x, (int64_t)n_in, x8, st);
        if (rc == FADB_OK) rc = pack_conv_weight_lo8(h, w_dev, Cout, Cin, ksize, nullptr, L.w8, &L.lo_scale, st);
    }
    L.bias = const_cast<float*>(bias_dev);
    if (rc == FADB_OK) {
        LayerIO io;
        io.in_hi = xh; io.in_lo = xl; io.B = B; io.H = H; io.W = W; io.Cin = Cin;
        io.taps = L.taps; io.relu = relu; io.pool = pool; io.out_f32 = out;
        io.in8 = x8;
        io.in8_wpad = wpad ? 1 : 0;
        rc = launch_gemm_layer(h, L, io, st);
    }
    cudaError_t e = cudaStreamSynchronize(st);
    if (xh) cudaFree(xh);
    if (xl) cudaFree(xl);
    if (L.w_hi) cudaFree(L.w_hi);
    if (L.w_lo) cudaFree(L.w_lo);
    if (L.w8) cudaFree(L.w8);
    if (x8) cudaFree(x8);
    if (rc != FADB_OK) return rc;
    if (e != cudaSuccess) { set_error("debug conv layer failed: %s", cudaGetErrorString(e)); return FADB_E_CUDA; }
    return check_device_flag(h);
}

}  // extern "C"
