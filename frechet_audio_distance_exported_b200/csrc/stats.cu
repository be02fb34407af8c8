// stats.cu — embedding statistics as a summable fp64 sufficient statistic.
//
// Replaces calculate_embd_statistics (fad.py:483-496: mu = np.mean, sigma = np.cov(rowvar=False), ddof = 1)
// — SURVEY.md §2.2 K10.  acc = { n, sum (x-K), sum (x-K)(x-K)^T } in fp64 (K = optional common shift),
// so partials from batches / GPUs add (NCCL allreduce between accumulate and finalize).
// The d x d second moment is an fp64 DFMA syrk over 128x128 upper-triangular tiles, rows split across
// CTAs, one fp64 atomicAdd flush per CTA tile.
#include <type_traits>

#include "common.cuh"

namespace fadb {

// ---------------------------------------------------------------- column sums + row count
template <typename TIn>
__global__ void __launch_bounds__(256) stats_colsum_kernel(const TIn* __restrict__ x, long long n, int d, long long ld,
                                                           const double* __restrict__ shift, double* __restrict__ acc,
                                                           int rows_per_cta) {
    const int c = blockIdx.x * 256 + threadIdx.x;
    const long long r0 = (long long)blockIdx.y * rows_per_cta;
    long long r1 = r0 + rows_per_cta;
    if (r1 > n) r1 = n;
    if (c < d && r0 < r1) {
        const double k = shift ? shift[c] : 0.0;
        // four independent chains: the loop is a latency chain of dependent loads + adds otherwise (ncu r02: 107 us for
        // 4000 x 128 floats = 19 GB/s with one chain and 1024 rows per CTA)
        double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
        long long r = r0;
        for (; r + 3 < r1; r += 4) {
            s0 += (double)__ldg(x + r * ld + c) - k;
            s1 += (double)__ldg(x + (r + 1) * ld + c) - k;
            s2 += (double)__ldg(x + (r + 2) * ld + c) - k;
            s3 += (double)__ldg(x + (r + 3) * ld + c) - k;
        }
        for (; r < r1; ++r) s0 += (double)__ldg(x + r * ld + c) - k;
        atomicAdd(acc + 1 + c, (s0 + s1) + (s2 + s3));
    }
    if (blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0) atomicAdd(acc, (double)n);
}

// ---------------------------------------------------------------- syrk: S += (X-K)^T (X-K), upper tiles
// fp64 DFMA, 128 x 128 output tile per CTA, 8 x 8 register tile per thread (256 threads = 16 x 16).  A thread owns
// rows {2ty, 2ty+1} + 32q and columns {2tx, 2tx+1} + 32q of the tile, so every operand load is one 16-byte word
// that is contiguous across the 16 lanes of a half warp: conflict-free, 8 LDS.128 per 64 DFMA.  (The first version
// used 64 x 64 tiles with 4 contiguous outputs per thread: its operand loads were 4-way bank conflicted and it
// reached a quarter of the fp64 rate.)  Rows are staged 16 at a time as (x - K) in fp64; the next chunk's global
// loads are in flight while the current one is multiplied.
constexpr int TS = 128;  // output tile
constexpr int KC = 16;   // rows per smem chunk

template <typename TIn>
__global__ void __launch_bounds__(256, 1) stats_syrk_kernel(const TIn* __restrict__ x, long long n, int d, long long ld,
                                                            const double* __restrict__ shift, double* __restrict__ S,
                                                            int ntile, long long rows_per_cta) {
    // decode upper-triangular tile pair
    int pair = blockIdx.x, ti = 0;
    while (pair >= ntile - ti) { pair -= ntile - ti; ++ti; }
    const int tj = ti + pair;
    const long long r0 = (long long)blockIdx.y * rows_per_cta;
    long long r1 = r0 + rows_per_cta;
    if (r1 > n) r1 = n;
    if (r0 >= r1) return;

    __shared__ __align__(16) double sa[KC][TS];
    __shared__ __align__(16) double sb[KC][TS];
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
    double acc[8][8];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = 0.0;

    const int ci0 = ti * TS, cj0 = tj * TS;
    // staging role: element e = tid + 256 i of the KC x TS chunk -> row e / TS, column e % TS (coalesced rows)
    const int sc = threadIdx.x & (TS - 1);
    const int sr0 = threadIdx.x / TS;                       // 0 or 1; rows sr0 + 2 i
    const double ka = (shift && ci0 + sc < d) ? shift[ci0 + sc] : 0.0;
    const double kb = (shift && cj0 + sc < d) ? shift[cj0 + sc] : 0.0;
    const bool ina = ci0 + sc < d, inb = cj0 + sc < d;
    double pa[KC / 2], pb[KC / 2];
    auto fetch = [&](long long rb) {
#pragma unroll
        for (int i = 0; i < KC / 2; ++i) {
            const long long r = rb + sr0 + 2 * i;
            pa[i] = (r < r1 && ina) ? (double)__ldg(x + r * ld + ci0 + sc) - ka : 0.0;
            pb[i] = (r < r1 && inb) ? (double)__ldg(x + r * ld + cj0 + sc) - kb : 0.0;
        }
    };
    fetch(r0);
    for (long long rb = r0; rb < r1; rb += KC) {
#pragma unroll
        for (int i = 0; i < KC / 2; ++i) {
            sa[sr0 + 2 * i][sc] = pa[i];
            sb[sr0 + 2 * i][sc] = pb[i];
        }
        __syncthreads();
        if (rb + KC < r1) fetch(rb + KC);                   // next chunk's loads overlap this chunk's FMAs
#pragma unroll 4
        for (int k = 0; k < KC; ++k) {
            double a[8], b[8];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const double2 av = *reinterpret_cast<const double2*>(&sa[k][2 * ty + 32 * q]);
                const double2 bv = *reinterpret_cast<const double2*>(&sb[k][2 * tx + 32 * q]);
                a[2 * q] = av.x; a[2 * q + 1] = av.y;
                b[2 * q] = bv.x; b[2 * q + 1] = bv.y;
            }
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < 8; ++j) acc[i][j] = fma(a[i], b[j], acc[i][j]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int gi = ci0 + 2 * ty + 32 * (i >> 1) + (i & 1), gj = cj0 + 2 * tx + 32 * (j >> 1) + (j & 1);
            if (gi < d && gj < d) atomicAdd(S + (size_t)gi * d + gj, acc[i][j]);
        }
}

// ---------------------------------------------------------------- syrk on the tensor cores (d >= 512)
// S += Y^T Y with Y = X - K as ONE launch of the tcgen05 implicit-GEMM kernel per chunk of <= 65536 rows: the
// "activations" and the "weights" of that linear layer are the same matrix Yt = Y^T [d][rows] (K-major for both MMA
// operands), split into IEEE fp16 hi + lo planes (y = hi + lo to 2^-22), three MMAs per product (hi hi, lo hi, hi lo;
// lo lo ~ 2^-24 is dropped), each pass its own accumulation segments, segment sums added in fp32 round-to-nearest by
// the epilogue (gemm_tc.cu).  Tiles strictly below the diagonal are skipped.  The chunk's fp32 product is then added
// to the fp64 statistic.  Replaces 2 N d^2 fp64 DFMA flops by 3 x N d^2 tensor flops: 634 ms -> ~30 ms for two sets of
// 1e6 x 2048 (BASELINE configs[4]); the fp64 kernel above stays for d < 512 and as the checker.
constexpr int kSyrkChunkRows = 65536;

// x [rows][ld] -> hi / lo [d][stride] fp16 (kpad <= stride columns written) of y = x - K', zero beyond the chunk's rows,
// and ysum[c] += sum of y over the chunk (fp64).  K' = the column mean of the chunk's FIRST rows0 rows (csum0[1 + c] /
// rows0): any centre near the mean keeps the products free of the mean^2 term (the fp32 chains inside the MMA truncate: a
// bias relative to the SECOND MOMENT, which the later subtraction of mu mu^T would amplify), and a provisional one lets
// the chunk be read once.  64 features x 64 samples per CTA through shared memory; every thread writes 16 consecutive
// samples of one feature (two 16-byte stores per plane), a warp covers 8 complete 128-byte rows.
template <typename TIn>
__global__ void __launch_bounds__(256) syrk_split_transpose_kernel(const TIn* __restrict__ x, long long rows, int d,
                                                                   long long ld, const double* __restrict__ csum0,
                                                                   long long rows0, int stride, int kpad,
                                                                   __half* __restrict__ hi, __half* __restrict__ lo,
                                                                   double* __restrict__ ysum) {
    __shared__ float tile[64][65];                          // [sample][feature]
    __shared__ double part[4][64];
    const int c0 = blockIdx.x * 64;
    const long long r0 = (long long)blockIdx.y * 64;
    {
        const int f = threadIdx.x & 63, c = c0 + f;
        const double kp = (c < d) ? csum0[1 + c] / (double)rows0 : 0.0;
        double acc = 0.0;
#pragma unroll 4
        for (int i = 0; i < 16; ++i) {
            const int sl = (threadIdx.x >> 6) + 4 * i;
            const long long r = r0 + sl;
            float v = 0.f;
            if (r < rows && c < d) v = (float)((double)__ldg(x + r * ld + c) - kp);
            tile[sl][f] = v;
            acc += (double)v;
        }
        part[threadIdx.x >> 6][f] = acc;
    }
    __syncthreads();
    if (threadIdx.x < 64 && c0 + threadIdx.x < d)
        atomicAdd(ysum + c0 + threadIdx.x, (part[0][threadIdx.x] + part[1][threadIdx.x]) + (part[2][threadIdx.x] + part[3][threadIdx.x]));
    const int f = threadIdx.x >> 2, pt = threadIdx.x & 3, c = c0 + f;
    if (c < d && r0 + 16 * pt < kpad) {
        uint32_t ph[8], pl[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const float v0 = fminf(fmaxf(tile[16 * pt + 2 * k][f], -65504.f), 65504.f);
            const float v1 = fminf(fmaxf(tile[16 * pt + 2 * k + 1][f], -65504.f), 65504.f);
            const __half2 h = __floats2half2_rn(v0, v1);
            const __half2 l = __floats2half2_rn(v0 - __low2float(h), v1 - __high2float(h));
            ph[k] = *reinterpret_cast<const uint32_t*>(&h);
            pl[k] = *reinterpret_cast<const uint32_t*>(&l);
        }
        uint4* dh = reinterpret_cast<uint4*>(hi + (size_t)c * stride + r0 + 16 * pt);
        uint4* dl = reinterpret_cast<uint4*>(lo + (size_t)c * stride + r0 + 16 * pt);
        dh[0] = make_uint4(ph[0], ph[1], ph[2], ph[3]); dh[1] = make_uint4(ph[4], ph[5], ph[6], ph[7]);
        dl[0] = make_uint4(pl[0], pl[1], pl[2], pl[3]); dl[1] = make_uint4(pl[4], pl[5], pl[6], pl[7]);
    }
}

// S (fp64, upper 128 x 128 tiles) += s delta^T + delta s^T + rows delta delta^T with s = sum y (ysum), delta = K' - K
// (provisional chunk centre minus the caller's common shift):
//   sum (x-K)(x-K)^T = sum (y+delta)(y+delta)^T = sum y y^T + s delta^T + delta s^T + rows delta delta^T,
// and sum y y^T is added by the GEMM's own epilogue.
__global__ void syrk_rank2_kernel(int d, const double* __restrict__ csum0, long long rows0, const double* __restrict__ ysum,
                                  long long rows, const double* __restrict__ shift, double* __restrict__ S) {
    const size_t total = (size_t)d * d;
    for (size_t e = blockIdx.x * (size_t)blockDim.x + threadIdx.x; e < total; e += (size_t)gridDim.x * blockDim.x) {
        const int i = (int)(e / d), j = (int)(e % d);
        if (i / TS <= j / TS) {
            const double di = csum0[1 + i] / (double)rows0 - (shift ? shift[i] : 0.0);
            const double dj = csum0[1 + j] / (double)rows0 - (shift ? shift[j] : 0.0);
            S[e] += ysum[i] * dj + di * ysum[j] + (double)rows * di * dj;
        }
    }
}

static int stats_syrk_tensor(fadb_handle* h, const float* emb, int64_t n, int d, int64_t ld, const double* shift,
                             double* S, cudaStream_t st) {
    // Row chunks sized so that both fp16 planes of a chunk (4 d bytes per row) stay in L2 (126 MB) while every tile of
    // the product re-reads them; the row stride gets 64 extra elements so that rows do not sit 2^k bytes apart.
    int64_t chunk = (int64_t)(64u << 20) / (4 * (int64_t)d);
    chunk = chunk / 64 * 64;
    if (chunk > kSyrkChunkRows) chunk = kSyrkChunkRows;
    if (chunk < 1024) chunk = 1024;
    if (n < chunk) chunk = ((n + 63) / 64) * 64;
    const int64_t stride = chunk + 64;
    const size_t plane = (size_t)d * stride;                                  // fp16 elements per plane
    const size_t vec_bytes = ((size_t)(2 + 2 * d) * sizeof(double) + 255) / 256 * 256;    // csum0 [1 + d] | ysum [d]
    FADB_CHECK(h->ws_syrk.reserve(2 * plane * sizeof(__half) + vec_bytes + 256));
    double* csum0 = h->ws_syrk.as<double>();
    double* ysum = csum0 + 1 + d;
    __half* hi = reinterpret_cast<__half*>(h->ws_syrk.as<char>() + vec_bytes);
    __half* lo = hi + plane;
    for (int64_t r0 = 0; r0 < n; r0 += chunk) {
        const int64_t rows = (n - r0 < chunk) ? n - r0 : chunk;
        const int64_t rows0 = rows < 512 ? rows : 512;
        const int kpad = (int)(((rows + 63) / 64) * 64);
        // provisional centre: column sums (fp64) of the chunk's first rows
        FADB_CUDA_CHECK(cudaMemsetAsync(csum0, 0, (size_t)(2 + 2 * d) * sizeof(double), st));
        {
            const int rows_per = 32;
            dim3 g((d + 255) / 256, (unsigned)((rows0 + rows_per - 1) / rows_per));
            stats_colsum_kernel<float><<<g, 256, 0, st>>>(emb + r0 * ld, rows0, d, ld, nullptr, csum0, rows_per);
        }
        dim3 grid((d + 63) / 64, (unsigned)(kpad / 64));
        syrk_split_transpose_kernel<float><<<grid, 256, 0, st>>>(emb + r0 * ld, rows, d, ld, csum0, rows0, (int)stride, kpad,
                                                                 hi, lo, ysum);
        PackedLayer L;
        L.N = d; L.K = kpad; L.Cin = kpad; L.taps = 1; L.f16 = 1;
        L.w_hi = reinterpret_cast<__nv_bfloat16*>(hi);
        L.w_lo = reinterpret_cast<__nv_bfloat16*>(lo);
        L.bias = nullptr;
        LayerIO io;
        io.in_hi = reinterpret_cast<const __nv_bfloat16*>(hi);
        io.in_lo = reinterpret_cast<const __nv_bfloat16*>(lo);
        io.B = 1; io.H = 1; io.W = d; io.Cin = kpad; io.taps = 1; io.relu = 0; io.pool = 0;
        io.out_f64 = S;
        io.row_stride = stride;
        io.syrk = 1;
        io.seg_blocks = 4;              // 16-MMA chains: truncation bias ~1e-6 of the centred second moment
        FADB_CHECK(launch_gemm_layer(h, L, io, st));
        int g = (int)(((size_t)d * d + 255) / 256);
        if (g > 148 * 16) g = 148 * 16;
        syrk_rank2_kernel<<<g, 256, 0, st>>>(d, csum0, rows0, ysum, rows, shift, S);
        h->launches += 3;
    }
    FADB_CUDA_CHECK(cudaGetLastError());
    return FADB_OK;
}

// ---------------------------------------------------------------- finalize
__global__ void stats_finalize_kernel(const double* __restrict__ acc, int d, const double* __restrict__ shift,
                                      double* __restrict__ mu, double* __restrict__ sigma) {
    const double n = acc[0];
    const double* s1 = acc + 1;
    const double* S = acc + 1 + d;
    const size_t total = (size_t)d * d;
    for (size_t e = blockIdx.x * (size_t)blockDim.x + threadIdx.x; e < total; e += (size_t)gridDim.x * blockDim.x) {
        const int i = (int)(e / d), j = (int)(e % d);
        const double sij = (i / TS <= j / TS) ? S[(size_t)i * d + j] : S[(size_t)j * d + i];
        const double di = s1[i] / n, dj = s1[j] / n;
        sigma[e] = (sij - n * di * dj) / (n - 1.0);
        if (j == 0 && mu) mu[i] = (shift ? shift[i] : 0.0) + di;
    }
}

template <typename TIn>
static int stats_accumulate_t(fadb_handle* h, const TIn* emb, int64_t n, int d, int64_t ld, const double* shift,
                              double* acc, cudaStream_t st) {
    FADB_REQUIRE(d > 0 && d <= 8192, "embedding dim %d out of range", d);
    FADB_REQUIRE(ld >= d, "row stride %lld < d", (long long)ld);
    if (n <= 0) return FADB_OK;
    {
        int rows_per = 128;
        long long ny = (n + rows_per - 1) / rows_per;
        while (ny > 32768) { rows_per *= 2; ny = (n + rows_per - 1) / rows_per; }
        dim3 grid((d + 255) / 256, (unsigned)ny);
        stats_colsum_kernel<TIn><<<grid, 256, 0, st>>>(emb, n, d, ld, shift, acc, rows_per);
        h->launches++;
    }
    if (std::is_same<TIn, float>::value && h->tc_syrk && d >= 512 && d % 128 == 0 && n >= 8192) {
        // second moments on the tensor cores (the packed linear-layer path needs d % 128 == 0)
        FADB_CHECK(stats_syrk_tensor(h, reinterpret_cast<const float*>(emb), n, d, ld, shift, acc + 1 + d, st));
    } else {
        const int ntile = (d + TS - 1) / TS;
        const int npair = ntile * (ntile + 1) / 2;
        long long want = (2LL * h->sm_count + npair - 1) / npair;           // row splits to fill the machine
        long long rows_per = (n + want - 1) / want;
        if (rows_per < 256) rows_per = 256;
        rows_per = (rows_per + KC - 1) / KC * KC;
        long long ny = (n + rows_per - 1) / rows_per;
        while (ny > 32768) { rows_per *= 2; ny = (n + rows_per - 1) / rows_per; }
        dim3 grid(npair, (unsigned)ny);
        stats_syrk_kernel<TIn><<<grid, 256, 0, st>>>(emb, n, d, ld, shift, acc + 1 + d, ntile, rows_per);
        h->launches++;
    }
    FADB_CUDA_CHECK(cudaGetLastError());
    return FADB_OK;
}

int launch_stats_accumulate(fadb_handle* h, const float* emb, int64_t n, int d, int64_t ld, const double* shift,
                            double* acc, cudaStream_t st) {
    return stats_accumulate_t<float>(h, emb, n, d, ld, shift, acc, st);
}
int launch_stats_accumulate_f64(fadb_handle* h, const double* emb, int64_t n, int d, int64_t ld, const double* shift,
                                double* acc, cudaStream_t st) {
    return stats_accumulate_t<double>(h, emb, n, d, ld, shift, acc, st);
}

int launch_stats_finalize(fadb_handle* h, const double* acc, int d, const double* shift, double* mu, double* sigma,
                          cudaStream_t st) {
    FADB_REQUIRE(d > 0 && d <= 8192, "embedding dim %d out of range", d);
    size_t total = (size_t)d * d;
    int grid = (int)((total + 255) / 256);
    if (grid > 2048) grid = 2048;
    stats_finalize_kernel<<<grid, 256, 0, st>>>(acc, d, shift, mu, sigma);
    h->launches++;
    FADB_CUDA_CHECK(cudaGetLastError());
    return FADB_OK;
}

}  // namespace fadb
