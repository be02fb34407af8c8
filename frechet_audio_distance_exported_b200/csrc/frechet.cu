// frechet.cu — Frechet distance between two Gaussians, fp64, entirely on the device.
//
// Replaces calculate_frechet_distance (fad.py:498-555): the reference forms the NON-symmetric
// product S1 S2 and takes a complex Schur square root (scipy.linalg.sqrtm; 22-30 s at d = 2048 on
// 8 cores).  Only the TRACE of that square root is used (fad.py:553), and
//     tr sqrtm(S1 S2) = sum_i sqrt(lambda_i(S1 S2)) = sum_i sqrt(lambda_i(L^T S2 L)),  S1 = L L^T,
// so this file computes a semi-definite Cholesky factor of S1, the symmetric PSD matrix
// M = L^T S2 L, reduces it to tridiagonal form with Householder reflections and gets all
// eigenvalues by Sturm-sequence bisection (values only — no eigenvectors, no complex arithmetic).
// Negative rounding-noise eigenvalues are clipped at 0.  The reference's "non-finite -> add eps*I"
// retry (fad.py:539-544) and imaginary-part check (fad.py:547-551) are unreachable here because
// nothing can become complex or non-finite for finite PSD inputs; a non-finite result raises the
// device error flag instead.
//
// Everything is latency/L2-bound fp64 (the 33.5 MB matrix at d = 2048 is L2-resident):
//   * Cholesky: blocked right-looking, NB = 64 (3 launches per panel)
//   * two d^3 DGEMMs (CUDA-core DFMA, 64x64 tiles)
//   * tridiagonalisation: ONE persistent cooperative kernel, a grid-wide barrier per Householder step — the
//     rank-2 update of step k is fused with the symmetric matrix-vector product of step k+1, so the trailing
//     matrix is read+written once per step; the O(d) vector algebra is recomputed by every CTA
//   * bisection: one thread per eigenvalue.
#include <cooperative_groups.h>

#include "common.cuh"

namespace fadb {

// ------------------------------------------------------------------------------------------------
// small utilities
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ double warp_max(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
// block-wide sum broadcast to all threads; red must hold >= 33 doubles
__device__ __forceinline__ double block_sum(double v, double* red) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    v = warp_sum(v);
    __syncthreads();
    if (lane == 0) red[warp] = v;
    __syncthreads();
    if (warp == 0) {
        double t = (lane < nw) ? red[lane] : 0.0;
        t = warp_sum(t);
        if (lane == 0) red[32] = t;
    }
    __syncthreads();
    return red[32];
}

// copy + symmetrise: out = (in + in^T) / 2 ; also max diagonal / traces
__global__ void frechet_prepare_kernel(const double* __restrict__ s1, const double* __restrict__ s2, int d,
                                       double* __restrict__ A, double* __restrict__ B, double* __restrict__ scal) {
    const size_t total = (size_t)d * d;
    for (size_t e = blockIdx.x * (size_t)blockDim.x + threadIdx.x; e < total; e += (size_t)gridDim.x * blockDim.x) {
        const int i = (int)(e / d), j = (int)(e % d);
        A[e] = 0.5 * (s1[e] + s1[(size_t)j * d + i]);
        B[e] = 0.5 * (s2[e] + s2[(size_t)j * d + i]);
    }
    if (blockIdx.x == 0) {
        __shared__ double red[33];
        double t1 = 0, t2 = 0, mx = 0;
        for (int i = threadIdx.x; i < d; i += blockDim.x) {
            const double a = s1[(size_t)i * d + i], b = s2[(size_t)i * d + i];
            t1 += a; t2 += b; mx = fmax(mx, a);
        }
        const double T1 = block_sum(t1, red);
        const double T2 = block_sum(t2, red);
        // max via sum trick is wrong; do a proper max reduction
        mx = warp_max(mx);
        __syncthreads();
        if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = mx;
        __syncthreads();
        if (threadIdx.x == 0) {
            double m = 0;
            for (int w = 0; w < (int)((blockDim.x + 31) >> 5); ++w) m = fmax(m, red[w]);
            scal[0] = T1 + T2;      // tr S1 + tr S2
            scal[1] = m;            // max diag S1 (pivot threshold scale)
            scal[2] = 0.0;          // tr sqrt accumulator
        }
    }
}

// ------------------------------------------------------------------------------------------------
// DGEMM  C = alpha * op(A) * op(B) + beta * C   (row-major, 64x64x16 tiles, 4x4 per thread)
// ------------------------------------------------------------------------------------------------
// dyn (optional, device): the rank r found by the Cholesky step; dyn_mask bit 0 -> M = r, bit 1 -> N = r (the grid is
// sized for the largest case and surplus tiles leave at once: no host round trip for a data-dependent dimension)
template <bool TA, bool TB>
__global__ void __launch_bounds__(256) dgemm_kernel(int M, int N, int K, double alpha, const double* __restrict__ A,
                                                    int lda, const double* __restrict__ B, int ldb, double beta,
                                                    double* __restrict__ C, int ldc, int lower_only,
                                                    const int* __restrict__ dyn, int dyn_mask) {
    if (dyn) {
        const int r = *dyn;
        if (dyn_mask & 1) M = r;
        if (dyn_mask & 2) N = r;
    }
    if ((int)blockIdx.y * 64 >= M || (int)blockIdx.x * 64 >= N) return;
    if (lower_only && blockIdx.x > blockIdx.y) return;      // tile strictly above the diagonal
    __shared__ __align__(16) double sA[16][64 + 2];
    __shared__ __align__(16) double sB[16][64 + 2];
    const int m0 = blockIdx.y * 64, n0 = blockIdx.x * 64;
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
    double acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.0;

    for (int k0 = 0; k0 < K; k0 += 16) {
        // A tile -> sA[k][m]
        if (!TA) {
            const int m = threadIdx.x >> 2, kq = (threadIdx.x & 3) * 4;
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const int gm = m0 + m, gk = k0 + kq + q;
                sA[kq + q][m] = (gm < M && gk < K) ? A[(size_t)gm * lda + gk] : 0.0;
            }
        } else {
            const int k = threadIdx.x >> 4, mq = (threadIdx.x & 15) * 4;
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const int gm = m0 + mq + q, gk = k0 + k;
                sA[k][mq + q] = (gm < M && gk < K) ? A[(size_t)gk * lda + gm] : 0.0;
            }
        }
        // B tile -> sB[k][n]
        if (!TB) {
            const int k = threadIdx.x >> 4, nq = (threadIdx.x & 15) * 4;
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const int gn = n0 + nq + q, gk = k0 + k;
                sB[k][nq + q] = (gn < N && gk < K) ? B[(size_t)gk * ldb + gn] : 0.0;
            }
        } else {
            const int n = threadIdx.x >> 2, kq = (threadIdx.x & 3) * 4;
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const int gn = n0 + n, gk = k0 + kq + q;
                sB[kq + q][n] = (gn < N && gk < K) ? B[(size_t)gn * ldb + gk] : 0.0;
            }
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < 16; ++k) {
            double a[4], b[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) a[i] = sA[k][ty * 4 + i];
#pragma unroll
            for (int j = 0; j < 4; ++j) b[j] = sB[k][tx * 4 + j];
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fma(a[i], b[j], acc[i][j]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int gm = m0 + ty * 4 + i, gn = n0 + tx * 4 + j;
            if (gm < M && gn < N) {
                double* c = C + (size_t)gm * ldc + gn;
                *c = (beta == 0.0) ? alpha * acc[i][j] : fma(alpha, acc[i][j], beta * (*c));
            }
        }
}

template <bool TA, bool TB>
static void dgemm(fadb_handle* h, int M, int N, int K, double alpha, const double* A, int lda, const double* B, int ldb,
                  double beta, double* C, int ldc, int lower_only, cudaStream_t st, const int* dyn = nullptr,
                  int dyn_mask = 0) {
    if (M <= 0 || N <= 0) return;
    dim3 grid((N + 63) / 64, (M + 63) / 64);
    dgemm_kernel<TA, TB><<<grid, 256, 0, st>>>(M, N, K, alpha, A, lda, B, ldb, beta, C, ldc, lower_only, dyn, dyn_mask);
    h->launches++;
}

// ------------------------------------------------------------------------------------------------
// Semi-definite Cholesky (lower), blocked NB = 64
// ------------------------------------------------------------------------------------------------
constexpr int NB = 64;

// factor the nb x nb diagonal block in shared memory; pivots <= tol zero their column
__global__ void __launch_bounds__(256) chol_diag_kernel(double* __restrict__ A, int lda, int nb,
                                                        const double* __restrict__ scal, int d) {
    __shared__ double s[NB][NB + 1];
    const double tol = 16.0 * d * 2.220446049250313e-16 * scal[1];
    for (int e = threadIdx.x; e < nb * nb; e += 256) s[e / nb][e % nb] = A[(size_t)(e / nb) * lda + (e % nb)];
    __syncthreads();
    for (int j = 0; j < nb; ++j) {
        const double piv = s[j][j];
        const bool ok = piv > tol;
        const double l = ok ? sqrt(piv) : 0.0;
        const double inv = ok ? 1.0 / l : 0.0;
        __syncthreads();
        if (threadIdx.x == 0) s[j][j] = l;
        for (int i = j + 1 + threadIdx.x; i < nb; i += 256) s[i][j] = ok ? s[i][j] * inv : 0.0;
        __syncthreads();
        // trailing update of the lower triangle: s[i][c] -= s[i][j] * s[c][j], i >= c > j
        const int rem = nb - j - 1;
        for (int e = threadIdx.x; e < rem * rem; e += 256) {
            const int i = j + 1 + e / rem, c = j + 1 + e % rem;
            if (i >= c) s[i][c] = fma(-s[i][j], s[c][j], s[i][c]);
        }
        __syncthreads();
    }
    for (int e = threadIdx.x; e < nb * nb; e += 256) {
        const int i = e / nb, c = e % nb;
        A[(size_t)i * lda + c] = (i >= c) ? s[i][c] : 0.0;     // upper part of the block zeroed
    }
}

// rows below the diagonal block: X <- X * Lkk^{-T}   (one thread per row, forward substitution)
__global__ void __launch_bounds__(128) chol_trsm_kernel(const double* __restrict__ Lkk, double* __restrict__ P, int lda,
                                                        int nb, int nrows) {
    __shared__ double sl[NB][NB + 1];
    for (int e = threadIdx.x; e < nb * nb; e += 128) sl[e / nb][e % nb] = Lkk[(size_t)(e / nb) * lda + (e % nb)];
    __syncthreads();
    const int r = blockIdx.x * 128 + threadIdx.x;
    if (r >= nrows) return;
    double* row = P + (size_t)r * lda;
    double x[NB];
#pragma unroll
    for (int j = 0; j < NB; ++j) x[j] = (j < nb) ? row[j] : 0.0;
#pragma unroll
    for (int j = 0; j < NB; ++j) {
        if (j < nb) {
            double v = x[j];
#pragma unroll
            for (int c = 0; c < j; ++c) v = fma(-x[c], sl[j][c], v);
            const double l = sl[j][j];
            x[j] = (l > 0.0) ? v / l : 0.0;
        }
    }
#pragma unroll
    for (int j = 0; j < NB; ++j)
        if (j < nb) row[j] = x[j];
}

// zero the strict upper triangle (Cholesky leaves the old symmetric values there)
__global__ void tril_kernel(double* __restrict__ A, int d) {
    const size_t total = (size_t)d * d;
    for (size_t e = blockIdx.x * (size_t)blockDim.x + threadIdx.x; e < total; e += (size_t)gridDim.x * blockDim.x) {
        const int i = (int)(e / d), j = (int)(e % d);
        if (j > i) A[e] = 0.0;
    }
}
// Rank compaction.  The semi-definite Cholesky leaves a ZERO column for every pivot below the tolerance, so a covariance of
// N < d embeddings (BASELINE configs[1]: N = 1000, d = 2048 -> rank 999) has only r = N - 1 non-zero columns, and
// L^T S2 L is non-zero only in those r rows / columns.  The columns are gathered to the left (Lc: d x r) and everything
// downstream works on r instead of d: the products shrink to d*d*r + r*d*r and the tridiagonalisation to r^3.
// rank_scan: one CTA; idx[c] = c-th column with a non-zero diagonal, meta[0] = r.
__global__ void __launch_bounds__(1024) rank_scan_kernel(const double* __restrict__ L, int d, int* __restrict__ idx,
                                                         int* __restrict__ meta) {
    __shared__ int s_cnt[1024];
    const int per = (d + 1023) / 1024;
    const int j0 = threadIdx.x * per;
    int c = 0;
    for (int j = j0; j < j0 + per && j < d; ++j) c += (L[(size_t)j * d + j] != 0.0);
    s_cnt[threadIdx.x] = c;
    __syncthreads();
    for (int o = 1; o < 1024; o <<= 1) {                    // inclusive scan
        const int v = (threadIdx.x >= o) ? s_cnt[threadIdx.x - o] : 0;
        __syncthreads();
        s_cnt[threadIdx.x] += v;
        __syncthreads();
    }
    int pos = s_cnt[threadIdx.x] - c;
    for (int j = j0; j < j0 + per && j < d; ++j)
        if (L[(size_t)j * d + j] != 0.0) idx[pos++] = j;
    if (threadIdx.x == 1023) meta[0] = s_cnt[1023];
}
// Lc[i][c] = L[i][idx[c]] for c < r (row-major, leading dimension d); columns >= r are left untouched (never read)
__global__ void rank_gather_kernel(const double* __restrict__ L, int d, const int* __restrict__ idx,
                                   const int* __restrict__ meta, double* __restrict__ Lc) {
    const int r = meta[0];
    const size_t total = (size_t)d * r;
    for (size_t e = blockIdx.x * (size_t)blockDim.x + threadIdx.x; e < total; e += (size_t)gridDim.x * blockDim.x) {
        const int i = (int)(e / r), c = (int)(e % r);
        const int j = idx[c];
        Lc[(size_t)i * d + c] = (j <= i) ? L[(size_t)i * d + j] : 0.0;      // lower triangle only (upper is stale)
    }
}

__global__ void symmetrize_kernel(double* __restrict__ A, int d, const int* __restrict__ meta) {
    const int n = meta ? meta[0] : d;                      // active size (leading dimension stays d)
    const size_t total = (size_t)n * n;
    for (size_t e = blockIdx.x * (size_t)blockDim.x + threadIdx.x; e < total; e += (size_t)gridDim.x * blockDim.x) {
        const int i = (int)(e / n), j = (int)(e % n);
        if (j > i) {
            const double v = 0.5 * (A[(size_t)i * d + j] + A[(size_t)j * d + i]);
            A[(size_t)i * d + j] = v;
            A[(size_t)j * d + i] = v;
        }
    }
}

// ------------------------------------------------------------------------------------------------
// Householder tridiagonalisation, one step (k = -1 .. d-3) per call of tridiag_step
//   in : A (rows/cols >= k+1 current), v_k, beta_k, p_k = beta_k A v_k        (k = -1: all zero)
//   out: A updated for rows/cols >= k+2, v_{k+1}, beta_{k+1}, p_{k+1}, diag[k+1], off[k+1]
// vec layout: [v (d) | p (d) | beta (1)] , ping-pong by step parity.
// ------------------------------------------------------------------------------------------------
// NOTE: no __restrict__ / read-only-cache loads on A and the vectors here — inside the persistent kernel they are
// written by other CTAs one grid barrier earlier, so they must be read through the coherent L2 path (__ldcg).
// d = active matrix size, ld = leading dimension of A (row stride)
__device__ __forceinline__ void tridiag_step(double* A, int d, int ld, int k, const double* vec_in, double* vec_out,
                                             double* diag, double* off, double* tsm, double* red) {
    double* sv = tsm;            // v_k          [d]
    double* sw = tsm + d;        // w_k          [d]
    double* sn = tsm + 2 * d;    // v_{k+1}      [d]
    const int tid = threadIdx.x, nt = blockDim.x;
    const int r = k + 1;                         // row that becomes final in this step
    const double* vin = vec_in;
    const double* pin = vec_in + d;
    const double beta = __ldcg(vec_in + 2 * d);

    // (a) K = beta (v.p)/2 ; w = p - K v
    double part = 0.0;
    for (int i = r + tid; i < d; i += nt) {
        const double v = (i >= r) ? __ldcg(vin + i) : 0.0;
        sv[i] = v;
        part += v * __ldcg(pin + i);
    }
    const double K = 0.5 * beta * block_sum(part, red);
    for (int i = r + tid; i < d; i += nt) sw[i] = __ldcg(pin + i) - K * sv[i];
    __syncthreads();

    // (b) updated row r -> diag[r], x = row[r+1:], next reflector
    const double vr = sv[r], wr = sw[r];
    const double* Ar = A + (size_t)r * ld;
    part = 0.0;
    for (int c = r + 1 + tid; c < d; c += nt) {
        const double x = fma(-vr, sw[c], fma(-wr, sv[c], __ldcg(Ar + c)));
        sn[c] = x;
        part += x * x;
    }
    const double nrm2 = block_sum(part, red);
    const int m = d - r - 1;                     // length of x
    double x0 = 0.0, alpha = 0.0, bnext = 0.0;
    if (m >= 1) x0 = sn[r + 1];
    const double tail2 = nrm2 - x0 * x0;
    const bool reflect = (m >= 2) && (tail2 > 0.0);
    if (reflect) {
        alpha = (x0 > 0.0) ? -sqrt(nrm2) : sqrt(nrm2);
        // v = x - alpha e1 ; v.v = 2 alpha (alpha - x0)
        bnext = 2.0 / (2.0 * alpha * (alpha - x0));
    } else {
        alpha = x0;
    }
    __syncthreads();
    if (tid == 0 && m >= 1) sn[r + 1] = reflect ? (x0 - alpha) : 0.0;
    if (tid == 1) sn[r] = 0.0;                   // the vector path of phase (c) may touch column r
    if (!reflect)
        for (int c = r + 2 + tid; c < d; c += nt) sn[c] = 0.0;
    __syncthreads();
    if (blockIdx.x == 0) {
        if (tid == 0) {
            diag[r] = fma(-2.0 * vr, wr, __ldcg(Ar + r));
            if (m >= 1) off[r] = alpha;
            vec_out[2 * d] = bnext;
        }
        for (int c = tid; c < d; c += nt) vec_out[c] = (c > r) ? sn[c] : 0.0;
    }

    // (c) rows >= r+1: rank-2 update with (v_k, w_k) fused with p_{k+1} = beta_{k+1} A v_{k+1}
    const int warp = tid >> 5, lane = tid & 31, nwarp = nt >> 5;
    const int first = r + 1;
    double* pout = vec_out + d;
    // fixed row ownership (row -> global warp row % W) so a row is always touched by the same SM; loads bypass
    // L1 anyway (__ldcg): L1 is not coherent across SMs and this runs inside a persistent kernel
    for (int row = blockIdx.x * nwarp + warp; row < d; row += gridDim.x * nwarp) {
        if (row < first) continue;
        double* Arow = A + (size_t)row * ld;
        const double vrow = sv[row], wrow = sw[row];
        double dot = 0.0;
        if (((d | ld) & 1) == 0) {
            // even d: rows are 16-byte aligned -> 128-bit loads, 8 x 512 B in flight per warp (the streaming part is
            // bound by bytes in flight per SM against the L2 round trip, not by L2 bandwidth).  Starts at the even
            // column <= first; the extra column (the finished column r) is updated harmlessly, sn[r] = 0.
            const int c0 = first & ~1;
            int c = c0 + 2 * lane;
            for (; c + 448 < d; c += 512) {
                double2 a[8];
#pragma unroll
                for (int u = 0; u < 8; ++u) a[u] = __ldcg(reinterpret_cast<const double2*>(Arow + c + 64 * u));
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    const int cc = c + 64 * u;
                    a[u].x = fma(-vrow, sw[cc], fma(-wrow, sv[cc], a[u].x));
                    a[u].y = fma(-vrow, sw[cc + 1], fma(-wrow, sv[cc + 1], a[u].y));
                    *reinterpret_cast<double2*>(Arow + cc) = a[u];
                    dot = fma(a[u].x, sn[cc], dot);
                    dot = fma(a[u].y, sn[cc + 1], dot);
                }
            }
            for (; c < d; c += 64) {
                double2 a = __ldcg(reinterpret_cast<const double2*>(Arow + c));
                a.x = fma(-vrow, sw[c], fma(-wrow, sv[c], a.x));
                a.y = fma(-vrow, sw[c + 1], fma(-wrow, sv[c + 1], a.y));
                *reinterpret_cast<double2*>(Arow + c) = a;
                dot = fma(a.x, sn[c], dot);
                dot = fma(a.y, sn[c + 1], dot);
            }
        } else {
            int c = first + lane;
            for (; c + 96 < d; c += 128) {
                const double a0 = __ldcg(Arow + c), a1 = __ldcg(Arow + c + 32), a2 = __ldcg(Arow + c + 64),
                             a3 = __ldcg(Arow + c + 96);
                const double b0 = fma(-vrow, sw[c], fma(-wrow, sv[c], a0));
                const double b1 = fma(-vrow, sw[c + 32], fma(-wrow, sv[c + 32], a1));
                const double b2 = fma(-vrow, sw[c + 64], fma(-wrow, sv[c + 64], a2));
                const double b3 = fma(-vrow, sw[c + 96], fma(-wrow, sv[c + 96], a3));
                Arow[c] = b0; Arow[c + 32] = b1; Arow[c + 64] = b2; Arow[c + 96] = b3;
                dot = fma(b0, sn[c], dot);
                dot = fma(b1, sn[c + 32], dot);
                dot = fma(b2, sn[c + 64], dot);
                dot = fma(b3, sn[c + 96], dot);
            }
            for (; c < d; c += 32) {
                const double a = fma(-vrow, sw[c], fma(-wrow, sv[c], __ldcg(Arow + c)));
                Arow[c] = a;
                dot = fma(a, sn[c], dot);
            }
        }
        dot = warp_sum(dot);
        if (lane == 0) pout[row] = bnext * dot;
    }
    if (blockIdx.x == 0)
        for (int c = tid; c <= r && c < d; c += nt) pout[c] = 0.0;
}

// All Householder steps in ONE cooperative launch: a grid-wide barrier separates step k (which leaves p_{k+1}
// complete in global memory) from step k+1.  One launch per step cost ~6 us of launch + ramp per step
// (2047 steps at d = 2048); the grid barrier costs ~2 us.
// ld = leading dimension (and size of the vec / diag buffers); the active size is *meta (the rank found by the Cholesky
// step, read on the device: every CTA sees the same value, so the grid-wide barriers stay matched)
__global__ void __launch_bounds__(512) tridiag_persistent_kernel(double* __restrict__ A, int ld, double* __restrict__ vec0,
                                                                 double* __restrict__ vec1, double* __restrict__ diag,
                                                                 double* __restrict__ off, const int* __restrict__ meta) {
    extern __shared__ __align__(16) double tsm[];
    __shared__ double red[33];
    cooperative_groups::grid_group grid = cooperative_groups::this_grid();
    const int d = meta ? meta[0] : ld;
    int step = 0;
    for (int k = -1; k <= d - 3; ++k, ++step) {
        // the vector buffers keep their [v (ld) | p (ld) | beta] layout
        const double* vin = (step & 1) ? vec1 : vec0;
        double* vout = (step & 1) ? vec0 : vec1;
        tridiag_step(A, d, ld, k, vin, vout, diag, off, tsm, red);
        grid.sync();
    }
    if (blockIdx.x == 0 && threadIdx.x == 0 && d >= 1) diag[d - 1] = A[(size_t)(d - 1) * ld + (d - 1)];
}

// ------------------------------------------------------------------------------------------------
// Sturm bisection: thread i finds the i-th smallest eigenvalue of the tridiagonal (diag, off);
// accumulates sum sqrt(max(lambda, 0)) into scal[2].
//
// The count "eigenvalues < x" is the number of sign changes of the leading-principal-minor sequence
//   p_0 = 1, p_1 = a_0 - x, p_{j+1} = (a_j - x) p_j - e_{j-1}^2 p_{j-1}
// (division-free: one dependent DFMA per step instead of an fp64 division; the e^2 p_{j-1} product is
// off the critical path).  The matrix is scaled to spectral radius <= 1 so |p| grows by at most 2.42x
// per step; every 8 steps the pair is renormalised by its exponent.  An exact zero takes the sign
// opposite to its predecessor (Wilkinson's convention), which also keeps decoupled blocks (e = 0) right.
// Fixed 56 halvings of [-1-delta, 1+delta] -> absolute accuracy ~4e-17 * radius, all that tr sqrt needs.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) bisect_kernel(const double* __restrict__ diag, const double* __restrict__ off,
                                                     int dmax, double* __restrict__ scal, double* __restrict__ eig_out,
                                                     const int* __restrict__ meta) {
    const int d = meta ? meta[0] : dmax;                   // active size (the rank), device-resident
    if (d <= 0) return;                                    // S1 = 0: tr sqrt(S1 S2) = 0 (scal[2] stays 0)
    extern __shared__ __align__(16) double bsm[];
    double* sa = bsm;          // [d]  a / radius
    double* se2 = bsm + d;     // [d]  (off / radius)^2 ; se2[j] couples j and j+1
    __shared__ double red[33];
    __shared__ double s_rad;
    double rad = 0.0;
    for (int i = threadIdx.x; i < d; i += blockDim.x) {
        const double el = (i > 0) ? fabs(off[i - 1]) : 0.0;
        const double er = (i < d - 1) ? fabs(off[i]) : 0.0;
        rad = fmax(rad, fabs(diag[i]) + el + er);                 // Gershgorin bound on the spectral radius
    }
    rad = warp_max(rad);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = rad;
    __syncthreads();
    if (threadIdx.x == 0) {
        double m = 0.0;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) m = fmax(m, red[w]);
        s_rad = (m > 0.0 && isfinite(m)) ? m : 1.0;
    }
    __syncthreads();
    const double radius = s_rad, inv = 1.0 / radius;
    for (int i = threadIdx.x; i < d; i += blockDim.x) {
        sa[i] = diag[i] * inv;
        const double e = (i < d - 1) ? off[i] * inv : 0.0;
        se2[i] = e * e;
    }
    __syncthreads();

    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    double contrib = 0.0;
    if (idx < d) {
        double lo = -1.0 - 1e-9, hi = 1.0 + 1e-9;
        for (int it = 0; it < 56; ++it) {
            const double x = 0.5 * (lo + hi);
            double pp = 1.0;
            double pc = sa[0] - x;
            if (pc == 0.0) pc = -1e-30;
            int cnt = (pc < 0.0);
            for (int j0 = 1; j0 < d; j0 += 8) {
                const int j1 = (j0 + 8 < d) ? j0 + 8 : d;
                for (int j = j0; j < j1; ++j) {
                    double pn = fma(sa[j] - x, pc, -(se2[j - 1] * pp));
                    if (pn == 0.0) pn = -pc * 7.888609052210118e-31;          // 2^-100, opposite sign
                    cnt += (__double2hiint(pn) ^ __double2hiint(pc)) < 0;     // sign change
                    pp = pc;
                    pc = pn;
                }
                // renormalise the pair by the exponent of |pc| (keeps everything far from over/underflow)
                const int ex = ((__double2hiint(pc) >> 20) & 0x7ff) - 1023;
                if (ex > 200 || ex < -200) {
                    const double sc = __hiloint2double((1023 - ex) << 20, 0);  // 2^-ex
                    pc *= sc;
                    pp *= sc;
                }
            }
            if (cnt > idx) hi = x; else lo = x;
        }
        const double lam = 0.5 * (lo + hi) * radius;
        if (eig_out) eig_out[idx] = lam;
        contrib = sqrt(fmax(lam, 0.0));
    }
    const double tot = block_sum(contrib, red);
    if (threadIdx.x == 0) atomicAdd(scal + 2, tot);
}

__global__ void frechet_combine_kernel(const double* __restrict__ mu1, const double* __restrict__ mu2, int d,
                                       const double* __restrict__ scal, double* __restrict__ out, int* err_flag) {
    __shared__ double red[33];
    double part = 0.0;
    for (int i = threadIdx.x; i < d; i += blockDim.x) {
        const double t = mu1[i] - mu2[i];
        part += t * t;
    }
    const double dd = block_sum(part, red);
    if (threadIdx.x == 0) {
        const double fad = dd + scal[0] - 2.0 * scal[2];      // fad.py:555
        out[0] = fad;
        out[1] = scal[2];
        out[2] = scal[0];
        out[3] = dd;
        // a non-finite result (NaN covariance from a one-row set, ...) is DATA, not a pipeline failure: it is returned
        // in out[0] and the caller raises (fad.py's ValueError); the handle stays usable
        (void)err_flag;
    }
}

// ------------------------------------------------------------------------------------------------
// per handle (= per device): kernel attributes are device state, a process may hold handles on several GPUs
int frechet_init(fadb_handle* h) {
    (void)h;
    FADB_CUDA_CHECK(cudaFuncSetAttribute(tridiag_persistent_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         3 * 8192 * (int)sizeof(double)));
    FADB_CUDA_CHECK(cudaFuncSetAttribute(bisect_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         2 * 8192 * (int)sizeof(double)));
    return FADB_OK;
}

int launch_frechet(fadb_handle* h, const double* mu1, const double* s1, const double* mu2, const double* s2, int d,
                   double* out, cudaStream_t st) {
    FADB_REQUIRE(d >= 1 && d <= 8192, "Frechet: d=%d out of range", d);
    const size_t dd = (size_t)d * d;
    // workspace: A (d*d) | B (d*d) | T (d*d) | vec ping-pong 2*(2d+1) | diag d | off d | eig d | scal 8 | idx d ints | meta
    const size_t need = (3 * dd + 2 * (2 * (size_t)d + 1) + 3 * (size_t)d + 8) * sizeof(double) + ((size_t)d + 4) * sizeof(int);
    FADB_CHECK(h->ws_frechet.reserve(need));
    double* A = h->ws_frechet.as<double>();
    double* B = A + dd;
    double* T = B + dd;
    double* vec0 = T + dd;
    double* vec1 = vec0 + (2 * (size_t)d + 1);
    double* diag = vec1 + (2 * (size_t)d + 1);
    double* off = diag + d;
    double* eig = off + d;
    double* scal = eig + d;

    int g = (int)((dd + 255) / 256);
    if (g > 1184) g = 1184;
    frechet_prepare_kernel<<<g, 256, 0, st>>>(s1, s2, d, A, B, scal);
    h->launches++;

    // ---- Cholesky of A (lower), semi-definite safe
    for (int k0 = 0; k0 < d; k0 += NB) {
        const int nb = (d - k0 < NB) ? d - k0 : NB;
        double* Akk = A + (size_t)k0 * d + k0;
        chol_diag_kernel<<<1, 256, 0, st>>>(Akk, d, nb, scal, d);
        h->launches++;
        const int rem = d - k0 - nb;
        if (rem > 0) {
            double* P = A + (size_t)(k0 + nb) * d + k0;
            chol_trsm_kernel<<<(rem + 127) / 128, 128, 0, st>>>(Akk, P, d, nb, rem);
            h->launches++;
            double* A22 = A + (size_t)(k0 + nb) * d + (k0 + nb);
            dgemm<false, true>(h, rem, rem, nb, -1.0, P, d, P, d, 1.0, A22, d, /*lower_only=*/1, st);
        }
    }
    // ---- rank compaction: Lc = the r non-zero columns of L, gathered to the left (r = d for a full-rank S1)
    int* idx = reinterpret_cast<int*>(scal + 8);
    int* meta = idx + d;
    rank_scan_kernel<<<1, 1024, 0, st>>>(A, d, idx, meta);
    rank_gather_kernel<<<g, 256, 0, st>>>(A, d, idx, meta, T);
    h->launches += 2;

    // ---- M = Lc^T (S2 Lc): r x r, leading dimension d
    dgemm<false, false>(h, d, d, d, 1.0, B, d, T, d, 0.0, A, d, 0, st, meta, 2);     // A = S2 Lc       (d x r)
    dgemm<true, false>(h, d, d, d, 1.0, T, d, A, d, 0.0, B, d, 0, st, meta, 3);      // B = Lc^T A      (r x r)
    symmetrize_kernel<<<g, 256, 0, st>>>(B, d, meta);
    h->launches++;

    // ---- tridiagonalise B (active size r)
    FADB_CUDA_CHECK(cudaMemsetAsync(vec0, 0, 2 * (2 * (size_t)d + 1) * sizeof(double), st));
    {
        const size_t smem = 3 * (size_t)d * sizeof(double);
        int grid = (d + 15) / 16;                        // 16 warps per CTA, one row per warp per sweep
        if (grid > h->sm_count) grid = h->sm_count;
        if (grid < 1) grid = 1;
        double* Bm = B;
        int dd_ = d;
        const int* mp = meta;
        void* args[] = {(void*)&Bm, (void*)&dd_, (void*)&vec0, (void*)&vec1, (void*)&diag, (void*)&off, (void*)&mp};
        FADB_CUDA_CHECK(cudaLaunchCooperativeKernel((const void*)tridiag_persistent_kernel, dim3(grid), dim3(512), args,
                                                    smem, st));
        h->launches++;
    }
    // ---- eigenvalues + trace of the square root
    bisect_kernel<<<(d + 127) / 128, 128, 2 * (size_t)d * sizeof(double), st>>>(diag, off, d, scal, eig, meta);
    h->launches++;
    frechet_combine_kernel<<<1, 256, 0, st>>>(mu1, mu2, d, scal, out, h->err_flag);
    h->launches++;
    FADB_CUDA_CHECK(cudaGetLastError());
    return FADB_OK;
}

}  // namespace fadb
