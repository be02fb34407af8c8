// frontend.cu — fused log-mel front-end kernel family (PCM -> framing -> window -> real FFT ->
// |.| or |.|^2 -> sparse triangular mel projection -> log -> patch / time-padded layout), one warp
// per STFT frame, everything between the PCM read and the feature write stays in shared memory.
//
// Replaces:
//   VGGish : models/vggish.py:102-141 (_frame, _periodic_hann, _stft_magnitude), :150-190 (HTK mel
//            matrix), :223-227 (log(mel + 0.01)), :268-277 (96-frame patches, trailing frames dropped)
//   PANN   : models/pann.py:104-139 (librosa.stft centred/reflect, power, Slaney mel, 10 log10 max(.,1e-10)),
//            fad.py:41-66 (zero rows up to T' = 32k-24)
//   CLAP   : models/clap.py:70-72 (int16 truncation), fad.py:356-359 (zero-pad to 480000), fad.py:69-91
//
// Arithmetic: windowing, FFT and the real-FFT split run in fp64 like the reference (numpy promotes
// float32 PCM x float64 window to float64; librosa evaluates the FFT in float64) — an fp32 FFT leaves a
// noise floor that breaks 1e-5 log-mel parity on tonal inputs (empty bins sit at log(0.01)).  From the
// spectrum on (|X|, mel projection, log) fp32 is enough: sums of positives, relative error ~1e-7.
// Real FFT of length NF = complex in-place Stockham FFT of length NF/2 (radix-4 passes + one radix-2
// pass when needed, inputs of a pass staged in registers) + split.
#include <algorithm>
#include <cmath>
#include <vector>

#include "common.cuh"
#include "conv1_dev.cuh"
#include "tc_ptx.cuh"

namespace fadb {

// FrontTables (common.cuh) live in the handle: one set per model, on the handle's own device

struct FrontParams {
    const void* pcm;    // fp32 samples, or int16 when pcm_i16
    int pcm_i16;
    long long pcm_stride;
    int n_samples;      // samples physically present per clip
    int logical_len;    // length used for reflect padding (CLAP: 480000)
    int frames_valid;   // frames computed (rest of rows_out are zero rows)
    int rows_out;       // rows written per clip
    int hop, win_len;
    int centered, power_db, quantize;
    const double2* tw;
    const double* win;
    const int* band_start;
    const int* band_len;
    const float* band_wt;   // transposed band weights (FrontTables::band_wt), staged in shared memory by the kernels
    int wt_rows, wt_off1;
    double win_rot_c, win_rot_s;   // cos / sin of 2 pi 64 / win_len: the Hann window advances 64 samples per register slot
    float* out;         // [n_clips][rows_out][64]
};

// one clip's samples: fp32, or int16 scaled by 2^-15 on load (exact)
struct PcmView {
    const float* f;
    const short* s;
    __device__ __forceinline__ float operator[](long long i) const {
        return s ? (float)__ldg(s + i) * 3.0517578125e-05f : __ldg(f + i);
    }
    // byte address of sample i (for prefetch instructions)
    __device__ __forceinline__ const char* addr(long long i) const {
        return s ? reinterpret_cast<const char*>(s + i) : reinterpret_cast<const char*>(f + i);
    }
    __device__ __forceinline__ int sample_bytes() const { return s ? 2 : 4; }
    // samples i, i+1 in one load; valid when pair_aligned(i)
    __device__ __forceinline__ bool pair_aligned(long long i) const {
        return s ? ((reinterpret_cast<uintptr_t>(s + i) & 3) == 0) : ((reinterpret_cast<uintptr_t>(f + i) & 7) == 0);
    }
    __device__ __forceinline__ float2 pair(long long i) const {
        if (s) {
            const short2 v = __ldg(reinterpret_cast<const short2*>(s + i));
            return make_float2((float)v.x * 3.0517578125e-05f, (float)v.y * 3.0517578125e-05f);
        }
        return __ldg(reinterpret_cast<const float2*>(f + i));
    }
};
__device__ __forceinline__ PcmView clip_view(const void* base, int i16, long long clip, long long stride) {
    PcmView v;
    v.f = i16 ? nullptr : static_cast<const float*>(base) + clip * stride;
    v.s = i16 ? static_cast<const short*>(base) + clip * stride : nullptr;
    return v;
}

__device__ __forceinline__ void prefetch_l1(const void* p) { asm volatile("prefetch.global.L1 [%0];" ::"l"(p)); }
__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

__device__ __forceinline__ double2 cmul(double2 a, double2 b) {
    return make_double2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}
__device__ __forceinline__ double2 cadd(double2 a, double2 b) { return make_double2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ double2 csub(double2 a, double2 b) { return make_double2(a.x - b.x, a.y - b.y); }

// FFT buffer index padding: one 16-byte slot every 8 complex values, so that the stride-4 / stride-16
// scatter of the early Stockham passes spreads over all shared-memory banks (ncu r01: the unpadded version
// was shared-memory-bandwidth bound, 83 % LSU wavefront utilisation, 63 M bank-conflict wavefronts).
__device__ __forceinline__ int pidx(int n) { return n + (n >> 3); }

// Per-pass twiddle tables in shared memory, laid out [pass][butterfly q][r-1][lane]: a warp's twiddle load is
// 32 consecutive 16-byte slots (conflict-free) instead of a strided gather from the exp(-2 pi i k/NF) table
// (16-way bank conflicts at stride NF/64).
template <int M>
struct FftPlan {
    // radix of every pass (the first pass has all twiddles = 1 and takes its input straight from the loader):
    // M = 128: 4,4,4,2 ; M = 256: 8,8,4 ; M = 512: 8,8,8.  Radix 8 = one butterfly of 8 points per lane and pass:
    // one shared-memory round trip less than radix 4 (ncu r01: this stage is shared-memory-wavefront bound).
    static constexpr int kNumPasses = (M == 128) ? 4 : 3;
    static constexpr int radix(int i) { return M == 128 ? (i == 3 ? 2 : 4) : (M == 256 ? (i == 2 ? 4 : 8) : 8); }
    static constexpr int per(int i) { return ((M / radix(i)) + 31) / 32; }
    static constexpr int offset(int i) {                                // twiddle slots before pass i >= 1, units of 32 double2
        int o = 0;
        for (int j = 1; j < i; ++j) o += per(j) * (radix(j) - 1);
        return o;
    }
    static constexpr int kSlots = offset(kNumPasses) * 32;              // double2 entries
};

// Per-lane constants of the front end, held in REGISTERS for the whole kernel.  ncu (profiles/r02_ncu_front.json) shows the
// kernel bound by shared-memory wavefronts (LSU data pipe 78 % busy, fp64 pipe 18 %); a fifth of those wavefronts were
// per-frame reloads of constants that depend on the lane only: the pass twiddles (7 + 6 double2 per lane), the split
// twiddles (5) and the window (8).  They are now derived per frame from one BASE value each with a few fp64
// multiplications: w^r = w^(r-1) * w for the radix-R butterfly inputs, w_(k+32) = w_k * W^32 for the split, and a
// rotation by 64 samples for the Hann window (cos / sin recurrence; at most 8 steps, ~1e-15 absolute).
template <int M>
struct LaneConsts {
    static constexpr int kBases = (M == 256) ? 3 : 4;       // sum over passes >= 1 of butterflies per lane
    double2 pb[kBases];     // pass twiddle bases: exp(-2 pi i (i & (pp-1)) tstep / NF) of butterfly i = lane + 32 q
    double2 sb;             // split twiddle base exp(-2 pi i lane / NF)
    double wc[2], ws[2];    // cos / sin of 2 pi n / win_len for n = 2 lane, 2 lane + 1
};
template <int NF> struct SplitStep;      // exp(-2 pi i 32 / NF)
template <> struct SplitStep<256> { static constexpr double c = 0.70710678118654752440, s = -0.70710678118654752440; };
template <> struct SplitStep<512> { static constexpr double c = 0.92387953251128675613, s = -0.38268343236508977173; };
template <> struct SplitStep<1024> { static constexpr double c = 0.98078528040323044913, s = -0.19509032201612826785; };

template <int M, int NF>
__device__ __forceinline__ void build_lane_consts(LaneConsts<M>& lc, const double2* __restrict__ tw_global, int win_len,
                                                  int lane) {
    using P = FftPlan<M>;
    int pp = P::radix(0), slot = 0;
#pragma unroll
    for (int ps = 1; ps < P::kNumPasses; ++ps) {
        const int R = P::radix(ps), PER = P::per(ps);
        const int tstep = NF / (pp * R);
#pragma unroll
        for (int q = 0; q < PER; ++q) {
            const int i = lane + 32 * q;
            lc.pb[slot++] = tw_global[(i & (pp - 1)) * tstep];
        }
        pp *= R;
    }
    lc.sb = tw_global[lane];
#pragma unroll
    for (int e = 0; e < 2; ++e) sincospi(2.0 * (double)(2 * lane + e) / (double)win_len, &lc.ws[e], &lc.wc[e]);
}

// One in-place Stockham pass of radix R over the warp's M-point buffer: every lane first pulls ALL of
// its butterfly inputs into registers, the warp syncs, then the outputs go back to the same buffer.
// The FIRST pass takes its inputs from registers instead: z[j] = input element lane + 32 j (windowed PCM straight from
// global memory).
template <int M, int R, bool FIRST>
__device__ __forceinline__ void fft_pass(double2* __restrict__ x, const double2* __restrict__ pbase, int pp, int lane,
                                         const double2 (&z)[M / 32]) {
    constexpr int T = M / R;                 // butterflies in the pass
    constexpr int PER = (T + 31) / 32;       // per lane
    double2 u[PER][R];
#pragma unroll
    for (int q = 0; q < PER; ++q) {
        const int i = lane + 32 * q;
        if (T >= 32 || i < T) {
            double2 pw = FIRST ? make_double2(1.0, 0.0) : pbase[q];      // w^r of this butterfly, built up by multiplication
#pragma unroll
            for (int r = 0; r < R; ++r) {
                double2 v;
                if (FIRST) {
                    v = z[q + r * PER];               // element i + r T = lane + 32 (q + r PER)
                } else {
                    v = x[pidx(i + r * T)];
                    if (r > 0) {
                        v = cmul(v, pw);
                        if (r + 1 < R) pw = cmul(pw, pbase[q]);
                    }
                }
                u[q][r] = v;
            }
        }
    }
    __syncwarp();
#pragma unroll
    for (int q = 0; q < PER; ++q) {
        const int i = lane + 32 * q;
        if (T >= 32 || i < T) {
            const int k = i & (pp - 1);
            const int j = (i - k) * R + k;
            if (R == 8) {
                constexpr double kH = 0.70710678118654752440;
                const double2 a0 = cadd(u[q][0], u[q][4]), a1 = csub(u[q][0], u[q][4]);
                const double2 a2 = cadd(u[q][2], u[q][6]), t26 = csub(u[q][2], u[q][6]);
                const double2 b0 = cadd(u[q][1], u[q][5]), b1 = csub(u[q][1], u[q][5]);
                const double2 b2 = cadd(u[q][3], u[q][7]), t37 = csub(u[q][3], u[q][7]);
                const double2 a3 = make_double2(t26.y, -t26.x), b3 = make_double2(t37.y, -t37.x);   // -i * (.)
                const double2 e0 = cadd(a0, a2), e2 = csub(a0, a2), e1 = cadd(a1, a3), e3 = csub(a1, a3);
                const double2 o0 = cadd(b0, b2), o2r = csub(b0, b2), o1r = cadd(b1, b3), o3r = csub(b1, b3);
                const double2 o1 = make_double2(kH * (o1r.x + o1r.y), kH * (o1r.y - o1r.x));        // * (1 - i)/sqrt2
                const double2 o2 = make_double2(o2r.y, -o2r.x);                                     // * -i
                const double2 o3 = make_double2(kH * (o3r.y - o3r.x), -kH * (o3r.x + o3r.y));       // * (-1 - i)/sqrt2
                x[pidx(j)] = cadd(e0, o0);
                x[pidx(j + pp)] = cadd(e1, o1);
                x[pidx(j + 2 * pp)] = cadd(e2, o2);
                x[pidx(j + 3 * pp)] = cadd(e3, o3);
                x[pidx(j + 4 * pp)] = csub(e0, o0);
                x[pidx(j + 5 * pp)] = csub(e1, o1);
                x[pidx(j + 6 * pp)] = csub(e2, o2);
                x[pidx(j + 7 * pp)] = csub(e3, o3);
            } else if (R == 4) {
                const double2 v0 = cadd(u[q][0], u[q][2]);
                const double2 v1 = csub(u[q][0], u[q][2]);
                const double2 v2 = cadd(u[q][1], u[q][3]);
                const double2 d = csub(u[q][1], u[q][3]);
                const double2 v3 = make_double2(d.y, -d.x);          // -i * d
                x[pidx(j)] = cadd(v0, v2);
                x[pidx(j + pp)] = cadd(v1, v3);
                x[pidx(j + 2 * pp)] = csub(v0, v2);
                x[pidx(j + 3 * pp)] = csub(v1, v3);
            } else {
                x[pidx(j)] = cadd(u[q][0], u[q][1]);
                x[pidx(j + pp)] = csub(u[q][0], u[q][1]);
            }
        }
    }
    __syncwarp();
}

// complex FFT of length M (128 / 256 / 512) of the sequence held as z[j] = element lane + 32 j, result in x (padded
// index pidx)
template <int M>
__device__ __forceinline__ void fft_inplace(double2* __restrict__ x, const LaneConsts<M>& lc, int lane,
                                            const double2 (&z)[M / 32]) {
    using P = FftPlan<M>;
    constexpr int R0 = P::radix(0), R1 = P::radix(1), R2 = P::radix(2);
    static_assert(M / P::radix(0) >= 32, "first pass: at least one butterfly per lane");
    static_assert(P::per(1) + P::per(2) + (P::kNumPasses == 4 ? P::per(3) : 0) == LaneConsts<M>::kBases, "base count");
    fft_pass<M, R0, true>(x, lc.pb, 1, lane, z);
    fft_pass<M, R1, false>(x, lc.pb, R0, lane, z);
    fft_pass<M, R2, false>(x, lc.pb + P::per(1), R0 * R1, lane, z);
    if constexpr (P::kNumPasses == 4) fft_pass<M, P::radix(3), false>(x, lc.pb + P::per(1) + P::per(2), R0 * R1 * R2, lane, z);
}

constexpr int kFrontWarps = 8;
constexpr int kFramesPerWarp = 8;

template <int NF>
struct FrontSmem {
    static constexpr int M = NF / 2;
    static constexpr int kBufSlots = M + M / 8 + 2;                       // padded complex buffer per warp
    static constexpr int kSpecPitch = M + 4;                              // floats per warp spectrum (M + 1 used)
    // (the twiddle and window tables of round 1 are gone from shared memory: LaneConsts keeps their bases in registers)
    static constexpr int kBufBytes = kFrontWarps * kBufSlots * 16;
    static constexpr int kSpecBytes = kFrontWarps * kSpecPitch * 4;
    static constexpr int kWtRows = NF / 16;                               // >= band-weight rows of every model (checked on the host)
    static constexpr int kWtBytes = kWtRows * 32 * 4;
    static constexpr int kTotal = kBufBytes + kSpecBytes + kWtBytes;
};

// One STFT frame -> 64 log-mel values, by one warp.  out_row[lane] and out_row[lane + 32] are written
// (global memory in the stand-alone kernel, the shared-memory patch tile in the fused VGGish kernel).
template <int NF, bool SPLIT = false>
__device__ __forceinline__ void frame_logmel(const FrontParams& p, const PcmView pcm, int row, int lane,
                                             double2* __restrict__ x, float* __restrict__ spec,
                                             const LaneConsts<NF / 2>& lc, const float* __restrict__ s_wt,
                                             const int (&bst)[2], const int (&bln)[2], float* __restrict__ out_row) {
    constexpr int M = NF / 2;
    // ---- load + window (fp32 PCM x fp64 Hann, like numpy's promotion): z[n] = x[2n] + i x[2n+1]
    const long long f0 = (long long)row * p.hop - (p.centered ? NF / 2 : 0);
    const bool interior = (f0 >= 0) && (f0 + p.win_len <= p.n_samples);
    // both samples of z[n] in one load when the pair is naturally aligned (8 B fp32 / 4 B int16) and inside the window
    const bool pair_ok = interior && ((p.win_len & 1) == 0) && pcm.pair_aligned(f0);
    {   // this warp's next frame (row + 1) starts one hop later: pull its lines into L1 while this frame computes
        // (ncu r01: a third of all warp time in this phase was long-scoreboard stall on the PCM loads)
        const long long nb = (f0 + p.hop) + (long long)lane * (128 / pcm.sample_bytes());
        if (nb >= 0 && nb < (long long)p.n_samples && nb < f0 + p.hop + p.win_len + (128 / pcm.sample_bytes()))
            prefetch_l1(pcm.addr(nb));
    }
    // The windowed frame goes to the first FFT pass in registers (no round trip through shared memory).  The common
    // case is a branch-free block of predicated 8-byte loads, so all of a frame's loads are in flight together; the
    // first version decided per element and serialised eight load -> convert round trips (20 % of the stall samples
    // of this phase sat on the first conversion after each load).
    constexpr int J = M / 32;
    float2 raw[J];
    if (pair_ok) {
        if (pcm.s) {
#pragma unroll
            for (int j = 0; j < J; ++j) {
                const int n = lane + 32 * j;
                short2 v = make_short2(0, 0);
                if (2 * n < p.win_len) v = __ldg(reinterpret_cast<const short2*>(pcm.s + f0 + 2 * n));
                raw[j] = make_float2((float)v.x * 3.0517578125e-05f, (float)v.y * 3.0517578125e-05f);
            }
        } else {
#pragma unroll
            for (int j = 0; j < J; ++j) {
                const int n = lane + 32 * j;
                raw[j] = make_float2(0.f, 0.f);
                if (2 * n < p.win_len) raw[j] = __ldg(reinterpret_cast<const float2*>(pcm.f + f0 + 2 * n));
            }
        }
    } else {
#pragma unroll
        for (int j = 0; j < J; ++j) {            // frame edges (reflect padding, clip ends) and odd alignments
            const int n = lane + 32 * j;
            float s01[2] = {0.f, 0.f};
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const int i = 2 * n + e;
                if (i < p.win_len) {
                    long long src = f0 + i;
                    if (p.centered) {            // np.pad(mode='reflect'): edge sample not repeated
                        if (src < 0) src = -src;
                        if (src >= p.logical_len) src = 2LL * (p.logical_len - 1) - src;
                    }
                    if (src >= 0 && src < p.n_samples) s01[e] = pcm[src];
                }
            }
            raw[j] = make_float2(s01[0], s01[1]);
        }
    }
    if (p.quantize) {                            // clap.py:70-72: (x*32767).astype(int16)/32767, trunc toward 0
#pragma unroll
        for (int j = 0; j < J; ++j) {
            raw[j].x = __fdiv_rn(truncf(__fmul_rn(raw[j].x, 32767.0f)), 32767.0f);
            raw[j].y = __fdiv_rn(truncf(__fmul_rn(raw[j].y, 32767.0f)), 32767.0f);
        }
    }
    double2 z[J];
    {   // fp32 PCM x fp64 periodic Hann 0.5 - 0.5 cos(2 pi n / win_len) (numpy's promotion), 0 past win_len; the cosine of
        // n = 2 (lane + 32 j) + e comes from the lane's base angle by j rotations of 64 samples
        double c0 = lc.wc[0], s0 = lc.ws[0], c1 = lc.wc[1], s1 = lc.ws[1];
#pragma unroll
        for (int j = 0; j < J; ++j) {
            const int n = 2 * (lane + 32 * j);
            const double w0 = (n < p.win_len) ? fma(-0.5, c0, 0.5) : 0.0;
            const double w1 = (n + 1 < p.win_len) ? fma(-0.5, c1, 0.5) : 0.0;
            z[j] = make_double2((double)raw[j].x * w0, (double)raw[j].y * w1);
            if (j + 1 < J) {
                const double t0 = c0 * p.win_rot_c - s0 * p.win_rot_s, t1 = c1 * p.win_rot_c - s1 * p.win_rot_s;
                s0 = s0 * p.win_rot_c + c0 * p.win_rot_s;
                s1 = s1 * p.win_rot_c + c1 * p.win_rot_s;
                c0 = t0;
                c1 = t1;
            }
        }
    }
    fft_inplace<M>(x, lc, lane, z);

    // ---- real-FFT split, then |X| (VGGish, vggish.py:141) or |X|^2 (PANN, pann.py:118) in fp32.
    // Bins k and M-k come from the same pair (Z[k], Z[M-k]): X[k] = ze + w_k zo, X[M-k] = conj(ze - w_k zo), so each
    // lane reads a pair once and writes both bins (k = 0 gives DC and Nyquist, k = M/2 pairs with itself).
    auto put = [&](int k, double re_d, double im_d) {
        const float re = (float)re_d, im = (float)im_d;
        const float pw = fmaf(re, re, im * im);
        spec[k] = p.power_db ? pw : sqrtf(pw);
    };
    double2 w = lc.sb;                           // exp(-2 pi i k / NF), k = lane + 32 q: advanced by W^32 per q
#pragma unroll
    for (int q = 0; q <= M / 64; ++q) {
        const int k = lane + 32 * q;
        if (q > 0) w = cmul(w, make_double2(SplitStep<NF>::c, SplitStep<NF>::s));
        if (k <= M / 2) {
            const double2 a = x[pidx(k)];
            const double2 bz = x[pidx((M - k) & (M - 1))];
            const double2 ze = make_double2(0.5 * (a.x + bz.x), 0.5 * (a.y - bz.y));
            const double2 zo = make_double2(0.5 * (a.y + bz.y), -0.5 * (a.x - bz.x));   // (a - conj b) / (2i)
            const double tx = w.x * zo.x - w.y * zo.y, ty = w.x * zo.y + w.y * zo.x;
            put(k, ze.x + tx, ze.y + ty);
            if (k != M / 2) put(M - k, ze.x - tx, ze.y - ty);
        }
    }
    __syncwarp();
    // ---- mel projection (sparse triangular bands) + log, fp32; weights transposed in shared memory (row i, column
    // lane) so a warp's weight load is one conflict-free wavefront
#pragma unroll
    for (int h2 = 0; h2 < 2; ++h2) {
        float acc = 0.f;
        const float* wt = s_wt + (h2 ? p.wt_off1 : 0) * 32 + lane;
        for (int i = 0; i < bln[h2]; ++i) acc = fmaf(wt[i * 32], spec[bst[h2] + i], acc);
        float o;
        if (p.power_db) o = 10.0f * log10f(fmaxf(acc, 1e-10f));      // pann.py:133-134
        else o = logf(acc + 0.01f);                                  // vggish.py:227
        if (SPLIT) {   // fused VGGish kernel: the tile holds {bf16 hi, bf16 lo} pairs, o = hi + lo to ~2^-17
            const __nv_bfloat16 hi = __float2bfloat16_rn(o);
            const __nv_bfloat16 lo = __float2bfloat16_rn(o - __bfloat162float(hi));
            reinterpret_cast<uint32_t*>(out_row)[lane + 32 * h2] =
                (uint32_t)__bfloat16_as_ushort(hi) | ((uint32_t)__bfloat16_as_ushort(lo) << 16);
        } else {
            out_row[lane + 32 * h2] = o;
        }
    }
    __syncwarp();
}

template <int NF>
__global__ void __launch_bounds__(kFrontWarps * 32, 2) fadb_frontend_kernel(const FrontParams p) {
    using S = FrontSmem<NF>;
    constexpr int M = NF / 2;
    extern __shared__ __align__(16) uint8_t fsm[];
    double2* s_buf = reinterpret_cast<double2*>(fsm);
    float* s_spec = reinterpret_cast<float*>(fsm + S::kBufBytes);
    float* s_wt = s_spec + S::kSpecBytes / 4;

    for (int i = threadIdx.x; i < p.wt_rows * 32; i += kFrontWarps * 32) s_wt[i] = p.band_wt[i];
    __syncthreads();

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    LaneConsts<M> lc;
    build_lane_consts<M, NF>(lc, p.tw, p.win_len, lane);
    double2* x = s_buf + warp * S::kBufSlots;
    float* spec = s_spec + warp * S::kSpecPitch;
    const int clip = blockIdx.y;
    const PcmView pcm = clip_view(p.pcm, p.pcm_i16, clip, p.pcm_stride);
    float* out = p.out + (size_t)clip * p.rows_out * 64;
    int bst[2], bln[2];                      // this lane's two mel bands
#pragma unroll
    for (int h2 = 0; h2 < 2; ++h2) {
        bst[h2] = p.band_start[lane + 32 * h2];
        bln[h2] = p.band_len[lane + 32 * h2];
    }
    for (int fi = 0; fi < kFramesPerWarp; ++fi) {
        const int row = blockIdx.x * (kFrontWarps * kFramesPerWarp) + warp * kFramesPerWarp + fi;
        if (row >= p.rows_out) break;
        if (row >= p.frames_valid) {                 // PANN time padding: literal zero rows (fad.py:61-64)
            out[(size_t)row * 64 + lane] = 0.f;
            out[(size_t)row * 64 + lane + 32] = 0.f;
            continue;
        }
        frame_logmel<NF>(p, pcm, row, lane, x, spec, lc, s_wt, bst, bln, out + (size_t)row * 64);
    }
}

// ------------------------------------------------------------------------------------------------
// Fused VGGish kernel.  Phase 1 (fp64 FFT, shared-memory / latency bound): 8 warps x 12 frames write the log-mel
// patch into a shared-memory tile, so the fp32 features never touch HBM.  Phase 2 runs conv3x3(1->64) as a tcgen05 GEMM:
//   D[pixel, cout] = sum_k A[pixel, k] * B[cout, k],  K = 32 = 27 used columns:
//     A = [x_hi(tap 0..7) | x_lo(tap 0..7) | x_hi(tap 0..7) | x_hi8, x_lo8, x_hi8, 0...]
//     B = [w_hi(tap 0..7) | w_hi(tap 0..7) | w_lo(tap 0..7) | w_hi8, w_hi8, w_lo8, 0...]
// i.e. the three significant terms of (x_hi + x_lo)(w_hi + w_lo) in ONE accumulation: conv1 keeps ~2^-17
// operand precision in both precision modes at 2 MMAs (M=128, N=64, K=16) per 128 pixels.  Phase 1 leaves the
// patch as {hi, lo} bf16 pairs, so building an A row is shared-memory words + byte permutes.
// The 2x2 max pool costs no shuffles: a step covers 128 pooling WINDOWS (4 pooled rows) as FOUR M=128 tiles, one
// per window position q = (dy, dx), accumulating into four TMEM column groups; TMEM lane = window, so a thread
// reads the 4 candidates of its window from columns q*64 + ch and pools in registers.
// A / B tiles are un-swizzled core-matrix tiles (tc_ptx.cuh make_nosw_desc); the A tiles overlay the FFT buffers,
// which are idle in phase 2.  CTAs are persistent (tables, B tile, TMEM set up once) and loop over patches.
// ------------------------------------------------------------------------------------------------
struct FusedTcSmem {
    using S = FrontSmem<512>;
    static constexpr int kTilePitch = 80;                                     // words (8-byte aligned rows)
    static constexpr int kTileBytes = 98 * kTilePitch * 4;
    static constexpr int kATileBytes = kC1ATileBytes;                         // 128 rows x K=32 bf16
    static constexpr int kBBytes = kC1BBytes;                                 // 64 couts x K=32 bf16
    static constexpr int kBiasBytes = 64 * 4;
    static constexpr int kCtlBytes = 32;                                      // mbarrier + TMEM slot
    static constexpr uint32_t kLbo = kC1Lbo, kSbo = kC1Sbo;                   // operand tile layout of conv1_dev.cuh
    static constexpr int kTotal = S::kTotal + kTileBytes + kBBytes + kBiasBytes + kCtlBytes;
    static_assert(4 * kATileBytes <= S::kBufBytes, "A tiles overlay the FFT buffers");
};

__global__ void __launch_bounds__(kFrontWarps * 32, 2) fadb_vggish_front_conv1_tc_kernel(
    const FrontParams p, const float* __restrict__ conv_w, const float* __restrict__ conv_b, int patches_per_clip,
    int total_patches, __nv_bfloat16* __restrict__ out_hi, __nv_bfloat16* __restrict__ out_lo, uint8_t* __restrict__ out8p,
    int f16, int* err_flag, int dbg) {
    constexpr int NF = 512, M = 256;
    using S = FrontSmem<NF>;
    using F = FusedTcSmem;
    constexpr uint32_t lbo = F::kLbo, sbo = F::kSbo;
    extern __shared__ __align__(16) uint8_t fsm[];
    uint8_t* s_bufb = fsm;
    double2* s_buf = reinterpret_cast<double2*>(s_bufb);
    float* s_spec = reinterpret_cast<float*>(s_bufb + S::kBufBytes);
    float* s_wt = s_spec + S::kSpecBytes / 4;
    uint32_t (*s_tile)[F::kTilePitch] = reinterpret_cast<uint32_t (*)[F::kTilePitch]>(fsm + S::kTotal);
    uint8_t* s_b = fsm + S::kTotal + F::kTileBytes;                                            // B operand tile
    float* s_bias = reinterpret_cast<float*>(s_b + F::kBBytes);
    uint64_t* bar = reinterpret_cast<uint64_t*>(s_b + F::kBBytes + F::kBiasBytes);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar + 1);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    for (int i = threadIdx.x; i < p.wt_rows * 32; i += kFrontWarps * 32) s_wt[i] = p.band_wt[i];
    // zero halo: rows 0 and 97, columns 0 and 65 (tile row r holds frame r-1, column c holds mel c-1)
    for (int i = threadIdx.x; i < 2 * F::kTilePitch; i += kFrontWarps * 32)
        s_tile[(i / F::kTilePitch) * 97][i % F::kTilePitch] = 0u;
    for (int i = threadIdx.x; i < 98; i += kFrontWarps * 32) { s_tile[i][0] = 0u; s_tile[i][65] = 0u; }
    if (threadIdx.x < 64) s_bias[threadIdx.x] = conv_b[threadIdx.x];
    c1tc_build_b(s_b, conv_w, threadIdx.x);              // B operand [64 couts][K=32] (conv1_dev.cuh)
    if (threadIdx.x == 0) {
        mbar_init(smem_u32(bar), 1);
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc(smem_u32(tmem_slot), 256);       // 4 window positions x 64 fp32 columns
    fence_proxy_async();                                        // B tile (generic-proxy stores) -> visible to the MMA
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    double2* x = s_buf + warp * S::kBufSlots;
    float* spec = s_spec + warp * S::kSpecPitch;
    int bst[2], bln[2];
#pragma unroll
    for (int h2 = 0; h2 < 2; ++h2) {
        bst[h2] = p.band_start[lane + 32 * h2];
        bln[h2] = p.band_len[lane + 32 * h2];
    }
    LaneConsts<M> lc;
    build_lane_consts<M, NF>(lc, p.tw, p.win_len, lane);
    // phase-2 roles: builder thread = (window w, dy), both dx; drain warp = (TMEM lane quarter, channel half)
    const int bw = threadIdx.x & 127, bdy = threadIdx.x >> 7;
    const uint32_t a_row_off = (uint32_t)(bw >> 3) * sbo + (uint32_t)(bw & 7) * 16;
    const int quarter = warp & 3, half = warp >> 2;
    const uint32_t a_base = smem_u32(s_bufb);
    const uint64_t db = make_nosw_desc(smem_u32(s_b), lbo, sbo);
    constexpr uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | (uint32_t(64 >> 3) << 17) | (uint32_t(128 >> 4) << 24);

#pragma unroll 1
    for (int pi = blockIdx.x; pi < total_patches; pi += gridDim.x) {
        const int clip = pi / patches_per_clip, patch = pi - clip * patches_per_clip;
        const PcmView pcm = clip_view(p.pcm, p.pcm_i16, clip, p.pcm_stride);
        {   // start pulling the NEXT patch of this CTA (62 KB of PCM) into L2 now; it is read one patch period later
            const int pn = pi + gridDim.x;
            if (pn < total_patches) {
                const int c2 = pn / patches_per_clip, p2 = pn - c2 * patches_per_clip;
                const PcmView nx = clip_view(p.pcm, p.pcm_i16, c2, p.pcm_stride);
                const long long first = (long long)p2 * 96 * p.hop, count = 95LL * p.hop + p.win_len;
                const int per_line = 128 / nx.sample_bytes();
                for (long long o = (long long)threadIdx.x * per_line; o < count; o += (long long)kFrontWarps * 32 * per_line)
                    if (first + o < p.n_samples) prefetch_l2(nx.addr(first + o));
            }
        }
        // ---- phase 1: 96 frames of this patch -> s_tile rows 1..96, columns 1..64 as {hi, lo} bf16 pairs
        for (int fi = 0; fi < ((dbg & 1) ? 0 : 12); ++fi) {
            const int fr = warp * 12 + fi;
            frame_logmel<NF, true>(p, pcm, patch * 96 + fr, lane, x, spec, lc, s_wt, bst, bln,
                                   reinterpret_cast<float*>(&s_tile[fr + 1][1]));
        }
        __syncthreads();

        // ---- phase 2: conv1 + bias + ReLU + maxpool on the tensor core, 4 pooled rows per step
#pragma unroll 1
        for (int s = (dbg & 2) ? 12 : 0; s < 12; ++s) {
            {
                const int y = 2 * (4 * s + (bw >> 5)) + bdy;        // image row; tile row y + ky = image row y - 1 + ky
                const int xx = 2 * (bw & 31);                        // image x of dx = 0; tile col xx + kx = image x - 1 + kx
                uint32_t t[3][4];
#pragma unroll
                for (int ky = 0; ky < 3; ++ky) {
                    const uint2 lo2 = *reinterpret_cast<const uint2*>(&s_tile[y + ky][xx]);
                    const uint2 hi2 = *reinterpret_cast<const uint2*>(&s_tile[y + ky][xx + 2]);
                    t[ky][0] = lo2.x; t[ky][1] = lo2.y; t[ky][2] = hi2.x; t[ky][3] = hi2.y;
                }
#pragma unroll
                for (int dx = 0; dx < 2; ++dx) {
                    uint32_t w[9];
#pragma unroll
                    for (int ky = 0; ky < 3; ++ky)
#pragma unroll
                        for (int kx = 0; kx < 3; ++kx) w[ky * 3 + kx] = t[ky][dx + kx];
                    c1tc_store_a_row(s_bufb + (bdy * 2 + dx) * F::kATileBytes + a_row_off, w);
                }
                fence_proxy_async();
            }
            tc_fence_before();
            __syncthreads();
            if (warp == 0) {
                if (elect_one()) {
                    tc_fence_after();
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        const uint64_t da = make_nosw_desc(a_base + q * F::kATileBytes, lbo, sbo);
                        umma_bf16(tmem_base + q * 64, da, db, idesc, 0u);
                        umma_bf16(tmem_base + q * 64, da + ((2 * lbo) >> 4), db + ((2 * lbo) >> 4), idesc, 1u);   // K 16..31
                    }
                    umma_commit(smem_u32(bar));
                }
                __syncwarp();
            }
            mbar_wait(smem_u32(bar), (uint32_t)s & 1u, err_flag);
            tc_fence_after();
            {
                const int prow = 4 * s + quarter;
                const size_t obase = (((size_t)pi * 48 + prow) * 32 + lane) * 64 + half * 32;
                const uint32_t taddr = tmem_base + (uint32_t(quarter * 32) << 16) + half * 32;
                uint4 q8 = make_uint4(0u, 0u, 0u, 0u);
#pragma unroll
                for (int sub = 0; sub < 2; ++sub) {
                    uint32_t r0[16], r1[16], r2[16], r3[16];
                    tmem_ld_32x32b_x16(taddr + sub * 16, r0);
                    tmem_ld_32x32b_x16(taddr + 64 + sub * 16, r1);
                    tmem_ld_32x32b_x16(taddr + 128 + sub * 16, r2);
                    tmem_ld_32x32b_x16(taddr + 192 + sub * 16, r3);
                    tmem_ld_wait();
                    float g[16];
#pragma unroll
                    for (int j = 0; j < 16; ++j) {      // relu(max(.) + b) == max(relu(. + b))
                        const float m = fmaxf(fmaxf(__uint_as_float(r0[j]), __uint_as_float(r1[j])),
                                              fmaxf(__uint_as_float(r2[j]), __uint_as_float(r3[j])));
                        g[j] = fmaxf(m + s_bias[half * 32 + sub * 16 + j], 0.f);
                    }
                    uint4 h0, h1;
                    h0.x = pack_act2(g[0], g[1], f16); h0.y = pack_act2(g[2], g[3], f16);
                    h0.z = pack_act2(g[4], g[5], f16); h0.w = pack_act2(g[6], g[7], f16);
                    h1.x = pack_act2(g[8], g[9], f16); h1.y = pack_act2(g[10], g[11], f16);
                    h1.z = pack_act2(g[12], g[13], f16); h1.w = pack_act2(g[14], g[15], f16);
                    st_global_256(out_hi + obase + sub * 16, h0, h1);
                    if (out8p) {    // e4m3 copy in the W-padded layout [P][48][34][64] (gemm_tc.cu GemmParams::c64)
                        uint4 q;
                        q.x = pack4_e4m3(g[0], g[1], g[2], g[3]);    q.y = pack4_e4m3(g[4], g[5], g[6], g[7]);
                        q.z = pack4_e4m3(g[8], g[9], g[10], g[11]);  q.w = pack4_e4m3(g[12], g[13], g[14], g[15]);
                        if (sub == 0) {
                            q8 = q;                 // both 16-channel groups of this thread go out as one 32-byte store
                        } else {
                            uint8_t* o8 = out8p + (((size_t)pi * 48 + prow) * 34 + lane + 1) * 64 + half * 32;
                            const uint4 z = make_uint4(0u, 0u, 0u, 0u);
                            st_global_256(o8, q8, q);
                            if (lane == 0) st_global_256(o8 - 64, z, z);        // zero column 0
                            if (lane == 31) st_global_256(o8 + 64, z, z);       // zero column 33
                        }
                    }
                    if (out_lo) {
#pragma unroll
                        for (int j = 0; j < 16; ++j) g[j] -= bf16_round(g[j]);
                        h0.x = pack_bf16x2(g[0], g[1]); h0.y = pack_bf16x2(g[2], g[3]);
                        h0.z = pack_bf16x2(g[4], g[5]); h0.w = pack_bf16x2(g[6], g[7]);
                        h1.x = pack_bf16x2(g[8], g[9]); h1.y = pack_bf16x2(g[10], g[11]);
                        h1.z = pack_bf16x2(g[12], g[13]); h1.w = pack_bf16x2(g[14], g[15]);
                        st_global_256(out_lo + obase + sub * 16, h0, h1);
                    }
                }
            }
            tc_fence_before();      // orders these TMEM reads before the next step's MMAs (via its __syncthreads)
        }
        __syncthreads();            // phase 1 of the next patch reuses the A-tile bytes and the patch tile
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 256);
    }
}

// ------------------------------------------------------------------------------------------------
// Host: tables
// ------------------------------------------------------------------------------------------------
static const double kPi = 3.14159265358979323846;

static void vggish_mel(int nbins, std::vector<std::vector<double>>& wcols) {
    // models/vggish.py:150-190 : HTK mel, 64 bands 125..7500 Hz over linspace(0, 8000, 257), DC row zeroed
    auto mel = [](double hz) { return 1127.0 * std::log(1.0 + hz / 700.0); };
    const int nb = 64;
    std::vector<double> binmel(nbins);
    for (int k = 0; k < nbins; ++k) binmel[k] = mel(8000.0 * k / (nbins - 1));
    const double lo = mel(125.0), hi = mel(7500.0);
    std::vector<double> edges(nb + 2);
    for (int i = 0; i < nb + 2; ++i) edges[i] = lo + (hi - lo) * i / (nb + 1);
    wcols.assign(nb, std::vector<double>(nbins, 0.0));
    for (int b = 0; b < nb; ++b) {
        const double l = edges[b], c = edges[b + 1], u = edges[b + 2];
        for (int k = 1; k < nbins; ++k) {
            const double up = (binmel[k] - l) / (c - l), dn = (u - binmel[k]) / (u - c);
            const double w = std::fmax(0.0, std::fmin(up, dn));
            wcols[b][k] = w;
        }
    }
}

static void slaney_mel(int sr, int nfft, double fmin, double fmax, std::vector<std::vector<double>>& wcols) {
    // semantics of librosa.filters.mel(htk=False, norm='slaney') as called at models/pann.py:121-127
    const double f_sp = 200.0 / 3.0, min_log_hz = 1000.0, min_log_mel = min_log_hz / f_sp;
    const double logstep = std::log(6.4) / 27.0;
    auto h2m = [&](double hz) { return hz >= min_log_hz ? min_log_mel + std::log(hz / min_log_hz) / logstep : hz / f_sp; };
    auto m2h = [&](double m) { return m >= min_log_mel ? min_log_hz * std::exp(logstep * (m - min_log_mel)) : f_sp * m; };
    const int nb = 64, nbins = nfft / 2 + 1;
    std::vector<double> f(nb + 2);
    const double ml = h2m(fmin), mh = h2m(fmax);
    for (int i = 0; i < nb + 2; ++i) f[i] = m2h(ml + (mh - ml) * i / (nb + 1));
    wcols.assign(nb, std::vector<double>(nbins, 0.0));
    for (int b = 0; b < nb; ++b) {
        const double enorm = 2.0 / (f[b + 2] - f[b]);
        for (int k = 0; k < nbins; ++k) {
            const double fk = (double)k * sr / nfft;
            const double lower = (fk - f[b]) / (f[b + 1] - f[b]);
            const double upper = (f[b + 2] - fk) / (f[b + 2] - f[b + 1]);
            const double w = std::fmax(0.0, std::fmin(lower, upper)) * enorm;
            wcols[b][k] = (double)(float)w;          // the reference filterbank is float32
        }
    }
}

static int build_tables(fadb_handle* h, int model) {
    FrontTables& t = h->front_tables[model];
    if (t.ready) return FADB_OK;
    int sr = 16000;
    double fmin = 50, fmax = 8000;
    switch (model) {
        case FADB_MODEL_VGGISH: t.nfft = 512; t.win_len = 400; t.hop = 160; break;
        case FADB_MODEL_PANN8K: t.nfft = 256; t.win_len = 256; t.hop = 80; sr = 8000; fmax = 4000; break;
        case FADB_MODEL_PANN16K: t.nfft = 512; t.win_len = 512; t.hop = 160; sr = 16000; fmax = 8000; break;
        case FADB_MODEL_PANN32K: t.nfft = 1024; t.win_len = 1024; t.hop = 320; sr = 32000; fmax = 14000; break;
        case FADB_MODEL_CLAP: t.nfft = 1024; t.win_len = 1024; t.hop = 480; sr = 48000; fmax = 14000; break;
        default: set_error("unknown model %d", model); return FADB_E_INVALID;
    }
    std::vector<double2> tw(t.nfft);
    for (int k = 0; k < t.nfft; ++k) {
        const double a = -2.0 * kPi * k / t.nfft;
        tw[k] = make_double2(std::cos(a), std::sin(a));
    }
    std::vector<double> win(t.win_len);
    for (int n = 0; n < t.win_len; ++n) win[n] = 0.5 - 0.5 * std::cos(2.0 * kPi / t.win_len * n);
    std::vector<std::vector<double>> wc;
    if (model == FADB_MODEL_VGGISH) vggish_mel(t.nfft / 2 + 1, wc);
    else slaney_mel(sr, t.nfft, fmin, fmax, wc);
    std::vector<int> bs(64), bl(64);
    for (int b = 0; b < 64; ++b) {
        int first = -1, last = -1;
        for (int k = 0; k < (int)wc[b].size(); ++k)
            if (wc[b][k] != 0.0) { if (first < 0) first = k; last = k; }
        if (first < 0) { first = 0; last = -1; }
        bs[b] = first; bl[b] = last - first + 1;
    }
    // transposed weights: row (h2 ? max0 : 0) + i, column lane = weight i of band lane + 32*h2 (0 past the band's end)
    int mx[2] = {0, 0};
    for (int b = 0; b < 64; ++b) mx[b >> 5] = std::max(mx[b >> 5], bl[b]);
    t.wt_off1 = mx[0];
    t.wt_rows = std::max(mx[0] + mx[1], 1);
    FADB_REQUIRE(t.wt_rows <= t.nfft / 16, "mel bands too wide for the shared-memory weight table (%d rows)", t.wt_rows);
    std::vector<float> wt((size_t)t.wt_rows * 32, 0.f);
    for (int b = 0; b < 64; ++b)
        for (int i = 0; i < bl[b]; ++i) wt[(size_t)((b >> 5) * mx[0] + i) * 32 + (b & 31)] = (float)wc[b][bs[b] + i];
    FADB_CUDA_CHECK(cudaMalloc(&t.tw, tw.size() * sizeof(double2)));
    FADB_CUDA_CHECK(cudaMalloc(&t.win, win.size() * sizeof(double)));
    FADB_CUDA_CHECK(cudaMalloc(&t.band_start, 64 * sizeof(int)));
    FADB_CUDA_CHECK(cudaMalloc(&t.band_len, 64 * sizeof(int)));
    FADB_CUDA_CHECK(cudaMalloc(&t.band_wt, wt.size() * sizeof(float)));
    FADB_CUDA_CHECK(cudaMemcpy(t.tw, tw.data(), tw.size() * sizeof(double2), cudaMemcpyHostToDevice));
    FADB_CUDA_CHECK(cudaMemcpy(t.win, win.data(), win.size() * sizeof(double), cudaMemcpyHostToDevice));
    FADB_CUDA_CHECK(cudaMemcpy(t.band_start, bs.data(), 64 * sizeof(int), cudaMemcpyHostToDevice));
    FADB_CUDA_CHECK(cudaMemcpy(t.band_len, bl.data(), 64 * sizeof(int), cudaMemcpyHostToDevice));
    FADB_CUDA_CHECK(cudaMemcpy(t.band_wt, wt.data(), wt.size() * sizeof(float), cudaMemcpyHostToDevice));
    t.ready = true;
    return FADB_OK;
}

template <int NF>
static constexpr int front_smem() { return FrontSmem<NF>::kTotal; }

int frontend_init(fadb_handle* h) {
    (void)h;
    FADB_CUDA_CHECK(cudaFuncSetAttribute(fadb_frontend_kernel<256>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         front_smem<256>()));
    FADB_CUDA_CHECK(cudaFuncSetAttribute(fadb_frontend_kernel<512>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         front_smem<512>()));
    FADB_CUDA_CHECK(cudaFuncSetAttribute(fadb_frontend_kernel<1024>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         front_smem<1024>()));
    FADB_CUDA_CHECK(cudaFuncSetAttribute(fadb_vggish_front_conv1_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         FusedTcSmem::kTotal));
    return FADB_OK;
}

void frontend_release(fadb_handle* h) {
    for (FrontTables& t : h->front_tables) {
        if (t.tw) cudaFree(t.tw);
        if (t.win) cudaFree(t.win);
        if (t.band_start) cudaFree(t.band_start);
        if (t.band_len) cudaFree(t.band_len);
        if (t.band_wt) cudaFree(t.band_wt);
        t = FrontTables();
    }
}

static int64_t pann_pad(int64_t t) {        // fad.py:53-59
    int64_t k = (t + 24 + 31) / 32;
    int64_t v = 32 * k - 24;
    if (v < t) v += 32;
    return v;
}

// rows (patches for VGGish, frames T' otherwise) per clip
int64_t frontend_rows(int model, int64_t n) {
    switch (model) {
        case FADB_MODEL_VGGISH: {
            if (n < 400) return 0;
            const int64_t f = 1 + (n - 400) / 160;
            return f < 96 ? 0 : 1 + (f - 96) / 96;
        }
        case FADB_MODEL_PANN8K: return pann_pad(1 + n / 80);
        case FADB_MODEL_PANN16K: return pann_pad(1 + n / 160);
        case FADB_MODEL_PANN32K: return pann_pad(1 + n / 320);
        case FADB_MODEL_CLAP: return 1001;
        default: return -1;
    }
}


// per-launch CUDA events of the front-end kernels when the profile hook is on (bench.py's front-end GB/s figure)
struct FrontProfile {
    fadb_handle* h;
    cudaStream_t st;
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    FrontProfile(fadb_handle* h_, cudaStream_t st_) : h(h_), st(st_) {
        if (h->profile && cudaEventCreate(&e0) == cudaSuccess && cudaEventCreate(&e1) == cudaSuccess) cudaEventRecord(e0, st);
    }
    ~FrontProfile() {
        if (e0 && e1) {
            cudaEventRecord(e1, st);
            h->prof_front.push_back(e0);
            h->prof_front.push_back(e1);
        }
    }
};

// PCM -> conv1 output [n_clips * patches, 48, 32, 64] bf16 (hi / optional lo) in one kernel (VGGish only)
int launch_vggish_front_conv1(fadb_handle* h, PcmSrc pcm, int64_t n_clips, int64_t n_samples, int64_t pcm_stride,
                              __nv_bfloat16* out_hi, __nv_bfloat16* out_lo, uint8_t* out8p, cudaStream_t st) {
    FADB_CHECK(build_tables(h, FADB_MODEL_VGGISH));
    const FrontTables& t = h->front_tables[FADB_MODEL_VGGISH];
    const int64_t patches = frontend_rows(FADB_MODEL_VGGISH, n_samples);
    if (n_clips <= 0 || patches <= 0) return FADB_OK;
    FADB_REQUIRE(n_clips <= 65535 && n_samples < (1LL << 30), "fused front end: clip count / length out of range");
    FrontParams p;
    p.pcm = pcm.ptr; p.pcm_i16 = pcm.i16; p.pcm_stride = pcm_stride;
    p.n_samples = (int)n_samples; p.logical_len = (int)n_samples;
    p.hop = t.hop; p.win_len = t.win_len;
    p.centered = 0; p.power_db = 0; p.quantize = 0;
    p.tw = t.tw; p.win = t.win;
    p.band_start = t.band_start; p.band_len = t.band_len;
    p.band_wt = t.band_wt; p.wt_rows = t.wt_rows; p.wt_off1 = t.wt_off1;
    p.win_rot_c = std::cos(2.0 * kPi * 64.0 / t.win_len); p.win_rot_s = std::sin(2.0 * kPi * 64.0 / t.win_len);
    p.out = nullptr;
    p.rows_out = (int)(patches * 96);
    p.frames_valid = p.rows_out;
    dim3 grid((unsigned)patches, (unsigned)n_clips);
    FrontProfile prof(h, st);
    static const int front_dbg = getenv("FADB_FRONT_DBG") ? atoi(getenv("FADB_FRONT_DBG")) : 0;   // profiling only: 1 / 2 skip a phase
    __nv_bfloat16* lo = h->precision == FADB_PREC_BF16X3 ? out_lo : nullptr;
    const int64_t total = patches * n_clips;
    const int64_t slots = 2LL * h->sm_count;
    fadb_vggish_front_conv1_tc_kernel<<<(unsigned)(total < slots ? total : slots), kFrontWarps * 32, FusedTcSmem::kTotal, st>>>(
        p, h->conv1_w, h->conv1_b, (int)patches, (int)total, out_hi, lo, out8p, (int)prec_is_f16(h->precision), h->err_flag, front_dbg);
    h->launches++;
    FADB_CUDA_CHECK(cudaGetLastError());
    return FADB_OK;
}

int launch_frontend(fadb_handle* h, int model, PcmSrc pcm, int64_t n_clips, int64_t n_samples,
                    int64_t pcm_stride, float* feats, cudaStream_t st) {
    FADB_REQUIRE(model >= 0 && model <= 4, "unknown model %d", model);
    FADB_REQUIRE(n_samples > 0 && n_samples < (1LL << 30), "n_samples out of range");
    FADB_CHECK(build_tables(h, model));
    const FrontTables& t = h->front_tables[model];
    if (n_clips <= 0) return FADB_OK;
    FrontParams p;
    p.pcm = pcm.ptr;
    p.pcm_i16 = pcm.i16;
    p.pcm_stride = pcm_stride;
    p.n_samples = (int)n_samples;
    p.logical_len = (int)n_samples;
    p.hop = t.hop;
    p.win_len = t.win_len;
    p.centered = (model != FADB_MODEL_VGGISH);
    p.power_db = (model != FADB_MODEL_VGGISH);
    p.quantize = (model == FADB_MODEL_CLAP) && h->clap_quantize;
    p.tw = t.tw; p.win = t.win;
    p.band_start = t.band_start; p.band_len = t.band_len;
    p.band_wt = t.band_wt; p.wt_rows = t.wt_rows; p.wt_off1 = t.wt_off1;
    p.win_rot_c = std::cos(2.0 * kPi * 64.0 / t.win_len); p.win_rot_s = std::sin(2.0 * kPi * 64.0 / t.win_len);
    p.out = feats;
    if (model == FADB_MODEL_VGGISH) {
        const int64_t patches = frontend_rows(model, n_samples);
        if (patches <= 0) return FADB_OK;
        p.rows_out = (int)(patches * 96);
        p.frames_valid = p.rows_out;
    } else if (model == FADB_MODEL_CLAP) {
        // fad.py:356-359 zero-pads the waveform to 480000 first; a longer clip keeps its samples and only the first
        // 1001 frames survive (_pad_to_clap_time truncates, fad.py:87-89)
        p.logical_len = n_samples > 480000 ? (int)n_samples : 480000;
        p.rows_out = 1001;
        p.frames_valid = 1001;
    } else {
        FADB_REQUIRE(n_samples > t.nfft / 2, "clip shorter than n_fft/2 cannot be reflect-padded");
        p.frames_valid = (int)(1 + n_samples / t.hop);
        p.rows_out = (int)pann_pad(p.frames_valid);
    }
    FADB_REQUIRE(n_clips <= 65535, "front end: at most 65535 clips per call");
    dim3 grid((p.rows_out + 63) / 64, (unsigned)n_clips);
    FrontProfile prof(h, st);
    if (t.nfft == 256) fadb_frontend_kernel<256><<<grid, 256, front_smem<256>(), st>>>(p);
    else if (t.nfft == 512) fadb_frontend_kernel<512><<<grid, 256, front_smem<512>(), st>>>(p);
    else fadb_frontend_kernel<1024><<<grid, 256, front_smem<1024>(), st>>>(p);
    h->launches++;
    FADB_CUDA_CHECK(cudaGetLastError());
    return FADB_OK;
}

}  // namespace fadb
