// gemm_tc.cu — the tensor-core layer of the FAD hot path: 3x3 convolution (pad 1) and linear layers
// as ONE persistent, warp-specialised tcgen05 implicit-GEMM kernel for sm_100a.
//
// Replaces the cuDNN/cuBLAS calls the reference reaches through nn.Conv2d / nn.Linear
// (models/vggish.py:44-51,71-78; models/pann.py:161-176,234) — SURVEY.md §2.2 K7/K8/K9.
//
//   D[pixel, cout] = sum_{tap, cin} A[pixel + tap offset, cin] * Wt[cout, tap, cin]
//
//   * A operand: NHWC bf16 activations.  A 4-D TMA tensor map (C, W, H, B) with box
//     (64 ch, BW, BH, BB), BW*BH*BB = 128, lands one 128-pixel x 64-channel K-major SWIZZLE_128B
//     tile per (tap, channel block); the 3x3 halo and the zero padding come for free from TMA
//     out-of-bounds zero fill (coordinates start at x0-1 / y0-1).  No im2col buffer exists.
//   * B operand: packed weights [Cout][tap][Cin] bf16 (K-major), 2-D tensor map, box (64, BN).
//   * MMA: tcgen05.mma.kind::f16, N = BN (64/128/256), K = 16, fp32 accumulators in TMEM, double-buffered
//     (2 x BN columns) so the epilogue of tile i overlaps the MMAs of tile i+1.  Large plain layers run as
//     CLUSTERS OF TWO CTAs: one cta_group::2 MMA (M = 256, 128 rows per SM) per K step issued by the leader,
//     each CTA holding its own A tile and half of the weight tile (PAIR instantiation; FADB_TWOCTA=0 selects
//     the older variant: two M = 128 MMAs with the weight tile multicast).  Everything else is cta_group::1.
//   * Halo mode (3x3 layers on maps >= 8 wide, >= 16 high): ONE 10 x 18-pixel activation box per channel block
//     serves all nine taps as shifted descriptor views; weights stream through their own ring (warp 3) or stay
//     resident in shared memory.
//   * split-bf16 ("bf16x3") mode: the K loop runs three passes (A_hi*B_hi, A_lo*B_hi, A_hi*B_lo)
//     into the same accumulator, giving ~2^-16 relative operand precision with the same kernel.
//   * Epilogue (8 warps): tcgen05.ld -> +bias (folded BN, staged in shared memory) -> ReLU -> optional 2x2 max /
//     avg pool as a channel-splitting shuffle butterfly (the pooled layers use boxes <= 16 pixels wide, so a
//     pooling window lives inside one warp) -> bf16 hi(/lo) or fp32 NHWC, 16-byte stores straight from registers.
//
// Warp roles (384 threads): warp 0 = TMA producer, warp 1 = MMA issuer, warp 2 = TMEM allocator, warp 3 = weight
// producer in halo mode, warps 4..11 = epilogue (TMEM lane quarter = warp % 4, two warps per quarter).
#include <cstdarg>
#include <cstdlib>
#include <cstring>
#include <type_traits>

#include "common.cuh"
#include "tc_ptx.cuh"

namespace fadb {

// ------------------------------------------------------------------------------------------------
// Kernel parameters (tensor maps live in the parameter/constant bank: __grid_constant__)
// ------------------------------------------------------------------------------------------------
struct GemmParams {
    CUtensorMap tmA[3];   // activations hi, lo, e4m3 : dims (C, W, H, B)
    CUtensorMap tmB[3];   // weights hi, lo, e4m3-lo  : dims (Ktot, N)
    CUtensorMap tmH;      // halo mode: activations hi with box (64 ch, 10 px, 18 rows, 1 image)
    CUtensorMap tmH8;     // halo mode, e4m3 activations: box (128 ch, 10 px, 18 rows, 1 image)
    CUtensorMap tmBh[3];  // cluster mode: weights hi, lo, e4m3-lo with box (., BN/2): each CTA of a pair loads half (and multicasts it)
    int W, H, B;
    int BW, BH, BB;       // box; BW*BH*BB == 128
    int tiles_w, tiles_h, tiles_b, tiles_n;
    int num_tiles;
    int cluster;          // 1, or 2 / 4 = clusters of CTAs walk groups of M tiles of the same N tile and share the weight loads
    int twocta;           // cluster == 2 only: 1 = ONE tcgen05.mma.cta_group::2 (M = 256) per K step for the CTA pair, issued by
                          // rank 0; each CTA keeps A (its 128 rows) and half of the weight tile, nothing is duplicated
    int num_units;        // work units of the static schedule: tiles, or (M-tile group, N tile) in cluster mode
    int cin_blocks;       // Cin / 64
    int taps;             // 9 or 1
    int npass;            // operand passes over K: 1 = A*B (bf16 / fp16); 2 = A*B_hi, A*B_lo (fp16x2: split weights);
                          // 3 = A_hi*B_hi, A_lo*B_hi, A_hi*B_lo (bf16x3)
    int N;                // Cout
    int stages;           // smem pipeline depth (runtime: sized from the handle's smem budget)
    int nkb;              // K blocks (128 bytes of K per row) per tile: npass * taps * cin_blocks, or kb0 + taps * cin_blocks1
    int lo8;              // fp16x2 with the LOW-ORDER weight pass in e4m3: pass 1 multiplies e4m3(activations) by
                          // e4m3((W - fp16(W)) * 2^s) with kind::f8f6f4 MMAs (twice the fp16 rate) over 128-channel K blocks;
                          // its segment sums are scaled by lo_scale = 2^-s when the epilogue adds them in
    int kb0;              // K blocks of pass 0 (taps * cin_blocks)
    int cin_blocks1;      // channel blocks of pass 1: Cin / 128 when lo8
    int c64;              // lo8 on a 64-CHANNEL 3x3 layer (halo mode only): tmH8 views the W-padded e4m3 tensor
                          // [B][H][W+2][64] with OVERLAPPING 128-byte rows [pixel x | pixel x+1], so one 128-byte K block is
                          // the taps (ky,0),(ky,1) of a filter row and a second block, half used, is the tap (ky,2):
                          // 6 weight K blocks and 18 e4m3 MMAs per tile instead of 9 blocks and 36 fp16 MMAs
    int nseg0;            // accumulation segments of pass 0 (lo8: segments never straddle the pass boundary)
    int seg_len1;         // segment length of pass 1
    float lo_scale;
    uint8_t* out8;        // e4m3 copy of the output (the next layer's pass-1 activations), or nullptr
    int out8_wpad;        // 1: out8 is [B][Ho][Wo+2][N] with a zero column left and right (input layout of a c64 layer)
    int nseg, seg_len;    // accumulation segments per tile and their length (K blocks; halo mode: channel blocks).
                          // Each segment is its own accumulation chain in tensor memory (the two TMEM accumulators
                          // alternate per SEGMENT); the epilogue warps add the segment sums in fp32 registers with
                          // round-to-nearest.  The fp32 accumulation inside tcgen05.mma truncates, a bias that grows with
                          // the chain length (5e-6 of the output at K = 4608) and shrinks every activation of every
                          // layer the same way: cutting the chains to <= 64 MMAs removes it (FAD: 2e-4 -> 1e-5).
    int upper_only;       // syrk: work units whose N tile lies strictly below the diagonal of their M rows are skipped
    int halo;             // 1 = halo mode: one activation tile per channel block feeds all 9 taps
    int b_stages;         // halo mode: depth of the separate B ring
    int resb;             // halo mode: 1 = all weights of the (single) N tile stay resident in smem
    int relu;
    int pool;             // 0 none, 1 max, 2 avg
    int Ho, Wo;           // output spatial dims (after pooling)
    const float* bias;    // [N] or nullptr
    __nv_bfloat16* out_hi;
    __nv_bfloat16* out_lo;
    float* out_f32;
    double* out_f64;      // syrk: the fp64 statistic itself; the epilogue ADDS the tile (read-modify-write, one owner per tile)
    int* err_flag;
};

constexpr int kThreads = 384;                     // 4 control warps + 8 epilogue warps
constexpr int kTileM = 128;
constexpr int kBlockK = 64;                       // one 128-byte swizzle atom of bf16
constexpr int kABytes = kTileM * kBlockK * 2;     // 16384
constexpr int kHaloW = 10;                        // halo tile: (8 + 2) pixels x 18 rows x 64 channels bf16
constexpr int kHaloBoxBytes = 18 * kHaloW * 128;  // 23040 B land per TMA box
constexpr int kHaloBytes = (kHaloBoxBytes + 1023) / 1024 * 1024;   // stage pitch: keeps every stage 1024-B aligned

template <int BN>
struct GemmCfg {
    static constexpr int kBBytes = BN * kBlockK * 2;
    static constexpr int kStageBytes = kABytes + kBBytes;
    static constexpr int kMaxStages = 8;
    // accumulators in tensor memory: 2 for BN = 256, 4 for BN <= 128.  With one accumulation segment per <= 64 MMAs the
    // MMA warp may run up to kAcc - 1 segments ahead of the epilogue's TMEM reads (measured on the statistics GEMM, 16-MMA
    // segments: the commit -> read -> hand-back round trip is ~4000 cycles, four times the MMAs of such a segment)
    static constexpr int kAcc = BN == 256 ? 2 : 4;
    static constexpr int kTmemCols = kAcc * BN;   // 256 / 512 / 512: powers of two >= 32
    static constexpr int kExtraBytes = 512 /*barriers*/ + 1024 /*alignment slack*/;
    static constexpr int kMaxSmemBytes = 232448;                      // 227 KB: the per-CTA opt-in maximum
    static int stages_for(int budget_bytes) {
        int s = budget_bytes / kStageBytes;
        if (s > kMaxStages) s = kMaxStages;
        if (s < 2) s = 2;
        return s;
    }
};

// ------------------------------------------------------------------------------------------------
// The kernel
// ------------------------------------------------------------------------------------------------
// PAIR = the cta_group::2 variant (clusters of two only).  It is a separate instantiation because a kernel that contains
// cta_group::2 instructions can only be launched with an even cluster size ("cluster misconfiguration" otherwise).
// F16 = activations and weights are IEEE fp16 instead of bf16 (same kind::f16 MMA at the same rate; 8x smaller operand
// rounding error, range +-65504 with saturation in the epilogue).
// SYRK = the statistics variant (stats.cu): the finished tile (fp32 segment sums of one short row chunk) is ADDED to the
// fp64 second-moment matrix; no bias / activation / pooling.  (Adding every segment in fp64 registers was measured:
// the 64 F2F conversions per thread and segment cost more than the MMAs of a 16-MMA segment.)
// LO8 = fp16x2 with the e4m3 low-order pass (GemmParams::lo8).  A template parameter so that the single-pass and bf16x3
// instantiations carry none of its branches: with a run-time flag alone they ran 8 % (VGGish) to 14 % (CNN14) slower.
template <int BN, bool PAIR, bool F16, bool SYRK = false, bool LO8 = false>
__global__ void __launch_bounds__(kThreads, 1) fadb_gemm_tc_kernel(const __grid_constant__ GemmParams p) {
    using Cfg = GemmCfg<BN>;
    const int kStages = p.stages;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw_addr = smem_u32(smem_raw);
    const uint32_t base = (raw_addr + 1023u) & ~1023u;               // SWIZZLE_128B tiles need 1024-B alignment
    uint8_t* smem = smem_raw + (base - raw_addr);

    // layout: [resident B: nkb x kBBytes (resb only)] [stages] ([halo mode: B ring]) [barriers]
    constexpr int kBTile = PAIR ? Cfg::kBBytes / 2 : Cfg::kBBytes;   // bytes of one weight K block held by THIS CTA
    const int res_bytes = p.resb ? p.nkb * kBTile : 0;
    const int stage_pitch = p.halo ? kHaloBytes : (PAIR ? kABytes + Cfg::kBBytes / 2 : Cfg::kStageBytes);
    const uint32_t stage_base = base + res_bytes;
    const uint32_t bring_base = stage_base + kStages * stage_pitch;           // halo mode only
    const int bring_bytes = (p.halo && !p.resb) ? p.b_stages * kBTile : 0;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + res_bytes + kStages * stage_pitch + bring_bytes);
    const uint32_t bar_full = smem_u32(bars);                        // [kStages]
    const uint32_t bar_empty = bar_full + 8 * kStages;               // [kStages]
    const uint32_t bar_tfull = bar_empty + 8 * kStages;              // [4]
    const uint32_t bar_tempty = bar_tfull + 32;                      // [4]
    const uint32_t bar_bres = bar_tempty + 32;                       // [1]
    const uint32_t bar_bfull = bar_bres + 8;                         // [8]  halo mode B ring
    const uint32_t bar_bempty = bar_bfull + 64;                      // [8]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * kStages + 9 + 16);
    // the layer's whole bias vector lives in shared memory (bars + 512 B): the epilogue reads it with broadcast
    // LDS instead of exposing a global-load round trip per 32-column chunk (ncu: 16 % of all stall samples sat on
    // the first FADD after the bias __ldg in the short-K layers)
    float* s_bias = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(bars) + 512);
    for (int i = threadIdx.x; i < p.N; i += kThreads) s_bias[i] = p.bias ? __ldg(p.bias + i) : 0.f;

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;

    if (warp == 0 && lane == 0) {
        prefetch_tmap(&p.tmA[0]);
        prefetch_tmap(&p.tmB[0]);
        if (p.npass > 1) {
            prefetch_tmap(&p.tmA[1]);
            prefetch_tmap(&p.tmB[1]);
        }
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < kStages; ++s) {
            mbar_init(bar_full + 8 * s, 1);
            mbar_init(bar_empty + 8 * s, PAIR ? 1 : p.cluster);  // multicast mode: both CTAs' MMA warps release a stage
        }
        for (int s = 0; s < 4; ++s) {
            mbar_init(bar_tfull + 8 * s, 1);
            mbar_init(bar_tempty + 8 * s, PAIR ? 16 : 8);        // one arrive per epilogue warp (of both CTAs in pair mode)
        }
        mbar_init(bar_bres, 1);
        for (int s = 0; s < 8; ++s) {
            mbar_init(bar_bfull + 8 * s, 1);
            mbar_init(bar_bempty + 8 * s, 1);
        }
        fence_barrier_init();
    }
    if (warp == 2) {
        if constexpr (PAIR) tmem_alloc_2sm(smem_u32(tmem_slot), Cfg::kTmemCols);
        else tmem_alloc(smem_u32(tmem_slot), Cfg::kTmemCols);
    }
    tc_fence_before();
    __syncthreads();
    if (p.cluster > 1) cluster_sync_all();       // the peer's barriers exist before anything multicasts to them
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    // Register re-allocation per warpgroup: the control warps (TMA / MMA issue) need few registers, the epilogue warps
    // hold BN/2 running fp32 sums per thread across the accumulation segments of a tile.  128 x 56 + 256 x 224 = 64512.
    // (the setmaxnreg instructions open the two role branches below)
    // Static schedule.  Plain mode: unit = tile, N tile fastest.  Cluster mode: unit = (pair of M tiles, N tile); CTA
    // rank r of the pair takes M tile 2*pair + r, so both CTAs need the SAME weight tile at every K step and each
    // loads half of it, multicast to both.  An M tile past the end (odd count) runs on zero-filled boxes and stores
    // nothing (its image index is out of range).
    const int crank = (p.cluster > 1) ? (int)cluster_ctarank() : 0;
    const int cshift = p.cluster >> 1;               // log2(cluster size) for 1, 2, 4
    const int unit0 = (int)blockIdx.x >> cshift;
    const int ustep = (int)gridDim.x >> cshift;
    const uint16_t cmask = (uint16_t)((1u << p.cluster) - 1u);
    auto unit_tile = [&](int unit, int& m, int& n0) {
        const int q = unit / p.tiles_n;
        n0 = (unit - q * p.tiles_n) * BN;
        m = (q << cshift) + crank;
    };
    // syrk (C = Y Y^T, linear-layer tiling: M tile = 128 rows): every role skips the same units, so the pipeline state
    // (stages, accumulator segments) stays in step
    auto skip_unit = [&](int unit) {
        if (!p.upper_only) return false;
        const int q = unit / p.tiles_n;
        const int nt = unit - q * p.tiles_n;
        return (nt + 1) * BN <= (q << cshift) * kTileM;
    };

    const int kb_per_pass = p.taps * p.cin_blocks;
    // which activation / weight plane a pass multiplies: npass 1: (0,0); 2: (0,0),(0,1); 3: (0,0),(1,0),(0,1)
    auto pass_a = [&](int pass) { return (p.npass == 3 && pass == 1) ? 1 : ((LO8 && pass == 1) ? 2 : 0); };
    auto pass_b = [&](int pass) { return (p.npass == 3) ? (pass == 2 ? 1 : 0) : (pass == 1 ? (LO8 ? 2 : 1) : 0); };
    // instruction descriptor: D = f32 (bit 4), A / B format bf16 = 1 or f16 = 0 (bits 7, 10), both K-major, N (bits 17..), M (bits 24..)
    constexpr uint32_t kFmtBits = F16 ? 0u : ((1u << 7) | (1u << 10));

    // pair mode (halo path): every load completes on the LEADER's barrier, which expects the bytes of both CTAs; each CTA
    // loads its own activations and its half of the weight rows; the leader alone issues and commits for both
    auto expect = [&](uint32_t bar, uint32_t bytes_per_cta) {
        if constexpr (PAIR) { if (crank == 0) mbar_arrive_expect_tx(bar, 2u * bytes_per_cta); }
        else mbar_arrive_expect_tx(bar, bytes_per_cta);
    };
    auto load_act = [&](const CUtensorMap* m, uint32_t bar, uint32_t dst, int c0, int c1, int c2, int c3) {
        if constexpr (PAIR) tma_load_4d_2sm(m, mapa_rank(bar, 0), dst, c0, c1, c2, c3);
        else tma_load_4d(m, bar, dst, c0, c1, c2, c3);
    };
    auto load_wgt = [&](uint32_t bar, uint32_t dst, int k0, int n0, int plane) {
        if constexpr (PAIR) tma_load_2d_2sm(&p.tmBh[plane], mapa_rank(bar, 0), dst, k0, n0 + crank * (BN / 2));
        else tma_load_2d(&p.tmB[plane], bar, dst, k0, n0);
    };
    auto mma = [&](uint32_t d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t acc) {
        if constexpr (PAIR) umma_bf16_2sm(d, da, db, idesc, acc);
        else umma_bf16(d, da, db, idesc, acc);
    };
    // e4m3 pass: same descriptors, same instruction-descriptor value with the format bits 0 (= E4M3), kind::f8f6f4
    auto mma8 = [&](uint32_t d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t acc) {
        if constexpr (PAIR) umma_f8_2sm(d, da, db, idesc & ~((7u << 7) | (7u << 10)), acc);
        else umma_f8(d, da, db, idesc & ~((7u << 7) | (7u << 10)), acc);
    };
    auto commit = [&](uint32_t bar) {
        if constexpr (PAIR) umma_commit_2sm(bar, (uint16_t)0x3);
        else umma_commit(bar);
    };

    if (warp < 4) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 56;");
    if (warp == 0) {
        // ======================= TMA producer (whole warp runs the loop; one elected lane issues) =====
        if (p.halo) {
            // halo mode: ONE activation load per (tile, channel block) serves all 9 taps
            int stage = 0;
            uint32_t phase = 0;
            if (p.resb && unit0 < p.num_units) {
                if (elect_one()) {
                    expect(bar_bres, (uint32_t)res_bytes);
                    for (int kbg = 0; kbg < p.nkb; ++kbg) {                  // resident layout: [plane][tap][channel block]
                        if (LO8 && kbg >= p.kb0) {                            // e4m3 lo plane: 128 channels per block
                            const int j = kbg - p.kb0;                        // c64: taps (ky,0..1) at 3 ky * 64, tap (ky,2) at (3 ky + 2) * 64
                            load_wgt(bar_bres, base + kbg * kBTile, p.c64 ? (3 * (j >> 1) + 2 * (j & 1)) * 64 : j * 128, 0, 2);
                        }
                        else {
                            const int plane = kbg / kb_per_pass;
                            load_wgt(bar_bres, base + kbg * kBTile, (kbg - plane * kb_per_pass) * kBlockK, 0, plane);
                        }
                    }
                }
                __syncwarp();
            }
            for (int unit = unit0; unit < p.num_units; unit += ustep) {
                if (skip_unit(unit)) continue;
                int m, n0;
                unit_tile(unit, m, n0);
                const int wt = m % p.tiles_w; m /= p.tiles_w;
                const int ht = m % p.tiles_h;
                const int bt = m / p.tiles_h;
                // lo8: the fp16 halo tiles of pass 0, then the e4m3 halo tiles (128 channels each) of pass 1
                const int a_tiles = p.cin_blocks + (LO8 ? p.cin_blocks1 : 0);
                for (int t = 0; t < a_tiles; ++t) {
                    const bool f8 = LO8 && t >= p.cin_blocks;
                    mbar_wait(bar_empty + 8 * stage, phase ^ 1u, p.err_flag);
                    if (elect_one()) {
                        expect(bar_full + 8 * stage, (uint32_t)kHaloBoxBytes);
                        // (c64: the e4m3 tensor carries a zero column left and right, so halo column -1 is row 0 of the view)
                        load_act(f8 ? &p.tmH8 : &p.tmH, bar_full + 8 * stage, stage_base + stage * kHaloBytes,
                                 f8 ? (t - p.cin_blocks) * 128 : t * kBlockK, wt * p.BW - 1 + ((f8 && p.c64) ? 1 : 0),
                                 ht * p.BH - 1, bt);
                    }
                    __syncwarp();
                    if (++stage == kStages) { stage = 0; phase ^= 1u; }
                }
            }
        } else {
            int stage = 0;
            uint32_t phase = 0;
            for (int unit = unit0; unit < p.num_units; unit += ustep) {
                if (skip_unit(unit)) continue;
                int m, n0;
                unit_tile(unit, m, n0);
                const int wt = m % p.tiles_w; m /= p.tiles_w;
                const int ht = m % p.tiles_h;
                const int bt = m / p.tiles_h;
                const int x0 = wt * p.BW, y0 = ht * p.BH, b0 = bt * p.BB;
                // K-block walk without integer divisions: this loop paces the whole pipeline (two runtime divisions
                // per K block in it cost 9 % of the layer time when measured), so (pass, tap, channel block) and the
                // tap offsets advance incrementally
                int pass = 0, kb = 0, tap = 0, cb = 0;
                int dy = 0, dx = 0;
                if (p.taps == 9) { dy = -1; dx = -1; }
                int cbs = p.cin_blocks, kbpp = kb_per_pass, kmul = kBlockK;   // of the current pass (lo8: pass 1 differs)
                // everything that only changes with the pass is kept out of the per-K-block path (this loop paces the
                // pipeline): tensor maps, and the K / channel coordinates as running sums
                const CUtensorMap* ta = &p.tmA[0];
                const CUtensorMap* tb = &p.tmB[0];
                const CUtensorMap* tbh = &p.tmBh[0];
                int ccoord = 0, kcoord = 0;
                const uint32_t b_off = kABytes + (p.cluster > 1 ? crank * (Cfg::kBBytes >> cshift) : 0);
                const int b_row = n0 + (p.cluster > 1 ? crank * (BN >> cshift) : 0);
                for (int kbg = 0; kbg < p.nkb; ++kbg) {
                    mbar_wait(bar_empty + 8 * stage, phase ^ 1u, p.err_flag);
                    const uint32_t sa = stage_base + stage * stage_pitch;
                    if constexpr (PAIR) {
                        // pair mode: both CTAs' loads complete on the LEADER's barrier, which expects all of them
                        if (elect_one()) {
                            const uint32_t lead_full = mapa_rank(bar_full + 8 * stage, 0);
                            if (crank == 0) mbar_arrive_expect_tx(bar_full + 8 * stage, 2u * (uint32_t)stage_pitch);
                            tma_load_4d_2sm(ta, lead_full, sa, ccoord, x0 + dx, y0 + dy, b0);
                            tma_load_2d_2sm(tbh, lead_full, sa + kABytes, kcoord, n0 + crank * (BN / 2));
                        }
                    } else {
                      if (elect_one()) {
                        mbar_arrive_expect_tx(bar_full + 8 * stage, (uint32_t)stage_pitch);
                        tma_load_4d(ta, bar_full + 8 * stage, sa, ccoord, x0 + dx, y0 + dy, b0);
                        if (p.cluster > 1)      // my slice of the weight tile, into every CTA of the cluster
                            tma_load_2d_multicast(tbh, bar_full + 8 * stage, sa + b_off, kcoord, b_row, cmask);
                        else
                            tma_load_2d(tb, bar_full + 8 * stage, sa + b_off, kcoord, b_row);
                      }
                    }
                    __syncwarp();
                    if (++stage == kStages) { stage = 0; phase ^= 1u; }
                    ++kb;
                    kcoord += kmul;
                    ccoord += kmul;
                    if (++cb == cbs) {
                        cb = 0;
                        ccoord = 0;
                        ++tap;
                        if (p.taps == 9 && ++dx == 2) { dx = -1; ++dy; }
                    }
                    if (kb == kbpp) {
                        kb = 0; tap = 0; cb = 0; ++pass;
                        kcoord = 0; ccoord = 0;
                        if (p.taps == 9) { dy = -1; dx = -1; }
                        if (LO8) { cbs = p.cin_blocks1; kbpp = p.taps * p.cin_blocks1; kmul = 128; }
                        ta = &p.tmA[pass_a(pass)];
                        tb = &p.tmB[pass_b(pass)];
                        tbh = &p.tmBh[pass_b(pass)];
                    }
                }
            }
        }
    } else if (warp == 3) {
        // ======================= halo mode: weight (B) producer on its own warp =======================
        if (p.halo && !p.resb) {
            int bs = 0;
            uint32_t bphase = 0;
            for (int unit = unit0; unit < p.num_units; unit += ustep) {
                if (skip_unit(unit)) continue;
                int m, n0;
                unit_tile(unit, m, n0);
                const int a_tiles = p.cin_blocks + (LO8 ? p.cin_blocks1 : 0);
                for (int t = 0; t < a_tiles; ++t) {
                    const bool f8 = LO8 && t >= p.cin_blocks;                        // lo8: pass-1 tiles come after all pass-0 tiles
                    const int cb = f8 ? t - p.cin_blocks : t;
                    const int planes = LO8 ? 1 : p.npass;
                    if (LO8 && f8 && p.c64) {
                        for (int j = 0; j < 6; ++j) {                         // K blocks (ky, kx 0..1) and (ky, kx 2) of the three filter rows
                            mbar_wait(bar_bempty + 8 * bs, bphase ^ 1u, p.err_flag);
                            if (elect_one()) {
                                expect(bar_bfull + 8 * bs, (uint32_t)kBTile);
                                load_wgt(bar_bfull + 8 * bs, bring_base + bs * kBTile, (3 * (j >> 1) + 2 * (j & 1)) * 64, n0, 2);
                            }
                            __syncwarp();
                            if (++bs == p.b_stages) { bs = 0; bphase ^= 1u; }
                        }
                        continue;
                    }
                    for (int tap = 0; tap < 9; ++tap) {
                        for (int plane = 0; plane < planes; ++plane) {        // fp16 lo pass: the hi and the lo weight tile of this tap
                            mbar_wait(bar_bempty + 8 * bs, bphase ^ 1u, p.err_flag);
                            if (elect_one()) {
                                expect(bar_bfull + 8 * bs, (uint32_t)kBTile);
                                if (f8) load_wgt(bar_bfull + 8 * bs, bring_base + bs * kBTile, (tap * p.cin_blocks1 + cb) * 128, n0, 2);
                                else load_wgt(bar_bfull + 8 * bs, bring_base + bs * kBTile, (tap * p.cin_blocks + cb) * kBlockK, n0, plane);
                            }
                            __syncwarp();
                            if (++bs == p.b_stages) { bs = 0; bphase ^= 1u; }
                        }
                    }
                }
            }
        }
    } else if (warp == 1 && p.halo) {
        // ======================= MMA issuer, halo mode =======================
        constexpr uint32_t idesc = (1u << 4) | kFmtBits | (uint32_t(BN >> 3) << 17) |
                                   (uint32_t((PAIR ? 2 * kTileM : kTileM) >> 4) << 24);
        int stage = 0, bs = 0, gs = 0;                          // gs = running segment count: accumulator gs % kAcc
        uint32_t phase = 0, bphase = 0;
        const bool issuer = !PAIR || crank == 0;                // pair mode: the leader issues for both CTAs
        if (issuer && p.resb && unit0 < p.num_units) {
            mbar_wait(bar_bres, 0, p.err_flag);
            tc_fence_after();
        }
        for (int unit = unit0; issuer && unit < p.num_units; unit += ustep) {
            if (skip_unit(unit)) continue;
            uint32_t tmem_d = 0;
            int seg_left = 0, as = 0;                           // channel blocks left in the open segment
            // The per-tap work is kept to a handful of uniform instructions: all 9 tap views are constant offsets
            // of ONE descriptor (fully unrolled), and with resident weights the 36 MMAs of a channel block are
            // issued from a single elected region.  (The first version rebuilt descriptors and re-elected per tap
            // in a rolled loop: ~500 cycles of issue overhead per tap, 4x the MMA time at N = 64.)
            const int a_tiles = p.cin_blocks + (LO8 ? p.cin_blocks1 : 0);
            for (int t = 0; t < a_tiles; ++t) {
                // lo8: tiles t >= cin_blocks are the e4m3 pass (128 channels per tile, one weight plane, f8f6f4 MMAs)
                const bool f8 = LO8 && t >= p.cin_blocks;
                const int cb = f8 ? t - p.cin_blocks : t;
                const int cbs = f8 ? p.cin_blocks1 : p.cin_blocks;
                const int planes = LO8 ? 1 : p.npass;
                const int slen = f8 ? p.seg_len1 : p.seg_len;
                if (seg_left == 0) {                                        // open the next accumulation segment
                    as = gs % Cfg::kAcc;
                    mbar_wait(bar_tempty + 8 * as, (((uint32_t)gs / Cfg::kAcc) & 1u) ^ 1u, p.err_flag);
                    tc_fence_after();
                    tmem_d = tmem_base + as * BN;
                    seg_left = slen;
                    ++gs;
                }
                const int first = (seg_left == slen) ? 0 : 1;               // 0: this channel block starts the chain
                mbar_wait(bar_full + 8 * stage, phase, p.err_flag);         // halo tile landed
                tc_fence_after();
                const uint64_t da0 = make_halo_desc(stage_base + stage * kHaloBytes);
                // The MMA issue sequence of a tile is kept free of data-dependent branches (the e4m3 / fp16 choice is made
                // once per tile, outside the unrolled loops): a branch between two tcgen05.mma issues costs more than the MMA.
                // c64 e4m3 tile: per filter row one full K block (taps kx = 0, 1: 4 MMAs of 32 bytes) at view column 0 and
                // one half-used block (tap kx = 2: 2 MMAs) at view column 2
                auto issue_tile_c64 = [&]() {
                    if (p.resb) {
                        if (elect_one()) {
                            const uint64_t db0 = make_sw128_desc(base + p.kb0 * kBTile);
#pragma unroll
                            for (int j = 0; j < 6; ++j) {
                                const uint64_t da = da0 + (uint64_t)(((j >> 1) * kHaloW + 2 * (j & 1)) * 8);
                                const uint64_t db = db0 + (uint64_t)(j * ((uint32_t)kBTile >> 4));
#pragma unroll
                                for (int k = 0; k < ((j & 1) ? 2 : 4); ++k)
                                    mma8(tmem_d, da + 2 * k, db + 2 * k, idesc, (first | j | k) != 0 ? 1u : 0u);
                            }
                            commit(bar_empty + 8 * stage);
                        }
                        __syncwarp();
                    } else {
#pragma unroll
                        for (int j = 0; j < 6; ++j) {
                            const uint64_t da = da0 + (uint64_t)(((j >> 1) * kHaloW + 2 * (j & 1)) * 8);
                            mbar_wait(bar_bfull + 8 * bs, bphase, p.err_flag);
                            tc_fence_after();
                            const uint64_t db = make_sw128_desc(bring_base + bs * kBTile);
                            if (elect_one()) {
#pragma unroll
                                for (int k = 0; k < ((j & 1) ? 2 : 4); ++k)
                                    mma8(tmem_d, da + 2 * k, db + 2 * k, idesc, (first | j | k) != 0 ? 1u : 0u);
                                commit(bar_bempty + 8 * bs);
                                if (j == 5) commit(bar_empty + 8 * stage);
                            }
                            __syncwarp();
                            if (++bs == p.b_stages) { bs = 0; bphase ^= 1u; }
                        }
                    }
                };
                auto issue_tile = [&](auto f8tag) {
                    constexpr bool F8 = decltype(f8tag)::value;
                    if (p.resb) {
                        const uint32_t bstep = (uint32_t)(cbs * kBTile) >> 4;   // next tap, same channel block
                        if (elect_one()) {
                            // (tap, plane) order like the weight-ring path below: a layer's accumulation order — and with
                            // it every bit of its output — is the same whichever path the batch size selects
                            const uint64_t db0 = make_sw128_desc(base + ((F8 ? p.kb0 : 0) + cb) * kBTile);
                            const uint32_t pstep = (uint32_t)(kb_per_pass * kBTile) >> 4;    // hi plane -> lo plane (fp16 lo pass)
#pragma unroll
                            for (int tap = 0; tap < 9; ++tap) {
                                const uint64_t da = da0 + (uint64_t)(((tap / 3) * kHaloW + (tap % 3)) * 8);   // (ky*10+kx)*128 B >> 4
                                for (int plane = 0; plane < planes; ++plane) {
                                    const uint64_t db = db0 + (uint64_t)(tap * bstep + plane * pstep);
#pragma unroll
                                    for (int k = 0; k < kBlockK / 16; ++k) {
                                        if constexpr (F8) mma8(tmem_d, da + 2 * k, db + 2 * k, idesc, (first | tap | k | plane) != 0 ? 1u : 0u);
                                        else mma(tmem_d, da + 2 * k, db + 2 * k, idesc, (first | tap | k | plane) != 0 ? 1u : 0u);
                                    }
                                }
                            }
                            commit(bar_empty + 8 * stage);                 // halo tile free again
                        }
                        __syncwarp();
                    } else {
#pragma unroll
                        for (int tap = 0; tap < 9; ++tap) {
                            const uint64_t da = da0 + (uint64_t)(((tap / 3) * kHaloW + (tap % 3)) * 8);
                            for (int plane = 0; plane < planes; ++plane) {
                                mbar_wait(bar_bfull + 8 * bs, bphase, p.err_flag);
                                tc_fence_after();
                                const uint64_t db = make_sw128_desc(bring_base + bs * kBTile);
                                if (elect_one()) {
#pragma unroll
                                    for (int k = 0; k < kBlockK / 16; ++k) {
                                        if constexpr (F8) mma8(tmem_d, da + 2 * k, db + 2 * k, idesc, (first | tap | k | plane) != 0 ? 1u : 0u);
                                        else mma(tmem_d, da + 2 * k, db + 2 * k, idesc, (first | tap | k | plane) != 0 ? 1u : 0u);
                                    }
                                    commit(bar_bempty + 8 * bs);
                                    if (tap == 8 && plane == planes - 1) commit(bar_empty + 8 * stage);
                                }
                                __syncwarp();
                                if (++bs == p.b_stages) { bs = 0; bphase ^= 1u; }
                            }
                        }
                    }
                };
                bool done = false;
                if constexpr (LO8) if (f8 && p.c64) { issue_tile_c64(); done = true; }
                if (!done) {
                    if (f8) issue_tile(std::true_type{});
                    else issue_tile(std::false_type{});
                }
                if (++stage == kStages) { stage = 0; phase ^= 1u; }
                // segment complete (also at the pass boundary and at the end of the tile) -> epilogue adds it in
                if (--seg_left == 0 || t == a_tiles - 1 || (LO8 && t == p.cin_blocks - 1)) {
                    seg_left = 0;
                    if (elect_one()) commit(bar_tfull + 8 * as);
                    __syncwarp();
                }
            }
        }
    } else if (warp == 1 && PAIR) {
        // ======================= MMA issuer, pair mode: rank 0 issues M = 256 MMAs for both CTAs =======
        if constexpr (PAIR) if (crank == 0) {
            constexpr uint32_t idesc = (1u << 4) | kFmtBits | (uint32_t(BN >> 3) << 17) |
                                       (uint32_t((2 * kTileM) >> 4) << 24);
            int stage = 0;
            uint32_t phase = 0;
            int gs = 0;
            for (int unit = unit0; unit < p.num_units; unit += ustep) {
                if (skip_unit(unit)) continue;
                for (int k0 = 0, k1 = 0; k0 < p.nkb; k0 = k1, ++gs) {       // one accumulation chain per segment
                    const int as = gs % Cfg::kAcc;
                    {   // lo8: segments end at the pass boundary kb0; pass 1 has its own segment length
                        const bool in1 = LO8 && k0 >= p.kb0;
                        const int lim = (LO8 && !in1) ? p.kb0 : p.nkb;
                        k1 = k0 + (in1 ? p.seg_len1 : p.seg_len);
                        if (k1 > lim) k1 = lim;
                    }
                    mbar_wait(bar_tempty + 8 * as, (((uint32_t)gs / Cfg::kAcc) & 1u) ^ 1u, p.err_flag);   // both CTAs' epilogues drained it
                    tc_fence_after();
                    const uint32_t tmem_d = tmem_base + as * BN;
                    const bool f8 = LO8 && k0 >= p.kb0;                    // whole segment: e4m3 pass or not
                    for (int kb = k0; kb < k1; ++kb) {
                        mbar_wait(bar_full + 8 * stage, phase, p.err_flag);     // both CTAs' TMA bytes landed
                        tc_fence_after();
                        const uint32_t sa = stage_base + stage * stage_pitch;
                        const uint64_t da = make_sw128_desc(sa);
                        const uint64_t db = make_sw128_desc(sa + kABytes);
                        if (elect_one()) {
                            if (f8) {
#pragma unroll
                                for (int k = 0; k < kBlockK / 16; ++k)
                                    mma8(tmem_d, da + 2 * k, db + 2 * k, idesc, (kb > k0 || k > 0) ? 1u : 0u);
                            } else {
#pragma unroll
                                for (int k = 0; k < kBlockK / 16; ++k)
                                    umma_bf16_2sm(tmem_d, da + 2 * k, db + 2 * k, idesc, (kb > k0 || k > 0) ? 1u : 0u);
                            }
                            umma_commit_2sm(bar_empty + 8 * stage, (uint16_t)0x3);   // stage free again, in both CTAs
                        }
                        __syncwarp();
                        if (++stage == kStages) { stage = 0; phase ^= 1u; }
                    }
                    if (elect_one()) umma_commit_2sm(bar_tfull + 8 * as, (uint16_t)0x3);   // both epilogues may add it in
                    __syncwarp();
                }
            }
        }
    } else if (warp == 1) {
        // ======================= MMA issuer (whole warp runs the loop; one elected lane issues) =======
        {
            constexpr uint32_t idesc = (1u << 4) | kFmtBits | (uint32_t(BN >> 3) << 17) |
                                       (uint32_t(kTileM >> 4) << 24);
            int stage = 0;
            uint32_t phase = 0;
            int gs = 0;
            for (int unit = unit0; unit < p.num_units; unit += ustep) {
                if (skip_unit(unit)) continue;
                for (int k0 = 0, k1 = 0; k0 < p.nkb; k0 = k1, ++gs) {       // one accumulation chain per segment
                    const int as = gs % Cfg::kAcc;
                    {   // lo8: segments end at the pass boundary kb0; pass 1 has its own segment length
                        const bool in1 = LO8 && k0 >= p.kb0;
                        const int lim = (LO8 && !in1) ? p.kb0 : p.nkb;
                        k1 = k0 + (in1 ? p.seg_len1 : p.seg_len);
                        if (k1 > lim) k1 = lim;
                    }
                    mbar_wait(bar_tempty + 8 * as, (((uint32_t)gs / Cfg::kAcc) & 1u) ^ 1u, p.err_flag);   // epilogue drained this accumulator
                    tc_fence_after();
                    const uint32_t tmem_d = tmem_base + as * BN;
                    for (int kb = k0; kb < k1; ++kb) {
                        mbar_wait(bar_full + 8 * stage, phase, p.err_flag);     // TMA bytes landed
                        tc_fence_after();
                        const uint32_t sa = stage_base + stage * stage_pitch;
                        const uint64_t da = make_sw128_desc(sa);
                        const uint64_t db = make_sw128_desc(sa + kABytes);
                        if (elect_one()) {
                            // advance 16 bf16 = 32 B inside the 128-B swizzle atom: +2 in the (addr >> 4) field
                            if (LO8 && k0 >= p.kb0) {                 // whole segment: e4m3 pass
#pragma unroll
                                for (int k = 0; k < kBlockK / 16; ++k)
                                    mma8(tmem_d, da + 2 * k, db + 2 * k, idesc, (kb > k0 || k > 0) ? 1u : 0u);
                            } else {
#pragma unroll
                                for (int k = 0; k < kBlockK / 16; ++k)
                                    umma_bf16(tmem_d, da + 2 * k, db + 2 * k, idesc, (kb > k0 || k > 0) ? 1u : 0u);
                            }
                            // frees the smem stage when the MMAs retire — in cluster mode on BOTH CTAs: the peer's
                            // producer multicasts into this stage too and must see it released
                            if (p.cluster > 1) umma_commit_multicast(bar_empty + 8 * stage, cmask);
                            else umma_commit(bar_empty + 8 * stage);
                        }
                        __syncwarp();
                        if (++stage == kStages) { stage = 0; phase ^= 1u; }
                    }
                    if (elect_one()) umma_commit(bar_tfull + 8 * as);       // segment complete -> epilogue adds it in
                    __syncwarp();
                }
            }
        }
    }
    } else {
        asm volatile("setmaxnreg.inc.sync.aligned.u32 224;");
        // ======================= epilogue =======================
        // Thread = one accumulator row (pixel); a chunk = 32 output channels held in registers.
        // 2x2 pooling never leaves the warp: the tile box is at most 16 pixels wide, so a warp holds an
        // even number of complete tile rows and the 4 pixels of a pooling window are lanes
        // {l, l^1, l^BW, l^BW^1} -> a shuffle butterfly, no shared-memory staging, no block barrier.
        // 8 epilogue warps: two per TMEM lane quarter (a warp may only touch lanes 32*(warp%4)..+31); the two
        // warps of a quarter take alternate 32-column chunks, which doubles the epilogue rate of the short-K
        // layers where the epilogue, not the MMA, sets the tile time.
        const int ew = (warp - 4) & 3;               // TMEM lane quarter
        const int grp = (warp - 4) >> 2;             // chunk parity handled by this warp
        const int row = ew * 32 + lane;              // accumulator row (= pixel within the tile)
        const int ww = row % p.BW;
        const int t2 = row / p.BW;
        const int hh = t2 % p.BH;
        const int bb = t2 / p.BH;
        constexpr int NC = BN / 64;                  // 32-column chunks per thread: chunk index c = grp + 2 * ci
        int gs = 0;                                  // running segment count (same sequence as the MMA warp's)
        for (int unit = unit0; unit < p.num_units; unit += ustep) {
            if (skip_unit(unit)) continue;
            int m, n0;
            unit_tile(unit, m, n0);
            const int wt = m % p.tiles_w; m /= p.tiles_w;
            const int ht = m % p.tiles_h;
            const int bt = m / p.tiles_h;
            const int x = wt * p.BW + ww, y = ht * p.BH + hh, b = bt * p.BB + bb;
            bool valid;
            size_t obase, o8base;                    // o8base: the same pixel in the (possibly W-padded) e4m3 copy
            int edge = 0;                            // W-padded e4m3 copy: -1 / +1 = this pixel also zeroes the column left / right of it
            {
                const int ox = p.pool ? (x >> 1) : x, oy = p.pool ? (y >> 1) : y;
                valid = ox < p.Wo && oy < p.Ho && b < p.B;          // pooling: all 4 lanes of a window store (8 channels each)
                obase = ((size_t(b) * p.Ho + oy) * p.Wo + ox) * p.N + n0;
                o8base = obase;
                if (F16 && p.out8_wpad) {
                    o8base = ((size_t(b) * p.Ho + oy) * (p.Wo + 2) + ox + 1) * p.N + n0;
                    edge = ox == 0 ? -1 : (ox == p.Wo - 1 ? 1 : 0);
                }
            }

            // ---- add the tile's accumulation segments in fp32 registers (round-to-nearest): run[ci][j] = this row's
            // column n0 + 32 * (grp + 2 ci) + j.  A segment's accumulator goes back to the MMA warp as soon as it has
            // been read, so the tensor pipe runs the next segment while this one is being added.
            float run[NC][32];
#pragma unroll
            for (int ci = 0; ci < NC; ++ci)
#pragma unroll
                for (int j = 0; j < 32; ++j) run[ci][j] = 0;
#pragma unroll 1
            for (int seg = 0; seg < p.nseg; ++seg, ++gs) {
                const int as = gs % Cfg::kAcc;
                const float sc = (LO8 && seg >= p.nseg0) ? p.lo_scale : 1.0f;
                mbar_wait(bar_tfull + 8 * as, ((uint32_t)gs / Cfg::kAcc) & 1u, p.err_flag);
                tc_fence_after();
                const uint32_t taddr = tmem_base + (uint32_t(ew * 32) << 16) + as * BN + grp * 32;
#pragma unroll
                for (int ci = 0; ci < NC; ++ci) {
                    uint32_t r[32];
                    tmem_ld_32x32b_x32(taddr + ci * 64, r);
                    tmem_ld_wait();
                    // (pass-1 segments of the e4m3 lo pass carry the factor 2^s of their weights; sc = 1 otherwise)
#pragma unroll
                    for (int j = 0; j < 32; ++j) {
                        if constexpr (LO8) run[ci][j] = __fmaf_rn(__uint_as_float(r[j]), sc, run[ci][j]);
                        else run[ci][j] = __fadd_rn(run[ci][j], __uint_as_float(r[j]));
                    }
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) {
                    if constexpr (PAIR) mbar_arrive_cluster(mapa_rank(bar_tempty + 8 * as, 0));   // the leader's MMA warp waits
                    else mbar_arrive(bar_tempty + 8 * as);
                }
            }

            if constexpr (SYRK) {
                // S[row][n0 + 32 c + j] += tile sum, for tiles on or above the diagonal (128-granular, like stats_finalize
                // reads them); every (M tile, N tile) has exactly one owner in a launch, launches are stream-ordered
                const bool keep = valid && (n0 / kTileM >= x / kTileM);
#pragma unroll
                for (int ci = 0; ci < NC; ++ci) {
                    // (8-byte accesses: the statistic sits 1 + d doubles into the caller's buffer, so rows are not
                    // 16-byte aligned in general)
                    double* dst = p.out_f64 + obase + (grp + 2 * ci) * 32;
                    if (keep) {
#pragma unroll
                        for (int j = 0; j < 32; ++j) dst[j] += (double)run[ci][j];
                    }
                }
            } else {
#pragma unroll
            for (int ci = 0; ci < NC; ++ci) {
                const int c = grp + 2 * ci;
                float (&v)[32] = run[ci];
                {
                    const float4* bp = reinterpret_cast<const float4*>(s_bias + n0 + c * 32);
#pragma unroll
                    for (int q = 0; q < 8; ++q) {
                        const float4 bv = bp[q];
                        v[4 * q + 0] += bv.x;
                        v[4 * q + 1] += bv.y;
                        v[4 * q + 2] += bv.z;
                        v[4 * q + 3] += bv.w;
                    }
                }
                if (p.relu) {
#pragma unroll
                    for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], 0.f);
                }
                if (p.pool) {
                    // 2x2 pooling as a two-step butterfly that also SPLITS the channels: after exchanging with the
                    // x-neighbour (lane ^ 1) the even lane owns pooled channels 0-15 and the odd lane 16-31; after
                    // the y-neighbour (lane ^ BW) each of the 4 lanes of a window owns 8 pooled channels.
                    // 24 shuffles per chunk instead of 64, and all 4 lanes take part in the store.
                    const bool odd = (lane & 1) != 0;
                    const bool up = (lane & p.BW) != 0;
                    float hx[16];
#pragma unroll
                    for (int j = 0; j < 16; ++j) {
                        const float recv = __shfl_xor_sync(0xffffffffu, odd ? v[j] : v[16 + j], 1);
                        const float mine = odd ? v[16 + j] : v[j];
                        hx[j] = p.pool == 1 ? fmaxf(mine, recv) : mine + recv;
                    }
                    float g8[8];
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const float recv = __shfl_xor_sync(0xffffffffu, up ? hx[j] : hx[8 + j], p.BW);
                        const float mine = up ? hx[8 + j] : hx[j];
                        g8[j] = p.pool == 1 ? fmaxf(mine, recv) : (mine + recv) * 0.25f;
                    }
                    if (valid) {
                        const size_t o = obase + c * 32 + (odd ? 16 : 0) + (up ? 8 : 0);
                        if (p.out_f32) {
                            st_global_256(p.out_f32 + o,
                                          make_uint4(__float_as_uint(g8[0]), __float_as_uint(g8[1]), __float_as_uint(g8[2]),
                                                     __float_as_uint(g8[3])),
                                          make_uint4(__float_as_uint(g8[4]), __float_as_uint(g8[5]), __float_as_uint(g8[6]),
                                                     __float_as_uint(g8[7])));
                        } else {
                            uint4 hi;
                            hi.x = pack_act2<F16>(g8[0], g8[1]); hi.y = pack_act2<F16>(g8[2], g8[3]);
                            hi.z = pack_act2<F16>(g8[4], g8[5]); hi.w = pack_act2<F16>(g8[6], g8[7]);
                            *reinterpret_cast<uint4*>(p.out_hi + o) = hi;
                            if (F16 && p.out8) {
                                uint8_t* o8 = p.out8 + o8base + (o - obase);
                                *reinterpret_cast<uint2*>(o8) = make_uint2(pack4_e4m3(g8[0], g8[1], g8[2], g8[3]),
                                                                           pack4_e4m3(g8[4], g8[5], g8[6], g8[7]));
                                if (edge != 0) *reinterpret_cast<uint2*>(o8 + edge * p.N) = make_uint2(0u, 0u);
                                if (edge < 0 && p.Wo == 1) *reinterpret_cast<uint2*>(o8 + p.N) = make_uint2(0u, 0u);
                            }
                            if (!F16 && p.out_lo) {
                                uint4 lo;
                                lo.x = pack_bf16x2(g8[0] - bf16_round(g8[0]), g8[1] - bf16_round(g8[1]));
                                lo.y = pack_bf16x2(g8[2] - bf16_round(g8[2]), g8[3] - bf16_round(g8[3]));
                                lo.z = pack_bf16x2(g8[4] - bf16_round(g8[4]), g8[5] - bf16_round(g8[5]));
                                lo.w = pack_bf16x2(g8[6] - bf16_round(g8[6]), g8[7] - bf16_round(g8[7]));
                                *reinterpret_cast<uint4*>(p.out_lo + o) = lo;
                            }
                        }
                    }
                    continue;
                }
                if (valid) {
                    const size_t o = obase + c * 32;
                    if (p.out_f32) {
#pragma unroll
                        for (int q = 0; q < 4; ++q)       // 128 contiguous bytes per lane: four full-sector stores
                            st_global_256(p.out_f32 + o + 8 * q,
                                          make_uint4(__float_as_uint(v[8 * q]), __float_as_uint(v[8 * q + 1]),
                                                     __float_as_uint(v[8 * q + 2]), __float_as_uint(v[8 * q + 3])),
                                          make_uint4(__float_as_uint(v[8 * q + 4]), __float_as_uint(v[8 * q + 5]),
                                                     __float_as_uint(v[8 * q + 6]), __float_as_uint(v[8 * q + 7])));
                    } else {
                        uint4 hi[4];
#pragma unroll
                        for (int q = 0; q < 4; ++q) {
                            hi[q].x = pack_act2<F16>(v[8 * q + 0], v[8 * q + 1]); hi[q].y = pack_act2<F16>(v[8 * q + 2], v[8 * q + 3]);
                            hi[q].z = pack_act2<F16>(v[8 * q + 4], v[8 * q + 5]); hi[q].w = pack_act2<F16>(v[8 * q + 6], v[8 * q + 7]);
                        }
                        st_global_256(p.out_hi + o, hi[0], hi[1]);          // 64 contiguous bytes per lane: two full sectors
                        st_global_256(p.out_hi + o + 16, hi[2], hi[3]);
                        if (F16 && p.out8) {
                            uint4 q0, q1;
                            q0.x = pack4_e4m3(v[0], v[1], v[2], v[3]);     q0.y = pack4_e4m3(v[4], v[5], v[6], v[7]);
                            q0.z = pack4_e4m3(v[8], v[9], v[10], v[11]);   q0.w = pack4_e4m3(v[12], v[13], v[14], v[15]);
                            q1.x = pack4_e4m3(v[16], v[17], v[18], v[19]); q1.y = pack4_e4m3(v[20], v[21], v[22], v[23]);
                            q1.z = pack4_e4m3(v[24], v[25], v[26], v[27]); q1.w = pack4_e4m3(v[28], v[29], v[30], v[31]);
                            uint8_t* o8 = p.out8 + o8base + c * 32;
                            st_global_256(o8, q0, q1);
                            if (edge != 0) {
                                const uint4 z = make_uint4(0u, 0u, 0u, 0u);
                                st_global_256(o8 + edge * p.N, z, z);
                                if (edge < 0 && p.Wo == 1) st_global_256(o8 + p.N, z, z);
                            }
                        }
                        if (!F16 && p.out_lo) {
                            uint4 lo[4];
#pragma unroll
                            for (int q = 0; q < 4; ++q) {
                                lo[q].x = pack_bf16x2(v[8 * q + 0] - bf16_round(v[8 * q + 0]), v[8 * q + 1] - bf16_round(v[8 * q + 1]));
                                lo[q].y = pack_bf16x2(v[8 * q + 2] - bf16_round(v[8 * q + 2]), v[8 * q + 3] - bf16_round(v[8 * q + 3]));
                                lo[q].z = pack_bf16x2(v[8 * q + 4] - bf16_round(v[8 * q + 4]), v[8 * q + 5] - bf16_round(v[8 * q + 5]));
                                lo[q].w = pack_bf16x2(v[8 * q + 6] - bf16_round(v[8 * q + 6]), v[8 * q + 7] - bf16_round(v[8 * q + 7]));
                            }
                            st_global_256(p.out_lo + o, lo[0], lo[1]);
                            st_global_256(p.out_lo + o + 16, lo[2], lo[3]);
                        }
                    }
                }
            }
            }   // !SYRK
        }
    }

    tc_fence_before();
    __syncthreads();
    if (p.cluster > 1) cluster_sync_all();       // no CTA leaves while its peer may still arrive on its barriers
    if (warp == 2) {
        tc_fence_after();
        if constexpr (PAIR) tmem_dealloc_2sm(tmem_base, Cfg::kTmemCols);
        else tmem_dealloc(tmem_base, Cfg::kTmemCols);
    }
}

constexpr int kSegmentBlocks = 16;  // K blocks (= 64 MMAs of K = 16) per accumulation chain in tensor memory

// ------------------------------------------------------------------------------------------------
// Host side
// ------------------------------------------------------------------------------------------------
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static PFN_encodeTiled g_encode = nullptr;

int gemm_init(fadb_handle* h) {
    (void)h;
    if (!g_encode) {
        void* fn = nullptr;
        cudaDriverEntryPointQueryResult qres;
        FADB_CUDA_CHECK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
        if (!fn || qres != cudaDriverEntryPointSuccess) {
            set_error("cuTensorMapEncodeTiled not available from the driver");
            return FADB_E_CUDA;
        }
        g_encode = reinterpret_cast<PFN_encodeTiled>(fn);
    }
    for (void (*k)(GemmParams) : {fadb_gemm_tc_kernel<64, false, false>, fadb_gemm_tc_kernel<128, false, false>,
                                  fadb_gemm_tc_kernel<256, false, false>, fadb_gemm_tc_kernel<64, true, false>,
                                  fadb_gemm_tc_kernel<128, true, false>, fadb_gemm_tc_kernel<256, true, false>,
                                  fadb_gemm_tc_kernel<64, false, true>, fadb_gemm_tc_kernel<128, false, true>,
                                  fadb_gemm_tc_kernel<256, false, true>, fadb_gemm_tc_kernel<64, true, true>,
                                  fadb_gemm_tc_kernel<128, true, true>, fadb_gemm_tc_kernel<256, true, true>,
                                  fadb_gemm_tc_kernel<128, false, true, true>, fadb_gemm_tc_kernel<128, true, true, true>,
                                  fadb_gemm_tc_kernel<64, false, true, false, true>, fadb_gemm_tc_kernel<128, false, true, false, true>,
                                  fadb_gemm_tc_kernel<256, false, true, false, true>, fadb_gemm_tc_kernel<64, true, true, false, true>,
                                  fadb_gemm_tc_kernel<128, true, true, false, true>, fadb_gemm_tc_kernel<256, true, true, false, true>})
        FADB_CUDA_CHECK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, GemmCfg<64>::kMaxSmemBytes));
    return FADB_OK;
}

// rs = row stride in elements (0 = C: rows are dense)
// kind: 0 = bf16, 1 = fp16, 2 = e4m3 bytes (128 channels per 128-byte box row instead of 64)
static int encode_act_map(CUtensorMap* tm, const void* ptr, int C, int W, int H, int B, int BW, int BH, int BB, int kind,
                          long long rs = 0) {
    if (rs <= 0) rs = C;
    const cuuint64_t es = kind == 2 ? 1 : 2;
    cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
    cuuint64_t strides[3] = {(cuuint64_t)rs * es, (cuuint64_t)W * rs * es, (cuuint64_t)H * W * rs * es};
    cuuint32_t box[4] = {(cuuint32_t)(kind == 2 ? 128 : kBlockK), (cuuint32_t)BW, (cuuint32_t)BH, (cuuint32_t)BB};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = g_encode(tm, kind == 2 ? CU_TENSOR_MAP_DATA_TYPE_UINT8
                              : (kind == 1 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16), 4,
                          const_cast<void*>(ptr), dims, strides, box, estr,
                          CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled(activations C=%d W=%d H=%d B=%d box %d,%d,%d) failed: %d", C, W, H, B, BW, BH,
                  BB, (int)r);
        return FADB_E_CUDA;
    }
    return FADB_OK;
}

// GemmParams::c64: the W-padded e4m3 tensor [B][H][W+2][64] as rows of 128 bytes that OVERLAP by 64 (row r = padded
// pixels r, r+1); W+1 rows per image row.  (tools/tma_overlap_probe.cu: a W stride below the innermost extent encodes
// and loads as expected.)
static int encode_act_map_c64(CUtensorMap* tm, const void* ptr, int W, int H, int B) {
    cuuint64_t dims[4] = {128, (cuuint64_t)(W + 1), (cuuint64_t)H, (cuuint64_t)B};
    cuuint64_t strides[3] = {64, (cuuint64_t)(W + 2) * 64, (cuuint64_t)H * (W + 2) * 64};
    cuuint32_t box[4] = {128, (cuuint32_t)kHaloW, 18, 1};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = g_encode(tm, CU_TENSOR_MAP_DATA_TYPE_UINT8, 4, const_cast<void*>(ptr), dims, strides, box, estr,
                          CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled(e4m3 activations, overlapping rows, W=%d H=%d B=%d) failed: %d", W, H, B, (int)r);
        return FADB_E_CUDA;
    }
    return FADB_OK;
}

static int encode_weight_map(CUtensorMap* tm, const void* ptr, int K, int N, int BN, int kind, long long rs = 0) {
    if (rs <= 0) rs = K;
    const cuuint64_t es = kind == 2 ? 1 : 2;
    cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)N};
    cuuint64_t strides[1] = {(cuuint64_t)rs * es};
    cuuint32_t box[2] = {(cuuint32_t)(kind == 2 ? 128 : kBlockK), (cuuint32_t)BN};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = g_encode(tm, kind == 2 ? CU_TENSOR_MAP_DATA_TYPE_UINT8
                              : (kind == 1 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16), 2,
                          const_cast<void*>(ptr), dims, strides, box, estr,
                          CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled(weights K=%d N=%d BN=%d) failed: %d", K, N, BN, (int)r);
        return FADB_E_CUDA;
    }
    return FADB_OK;
}

int launch_gemm_layer(fadb_handle* h, const PackedLayer& L, const LayerIO& io, cudaStream_t st) {
    FADB_REQUIRE(g_encode != nullptr, "gemm layer used before gemm_init");
    FADB_REQUIRE(io.Cin % kBlockK == 0, "Cin=%d must be a multiple of 64", io.Cin);
    FADB_REQUIRE(io.Cin == L.Cin && io.taps == L.taps, "layer/IO mismatch (Cin %d vs %d, taps %d vs %d)", io.Cin,
                 L.Cin, io.taps, L.taps);
    FADB_REQUIRE(io.taps == 9 || io.taps == 1, "taps must be 1 or 9");
    const bool f16 = io.syrk ? true : prec_is_f16(h->precision);
    FADB_REQUIRE(L.f16 == (int)f16, "layer weights were packed for %s but the handle's precision wants %s: commit the weights "
                 "again after fadb_set_precision", L.f16 ? "fp16" : "bf16", f16 ? "fp16" : "bf16");
    // operand passes over K (see GemmParams::npass)
    const int npass = io.syrk ? 3
                      : (h->precision == FADB_PREC_BF16X3 && io.in_lo && L.w_lo) ? 3
                      : (h->precision == FADB_PREC_FP16X2 && L.w_lo && io.use_lo_weights) ? 2 : 1;
    const int BN = (L.N % 256 == 0 && !io.syrk) ? 256 : (L.N % 128 == 0 ? 128 : 64);
    if (io.syrk) FADB_REQUIRE(io.in_lo && L.w_lo && io.taps == 1 && io.H == 1 && io.B == 1 && io.out_f64 && BN == 128,
                              "syrk launch: needs both lo planes, linear-layer shape, fp64 output and N %% 128 == 0");
    FADB_REQUIRE(L.N % BN == 0 && L.N >= 64, "Cout=%d must be a multiple of 64", L.N);
    FADB_REQUIRE(io.B > 0 && io.H > 0 && io.W > 0, "empty layer input");

    // halo mode (single-pass 3x3 layers on maps at least 8 wide / 16 high): 8 x 16 pixel tiles whose input is ONE
    // 10 x 18-pixel halo box per channel block instead of nine shifted boxes -> 6x less L2 -> SM traffic for
    // the A operand.  Only used when H is large against the 16-row tile (few wasted rows).
    int halo = 0;
    int BW = 1, BH = 1, BB = 1;
    if (h->halo && io.taps == 9 && npass <= 2 && io.W % 8 == 0 && io.H >= 16 &&
        double((io.H + 15) / 16 * 16) / io.H <= 1.13) {
        halo = 1;
        BW = 8; BH = 16; BB = 1;
    } else if (io.W >= kTileM || (io.taps == 1 && io.H == 1 && io.B == 1)) {
        BW = kTileM;      // rows of a linear layer: the box may overhang the tensor, TMA zero-fills
    } else {
        // pick the 128-pixel box: full rows first, then rows, then images
        BW = 1;
        while (BW * 2 <= io.W && BW * 2 <= kTileM) BW *= 2;        // largest power of two <= W
        FADB_REQUIRE(BW == io.W, "W=%d must be a power of two below 128 or >= 128", io.W);
        if (io.pool && BW > 16) BW = 16;                           // pooling window must live inside one warp
        BH = kTileM / BW;
        if (BH > io.H) {
            // fewer rows than the box: pick the power-of-two row count that wastes the fewest box rows
            // (ties -> larger), and fill the rest of the 128 pixels with further images
            int best = io.pool ? 2 : 1;
            double best_waste = 1e30;
            for (int c = best; c <= BH; c *= 2) {
                const int tiles = (io.H + c - 1) / c;
                const double waste = double(tiles) * c / io.H;
                if (waste < best_waste - 1e-9 || (waste < best_waste + 1e-9 && c > best)) {
                    best = c;
                    best_waste = waste;
                }
            }
            BH = best;
            BB = kTileM / (BW * BH);
        }
    }
    FADB_REQUIRE(BW * BH * BB == kTileM, "cannot tile W=%d H=%d into 128-pixel boxes", io.W, io.H);
    if (io.pool)
        FADB_REQUIRE(BW % 2 == 0 && BH % 2 == 0 && BW <= 16 && io.W % 2 == 0,
                     "pooling needs an even box at most 16 wide (BW=%d BH=%d W=%d)", BW, BH, io.W);

    GemmParams p;
    memset(&p, 0, sizeof(p));
    // fp16x2: the low-order weight pass in e4m3 (twice the MMA rate, half the operand bytes) where the layer has whole
    // 128-channel blocks and both e4m3 operands exist
    // 64-channel 3x3 layers in halo mode: the same with two taps per 128-byte K block (GemmParams::c64); their e4m3 input
    // is the W-padded layout (io.in8_wpad)
    const bool c64 = npass == 2 && !io.syrk && h->lo_fp8 && io.in8 && io.in8_wpad && L.w8 && io.Cin == 64 && halo;
    const bool lo8 = c64 || (npass == 2 && !io.syrk && h->lo_fp8 && io.in8 && !io.in8_wpad && L.w8 && io.Cin % 128 == 0);
    FADB_CHECK(encode_act_map(&p.tmA[0], io.in_hi, io.Cin, io.W, io.H, io.B, BW, BH, BB, f16, io.row_stride));
    if (halo) FADB_CHECK(encode_act_map(&p.tmH, io.in_hi, io.Cin, io.W, io.H, io.B, kHaloW, 18, 1, f16));
    else p.tmH = p.tmA[0];
    p.halo = halo;
    FADB_CHECK(encode_weight_map(&p.tmB[0], L.w_hi, L.K, L.N, BN, f16, io.row_stride));
    p.tmA[1] = p.tmA[0];
    p.tmB[1] = p.tmB[0];
    if (npass == 3) FADB_CHECK(encode_act_map(&p.tmA[1], io.in_lo, io.Cin, io.W, io.H, io.B, BW, BH, BB, f16, io.row_stride));
    if (npass >= 2 && !lo8) FADB_CHECK(encode_weight_map(&p.tmB[1], L.w_lo, L.K, L.N, BN, f16, io.row_stride));
    p.tmA[2] = p.tmA[0];
    p.tmB[2] = p.tmB[0];
    p.tmH8 = p.tmH;
    if (c64) {
        FADB_CHECK(encode_act_map_c64(&p.tmH8, io.in8, io.W, io.H, io.B));
        FADB_CHECK(encode_weight_map(&p.tmB[2], L.w8, L.K, L.N, BN, 2));
    } else if (lo8) {
        FADB_CHECK(encode_act_map(&p.tmA[2], io.in8, io.Cin, io.W, io.H, io.B, BW, BH, BB, 2));
        if (halo) FADB_CHECK(encode_act_map(&p.tmH8, io.in8, io.Cin, io.W, io.H, io.B, kHaloW, 18, 1, 2));
        FADB_CHECK(encode_weight_map(&p.tmB[2], L.w8, L.K, L.N, BN, 2));
    }
    p.lo8 = lo8 ? 1 : 0;
    p.c64 = c64 ? 1 : 0;
    p.lo_scale = lo8 ? L.lo_scale : 1.f;
    p.W = io.W; p.H = io.H; p.B = io.B;
    p.BW = BW; p.BH = BH; p.BB = BB;
    p.tiles_w = (io.W + BW - 1) / BW;
    p.tiles_h = (io.H + BH - 1) / BH;
    p.tiles_b = (io.B + BB - 1) / BB;
    p.tiles_n = L.N / BN;
    const long long nt = 1LL * p.tiles_w * p.tiles_h * p.tiles_b * p.tiles_n;
    FADB_REQUIRE(nt < (1LL << 31), "too many tiles");
    p.num_tiles = (int)nt;
    p.cin_blocks = io.Cin / kBlockK;
    p.taps = io.taps;
    p.npass = npass;
    p.N = L.N;
    p.relu = io.relu;
    p.pool = io.pool;
    p.Ho = io.pool ? io.H / 2 : io.H;
    p.Wo = io.pool ? io.W / 2 : io.W;
    p.bias = L.bias;
    p.out_hi = io.out_hi;
    p.out_lo = (h->precision == FADB_PREC_BF16X3 && !io.syrk) ? io.out_lo : nullptr;
    p.upper_only = io.syrk;
    p.out_f32 = io.out_f32;
    p.out_f64 = io.out_f64;
    p.out8 = (h->precision == FADB_PREC_FP16X2 && h->lo_fp8 && !io.out_f32 && (L.N % 128 == 0 || io.out8_wpad)) ? io.out8 : nullptr;
    p.out8_wpad = (p.out8 && io.out8_wpad) ? 1 : 0;
    p.err_flag = h->err_flag;
    FADB_REQUIRE(p.out_f32 || p.out_hi || p.out_f64, "layer has no output buffer");

    const int grid = p.num_tiles < h->sm_count ? p.num_tiles : h->sm_count;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    if (h->profile) {
        FADB_CUDA_CHECK(cudaEventCreate(&ev0));
        FADB_CUDA_CHECK(cudaEventCreate(&ev1));
        FADB_CUDA_CHECK(cudaEventRecord(ev0, st));
    }
    const int bias_bytes = ((L.N * 4 + 127) / 128) * 128;
    const int budget = h->gemm_smem_budget - bias_bytes;
    FADB_REQUIRE(L.N <= 8192, "Cout=%d too large for the shared-memory bias vector", L.N);
    cudaError_t launch_err = cudaSuccess;
    auto launch = [&](const GemmParams& q) {
        GemmParams pp = q;
        // shared-memory plan for a weight K block of `b_bytes` per CTA (pair mode: half of BN rows)
        struct Plan { int halo, resb, b_stages, stages, smem; };
        auto plan = [&](int b_bytes) {
            Plan r;
            const int res = pp.nkb * b_bytes;
            if (pp.halo) {
                r.halo = 1;
                r.resb = (h->resident_b && pp.tiles_n == 1 && res + 2 * kHaloBytes <= budget) ? 1 : 0;
                if (r.resb) {
                    r.b_stages = 0;
                    r.stages = (budget - res) / kHaloBytes;
                    if (r.stages > 4) r.stages = 4;
                    r.smem = res + r.stages * kHaloBytes + GemmCfg<64>::kExtraBytes + bias_bytes;
                } else {
                    r.b_stages = (BN == 256) ? 4 : (BN == 128 ? 6 : 8);
                    r.stages = (budget - r.b_stages * b_bytes) / kHaloBytes;
                    if (r.stages > 4) r.stages = 4;
                    r.smem = r.stages * kHaloBytes + r.b_stages * b_bytes + GemmCfg<64>::kExtraBytes + bias_bytes;
                }
            } else {
                r.halo = 0;
                r.resb = 0;
                r.b_stages = 0;
                r.stages = budget / (kABytes + b_bytes);
                if (r.stages > 8) r.stages = 8;
                if (r.stages < 2) r.stages = 2;
                r.smem = r.stages * (kABytes + b_bytes) + GemmCfg<64>::kExtraBytes + bias_bytes;
            }
            return r;
        };
        auto apply = [&](const Plan& r) { pp.halo = r.halo; pp.resb = r.resb; pp.b_stages = r.b_stages; pp.stages = r.stages; };
        const int b_full = BN * kBlockK * 2;
        const Plan plain = plan(b_full), pair = plan(b_full / 2);
        apply(plain);
        int smem = plain.smem;
        // cluster mode (plain single-pass layers): CTA pairs share every weight tile through TMA multicast, which
        // halves the L2 -> SM weight traffic of the layers whose operand fetch, not the MMA, limits them
        pp.cluster = 1;
        pp.num_units = pp.num_tiles;
        int g = grid;
        const int num_m = pp.tiles_w * pp.tiles_h * pp.tiles_b;
        const int cs = (h->gemm_cluster_size == 4 && BN >= 128) ? 4 : 2;
        const int pair_units = ((num_m + cs - 1) / cs) * pp.tiles_n;
        // gemm_cluster = 1: when every CTA pair gets work; 2 (tests): whenever the layer has two M tiles;
        // 3: only layers with many M tiles per CTA
        const bool pair_mode = h->gemm_twocta && cs == 2;
        if (h->gemm_cluster && (!pp.halo || (pair_mode && h->gemm_pair_halo)) && num_m >= 2 &&
            (h->gemm_cluster == 2 || (h->gemm_cluster == 1 && pair_units >= h->sm_count / cs) ||
             (h->gemm_cluster == 3 && num_m >= 2 * h->sm_count))) {
            pp.cluster = cs;
            pp.num_units = pair_units;
            if (pair_mode) { pp.twocta = 1; apply(pair); smem = pair.smem; }
            g = cs * (h->sm_count / cs);
            if (encode_weight_map(&pp.tmBh[0], L.w_hi, L.K, L.N, BN / cs, f16, io.row_stride) != FADB_OK ||
                (npass >= 2 && !pp.lo8 && encode_weight_map(&pp.tmBh[1], L.w_lo, L.K, L.N, BN / cs, f16, io.row_stride) != FADB_OK) ||
                (pp.lo8 && encode_weight_map(&pp.tmBh[2], L.w8, L.K, L.N, BN / cs, 2) != FADB_OK)) {
                pp.cluster = 1; pp.num_units = pp.num_tiles; g = grid; pp.twocta = 0; apply(plain); smem = plain.smem;
            }
            if (npass < 2 || pp.lo8) pp.tmBh[1] = pp.tmBh[0];
            if (!pp.lo8) pp.tmBh[2] = pp.tmBh[0];
        }
        void (*kern)(GemmParams);
        void (*kern_pair)(GemmParams);
        if (f16) {
            kern = (BN == 256) ? fadb_gemm_tc_kernel<256, false, true>
                               : (BN == 128 ? fadb_gemm_tc_kernel<128, false, true> : fadb_gemm_tc_kernel<64, false, true>);
            kern_pair = (BN == 256) ? fadb_gemm_tc_kernel<256, true, true>
                                    : (BN == 128 ? fadb_gemm_tc_kernel<128, true, true> : fadb_gemm_tc_kernel<64, true, true>);
        } else {
            kern = (BN == 256) ? fadb_gemm_tc_kernel<256, false, false>
                               : (BN == 128 ? fadb_gemm_tc_kernel<128, false, false> : fadb_gemm_tc_kernel<64, false, false>);
            kern_pair = (BN == 256) ? fadb_gemm_tc_kernel<256, true, false>
                                    : (BN == 128 ? fadb_gemm_tc_kernel<128, true, false> : fadb_gemm_tc_kernel<64, true, false>);
        }
        if (io.syrk) {
            kern = fadb_gemm_tc_kernel<128, false, true, true>;
            kern_pair = fadb_gemm_tc_kernel<128, true, true, true>;
        }
        if (pp.lo8) {
            kern = (BN == 256) ? fadb_gemm_tc_kernel<256, false, true, false, true>
                               : (BN == 128 ? fadb_gemm_tc_kernel<128, false, true, false, true>
                                            : fadb_gemm_tc_kernel<64, false, true, false, true>);
            kern_pair = (BN == 256) ? fadb_gemm_tc_kernel<256, true, true, false, true>
                                    : (BN == 128 ? fadb_gemm_tc_kernel<128, true, true, false, true>
                                                 : fadb_gemm_tc_kernel<64, true, true, false, true>);
        }
        void (*kern_plain)(GemmParams) = kern;
        if (pp.twocta) kern = kern_pair;
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3((unsigned)g);
        cfg.blockDim = dim3(kThreads);
        cfg.dynamicSmemBytes = (size_t)smem;
        cfg.stream = st;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = (unsigned)pp.cluster;
        attr[0].val.clusterDim.y = 1;
        attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
        if (pp.cluster > 1) {
            // a persistent kernel wants every cluster co-resident: size the grid to what fits (a GPC with an odd
            // number of free SMs leaves one without a partner)
            // (the answer depends only on the kernel, the cluster size and the shared-memory size: cache it, the query
            // costs tens of microseconds per launch)
            struct Occ { const void* k; int device, cluster, smem, n; };
            static thread_local Occ occ_cache[16];
            static thread_local int occ_used = 0;
            int nclusters = -1;
            for (int i = 0; i < occ_used; ++i)
                if (occ_cache[i].k == (const void*)kern && occ_cache[i].device == h->device && occ_cache[i].cluster == pp.cluster &&
                    occ_cache[i].smem == smem)
                    nclusters = occ_cache[i].n;
            if (nclusters < 0) {
                nclusters = 0;
                if (cudaOccupancyMaxActiveClusters(&nclusters, kern, &cfg) != cudaSuccess) { nclusters = 0; (void)cudaGetLastError(); }
                if (occ_used < 16) occ_cache[occ_used++] = Occ{(const void*)kern, h->device, pp.cluster, smem, nclusters};
            }
            if (nclusters * pp.cluster * 10 >= h->sm_count * 9) {
                if (nclusters > pp.num_units) nclusters = pp.num_units;
                cfg.gridDim = dim3((unsigned)(pp.cluster * nclusters));
            } else {
                // no answer, or clusters would leave more than a tenth of the SMs idle: plain launch
                (void)cudaGetLastError();
                pp.cluster = 1;
                if (pp.twocta) { pp.twocta = 0; apply(plain); cfg.dynamicSmemBytes = (size_t)plain.smem; kern = kern_plain; }
                pp.num_units = pp.num_tiles;
                attr[0].val.clusterDim.x = 1;
                cfg.gridDim = dim3((unsigned)grid);
            }
            static const bool verbose = getenv("FADB_VERBOSE") != nullptr;
            if (verbose)
                fprintf(stderr, "[fadb] gemm cluster=%d: %d co-resident clusters -> grid %u (units %d, BN %d, smem %d)\n",
                        pp.cluster, nclusters, cfg.gridDim.x, pp.num_units, BN, smem);
        }
        launch_err = cudaLaunchKernelEx(&cfg, kern, pp);
        h->launches++;
    };
    // accumulation segments (GemmParams::nseg): chains of at most kSegmentBlocks K blocks (64 MMAs); in halo mode a
    // segment is a whole number of channel blocks (9 * npass K blocks each)
    p.kb0 = io.taps * p.cin_blocks;
    p.cin_blocks1 = c64 ? 1 : (lo8 ? io.Cin / 128 : p.cin_blocks);
    p.nkb = c64 ? p.kb0 + 6 : (lo8 ? p.kb0 + io.taps * p.cin_blocks1 : npass * io.taps * p.cin_blocks);
    auto split = [](int n, int kseg, int& nseg, int& len) {          // n K blocks into chains of <= kseg, evenly
        nseg = (n + kseg - 1) / kseg;
        len = (n + nseg - 1) / nseg;
        nseg = (n + len - 1) / len;
    };
    if (halo) {
        const int planes = lo8 ? 1 : npass;                          // weight planes multiplied per activation tile
        p.seg_len = kSegmentBlocks / (9 * planes) > 0 ? kSegmentBlocks / (9 * planes) : 1;
        p.nseg0 = (p.cin_blocks + p.seg_len - 1) / p.seg_len;
        p.seg_len1 = 1;
        p.nseg = p.nseg0 + (lo8 ? p.cin_blocks1 : 0);
        if (!lo8) p.nseg0 = p.nseg;
    } else {
        const int kseg = io.seg_blocks > 0 ? io.seg_blocks : kSegmentBlocks;
        if (lo8) {
            int n1 = 0;
            split(p.kb0, kseg, p.nseg0, p.seg_len);
            split(p.nkb - p.kb0, kseg, n1, p.seg_len1);
            p.nseg = p.nseg0 + n1;
        } else {
            split(p.nkb, kseg, p.nseg, p.seg_len);
            p.nseg0 = p.nseg;
            p.seg_len1 = p.seg_len;
        }
    }
    launch(p);
    FADB_CUDA_CHECK(launch_err);
    if (h->profile) {
        FADB_CUDA_CHECK(cudaEventRecord(ev1, st));
        h->prof_events.push_back(ev0);
        h->prof_events.push_back(ev1);
        h->prof_flops += 2.0 * double(io.B) * io.H * io.W * double(L.N) * double(L.K);   // algorithmic (1 pass)
    }
    FADB_CUDA_CHECK(cudaGetLastError());
    return FADB_OK;
}

}  // namespace fadb
