// Probe: does a tiled tensor map whose W stride (64 B) is SMALLER than its innermost extent (128 B) encode and load?
// The e4m3 low-order pass of 64-channel layers wants halo-tile rows of 128 B = [pixel x | pixel x+1] (two taps of a
// 3x3 filter in one 128-byte K block), i.e. overlapping rows of a W-padded [B][H][W+2][64] byte tensor.
//   nvcc -gencode arch=compute_100a,code=sm_100a -I frechet_audio_distance_exported_b200/csrc tools/tma_overlap_probe.cu -o build/tma_probe -lcuda
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <vector>

#include "tc_ptx.cuh"

using namespace fadb;

__global__ void probe_kernel(const __grid_constant__ CUtensorMap tm, uint8_t* out, int bytes, int x0, int y0) {
    extern __shared__ uint8_t raw[];
    const uint32_t a = (smem_u32(raw) + 1023u) & ~1023u;
    uint8_t* sm = raw + (a - smem_u32(raw));
    uint64_t* bar = reinterpret_cast<uint64_t*>(sm + bytes);
    if (threadIdx.x == 0) {
        mbar_init(smem_u32(bar), 1);
        fence_barrier_init();
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        mbar_arrive_expect_tx(smem_u32(bar), (uint32_t)bytes);
        tma_load_4d(&tm, smem_u32(bar), a, 0, x0, y0, 0);
    }
    mbar_wait(smem_u32(bar), 0, nullptr);
    for (int i = threadIdx.x; i < bytes; i += blockDim.x) out[i] = sm[i];
}

typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main() {
    const int W = 16, H = 20, WP = W + 2;
    std::vector<uint8_t> host((size_t)H * WP * 64 + 64);
    auto val = [&](int y, int xp, int c) { return (uint8_t)((y * 37 + xp * 11 + c) & 0xff); };
    for (int y = 0; y < H; ++y)
        for (int xp = 0; xp < WP; ++xp)
            for (int c = 0; c < 64; ++c) host[((size_t)y * WP + xp) * 64 + c] = (xp == 0 || xp == WP - 1) ? 0 : val(y, xp, c);
    uint8_t *d_in, *d_out;
    const int BX = 10, BY = 18, bytes = BX * BY * 128;
    cudaMalloc(&d_in, host.size());
    cudaMalloc(&d_out, bytes);
    cudaMemcpy(d_in, host.data(), host.size(), cudaMemcpyHostToDevice);
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
    PFN_encodeTiled enc = (PFN_encodeTiled)fn;
    CUtensorMap tm;
    cuuint64_t dims[4] = {128, (cuuint64_t)(W + 1), (cuuint64_t)H, 1};
    cuuint64_t strides[3] = {64, (cuuint64_t)WP * 64, (cuuint64_t)H * WP * 64};
    cuuint32_t box[4] = {128, BX, BY, 1};
    cuuint32_t es[4] = {1, 1, 1, 1};
    CUresult r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_UINT8, 4, d_in, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("encode (stride 64 < extent 128): %d\n", (int)r);
    if (r != CUDA_SUCCESS) return 1;
    cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes + 2048);
    int bad_total = 0;
    for (int tcase = 0; tcase < 3; ++tcase) {
        const int x0 = tcase == 0 ? 0 : 8, y0 = tcase == 2 ? 15 : -1;     // halo boxes of tiles (0,0), (1,0), (1,1)
        cudaMemset(d_out, 0xEE, bytes);
        probe_kernel<<<1, 128, bytes + 2048>>>(tm, d_out, bytes, x0, y0);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("kernel: %s\n", cudaGetErrorString(e)); return 1; }
        std::vector<uint8_t> got(bytes);
        cudaMemcpy(got.data(), d_out, bytes, cudaMemcpyDeviceToHost);
        int bad = 0;
        for (int j = 0; j < BY; ++j)
            for (int i = 0; i < BX; ++i) {
                const int row = j * BX + i;                      // smem row (128 B), swizzle: 16-B chunk ^= row & 7
                const int y = y0 + j, r0 = x0 + i;               // map row r0 = padded pixels r0, r0 + 1
                for (int b = 0; b < 128; ++b) {
                    const int chunk = (b >> 4) ^ (row & 7);
                    const uint8_t g = got[row * 128 + chunk * 16 + (b & 15)];
                    uint8_t want = 0;
                    const int xp = r0 + (b >> 6), c = b & 63;
                    if (y >= 0 && y < H && r0 <= W && xp > 0 && xp < WP - 1) want = val(y, xp, c);
                    if (g != want && bad++ < 5) printf("case %d row (%d,%d) byte %d: got %d want %d\n", tcase, j, i, b, g, want);
                }
            }
        printf("case %d (x0 %d, y0 %d): %d mismatching bytes\n", tcase, x0, y0, bad);
        bad_total += bad;
    }
    printf(bad_total ? "PROBE FAILED\n" : "PROBE OK\n");
    return bad_total ? 1 : 0;
}
