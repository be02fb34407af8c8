// resample.cu — band-limited sinc resampling on the device, with the arithmetic of resample.py (the host restatement
// of `resampy.resample(x, sr_orig, sr_new)`, filter kaiser_best) replayed operation by operation in fp64, so the two
// agree bit for bit.  Replaces the resampy calls of the reference (fad.py:159, models/vggish.py:250,
// models/pann.py:101) for clips that are already on the GPU.  PARITY UNPINNED against resampy itself (un-vendored,
// unpinned dependency, not installed here): see resample.py.
//
// One thread per output sample: time register t = j / ratio, left wing then right wing of the interpolated filter,
// table value = win[idx] + eta * (win[idx + 1] - win[idx]), every product and sum rounded separately like NumPy does.
#include "common.cuh"

namespace fadb {

__global__ void __launch_bounds__(256) resample_kernel(const float* __restrict__ in, long long n_in, long long in_stride,
                                                       double ratio, const double* __restrict__ win, int nwin,
                                                       int num_table, float* __restrict__ out, long long n_out,
                                                       long long out_stride) {
    const long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n_out) return;
    const float* x = in + (long long)blockIdx.y * in_stride;
    const double scale = ratio < 1.0 ? ratio : 1.0;
    const int index_step = (int)(scale * (double)num_table);
    const double t = __ddiv_rn((double)j, ratio);
    const long long n = (long long)t;
    const double frac0 = __dmul_rn(scale, __dsub_rn(t, (double)n));
    double acc = 0.0;
#pragma unroll
    for (int side = 0; side < 2; ++side) {
        const double frac = side == 0 ? frac0 : __dsub_rn(scale, frac0);
        const double index_frac = __dmul_rn(frac, (double)num_table);
        const long long offset = (long long)index_frac;
        const double eta = __dsub_rn(index_frac, (double)offset);
        const long long taps = ((long long)nwin - offset) / index_step;
        long long limit = side == 0 ? n + 1 : n_in - n - 1;
        if (taps < limit) limit = taps;
        for (long long i = 0; i < limit; ++i) {
            const long long idx = offset + i * index_step;
            const double w0 = __ldg(win + idx);
            const double d = (idx + 1 < nwin) ? __dsub_rn(__ldg(win + idx + 1), w0) : 0.0;
            const double w = __dadd_rn(w0, __dmul_rn(eta, d));
            const double xv = (double)__ldg(x + (side == 0 ? n - i : n + i + 1));
            acc = __dadd_rn(acc, __dmul_rn(w, xv));
        }
    }
    out[(long long)blockIdx.y * out_stride + j] = (float)acc;
}

int launch_resample(fadb_handle* h, const float* in, int64_t n_clips, int64_t n_in, int64_t in_stride, double ratio,
                    const double* win, int nwin, int num_table, float* out, int64_t n_out, int64_t out_stride,
                    cudaStream_t st) {
    FADB_REQUIRE(ratio > 0.0 && nwin > 1 && num_table > 0, "resample: bad ratio / filter table");
    FADB_REQUIRE((int)((ratio < 1.0 ? ratio : 1.0) * num_table) >= 1, "resample: ratio %g too small for the table", ratio);
    FADB_REQUIRE(n_clips <= 65535, "resample: at most 65535 clips per call");
    if (n_clips <= 0 || n_out <= 0) return FADB_OK;
    dim3 grid((unsigned)((n_out + 255) / 256), (unsigned)n_clips);
    resample_kernel<<<grid, 256, 0, st>>>(in, n_in, in_stride, ratio, win, nwin, num_table, out, n_out, out_stride);
    h->launches++;
    FADB_CUDA_CHECK(cudaGetLastError());
    return FADB_OK;
}

}  // namespace fadb
