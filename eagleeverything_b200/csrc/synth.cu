// Synthetic genotype images (bench / test utility, not part of the reference's path).
// Writes a byte-exact M.ascii image (CreateASCIInospace.cpp:119-122 format) straight into device
// memory: g(i,j) ~ Binomial(2, p_j), p_j = 0.05 + 0.45 u_j, from a counter-based splitmix64 hash
// so that eagleeverything_b200/synth.py (numpy) produces the same bytes.
#include "common.cuh"

namespace eg {
__host__ __device__ __forceinline__ uint64_t splitmix64(uint64_t x) {
    x += 0x9E3779B97F4A7C15ULL;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ULL;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBULL;
    return x ^ (x >> 31);
}
__host__ __device__ __forceinline__ uint32_t marker_threshold(uint64_t seed, uint64_t j) {
    const uint64_t z = splitmix64(seed ^ 0xA5A5A5A5A5A5A5A5ULL ^ (j * 0xD1342543DE82EF95ULL));
    const double u = (double)(z >> 11) * (1.0 / 9007199254740992.0);
    const double p = 0.05 + 0.45 * u;
    return (uint32_t)(p * 4294967296.0);
}
// image: rows x (cols+1) bytes; marker (column) index of local column c is col_offset + c;
// the hash counter is j*n_total + i so shards of one data set agree with the whole.
__global__ void __launch_bounds__(256) synth_ascii_kernel(uint8_t* img, int64_t rows, int64_t cols, int64_t col_offset,
                                                          int64_t n_total, int64_t row_offset, uint64_t seed) {
    const int64_t total = rows * (cols + 1);
    for (int64_t o = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; o < total; o += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = o / (cols + 1), c = o - r * (cols + 1);
        uint8_t b = '\n';
        if (c < cols) {
            const uint64_t j = (uint64_t)(col_offset + c), i = (uint64_t)(row_offset + r);
            const uint32_t thr = marker_threshold(seed, j);
            const uint64_t z = splitmix64(seed + (j * (uint64_t)n_total + i) * 0x2545F4914F6CDD1DULL);
            b = (uint8_t)('0' + ((uint32_t)z < thr) + ((uint32_t)(z >> 32) < thr));
        }
        img[o] = b;
    }
}
}  // namespace eg

extern "C" int eg_dev_synth_ascii(uint8_t* d_img, int64_t rows, int64_t cols, int64_t col_offset, int64_t n_total,
                                  int64_t row_offset, uint64_t seed, void* stream) {
    using namespace eg;
    if (!d_img || rows <= 0 || cols <= 0) return set_error(EG_ERR_ARG, "eg_dev_synth_ascii: bad argument");
    synth_ascii_kernel<<<num_sms() * 16, 256, 0, (cudaStream_t)stream>>>(d_img, rows, cols, col_offset, n_total,
                                                                        row_offset, seed);
    return check_launch("synth_ascii_kernel");
}
