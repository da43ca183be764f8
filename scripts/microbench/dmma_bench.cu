// DMMA.8x8x4 throughput microbenchmark: what does the FP64 tensor pipe need to stay busy?
// Variants: warps per CTA (1 CTA/SM), independent accumulators per warp, operand pattern.
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void dmma(double& d0, double& d1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};"
                 : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}

// MODE 0: one (a,b) pair for everything.  MODE 1: 4 a's x 4 b's (like a 32x32 warp tile).
// MODE 2: like 1 but a is re-derived with integer ops before each group (conversion cost).
template <int NACC, int MODE>
__global__ void __launch_bounds__(1024, 1) k(double* out, int iters, unsigned seed) {
    double acc[NACC][2];
#pragma unroll
    for (int j = 0; j < NACC; j++) acc[j][0] = acc[j][1] = 0.0;
    double a[4], b[4];
#pragma unroll
    for (int j = 0; j < 4; j++) { a[j] = 1.0 + threadIdx.x * 1e-3 + j; b[j] = 0.5 + j * 0.25; }
    unsigned w = seed + threadIdx.x;
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int j = 0; j < NACC; j++) {
            if (MODE == 0) dmma(acc[j][0], acc[j][1], a[0], b[0]);
            else {
                if (MODE == 2 && (j & 3) == 0) {
                    unsigned g = __byte_perm(w, 0u, 0x8880u | ((j >> 2) & 3));
                    unsigned hi = (g & 0x80000000u) | ((g & 1u) * 0x3FF00000u);
                    a[(j >> 2) & 3] = __hiloint2double((int)hi, 0);
                    w = w * 1664525u + 1013904223u;
                }
                dmma(acc[j][0], acc[j][1], a[(j >> 2) & 3], b[j & 3]);
            }
        }
    }
    double s = 0;
#pragma unroll
    for (int j = 0; j < NACC; j++) s += acc[j][0] + acc[j][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int NACC, int MODE>
void run(int warps, int ctas_per_sm, double* d_out) {
    int dev = 0, sms = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int iters = 20000 / NACC * 16;
    dim3 grid(sms * ctas_per_sm), block(32 * warps);
    k<NACC, MODE><<<grid, block>>>(d_out, 10, 1);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    k<NACC, MODE><<<grid, block>>>(d_out, iters, 1);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    double flops = (double)grid.x * warps * iters * NACC * 512.0;
    printf("mode %d nacc %2d warps/cta %2d ctas/sm %d : %7.2f TFLOP/s  (%.3f ms) %s\n", MODE, NACC, warps, ctas_per_sm,
           flops / (ms * 1e-3) / 1e12, ms, cudaGetErrorString(cudaGetLastError()));
}

int main() {
    double* d_out;
    cudaMalloc(&d_out, 148 * 4 * 1024 * sizeof(double));
    for (int warps : {4, 8, 16, 32}) {
        run<16, 0>(warps, 1, d_out);
        run<16, 1>(warps, 1, d_out);
        run<16, 2>(warps, 1, d_out);
        run<32, 1>(warps, 1, d_out);
    }
    run<8, 1>(16, 1, d_out);
    run<4, 1>(16, 1, d_out);
    run<16, 1>(8, 2, d_out);
    run<16, 1>(4, 4, d_out);
    run<64, 1>(8, 1, d_out);
    return 0;
}
