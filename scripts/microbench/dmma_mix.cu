// Which instruction types steal DMMA.8x8x4 issue bandwidth?  4 DMMAs + NOPS extra ops of one type per group.
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

__device__ __forceinline__ void dmma(double& d0, double& d1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};"
                 : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}
enum { NONE, LOP, IMADMUL, PRMT, MOV, IADD, SHF, SEL, LDS, FMUL, DADD, LDS64CVT, DEPCVT, PIPECVT, DEPCVT_ALU, DEPMOV };

template <int KIND, int NOPS>
__global__ void __launch_bounds__(512, 1) k(double* out, int iters, unsigned seed) {
    __shared__ double tab[1024];  // 8 KB
    tab[threadIdx.x] = 1.0 + threadIdx.x; tab[threadIdx.x + 512] = 2.0;
    __syncthreads();
    double acc[16][2];
#pragma unroll
    for (int j = 0; j < 16; j++) acc[j][0] = acc[j][1] = 0.0;
    double a[4], b[4];
#pragma unroll
    for (int j = 0; j < 4; j++) { a[j] = 1.0 + threadIdx.x * 1e-3 + j; b[j] = 0.5 + j * 0.25; }
    unsigned x = seed + threadIdx.x, y = seed * 3 + 1, z = 12345;
    float f = 1.0f + threadIdx.x;
    double dd = 1.0;
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int g = 0; g < 4; g++) {
#pragma unroll
            for (int o = 0; o < NOPS; o++) {
                if (KIND == LOP) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(x) : "r"(y), "r"(z));
                if (KIND == IMADMUL) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(x) : "r"(y), "r"(z));
                if (KIND == PRMT) asm volatile("prmt.b32 %0, %0, %1, 0x8880;" : "+r"(x) : "r"(y));
                if (KIND == MOV) asm volatile("mov.b32 %0, %1;" : "=r"(x) : "r"(y + o));
                if (KIND == IADD) asm volatile("add.u32 %0, %0, %1;" : "+r"(x) : "r"(y));
                if (KIND == SHF) asm volatile("shf.r.wrap.b32 %0, %0, %1, 5;" : "+r"(x) : "r"(y));
                if (KIND == SEL) asm volatile("{.reg .pred p; setp.ne.u32 p, %0, 1; selp.u32 %0, %1, %2, p;}" : "+r"(x) : "r"(y), "r"(z));
                if (KIND == LDS) x += ((volatile unsigned*)tab)[(threadIdx.x + o * 32 + i) & 2047];
                if (KIND == FMUL) asm volatile("fma.rn.f32 %0, %0, %1, %1;" : "+f"(f) : "f"(1.0001f));
                if (KIND == DADD) asm volatile("add.f64 %0, %0, %1;" : "+d"(dd) : "d"(1.0));
            }
            if (KIND == LDS64CVT) {  // fetch the converted operand from shared memory instead of computing it
                a[g] = ((volatile double*)tab)[(threadIdx.x + g * 32 + i) & 1023];
            }
            if (KIND == DEPCVT) {    // convert right before use (what the scan kernel does): PRMT, LOP, IMAD, LOP + pair build
                unsigned gg = __byte_perm(x, 0u, 0x8880u | (g & 3));
                unsigned hi = (gg & 0x80000000u) | ((gg & 1u) * 0x3FF00000u);
                a[g] = __hiloint2double((int)hi, 0);
                x = x * 1664525u + 1013904223u;
            }
            if (KIND == DEPCVT_ALU) {  // same dependency, ALU-pipe ops only (no IMAD): select via LOP3
                unsigned gg = __byte_perm(x, 0u, 0x8880u | (g & 3));
                unsigned m = (unsigned)(-(int)(gg & 1u));
                unsigned hi = (gg & 0x80000000u) | (m & 0x3FF00000u);
                a[g] = __hiloint2double((int)hi, 0);
                x = (x << 3) ^ (x >> 5) ^ 0x9E3779B9u;
            }
            if (KIND == DEPMOV) {    // minimal dependency: only the high word is rewritten by one LOP3
                unsigned hi = (x & 0x80000000u) | 0x3FF00000u;
                a[g] = __hiloint2double((int)hi, __double2loint(a[g]));
                x = (x << 3) ^ (x >> 5) ^ 0x9E3779B9u;
            }
            if (KIND == PIPECVT) {   // software pipelined: use the value converted one group earlier
                a[g] = dd;
                unsigned gg = __byte_perm(x, 0u, 0x8880u | (g & 3));
                unsigned hi = (gg & 0x80000000u) | ((gg & 1u) * 0x3FF00000u);
                dd = __hiloint2double((int)hi, 0);
                x = x * 1664525u + 1013904223u;
            }
#pragma unroll
            for (int t = 0; t < 4; t++) dmma(acc[g * 4 + t][0], acc[g * 4 + t][1], a[g], b[t]);
        }
    }
    double s = (double)x + f + dd;
#pragma unroll
    for (int j = 0; j < 16; j++) s += acc[j][0] + acc[j][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int KIND, int NOPS>
void run(const char* name, double* d_out) {
    const int iters = 20000;
    dim3 grid(148), block(512);
    k<KIND, NOPS><<<grid, block>>>(d_out, 10, 1);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("%s x%d warm-up failed: %s\n", name, NOPS, cudaGetErrorString(e)); fflush(stdout); exit(1); }
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    k<KIND, NOPS><<<grid, block>>>(d_out, iters, 1);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    double flops = (double)grid.x * 16 * iters * 16 * 512.0;
    printf("%-10s x%d per 4 DMMA : %7.2f TFLOP/s  %s\n", name, NOPS, flops / (ms * 1e-3) / 1e12, cudaGetErrorString(cudaGetLastError()));
}

int main() {
    setvbuf(stdout, nullptr, _IONBF, 0);
    double* d_out;
    cudaMalloc(&d_out, 148 * 1024 * sizeof(double));
    run<NONE, 0>("none", d_out);
    run<LOP, 2>("lop3", d_out);      run<LOP, 6>("lop3", d_out);
    run<IMADMUL, 2>("imad", d_out);  run<IMADMUL, 6>("imad", d_out);
    run<PRMT, 2>("prmt", d_out);     run<PRMT, 6>("prmt", d_out);
    run<MOV, 2>("mov", d_out);       run<MOV, 6>("mov", d_out);
    run<IADD, 2>("iadd", d_out);     run<IADD, 6>("iadd", d_out);
    run<SHF, 2>("shf", d_out);       run<SHF, 6>("shf", d_out);
    run<SEL, 2>("setp+sel", d_out);  run<SEL, 6>("setp+sel", d_out);
    run<LDS, 2>("lds", d_out);       run<LDS, 6>("lds", d_out);
    run<DEPCVT, 0>("dep cvt", d_out);
    run<DEPCVT_ALU, 0>("dep cvt alu", d_out);
    run<DEPMOV, 0>("dep hi-only", d_out);
    run<PIPECVT, 0>("piped cvt", d_out);
    run<FMUL, 2>("ffma", d_out);     run<FMUL, 6>("ffma", d_out);
    run<DADD, 1>("dadd", d_out);     run<DADD, 2>("dadd", d_out);
    run<LDS64CVT, 0>("lds64 a", d_out);
    return 0;
}
