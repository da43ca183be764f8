// SURVEY.md section 8(f), rank 1 -- the n x n FP64 algebra that feeds the scan each forward iteration:
//   calculateMMt_sqrt_and_sqrtinv   R/calculateMMt_sqrt_and_sqrtinv.R:14-31   K^(1/2) = U diag(sqrt(l)) U^T, K^(-1/2) = chol2inv(chol(K^(1/2)))
//   calculateH                      R/calculateH.R:36                         H = varE I + varG K
//   calculateP                      R/calculateP.R:27-28                      P = Hinv - Hinv X (X^T Hinv X)^-1 X^T Hinv, Hinv = chol2inv(chol(H))
//   calculate_reduced_a             R/calculate_reduced_a.R:31                a = varG K^(1/2) P y
//   calculate_reduced_vara          R/calculate_reduced_vara.R:21-35          V = varG I - (D1 + D1 C (A - B D1 C)^-1 B D1)
// In the reference these are base-R LAPACK / BLAS calls on the host; once the scan takes a fraction of a second they
// dominate an AM() iteration at n >= 10k.  Here: dense factorizations are cuSOLVER library calls (dsyevd, dpotrf,
// dpotri, dgetrf/dgetrs for the q x q pieces), products are cuBLAS, and everything stays on the device between them.
// Library calls on purpose: these are plain LAPACK-shaped operations, not the data-parallel hot path.
#include <cublas_v2.h>
#include <cusolverDn.h>

#include <cmath>
#include <cstring>
#include <vector>

#include "common.cuh"
#include "hostio.cuh"

namespace eg {

int ensure_init_pub();
cudaStream_t ctx_stream();
cublasHandle_t ctx_cublas();

struct AlgCtx {
    cusolverDnHandle_t solver = nullptr;
    double* work = nullptr;
    size_t work_cap = 0;  // doubles
    int* d_info = nullptr;
};
static thread_local AlgCtx g_alg;

void algebra_release() {
    if (g_alg.solver) cusolverDnDestroy(g_alg.solver);
    cudaFree(g_alg.work);
    cudaFree(g_alg.d_info);
    g_alg = AlgCtx();
}
static int alg_init(cudaStream_t st) {
    if (!g_alg.solver) {
        if (cusolverDnCreate(&g_alg.solver) != CUSOLVER_STATUS_SUCCESS) return set_error(EG_ERR_CUDA, "cusolverDnCreate failed");
        EG_CUDA(cudaMalloc(&g_alg.d_info, sizeof(int)));
    }
    if (cusolverDnSetStream(g_alg.solver, st) != CUSOLVER_STATUS_SUCCESS) return set_error(EG_ERR_CUDA, "cusolverDnSetStream");
    if (cublasSetStream(ctx_cublas(), st) != CUBLAS_STATUS_SUCCESS) return set_error(EG_ERR_CUDA, "cublasSetStream");
    return EG_OK;
}
static int alg_work(size_t doubles) {
    if (doubles <= g_alg.work_cap && g_alg.work) return EG_OK;
    cudaFree(g_alg.work);
    g_alg.work = nullptr;
    g_alg.work_cap = 0;
    if (malloc_retry((void**)&g_alg.work, (doubles ? doubles : 1) * sizeof(double)) != cudaSuccess) {
        cudaGetLastError();
        return set_error(EG_ERR_ALLOC, "out of device memory for the LAPACK workspace (%zu bytes)", doubles * 8);
    }
    g_alg.work_cap = doubles;
    return EG_OK;
}
static int alg_info(const char* what, cudaStream_t st) {
    int h = 0;
    EG_CUDA(cudaMemcpyAsync(&h, g_alg.d_info, sizeof(int), cudaMemcpyDeviceToHost, st));
    EG_CUDA(cudaStreamSynchronize(st));
    if (h != 0) return set_error(EG_ERR_ARG, "%s failed (LAPACK info = %d)", what, h);
    return EG_OK;
}
#define EG_BLAS(expr)                                                                    \
    do {                                                                                 \
        if ((expr) != CUBLAS_STATUS_SUCCESS) return set_error(EG_ERR_CUDA, "%s failed", #expr); \
    } while (0)
#define EG_SOLVER(expr)                                                                    \
    do {                                                                                   \
        if ((expr) != CUSOLVER_STATUS_SUCCESS) return set_error(EG_ERR_CUDA, "%s failed", #expr); \
    } while (0)

// ------------------------------------------------------------------ small kernels
// T[:, j] = U[:, j] * sqrt(w[j])
__global__ void __launch_bounds__(256) scale_cols_sqrt_kernel(const double* __restrict__ U, const double* __restrict__ w, int64_t n,
                                                              double* __restrict__ T) {
    const int64_t j = blockIdx.y;
    const double s = sqrt(w[j]);
    for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (int64_t)gridDim.x * 256) T[i + j * n] = U[i + j * n] * s;
}
// copy the upper triangle of a column-major matrix onto its lower triangle (what chol2inv returns to R)
__global__ void __launch_bounds__(256) mirror_upper_kernel(double* __restrict__ A, int64_t n) {
    __shared__ double tile[32][33];
    const int bx = blockIdx.x, by = blockIdx.y;  // source tile: rows by*32.., cols bx*32.. with bx >= by
    if (bx < by) return;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int64_t r0 = (int64_t)by * 32, c0 = (int64_t)bx * 32;
    for (int i = ty; i < 32; i += 8) {
        const int64_t r = r0 + tx, c = c0 + i;
        tile[i][tx] = (r < n && c < n) ? A[r + c * n] : 0.0;
    }
    __syncthreads();
    for (int i = ty; i < 32; i += 8) {
        const int64_t r = c0 + tx, c = r0 + i;  // element (r, c) of the mirrored tile = source (c, r)
        if (r < n && c < n && r > c) A[r + c * n] = tile[tx][i];
    }
}
// out = alpha * A + beta * I   (calculateH.R:36 with A = K;   -D1 + varG I)
__global__ void __launch_bounds__(256) axpby_eye_kernel(const double* __restrict__ A, int64_t n, double alpha, double beta,
                                                        double* __restrict__ out) {
    const int64_t j = blockIdx.y;
    for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (int64_t)gridDim.x * 256)
        out[i + j * n] = __dadd_rn(__dmul_rn(alpha, A[i + j * n]), i == j ? beta : 0.0);  // two roundings, as R's varE*I + varG*K
}
__global__ void __launch_bounds__(256) eye_kernel(double* __restrict__ A, int q) {
    const int t = blockIdx.x * 256 + threadIdx.x;
    if (t < q * q) A[t] = (t % q == t / q) ? 1.0 : 0.0;
}
// eigenvalue screen of matrixcalc::is.positive.definite: |ev| < tol counts as 0, all must be > 0
__global__ void pd_screen_kernel(const double* __restrict__ w, int64_t n, double tol, int* __restrict__ not_pd) {
    int bad = 0;
    for (int64_t i = threadIdx.x; i < n; i += blockDim.x) {
        const double v = fabs(w[i]) < tol ? 0.0 : w[i];
        if (!(v > 0.0)) bad = 1;
    }
    if (bad) atomicExch(not_pd, 1);
}

// eigen(): LAPACK returns ascending eigenvalues, R descending: reverse values and the columns of the vectors in place
__global__ void __launch_bounds__(256) reverse_cols_kernel(double* __restrict__ U, double* __restrict__ w, int64_t n) {
    const int64_t j = blockIdx.y;  // 0 .. n/2 - 1
    const int64_t jj = n - 1 - j;
    for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (int64_t)gridDim.x * 256) {
        const double a = U[i + j * n], b = U[i + jj * n];
        U[i + j * n] = b;
        U[i + jj * n] = a;
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        const double a = w[j], b = w[jj];
        w[j] = b;
        w[jj] = a;
    }
}
__global__ void __launch_bounds__(256) add_scalar_kernel(double* __restrict__ v, int64_t n, double c) {
    const int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x;
    if (i < n) v[i] += c;
}

static dim3 grid_cols(int64_t n) { return dim3((unsigned)((n + 255) / 256 < 64 ? (n + 255) / 256 : 64), (unsigned)n); }

// A (n x n, SPD, column-major) -> its inverse, full symmetric, in place: chol2inv(chol(A))
static int spd_inverse_inplace(double* d_A, int64_t n, cudaStream_t st, const char* what) {
    int lw1 = 0, lw2 = 0;
    EG_SOLVER(cusolverDnDpotrf_bufferSize(g_alg.solver, CUBLAS_FILL_MODE_UPPER, (int)n, d_A, (int)n, &lw1));
    EG_SOLVER(cusolverDnDpotri_bufferSize(g_alg.solver, CUBLAS_FILL_MODE_UPPER, (int)n, d_A, (int)n, &lw2));
    const int lw = lw1 > lw2 ? lw1 : lw2;
    EG_TRY(alg_work((size_t)lw));
    EG_SOLVER(cusolverDnDpotrf(g_alg.solver, CUBLAS_FILL_MODE_UPPER, (int)n, d_A, (int)n, g_alg.work, lw, g_alg.d_info));
    {
        int h = 0;
        EG_CUDA(cudaMemcpyAsync(&h, g_alg.d_info, sizeof(int), cudaMemcpyDeviceToHost, st));
        EG_CUDA(cudaStreamSynchronize(st));
        if (h != 0)  // R: "the leading minor of order k is not positive definite"
            return set_error(EG_ERR_ARG, "%s: the leading minor of order %d is not positive definite", what, h);
    }
    EG_SOLVER(cusolverDnDpotri(g_alg.solver, CUBLAS_FILL_MODE_UPPER, (int)n, d_A, (int)n, g_alg.work, lw, g_alg.d_info));
    EG_TRY(alg_info("dpotri", st));
    const unsigned nb = (unsigned)((n + 31) / 32);
    mirror_upper_kernel<<<dim3(nb, nb), 256, 0, st>>>(d_A, n);
    return check_launch("mirror_upper_kernel");
}
// q x q general inverse in place (R's solve(A)): LU with partial pivoting
static int small_inverse(double* d_A, int q, double* d_out, cudaStream_t st, const char* what) {
    int lw = 0;
    EG_SOLVER(cusolverDnDgetrf_bufferSize(g_alg.solver, q, q, d_A, q, &lw));
    EG_TRY(alg_work((size_t)lw + (size_t)q + 8));
    int* d_piv = reinterpret_cast<int*>(g_alg.work + lw);
    EG_SOLVER(cusolverDnDgetrf(g_alg.solver, q, q, d_A, q, g_alg.work, d_piv, g_alg.d_info));
    {
        int h = 0;
        EG_CUDA(cudaMemcpyAsync(&h, g_alg.d_info, sizeof(int), cudaMemcpyDeviceToHost, st));
        EG_CUDA(cudaStreamSynchronize(st));
        if (h != 0) return set_error(EG_ERR_ARG, "%s: Lapack routine dgesv: system is exactly singular: U[%d,%d] = 0", what, h, h);
    }
    eye_kernel<<<(q * q + 255) / 256, 256, 0, st>>>(d_out, q);
    EG_TRY(check_launch("eye_kernel"));
    EG_SOLVER(cusolverDnDgetrs(g_alg.solver, CUBLAS_OP_N, q, q, d_A, q, d_piv, d_out, q, g_alg.d_info));
    return alg_info("dgetrs", st);
}

}  // namespace eg

using namespace eg;

// ================================================================== device level
// d_K: n x n column-major, exactly symmetric.  d_sqrt, d_invsqrt: outputs.  d_tmp: n*n doubles of scratch.
// *not_pd = 1 (and nothing else written) when K fails matrixcalc::is.positive.definite's eigenvalue screen.
// *trace_check = trace(sqrt * invsqrt), the quantity R's checkres compares with n.
extern "C" int eg_dev_sqrt_and_sqrtinv(const double* d_K, int64_t n, double* d_sqrt, double* d_invsqrt, double* d_tmp,
                                       int* not_pd, double* trace_check, void* stream) {
    if (!d_K || !d_sqrt || !d_invsqrt || !d_tmp || !not_pd || n <= 0 || n > 46000)
        return set_error(EG_ERR_ARG, "eg_dev_sqrt_and_sqrtinv: bad argument");
    EG_TRY(ensure_init_pub());
    cudaStream_t st = (cudaStream_t)stream;
    EG_TRY(alg_init(st));
    *not_pd = 0;
    double mabs = 0, masym = 0;
    EG_TRY(eg_dev_symmetry(d_K, n, &mabs, &masym, stream));
    if (masym != 0.0) return set_error(EG_ERR_ARG, "argument x is not a symmetric matrix");  // matrixcalc's stop()
    // eigen(K, symmetric = TRUE): vectors in d_invsqrt (used as U), values (ascending) at the head of the workspace
    double* U = d_invsqrt;
    EG_CUDA(cudaMemcpyAsync(U, d_K, (size_t)n * n * 8, cudaMemcpyDeviceToDevice, st));
    int lw = 0;
    EG_SOLVER(cusolverDnDsyevd_bufferSize(g_alg.solver, CUSOLVER_EIG_MODE_VECTOR, CUBLAS_FILL_MODE_LOWER, (int)n, U, (int)n,
                                          nullptr, &lw));
    EG_TRY(alg_work((size_t)lw + (size_t)n + 8));
    double* w = g_alg.work + lw;
    EG_SOLVER(cusolverDnDsyevd(g_alg.solver, CUSOLVER_EIG_MODE_VECTOR, CUBLAS_FILL_MODE_LOWER, (int)n, U, (int)n, w, g_alg.work,
                               lw, g_alg.d_info));
    EG_TRY(alg_info("dsyevd", st));
    int* d_flag = g_alg.d_info;
    EG_CUDA(cudaMemsetAsync(d_flag, 0, sizeof(int), st));
    pd_screen_kernel<<<1, 1024, 0, st>>>(w, n, 1e-8, d_flag);
    EG_TRY(check_launch("pd_screen_kernel"));
    EG_CUDA(cudaMemcpyAsync(not_pd, d_flag, sizeof(int), cudaMemcpyDeviceToHost, st));
    EG_CUDA(cudaStreamSynchronize(st));
    if (*not_pd) return EG_OK;
    // sqrt = (U diag(sqrt w)) U^T      (the order of the eigenpairs does not matter for this product)
    scale_cols_sqrt_kernel<<<grid_cols(n), 256, 0, st>>>(U, w, n, d_tmp);
    EG_TRY(check_launch("scale_cols_sqrt_kernel"));
    const double one = 1.0, zero = 0.0;
    EG_BLAS(cublasDgemm(ctx_cublas(), CUBLAS_OP_N, CUBLAS_OP_T, (int)n, (int)n, (int)n, &one, d_tmp, (int)n, U, (int)n, &zero,
                        d_sqrt, (int)n));
    // invsqrt = chol2inv(chol(sqrt))
    EG_CUDA(cudaMemcpyAsync(d_invsqrt, d_sqrt, (size_t)n * n * 8, cudaMemcpyDeviceToDevice, st));
    EG_TRY(spd_inverse_inplace(d_invsqrt, n, st, "chol(sqrt_MMt)"));
    if (trace_check) {
        // trace(A B) = sum_ij A_ij B_ji; B is exactly symmetric after the mirror
        double t = 0.0;
        double acc = 0.0;
        const int64_t total = n * n, chunk = (int64_t)1 << 30;
        for (int64_t o = 0; o < total; o += chunk) {
            const int64_t len = total - o < chunk ? total - o : chunk;
            EG_BLAS(cublasDdot(ctx_cublas(), (int)len, d_sqrt + o, 1, d_invsqrt + o, 1, &t));
            acc += t;
        }
        *trace_check = acc;
    }
    return EG_OK;
}

extern "C" int eg_dev_calculateH(const double* d_K, int64_t n, double varE, double varG, double* d_H, void* stream) {
    if (!d_K || !d_H || n <= 0) return set_error(EG_ERR_ARG, "eg_dev_calculateH: bad argument");
    axpby_eye_kernel<<<grid_cols(n), 256, 0, (cudaStream_t)stream>>>(d_K, n, varG, varE, d_H);
    return check_launch("axpby_eye_kernel");
}

// d_X: n x q column-major.  d_P: n x n output.  d_small: at least 2*n*q + 2*q*q doubles of scratch.
extern "C" int eg_dev_calculateP(const double* d_H, const double* d_X, int64_t n, int q, double* d_P, double* d_small,
                                 void* stream) {
    if (!d_H || !d_X || !d_P || !d_small || n <= 0 || q <= 0 || n > 46000)
        return set_error(EG_ERR_ARG, "eg_dev_calculateP: bad argument");
    EG_TRY(ensure_init_pub());
    cudaStream_t st = (cudaStream_t)stream;
    EG_TRY(alg_init(st));
    const double one = 1.0, zero = 0.0, minus = -1.0;
    double *B = d_small, *T1 = B + (size_t)n * q, *Cq = T1 + (size_t)n * q, *Ci = Cq + (size_t)q * q;
    // Hinv = chol2inv(chol(H))   (in d_P)
    EG_CUDA(cudaMemcpyAsync(d_P, d_H, (size_t)n * n * 8, cudaMemcpyDeviceToDevice, st));
    EG_TRY(spd_inverse_inplace(d_P, n, st, "chol(H)"));
    // B = Hinv X;  C = X^T B;  P = Hinv - B C^-1 B^T
    EG_BLAS(cublasDgemm(ctx_cublas(), CUBLAS_OP_N, CUBLAS_OP_N, (int)n, q, (int)n, &one, d_P, (int)n, d_X, (int)n, &zero, B, (int)n));
    EG_BLAS(cublasDgemm(ctx_cublas(), CUBLAS_OP_T, CUBLAS_OP_N, q, q, (int)n, &one, d_X, (int)n, B, (int)n, &zero, Cq, q));
    EG_TRY(small_inverse(Cq, q, Ci, st, "solve(t(X) %*% Hinv %*% X)"));
    EG_BLAS(cublasDgemm(ctx_cublas(), CUBLAS_OP_N, CUBLAS_OP_N, (int)n, q, q, &one, B, (int)n, Ci, q, &zero, T1, (int)n));
    EG_BLAS(cublasDgemm(ctx_cublas(), CUBLAS_OP_N, CUBLAS_OP_T, (int)n, (int)n, q, &minus, T1, (int)n, B, (int)n, &one, d_P, (int)n));
    return EG_OK;
}

// out = varG * sqrt * (P * y): the same vector as R's left-associative (varG * sqrt %*% P) %*% y without its n^3 product
extern "C" int eg_dev_calculate_reduced_a(double varG, const double* d_P, const double* d_sqrt, const double* d_y, int64_t n,
                                          double* d_tmp_n, double* d_out, void* stream) {
    if (!d_P || !d_sqrt || !d_y || !d_tmp_n || !d_out || n <= 0) return set_error(EG_ERR_ARG, "eg_dev_calculate_reduced_a: bad argument");
    EG_TRY(ensure_init_pub());
    cudaStream_t st = (cudaStream_t)stream;
    EG_TRY(alg_init(st));
    const double one = 1.0, zero = 0.0;
    EG_BLAS(cublasDgemv(ctx_cublas(), CUBLAS_OP_N, (int)n, (int)n, &one, d_P, (int)n, d_y, 1, &zero, d_tmp_n, 1));
    EG_BLAS(cublasDgemv(ctx_cublas(), CUBLAS_OP_N, (int)n, (int)n, &varG, d_sqrt, (int)n, d_tmp_n, 1, &zero, d_out, 1));
    return EG_OK;
}

// d_V: n x n output.  d_D: n x n scratch.  d_small: at least 4*n*q + 3*q*q doubles.
extern "C" int eg_dev_calculate_reduced_vara(const double* d_X, int q, double varE, double varG, const double* d_sqrt, int64_t n,
                                             double* d_V, double* d_D, double* d_small, void* stream) {
    if (!d_X || !d_sqrt || !d_V || !d_D || !d_small || n <= 0 || q <= 0 || n > 46000 || !(varE != 0.0) || !(varG != 0.0))
        return set_error(EG_ERR_ARG, "eg_dev_calculate_reduced_vara: bad argument");
    EG_TRY(ensure_init_pub());
    cudaStream_t st = (cudaStream_t)stream;
    EG_TRY(alg_init(st));
    const double one = 1.0, zero = 0.0, minus = -1.0, r1 = 1.0 / varE, g1 = 1.0 / varG;
    double *Bm = d_small, *Cm = Bm + (size_t)n * q, *E = Cm + (size_t)n * q, *T2 = E + (size_t)n * q, *Aq = T2 + (size_t)n * q,
           *Sq = Aq + (size_t)q * q, *Si = Sq + (size_t)q * q;
    // A = X^T R1 X (q x q);  B = X^T R1 Ze (q x n);  C = Ze^T R1 X (n x q);  D = Ze^T R1 Ze + G1
    EG_BLAS(cublasDgemm(ctx_cublas(), CUBLAS_OP_T, CUBLAS_OP_N, q, q, (int)n, &r1, d_X, (int)n, d_X, (int)n, &zero, Aq, q));
    EG_BLAS(cublasDgemm(ctx_cublas(), CUBLAS_OP_T, CUBLAS_OP_N, q, (int)n, (int)n, &r1, d_X, (int)n, d_sqrt, (int)n, &zero, Bm, q));
    EG_BLAS(cublasDgemm(ctx_cublas(), CUBLAS_OP_T, CUBLAS_OP_N, (int)n, q, (int)n, &r1, d_sqrt, (int)n, d_X, (int)n, &zero, Cm, (int)n));
    EG_BLAS(cublasDgemm(ctx_cublas(), CUBLAS_OP_T, CUBLAS_OP_N, (int)n, (int)n, (int)n, &r1, d_sqrt, (int)n, d_sqrt, (int)n, &zero,
                        d_V, (int)n));
    axpby_eye_kernel<<<grid_cols(n), 256, 0, st>>>(d_V, n, 1.0, g1, d_D);
    EG_TRY(check_launch("axpby_eye_kernel"));
    // D1 = solve(D): D is symmetric positive definite (a Gram matrix plus a positive diagonal)
    EG_TRY(spd_inverse_inplace(d_D, n, st, "solve(D)"));
    // E = D1 C;  S = A - B E;  V = varG I - D1 - E S^-1 (B D1)
    EG_BLAS(cublasDgemm(ctx_cublas(), CUBLAS_OP_N, CUBLAS_OP_N, (int)n, q, (int)n, &one, d_D, (int)n, Cm, (int)n, &zero, E, (int)n));
    EG_CUDA(cudaMemcpyAsync(Sq, Aq, (size_t)q * q * 8, cudaMemcpyDeviceToDevice, st));
    EG_BLAS(cublasDgemm(ctx_cublas(), CUBLAS_OP_N, CUBLAS_OP_N, q, q, (int)n, &minus, Bm, q, E, (int)n, &one, Sq, q));
    EG_TRY(small_inverse(Sq, q, Si, st, "solve(A - B %*% D1 %*% C)"));
    EG_BLAS(cublasDgemm(ctx_cublas(), CUBLAS_OP_N, CUBLAS_OP_N, (int)n, q, q, &one, E, (int)n, Si, q, &zero, T2, (int)n));
    // F = B D1 (q x n), kept in Cm's place as F^T would need a transpose: compute F directly
    double* F = Cm;  // q x n, ld q   (Cm is no longer needed)
    EG_BLAS(cublasDgemm(ctx_cublas(), CUBLAS_OP_N, CUBLAS_OP_N, q, (int)n, (int)n, &one, Bm, q, d_D, (int)n, &zero, F, q));
    axpby_eye_kernel<<<grid_cols(n), 256, 0, st>>>(d_D, n, -1.0, varG, d_V);
    EG_TRY(check_launch("axpby_eye_kernel"));
    EG_BLAS(cublasDgemm(ctx_cublas(), CUBLAS_OP_N, CUBLAS_OP_N, (int)n, (int)n, q, &minus, T2, (int)n, F, q, &one, d_V, (int)n));
    return EG_OK;
}

// eigen(A, symmetric = TRUE) as R returns it: values in DEcreasing order, vectors in the columns of d_A (in place; only
// the lower triangle of A is read).  d_values: n doubles.
extern "C" int eg_dev_eigen_sym(double* d_A, int64_t n, double* d_values, void* stream) {
    if (!d_A || !d_values || n <= 0 || n > 46000) return set_error(EG_ERR_ARG, "eg_dev_eigen_sym: bad argument");
    EG_TRY(ensure_init_pub());
    cudaStream_t st = (cudaStream_t)stream;
    EG_TRY(alg_init(st));
    int lw = 0;
    EG_SOLVER(cusolverDnDsyevd_bufferSize(g_alg.solver, CUSOLVER_EIG_MODE_VECTOR, CUBLAS_FILL_MODE_LOWER, (int)n, d_A, (int)n,
                                          d_values, &lw));
    EG_TRY(alg_work((size_t)lw));
    EG_SOLVER(cusolverDnDsyevd(g_alg.solver, CUSOLVER_EIG_MODE_VECTOR, CUBLAS_FILL_MODE_LOWER, (int)n, d_A, (int)n, d_values,
                               g_alg.work, lw, g_alg.d_info));
    EG_TRY(alg_info("dsyevd", st));
    if (n > 1) {
        reverse_cols_kernel<<<dim3((unsigned)((n + 255) / 256 < 64 ? (n + 255) / 256 : 64), (unsigned)(n / 2)), 256, 0, st>>>(d_A, d_values, n);
        EG_TRY(check_launch("reverse_cols_kernel"));
    }
    return EG_OK;
}

// d_out = S (K + I) S with S = I - X (X^T X)^-1 X^T   (R/emma_eigen_R_wo_Z.R:7-13).  d_tmp: n*n, d_small: 2*n*q + 2*q*q.
extern "C" int eg_dev_emma_SKS(const double* d_K, const double* d_X, int64_t n, int q, double* d_out, double* d_tmp,
                               double* d_small, void* stream) {
    if (!d_K || !d_X || !d_out || !d_tmp || !d_small || n <= 0 || q <= 0) return set_error(EG_ERR_ARG, "eg_dev_emma_SKS: bad argument");
    EG_TRY(ensure_init_pub());
    cudaStream_t st = (cudaStream_t)stream;
    EG_TRY(alg_init(st));
    const double one = 1.0, zero = 0.0, minus = -1.0;
    double *B = d_small, *XtX = B + (size_t)n * q, *Xi = XtX + (size_t)q * q;
    EG_BLAS(cublasDgemm(ctx_cublas(), CUBLAS_OP_T, CUBLAS_OP_N, q, q, (int)n, &one, d_X, (int)n, d_X, (int)n, &zero, XtX, q));
    EG_TRY(small_inverse(XtX, q, Xi, st, "solve(crossprod(X, X))"));
    EG_BLAS(cublasDgemm(ctx_cublas(), CUBLAS_OP_N, CUBLAS_OP_N, (int)n, q, q, &one, d_X, (int)n, Xi, q, &zero, B, (int)n));
    // S in d_out:  I - B X^T
    EG_CUDA(cudaMemsetAsync(d_tmp, 0, (size_t)n * n * 8, st));
    axpby_eye_kernel<<<grid_cols(n), 256, 0, st>>>(d_tmp, n, 0.0, 1.0, d_out);
    EG_TRY(check_launch("axpby_eye_kernel"));
    EG_BLAS(cublasDgemm(ctx_cublas(), CUBLAS_OP_N, CUBLAS_OP_T, (int)n, (int)n, q, &minus, B, (int)n, d_X, (int)n, &one, d_out, (int)n));
    // d_tmp = K + I;  then (S (K+I)) S with S kept in d_out until the last product
    axpby_eye_kernel<<<grid_cols(n), 256, 0, st>>>(d_K, n, 1.0, 1.0, d_tmp);
    EG_TRY(check_launch("axpby_eye_kernel"));
    return EG_OK;
}

// Device-resident emma.eigen.R.wo.Z (R/emma_eigen_R_wo_Z.R:4-20) plus the projection EMMA takes of it: d_U (n x n)
// receives the eigenvectors of S (K + I) S in its columns (all n; the first n - q are the ones R keeps), d_values[0, n-q)
// the eigenvalues minus 1, and, when d_y / d_etas are given, d_etas[0, n-q) = U[, 1:(n-q)]^T y (R/emma_REMLE.R:40).
// d_w1, d_w2: n*n doubles of scratch each; d_small: 2*n*q + 2*q*q doubles.
extern "C" int eg_dev_emma_eigen_R_wo_Z(const double* d_K, const double* d_X, const double* d_y, int64_t n, int q, double* d_values,
                                        double* d_etas, double* d_U, double* d_w1, double* d_w2, double* d_small, void* stream) {
    if (!d_K || !d_X || !d_values || !d_U || !d_w1 || !d_w2 || !d_small || n <= 0 || q <= 0 || q >= n || (!d_y != !d_etas))
        return set_error(EG_ERR_ARG, "eg_dev_emma_eigen_R_wo_Z: bad argument");
    cudaStream_t st = (cudaStream_t)stream;
    EG_TRY(ensure_init_pub());
    EG_TRY(alg_init(st));
    // S A S with S = I - B X^T, B = X (X^T X)^-1 and A = K + I, expanded:  A - B (A X)^T - (A X) B^T + B (X^T A X) B^T.
    // Four products with inner dimension q instead of the two n^3 products of R's S %*% (K + I) %*% S (4 n^3 FP64 flops per
    // iteration on a GPU whose FP64 rate is the scarce resource); the host-level eg_emma_eigen_R_wo_Z keeps R's form.
    const double one = 1.0, zero = 0.0, minus = -1.0;
    // scratch: d_small = B (n q) | XtX (q q) | Xi (q q) | AX (n q);  C (q q) in d_w1, BC (n q) in d_w2 -- each inside its
    // documented size for every q < n (a layout with all three in d_w1 overran it from q ~ 0.41 n)
    double *B = d_small, *XtX = B + (size_t)n * q, *Xi = XtX + (size_t)q * q;
    double *AX = Xi + (size_t)q * q, *C = d_w1, *BC = d_w2;
    EG_BLAS(cublasDgemm(ctx_cublas(), CUBLAS_OP_T, CUBLAS_OP_N, q, q, (int)n, &one, d_X, (int)n, d_X, (int)n, &zero, XtX, q));
    EG_TRY(small_inverse(XtX, q, Xi, st, "solve(crossprod(X, X))"));
    EG_BLAS(cublasDgemm(ctx_cublas(), CUBLAS_OP_N, CUBLAS_OP_N, (int)n, q, q, &one, d_X, (int)n, Xi, q, &zero, B, (int)n));
    axpby_eye_kernel<<<grid_cols(n), 256, 0, st>>>(d_K, n, 1.0, 1.0, d_U);  // A = K + I
    EG_TRY(check_launch("axpby_eye_kernel"));
    EG_BLAS(cublasDgemm(ctx_cublas(), CUBLAS_OP_N, CUBLAS_OP_N, (int)n, q, (int)n, &one, d_U, (int)n, d_X, (int)n, &zero, AX, (int)n));
    EG_BLAS(cublasDgemm(ctx_cublas(), CUBLAS_OP_T, CUBLAS_OP_N, q, q, (int)n, &one, d_X, (int)n, AX, (int)n, &zero, C, q));
    EG_BLAS(cublasDgemm(ctx_cublas(), CUBLAS_OP_N, CUBLAS_OP_N, (int)n, q, q, &one, B, (int)n, C, q, &zero, BC, (int)n));
    EG_BLAS(cublasDgemm(ctx_cublas(), CUBLAS_OP_N, CUBLAS_OP_T, (int)n, (int)n, q, &minus, B, (int)n, AX, (int)n, &one, d_U, (int)n));
    EG_BLAS(cublasDgemm(ctx_cublas(), CUBLAS_OP_N, CUBLAS_OP_T, (int)n, (int)n, q, &minus, AX, (int)n, B, (int)n, &one, d_U, (int)n));
    EG_BLAS(cublasDgemm(ctx_cublas(), CUBLAS_OP_N, CUBLAS_OP_T, (int)n, (int)n, q, &one, BC, (int)n, B, (int)n, &one, d_U, (int)n));
    EG_TRY(eg_dev_eigen_sym(d_U, n, d_values, st));
    add_scalar_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(d_values, n - q, -1.0);
    EG_TRY(check_launch("add_scalar_kernel"));
    if (d_y)
        EG_BLAS(cublasDgemv(ctx_cublas(), CUBLAS_OP_T, (int)n, (int)(n - q), &one, d_U, (int)n, d_y, 1, &zero, d_etas, 1));
    return EG_OK;
}

// ================================================================== host level (what an R-facing glue binds)
namespace {
struct HostDev {
    double* p = nullptr;
    ~HostDev() { cudaFree(p); }
    int alloc(size_t doubles, const char* what) {
        if (malloc_retry((void**)&p, (doubles ? doubles : 1) * 8) != cudaSuccess) {
            cudaGetLastError();
            p = nullptr;
            return set_error(EG_ERR_ALLOC, "out of device memory allocating %zu bytes for %s", doubles * 8, what);
        }
        return EG_OK;
    }
};
// R-owned matrices are pageable: staged through the page-locked ring with parallel copier threads (hostio.cu)
int up(double* d, const double* h, size_t doubles, cudaStream_t st) { return h2d_staged(d, h, doubles * 8, st); }
int down(double* h, const double* d, size_t doubles, cudaStream_t st) { return d2h_staged(h, d, doubles * 8, st); }
void say(eg_message_fn message, void* ctx, const char* text) {
    if (message) message(ctx, text);
}
}  // namespace

extern "C" int eg_calculateMMt_sqrt_and_sqrtinv(const double* MMt, int64_t n, int checkres, eg_message_fn message,
                                                void* message_ctx, double* out_sqrt, double* out_invsqrt, int* ok) {
    if (!MMt || !out_sqrt || !out_invsqrt || !ok || n <= 0) return set_error(EG_ERR_ARG, "calculateMMt_sqrt_and_sqrtinv: bad argument");
    EG_TRY(ensure_init_pub());
    cudaStream_t st = ctx_stream();
    const size_t nn = (size_t)n * n;
    HostDev K, S, I, T;
    EG_TRY(K.alloc(nn, "MMt")); EG_TRY(S.alloc(nn, "sqrt")); EG_TRY(I.alloc(nn, "invsqrt")); EG_TRY(T.alloc(nn, "scratch"));
    EG_TRY(up(K.p, MMt, nn, st));
    int not_pd = 0;
    double tr = 0.0;
    EG_TRY(eg_dev_sqrt_and_sqrtinv(K.p, n, S.p, I.p, T.p, &not_pd, checkres ? &tr : nullptr, st));
    *ok = !not_pd;
    if (not_pd) {  // calculateMMt_sqrt_and_sqrtinv.R:15-21: messages, then NULL
        say(message, message_ctx, " Error: the matrix multiplication M %*% t(M) is not positive definite. \n");
        say(message, message_ctx, "        This can occur if there are individuals with identical marker \n");
        say(message, message_ctx, "        information. Please remove individuals with identical marker \n");
        say(message, message_ctx, "        information, remembering also to remove their associated phenotype \n");
        say(message, message_ctx, "        information as well. \n");
        say(message, message_ctx, " Internal function: calculateMMt_sqrt_and_sqrtinv has terminated with errors");
        return EG_OK;
    }
    if (checkres && std::trunc(tr) != (double)n) {  // :35-44
        char buf[512];
        say(message, message_ctx, " \n\n\nWARNING: these results may be unstable.\n");
        snprintf(buf, sizeof(buf), " The sum of the diagonal elements of the square root of M %%*%% t(M) and its inverse is %.15g where \n", tr);
        say(message, message_ctx, buf);
        snprintf(buf, sizeof(buf), "  it should have been %lld\n", (long long)n);
        say(message, message_ctx, buf);
        say(message, message_ctx, "  This can occur if the genotype file contains near identical rows and/or columns.  Please check.\n\n");
    }
    EG_TRY(down(out_sqrt, S.p, nn, st));
    return down(out_invsqrt, I.p, nn, st);
}

extern "C" int eg_calculateH(const double* MMt, int64_t n, double varE, double varG, eg_message_fn message, void* message_ctx,
                             double* out_H, int* ok) {
    if (!MMt || !out_H || !ok || n <= 0) return set_error(EG_ERR_ARG, "calculateH: bad argument");
    *ok = 0;
    if (varE < 0) { say(message, message_ctx, " VarE cannot be negative."); return EG_OK; }  // calculateH.R:22-30
    if (varG < 0) { say(message, message_ctx, " VarG cannot be negative."); return EG_OK; }
    EG_TRY(ensure_init_pub());
    cudaStream_t st = ctx_stream();
    const size_t nn = (size_t)n * n;
    HostDev K, H;
    EG_TRY(K.alloc(nn, "MMt")); EG_TRY(H.alloc(nn, "H"));
    EG_TRY(up(K.p, MMt, nn, st));
    EG_TRY(eg_dev_calculateH(K.p, n, varE, varG, H.p, st));
    *ok = 1;
    return down(out_H, H.p, nn, st);
}

extern "C" int eg_calculateP(const double* H, const double* X, int64_t n, int q, double* out_P) {
    if (!H || !X || !out_P || n <= 0 || q <= 0) return set_error(EG_ERR_ARG, "calculateP: bad argument");
    EG_TRY(ensure_init_pub());
    cudaStream_t st = ctx_stream();
    const size_t nn = (size_t)n * n;
    HostDev dH, dX, dP, dS;
    EG_TRY(dH.alloc(nn, "H")); EG_TRY(dX.alloc((size_t)n * q, "X")); EG_TRY(dP.alloc(nn, "P"));
    EG_TRY(dS.alloc(2 * (size_t)n * q + 2 * (size_t)q * q, "scratch"));
    EG_TRY(up(dH.p, H, nn, st)); EG_TRY(up(dX.p, X, (size_t)n * q, st));
    EG_TRY(eg_dev_calculateP(dH.p, dX.p, n, q, dP.p, dS.p, st));
    return down(out_P, dP.p, nn, st);
}

extern "C" int eg_calculate_reduced_a(double varG, const double* P, const double* MMtsqrt, const double* y, int64_t n,
                                      double* out_a) {
    if (!P || !MMtsqrt || !y || !out_a || n <= 0) return set_error(EG_ERR_ARG, "calculate_reduced_a: bad argument");
    EG_TRY(ensure_init_pub());
    cudaStream_t st = ctx_stream();
    const size_t nn = (size_t)n * n;
    HostDev dP, dS, dv;
    EG_TRY(dP.alloc(nn, "P")); EG_TRY(dS.alloc(nn, "MMtsqrt")); EG_TRY(dv.alloc(3 * (size_t)n, "vectors"));
    EG_TRY(up(dP.p, P, nn, st)); EG_TRY(up(dS.p, MMtsqrt, nn, st)); EG_TRY(up(dv.p, y, (size_t)n, st));
    EG_TRY(eg_dev_calculate_reduced_a(varG, dP.p, dS.p, dv.p, n, dv.p + n, dv.p + 2 * n, st));
    return down(out_a, dv.p + 2 * n, (size_t)n, st);
}

extern "C" int eg_calculate_reduced_vara(const double* X, int64_t n, int q, double varE, double varG, const double* MMtsqrt,
                                         double* out_V) {
    if (!X || !MMtsqrt || !out_V || n <= 0 || q <= 0) return set_error(EG_ERR_ARG, "calculate_reduced_vara: bad argument");
    EG_TRY(ensure_init_pub());
    cudaStream_t st = ctx_stream();
    const size_t nn = (size_t)n * n;
    HostDev dX, dS, dV, dD, dW;
    EG_TRY(dX.alloc((size_t)n * q, "X")); EG_TRY(dS.alloc(nn, "MMtsqrt")); EG_TRY(dV.alloc(nn, "V")); EG_TRY(dD.alloc(nn, "D"));
    EG_TRY(dW.alloc(4 * (size_t)n * q + 3 * (size_t)q * q, "scratch"));
    EG_TRY(up(dX.p, X, (size_t)n * q, st)); EG_TRY(up(dS.p, MMtsqrt, nn, st));
    EG_TRY(eg_dev_calculate_reduced_vara(dX.p, q, varE, varG, dS.p, n, dV.p, dD.p, dW.p, st));
    return down(out_V, dV.p, nn, st);
}

// R/emma_eigen_L_wo_Z.R:9  eigen(K, symmetric = TRUE): values (decreasing) and, when out_vectors is not NULL, vectors
extern "C" int eg_emma_eigen_L_wo_Z(const double* K, int64_t n, double* out_values, double* out_vectors) {
    if (!K || !out_values || n <= 0) return set_error(EG_ERR_ARG, "emma.eigen.L.wo.Z: bad argument");
    EG_TRY(ensure_init_pub());
    cudaStream_t st = ctx_stream();
    const size_t nn = (size_t)n * n;
    HostDev A, w;
    EG_TRY(A.alloc(nn, "K")); EG_TRY(w.alloc((size_t)n, "eigenvalues"));
    EG_TRY(up(A.p, K, nn, st));
    EG_TRY(eg_dev_eigen_sym(A.p, n, w.p, st));
    EG_TRY(down(out_values, w.p, (size_t)n, st));
    return out_vectors ? down(out_vectors, A.p, nn, st) : EG_OK;
}

// R/emma_eigen_R_wo_Z.R:4-20: eigen(S (K + I) S); values[1:(n-q)] - 1 and the first n-q vectors (n x (n-q), column-major)
extern "C" int eg_emma_eigen_R_wo_Z(const double* K, const double* X, int64_t n, int q, double* out_values, double* out_vectors) {
    if (!K || !X || !out_values || !out_vectors || n <= 0 || q <= 0 || q >= n) return set_error(EG_ERR_ARG, "emma.eigen.R.wo.Z: bad argument");
    EG_TRY(ensure_init_pub());
    cudaStream_t st = ctx_stream();
    const size_t nn = (size_t)n * n;
    HostDev dK, dX, S, T, M, sm, w;
    EG_TRY(dK.alloc(nn, "K")); EG_TRY(dX.alloc((size_t)n * q, "X")); EG_TRY(S.alloc(nn, "S")); EG_TRY(T.alloc(nn, "K + I"));
    EG_TRY(M.alloc(nn, "S (K + I) S")); EG_TRY(sm.alloc(2 * (size_t)n * q + 2 * (size_t)q * q, "scratch")); EG_TRY(w.alloc((size_t)n, "eigenvalues"));
    EG_TRY(up(dK.p, K, nn, st)); EG_TRY(up(dX.p, X, (size_t)n * q, st));
    EG_TRY(eg_dev_emma_SKS(dK.p, dX.p, n, q, S.p, T.p, sm.p, st));   // S, and K + I in T
    const double one = 1.0, zero = 0.0;
    EG_BLAS(cublasDgemm(ctx_cublas(), CUBLAS_OP_N, CUBLAS_OP_N, (int)n, (int)n, (int)n, &one, S.p, (int)n, T.p, (int)n, &zero, M.p, (int)n));
    EG_BLAS(cublasDgemm(ctx_cublas(), CUBLAS_OP_N, CUBLAS_OP_N, (int)n, (int)n, (int)n, &one, M.p, (int)n, S.p, (int)n, &zero, T.p, (int)n));
    EG_TRY(eg_dev_eigen_sym(T.p, n, w.p, st));
    add_scalar_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(w.p, n - q, -1.0);
    EG_TRY(check_launch("add_scalar_kernel"));
    EG_TRY(down(out_values, w.p, (size_t)(n - q), st));
    return down(out_vectors, T.p, (size_t)n * (size_t)(n - q), st);
}
