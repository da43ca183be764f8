// The n x n algebra of one AM() forward iteration in the basis of eigen(K)   (SURVEY.md section 8(f) rank 1).
//
// K = MMt / max(MMt) + 0.95 I is fixed after the first iteration (reference: R/AM.R:414-423), yet the reference
// re-derives everything from it with dense LAPACK calls in every iteration: eigen(S (K+I) S) for EMMA
// (R/emma_eigen_R_wo_Z.R:7-20), chol2inv(chol(H)) (R/calculateP.R:27), solve(D) (R/calculate_reduced_vara.R:33), and ~10
// n^3 products.  With K = U diag(xi) U^T computed ONCE,
//     H^-1   = U diag(1 / (varE + varG xi)) U^T                              (R/calculateH.R:36)
//     K^+-1/2 = U diag(xi^+-1/2) U^T                                         (R/calculateMMt_sqrt_and_sqrtinv.R:26-31)
//     D^-1   = U diag(1 / (xi / varE + 1 / varG)) U^T                        (R/calculate_reduced_vara.R:30-33)
// are all diagonal in U, P (R/calculateP.R:28) and V (R/calculate_reduced_vara.R:35) are diagonal + rank q, and the
// scan's right-hand side  W = K^-1/2 V K^-1/2  (src/calculate_a_and_vara_rcpp.cpp:97-98) collapses to
//     W = U diag(w) U^T - E E^T,   w = varG^2 / (varE + varG xi),   E = U (varG Dh Xt) L^-T  (n x q),   L L^T = Xt^T Dh Xt
// (Xt = U^T X, Dh = diag(1 / (varE + varG xi))): ONE n^3 product per iteration -- run on the int8 tensor cores as a
// digit-slice SYRK (csrc/prep_i8.cu) -- plus O(n^2 q) work, instead of two Cholesky inversions, a dense inverse and
// ten FP64 GEMMs; and v = K^-1/2 a_hat = varG P y is two matrix-vector products.  EMMA's per-iteration
// eigendecomposition becomes the O(q n^2) secular solve of csrc/secular.cuh, whose three data-parallel pieces are the
// kernels below.
#include <cublas_v2.h>

#include <time.h>

#include <cstring>
#include <vector>

#include "common.cuh"
#include "secular.cuh"

namespace eg {

int ensure_init_pub();
cudaStream_t ctx_stream();
cublasHandle_t ctx_cublas();
int launch_prepare_eig_i8(const double* d_Ut, const double* d_rs, int64_t n, double* d_Wp, int64_t Kpad, cudaStream_t st,
                          bool* done);
int launch_project_i8(const int8_t* d_Mt, int64_t L, int64_t n, int64_t pitch, const double* d_U, int64_t ld, double* d_B,
                      int64_t ldb, cudaStream_t st);

// ------------------------------------------------------------------ block-wide reductions (fixed tree: deterministic)
constexpr int SEC_THREADS = 128;

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ double warp_prod(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v *= __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
// every thread of the block receives the block-wide sums of vals[0..K)
template <int K>
__device__ __forceinline__ void block_sum(double (&vals)[K], double* sh /* [K * 4] */) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int k = 0; k < K; k++) vals[k] = warp_sum(vals[k]);
    __syncthreads();  // sh may still be read by the previous call
    if (lane == 0) {
#pragma unroll
        for (int k = 0; k < K; k++) sh[k * 4 + warp] = vals[k];
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < K; k++) vals[k] = (sh[k * 4 + 0] + sh[k * 4 + 1]) + (sh[k * 4 + 2] + sh[k * 4 + 3]);
}

// one CTA per root: f(lambda) = sum_i z2_i / (d_i - lambda) in the gap (d_j, d_j+1)
__global__ void __launch_bounds__(SEC_THREADS) sec_roots_kernel(int m, const double* __restrict__ d, const double* __restrict__ z,
                                                                int* __restrict__ origin, double* __restrict__ mu,
                                                                int* __restrict__ max_iters) {
    __shared__ double sh[5 * 4];
    const int j = blockIdx.x;
    auto eval = [&](int o, double x) {
        sec::Sums s = sec::sums_zero();
        const double dorg = d[o];
        for (int i = threadIdx.x; i < m; i += SEC_THREADS) {
            const double zi = z[i];
            sec::sums_term(s, i <= j, zi * zi, d[i] - dorg, x);
        }
        double v[5] = {s.psi, s.dpsi, s.phi, s.dphi, s.asum};
        block_sum<5>(v, sh);
        return sec::Sums{v[0], v[1], v[2], v[3], v[4]};
    };
    int o = 0, it = 0;
    double x = 0.0;
    sec::find_root(j, d[j + 1] - d[j], eval, o, x, it);
    if (threadIdx.x == 0) {
        origin[j] = o;
        mu[j] = x;
        atomicMax(max_iters, it);
    }
}

// one CTA per pole: Loewner weight zhat_i (sign of z_i)
__global__ void __launch_bounds__(SEC_THREADS) sec_lowner_kernel(int m, const double* __restrict__ d, const double* __restrict__ z,
                                                                 const int* __restrict__ origin, const double* __restrict__ mu,
                                                                 double* __restrict__ zhat) {
    __shared__ double sh[4];
    const int i = blockIdx.x;
    double p = 1.0;
    for (int j = threadIdx.x; j < m - 1; j += SEC_THREADS) p *= sec::lowner_factor(i, j, d, origin, mu);
    p = warp_prod(p);
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = p;
    __syncthreads();
    if (threadIdx.x == 0) {
        p = (sh[0] * sh[1]) * (sh[2] * sh[3]);
        zhat[i] = z[i] < 0.0 ? -sqrt(p) : sqrt(p);
    }
}

// one CTA per root: row j of Vout = x_j^T V,  x_j = (D - lambda_j)^-1 zhat / |.|;  columns in chunks of SEC_RC
constexpr int SEC_RC = 8;
__global__ void __launch_bounds__(SEC_THREADS) sec_transform_kernel(int m, int r, const double* __restrict__ d,
                                                                    const double* __restrict__ zhat, const int* __restrict__ origin,
                                                                    const double* __restrict__ mu, const double* __restrict__ V,
                                                                    double* __restrict__ Vout) {
    __shared__ double sh[(SEC_RC + 1) * 4];
    const int j = blockIdx.x;
    const double dorg = d[origin[j]], x0 = mu[j];
    for (int c0 = 0; c0 < r; c0 += SEC_RC) {
        double acc[SEC_RC + 1];
#pragma unroll
        for (int k = 0; k <= SEC_RC; k++) acc[k] = 0.0;
        for (int i = threadIdx.x; i < m; i += SEC_THREADS) {
            const double x = sec::vec_comp(zhat[i], d[i], dorg, x0);
            acc[SEC_RC] += x * x;
#pragma unroll
            for (int k = 0; k < SEC_RC; k++)
                if (c0 + k < r) acc[k] += x * V[i + (size_t)(c0 + k) * m];
        }
        block_sum<SEC_RC + 1>(acc, sh);
        if (threadIdx.x == 0) {
            const double inv = 1.0 / sqrt(acc[SEC_RC]);
#pragma unroll
            for (int k = 0; k < SEC_RC; k++)
                if (c0 + k < r) Vout[j + (size_t)(c0 + k) * (m - 1)] = acc[k] * inv;
        }
    }
}

// ------------------------------------------------------------------ the CUDA back end of sec::compress
struct SecWorkspace {
    double* buf = nullptr;  // d | z | zhat | mu | V | Vout
    int* ibuf = nullptr;    // origin | max_iters
    size_t cap = 0, icap = 0;
};
static thread_local SecWorkspace g_sec;
void eigbasis_release() {
    cudaFree(g_sec.buf);
    cudaFree(g_sec.ibuf);
    g_sec = SecWorkspace();
}

static double now_s() {
    timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}
static thread_local double g_sec_times[2] = {0.0, 0.0};   // seconds of the last solve: inside the device back end, in total

struct CudaBackend {
    cudaStream_t st;
    int err = EG_OK;
    double busy = 0.0;
    int solve(int m, const double* d, const double* z, const double* V, int r, int* origin, double* mu, double* Vout,
              int* max_iters) {
        const double t_in = now_s();
        const int rc = solve_(m, d, z, V, r, origin, mu, Vout, max_iters);
        busy += now_s() - t_in;
        return rc;
    }
    int solve_(int m, const double* d, const double* z, const double* V, int r, int* origin, double* mu, double* Vout,
               int* max_iters) {
        // sized for 64 carried columns at once: growing the buffers call by call (the design matrix gains a column per
        // forward iteration) meant a cudaFree + cudaMalloc per iteration, 0.4 - 0.8 s each next to the memory pools
        const size_t rcap = (size_t)(r < 64 ? 64 : r);
        const size_t need = (size_t)m * 4 + (size_t)m * rcap * 2, ineed = (size_t)m + 1;
        if (need > g_sec.cap) {
            cudaFree(g_sec.buf);
            g_sec.buf = nullptr;
            g_sec.cap = 0;
            if (malloc_retry((void**)&g_sec.buf, need * 8) != cudaSuccess) { cudaGetLastError(); err = set_error(EG_ERR_ALLOC, "secular solve: out of device memory"); return 1; }
            g_sec.cap = need;
        }
        if (ineed > g_sec.icap) {
            cudaFree(g_sec.ibuf);
            g_sec.ibuf = nullptr;
            g_sec.icap = 0;
            if (cudaMalloc(&g_sec.ibuf, ineed * 4) != cudaSuccess) { cudaGetLastError(); err = set_error(EG_ERR_ALLOC, "secular solve: out of device memory"); return 1; }
            g_sec.icap = ineed;
        }
        double *dd = g_sec.buf, *dz = dd + m, *dzh = dz + m, *dmu = dzh + m, *dV = dmu + m, *dVo = dV + (size_t)m * r;
        int *dor = g_sec.ibuf, *dit = dor + m;
        auto up = [&](void* dst, const void* src, size_t bytes) { return cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, st); };
        if (up(dd, d, (size_t)m * 8) != cudaSuccess || up(dz, z, (size_t)m * 8) != cudaSuccess ||
            up(dV, V, (size_t)m * r * 8) != cudaSuccess || cudaMemsetAsync(dit, 0, 4, st) != cudaSuccess) {
            err = check_cuda(cudaGetLastError(), "secular solve: H2D");
            return 1;
        }
        sec_roots_kernel<<<m - 1, SEC_THREADS, 0, st>>>(m, dd, dz, dor, dmu, dit);
        if ((err = check_launch("sec_roots_kernel")) != EG_OK) return 1;
        sec_lowner_kernel<<<m, SEC_THREADS, 0, st>>>(m, dd, dz, dor, dmu, dzh);
        if ((err = check_launch("sec_lowner_kernel")) != EG_OK) return 1;
        sec_transform_kernel<<<m - 1, SEC_THREADS, 0, st>>>(m, r, dd, dzh, dor, dmu, dV, dVo);
        if ((err = check_launch("sec_transform_kernel")) != EG_OK) return 1;
        cudaMemcpyAsync(origin, dor, (size_t)(m - 1) * 4, cudaMemcpyDeviceToHost, st);
        cudaMemcpyAsync(mu, dmu, (size_t)(m - 1) * 8, cudaMemcpyDeviceToHost, st);
        cudaMemcpyAsync(Vout, dVo, (size_t)(m - 1) * r * 8, cudaMemcpyDeviceToHost, st);
        cudaMemcpyAsync(max_iters, dit, 4, cudaMemcpyDeviceToHost, st);
        if ((err = check_cuda(cudaStreamSynchronize(st), "secular solve")) != EG_OK) return 1;
        return 0;
    }
};

// ------------------------------------------------------------------ small kernels of the eigenbasis route
// out (n x n column-major) = in^T; 32 x 32 tiles through shared memory
__global__ void __launch_bounds__(256) transpose_f64_kernel(const double* __restrict__ in, int64_t n, double* __restrict__ out) {
    __shared__ double tile[32][33];
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int64_t r0 = (int64_t)blockIdx.y * 32, c0 = (int64_t)blockIdx.x * 32;
    for (int i = ty; i < 32; i += 8) {
        const int64_t r = r0 + tx, c = c0 + i;
        tile[i][tx] = (r < n && c < n) ? in[r + c * n] : 0.0;
    }
    __syncthreads();
    for (int i = ty; i < 32; i += 8) {
        const int64_t r = c0 + tx, c = r0 + i;  // out(r, c) = in(c, r)
        if (r < n && c < n) out[r + c * n] = tile[tx][i];
    }
}
// T[:, j] = U[:, j] * rs[j]   (cuBLAS fallback of the product)
__global__ void __launch_bounds__(256) scale_cols_kernel(const double* __restrict__ U, const double* __restrict__ rs, int64_t n,
                                                         double* __restrict__ T) {
    const int64_t j = blockIdx.y;
    const double s = rs[j];
    for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (int64_t)gridDim.x * 256) T[i + j * n] = U[i + j * n] * s;
}
__global__ void eig_sqrt_kernel(const double* __restrict__ w, int64_t n, double* __restrict__ out) {
    const int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x;
    if (i < n) out[i] = sqrt(w[i]);
}
// In place on Wp (column-major, ld): the upper triangle holds W0 = U diag(w) U^T.  W = W0 - E E^T (E: n x q, ld n) is
// folded into the form the scan contracts,  diag(W) + 2 strict_upper(W)  (W symmetric; scan_i8.cu / scan_f64.cu), the
// strict lower triangle is zeroed.
__global__ void __launch_bounds__(256) eig_fold_kernel(double* __restrict__ Wp, int64_t n, int64_t ld, const double* __restrict__ E,
                                                       int q) {
    const int bx = blockIdx.x, by = blockIdx.y;  // tile rows by*32.., cols bx*32..
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int64_t r0 = (int64_t)by * 32, c0 = (int64_t)bx * 32;
    if (bx < by) {
        for (int i = ty; i < 32; i += 8) {
            const int64_t r = r0 + tx, c = c0 + i;
            if (r < n && c < n) Wp[r + c * ld] = 0.0;
        }
        return;
    }
    extern __shared__ double sh[];  // Er[q][32] | Ec[q][32]
    double* Er = sh;
    double* Ec = sh + (size_t)q * 32;
    for (int t = threadIdx.x; t < q * 32; t += 256) {
        const int c = t >> 5, l = t & 31;
        Er[t] = r0 + l < n ? E[r0 + l + (int64_t)c * n] : 0.0;
        Ec[t] = c0 + l < n ? E[c0 + l + (int64_t)c * n] : 0.0;
    }
    __syncthreads();
    for (int i = ty; i < 32; i += 8) {
        const int64_t r = r0 + tx, c = c0 + i;
        if (r < n && c < n) {
            if (r > c) { Wp[r + c * ld] = 0.0; continue; }
            double corr = 0.0;
            for (int k = 0; k < q; k++) corr += Er[k * 32 + tx] * Ec[k * 32 + i];
            const double w = Wp[r + c * ld] - corr;
            Wp[r + c * ld] = r == c ? w : w + w;
        }
    }
}

// s_j = sum_k w_k B_jk^2 for every marker row of B (L x n doubles, row pitch ldb): one warp per row, lane-strided
// partial sums over 4 independent accumulators, then a fixed shuffle tree -- the same order for every row, so identical
// markers give identical bits.  HBM-bound: 8 n bytes per marker.
__global__ void __launch_bounds__(256) bscan_kernel(const double* __restrict__ B, int64_t L, int64_t n, int64_t ldb,
                                                    const double* __restrict__ w, double* __restrict__ out) {
    const int lane = threadIdx.x & 31;
    const int64_t warps = (int64_t)gridDim.x * (blockDim.x >> 5);
    for (int64_t j = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); j < L; j += warps) {
        const double* row = B + j * ldb;
        double acc[4] = {0.0, 0.0, 0.0, 0.0};
        int64_t k = 2 * lane;
        for (; k + 192 + 1 < n; k += 256) {   // 4 x (2 doubles per lane): 16-byte loads, 512 B per warp and step
#pragma unroll
            for (int u = 0; u < 4; u++) {
                const double2 b = *reinterpret_cast<const double2*>(row + k + 64 * u);
                const double2 ww = *reinterpret_cast<const double2*>(w + k + 64 * u);
                acc[u] = fma(ww.x * b.x, b.x, acc[u]);
                acc[u] = fma(ww.y * b.y, b.y, acc[u]);
            }
        }
        for (; k < n; k += 64) {              // tail, same lane -> column assignment
            const double b0 = row[k], b1 = k + 1 < n ? row[k + 1] : 0.0;
            const double w0 = w[k], w1 = k + 1 < n ? w[k + 1] : 0.0;
            acc[0] = fma(w0 * b0, b0, acc[0]);
            acc[0] = fma(w1 * b1, b1, acc[0]);
        }
        double s = (acc[0] + acc[1]) + (acc[2] + acc[3]);
        s = warp_sum(s);
        if (lane == 0) out[j] = s;
    }
}
// vara_j = s_j - sum_c e_cj^2   (e: q vectors of length L, row-major q x L).
// A marker that lies in the span of the fixed effects (a locus already in the model, or an identical copy of one) has
// var(a) = 0 and a = 0 exactly: tsq = 0 / 0, which R's which(tsq == max(tsq, na.rm = TRUE)) ignores.  The dense route
// returns rounding noise there (|var(a)| ~ 1e-13 of its scale, tsq ~ 1e-14: never the maximum); this route cancels two
// accurately computed sums and can land on +-1e-17 or on 0, i.e. on a huge or infinite tsq.  So a difference below
// 1e-10 of s_j -- fifty times the rounding level n eps, a squared multiple correlation with the model above
// 1 - 1e-10 -- is reported as what it is in exact arithmetic: NaN (not testable).
__global__ void __launch_bounds__(256) bscan_combine_kernel(const double* __restrict__ s, const double* __restrict__ e, int q, int64_t L,
                                                            double* __restrict__ vara) {
    const int64_t j = (int64_t)blockIdx.x * 256 + threadIdx.x;
    if (j >= L) return;
    double c = 0.0;
    for (int k = 0; k < q; k++) c = fma(e[(int64_t)k * L + j], e[(int64_t)k * L + j], c);
    const double v = s[j] - c;
    vara[j] = (q > 0 && s[j] > 0.0 && v <= 1e-10 * s[j]) ? __longlong_as_double(0x7FF8000000000000LL) : v;
}

}  // namespace eg

using namespace eg;

#define EGB_BLAS(expr)                                                                     \
    do {                                                                                   \
        if ((expr) != CUBLAS_STATUS_SUCCESS) return set_error(EG_ERR_CUDA, "%s failed", #expr); \
    } while (0)

// ================================================================== EMMA's eigendecomposition in the basis of eigen(K)
extern "C" int eg_emma_eigen_R_wo_Z_eigbasis(const double* xi, const double* Xt, const double* yt, int64_t n, int q,
                                             double* out_values, double* out_etas, int64_t* stats4) {
    if (!xi || !Xt || !yt || !out_values || !out_etas || n <= 0 || q <= 0 || q >= n || n > 0x7fffffff)
        return set_error(EG_ERR_ARG, "emma.eigen.R.wo.Z (eigenbasis form): bad argument");
    EG_TRY(ensure_init_pub());
    CudaBackend be{ctx_stream()};
    sec::Stats st;
    const double t_all = now_s();
    const int rc = sec::compress(be, n, q, xi, Xt, yt, out_values, out_etas, &st);
    g_sec_times[0] = be.busy;
    g_sec_times[1] = now_s() - t_all;
    if (stats4) {
        stats4[0] = st.steps;
        stats4[1] = st.deflated;
        stats4[2] = st.max_iters;
        stats4[3] = st.roots;
    }
    if (rc >= 100) return be.err != EG_OK ? be.err : set_error(EG_ERR_CUDA, "secular solve failed");
    if (rc == 1) return set_error(EG_ERR_ARG, "emma.eigen.R.wo.Z: the design matrix X is rank deficient");
    if (rc) return set_error(EG_ERR_ARG, "emma.eigen.R.wo.Z: secular solve lost the ordering of its poles (%d)", rc);
    return EG_OK;
}

// seconds of the calling thread's last eg_emma_eigen_R_wo_Z_eigbasis: [0] uploads + kernels + downloads, [1] the whole call
extern "C" int eg_last_secular_times(double* out2) {
    if (!out2) return set_error(EG_ERR_ARG, "eg_last_secular_times: null");
    out2[0] = g_sec_times[0];
    out2[1] = g_sec_times[1];
    return EG_OK;
}

// ================================================================== device level
extern "C" int eg_dev_transpose_f64(const double* d_in, int64_t n, double* d_out, void* stream) {
    if (!d_in || !d_out || n <= 0 || d_in == d_out) return set_error(EG_ERR_ARG, "eg_dev_transpose_f64: bad argument");
    const unsigned nb = (unsigned)((n + 31) / 32);
    transpose_f64_kernel<<<dim3(nb, nb), 256, 0, (cudaStream_t)stream>>>(d_in, n, d_out);
    return check_launch("transpose_f64_kernel");
}

// d_out (n x r) = U^T d_in   (into the eigenbasis);   transpose_u == 0:  d_out = U d_in   (back)
extern "C" int eg_dev_eigbasis_apply(const double* d_U, int64_t n, const double* d_in, int r, int to_eigenbasis, double* d_out,
                                     void* stream) {
    if (!d_U || !d_in || !d_out || n <= 0 || r <= 0) return set_error(EG_ERR_ARG, "eg_dev_eigbasis_apply: bad argument");
    EG_TRY(ensure_init_pub());
    cudaStream_t st = (cudaStream_t)stream;
    EGB_BLAS(cublasSetStream(ctx_cublas(), st));
    const double one = 1.0, zero = 0.0;
    if (r == 1)
        EGB_BLAS(cublasDgemv(ctx_cublas(), to_eigenbasis ? CUBLAS_OP_T : CUBLAS_OP_N, (int)n, (int)n, &one, d_U, (int)n, d_in, 1, &zero, d_out, 1));
    else
        EGB_BLAS(cublasDgemm(ctx_cublas(), to_eigenbasis ? CUBLAS_OP_T : CUBLAS_OP_N, CUBLAS_OP_N, (int)n, r, (int)n, &one, d_U, (int)n,
                             d_in, (int)n, &zero, d_out, (int)n));
    return EG_OK;
}

// The scan's right-hand side from eigenbasis quantities:  W = U diag(w) U^T - E E^T  folded into d_Wp, v into its column n.
//   d_U, d_Ut: eigenvectors of K (columns) and their transpose, n x n column-major;  d_w[n] > 0;
//   d_Et: n x q column-major, E = U Et  (eigenbasis coordinates of the rank-q part; q may be 0);  d_vt[n]: v = U vt;
//   d_work: n * max(q, 1) doubles (receives E);  when the digit slices do not fit (or n < 1024) also n * n doubles are
//   taken from d_work2 (may be NULL when the int8 path is certain).
extern "C" int eg_dev_scan_prepare_eig(const double* d_U, const double* d_Ut, int64_t n, const double* d_w, const double* d_Et, int q,
                                       const double* d_vt, double* d_work, double* d_work2, double* d_Wp, void* stream) {
    if (!d_U || !d_Ut || !d_w || !d_vt || !d_work || !d_Wp || n <= 0 || q < 0 || (q > 0 && !d_Et) || q > 64)
        return set_error(EG_ERR_ARG, "eg_dev_scan_prepare_eig: bad argument");
    EG_TRY(ensure_init_pub());
    cudaStream_t st = (cudaStream_t)stream;
    const int64_t Kpad = round_up(n, 32);
    EGB_BLAS(cublasSetStream(ctx_cublas(), st));
    EG_CUDA(cudaMemsetAsync(d_Wp, 0, (size_t)eg_scan_wp_elems(n) * 8, st));
    // rs = sqrt(w): A = U diag(rs), W0 = A A^T
    double* d_rs = d_Wp + n * Kpad;  // column n of Wp as scratch until v is written there
    eig_sqrt_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(d_w, n, d_rs);
    EG_TRY(check_launch("eig_sqrt_kernel"));
    const bool want_i8 = eg_prep_uses_i8(n) != 0;
    bool done = false;
    if (want_i8) EG_TRY(launch_prepare_eig_i8(d_Ut, d_rs, n, d_Wp, Kpad, st, &done));
    const double one = 1.0, zero = 0.0;
    if (!done) {
        if (!d_work2) return set_error(EG_ERR_ARG, "eg_dev_scan_prepare_eig: the FP64 path needs d_work2 (n*n doubles)");
        scale_cols_kernel<<<dim3((unsigned)((n + 255) / 256 < 64 ? (n + 255) / 256 : 64), (unsigned)n), 256, 0, st>>>(d_U, d_rs, n, d_work2);
        EG_TRY(check_launch("scale_cols_kernel"));
        EGB_BLAS(cublasDsyrk(ctx_cublas(), CUBLAS_FILL_MODE_UPPER, CUBLAS_OP_N, (int)n, (int)n, &one, d_work2, (int)n, &zero, d_Wp, (int)Kpad));
    }
    if (q > 0)
        EGB_BLAS(cublasDgemm(ctx_cublas(), CUBLAS_OP_N, CUBLAS_OP_N, (int)n, q, (int)n, &one, d_U, (int)n, d_Et, (int)n, &zero, d_work, (int)n));
    const unsigned nb = (unsigned)((n + 31) / 32);
    eig_fold_kernel<<<dim3(nb, nb), 256, (size_t)(q > 0 ? q : 1) * 64 * sizeof(double), st>>>(d_Wp, n, Kpad, d_work, q);
    EG_TRY(check_launch("eig_fold_kernel"));
    // v = U vt -> column n of Wp   (src/calculate_a_and_vara_rcpp.cpp:90: v = inv_MMt_sqrt * a_hat)
    EGB_BLAS(cublasDgemv(ctx_cublas(), CUBLAS_OP_N, (int)n, (int)n, &one, d_U, (int)n, d_vt, 1, &zero, d_Wp + n * Kpad, 1));
    return EG_OK;
}

// ================================================================== the scan from a cached projection B = M^T U
// (am.AM_resident, bcache route).  K is fixed for the whole search, so B = M^T U (L x n doubles: 80 GB at config 3) is
// computed ONCE, by the scan's own int8 digit-slice contraction in projection mode; after that
//     var(a)_j = m_j^T W m_j = sum_k w_k B_jk^2 - sum_c (E_c^T m_j)^2,      a_j = m_j^T v
// is one HBM-bound pass over B plus q + 1 exact int8 matrix-vector products with Mt, instead of the n^2 L contraction of
// src/calculate_a_and_vara_rcpp.cpp:103-112 in every forward iteration.
extern "C" int eg_dev_project_i8(const int8_t* d_Mt, int64_t L, int64_t n, int64_t pitch, const double* d_U, double* d_B,
                                 int64_t ldb, void* stream) {
    if (!d_Mt || !d_U || !d_B || L <= 0 || n <= 0 || ldb < n || (ldb & 1))
        return set_error(EG_ERR_ARG, "eg_dev_project_i8: bad argument (ldb >= n and even)");
    EG_TRY(ensure_init_pub());
    return launch_project_i8(d_Mt, L, n, pitch, d_U, n, d_B, ldb, (cudaStream_t)stream);
}
extern "C" int eg_dev_bscan(const double* d_B, int64_t L, int64_t n, int64_t ldb, const double* d_w, const double* d_e, int q,
                            double* d_tmp_L, double* d_vara, void* stream) {
    if (!d_B || !d_w || !d_tmp_L || !d_vara || L <= 0 || n <= 0 || ldb < n || (ldb & 1) || q < 0 || (q > 0 && !d_e))
        return set_error(EG_ERR_ARG, "eg_dev_bscan: bad argument");
    cudaStream_t st = (cudaStream_t)stream;
    const int64_t blocks = (L + 7) / 8;
    bscan_kernel<<<(unsigned)(blocks < 148 * 16 ? blocks : 148 * 16), 256, 0, st>>>(d_B, L, n, ldb, d_w, d_tmp_L);
    EG_TRY(check_launch("bscan_kernel"));
    bscan_combine_kernel<<<(unsigned)((L + 255) / 256), 256, 0, st>>>(d_tmp_L, d_e, q, L, d_vara);
    return check_launch("bscan_combine_kernel");
}
