// K3 -- per-marker BLUP and variance scan on the FP64 tensor cores.
//
// Replaces, for all markers j (rows of Mt), reference src/calculate_a_and_vara_rcpp.cpp:
//     a_j    = Mt_j . v              v = inv_MMt_sqrt * a_hat                     (:90-91)
//     T      = Mt * W                W = inv_MMt_sqrt * (dim_reduced_vara * inv_MMt_sqrt) (:97-98, :103)
//     vara_j = T_j . Mt_j                                                          (:107-112)
// without ever materialising T (L x n doubles in the reference).  v is carried as column n of
// the packed right-hand side Wp, so a_j falls out of the same GEMM.
//
// Wp layout (built by eg_dev_scan_prepare): Npad columns x Kpad rows, column-major, ld = Kpad,
// Kpad = round_up(n,32), Npad = round_up(n+1,128); zero outside W and v.
// Mt: int8 store, L x n, row-major, pitch >= Npad, zero padded (so pad columns add exact zeros).
//
// One CTA owns a block of 128 markers and sweeps all column panels of Wp (128 wide) in the
// same order on every CTA, so the panels are shared through L2.  Inside: 8 warps (2 x 4), warp
// tile 64 x 32, DMMA.8x8x4 (mma.sync.m8n8k4.f64), int8 -> f64 conversion in registers with
// integer ops only, 3-stage cp.async ring.  The row-dot is fused into the panel epilogue with
// a fixed reduction order (registers -> 4-lane shuffle -> 4 warps through shared memory ->
// panel order), so a marker's a/vara depend only on its genotypes, never on its position,
// tile or GPU: identical marker rows give bit-identical results (tie rule of find_qtl.R:76-80).
#include "common.cuh"
#include "ptx.cuh"

namespace eg {

constexpr int SC_BM = 128;
constexpr int SC_BN = 128;
constexpr int SC_BK = 32;
constexpr int SC_STAGES = 3;
constexpr int SC_THREADS = 256;
constexpr int SC_B_STRIDE = SC_BK + 4;               // doubles per staged W column (== 4 mod 16: conflict-free)
constexpr int SC_A_STRIDE = 48;                      // bytes per staged Mt row (12 words: conflict-free)
constexpr int SC_B_BYTES = SC_BN * SC_B_STRIDE * 8;  // 36864
constexpr int SC_A_BYTES = SC_BM * SC_A_STRIDE;      // 6144
constexpr int SC_STAGE_BYTES = SC_B_BYTES + SC_A_BYTES;
constexpr int SC_SMEM_BYTES = SC_STAGES * SC_STAGE_BYTES + 4 * SC_BM * 8;

struct ScanParams {
    const int8_t* Mt;
    int64_t L, n, pitch;
    const double* Wp;
    int64_t Kpad;
    int32_t NP;        // column panels
    int32_t KT;        // k tiles per panel
    const int64_t* zero_rows;
    int32_t n_zero;
    double* out_a;
    double* out_vara;
    int64_t num_blocks;
};

__device__ __forceinline__ double s8_to_f64(int g) {  // g in {-1,0,1}; no FP64-pipe conversion
    const uint32_t hi = ((uint32_t)g & 0x80000000u) | ((uint32_t)(-(g & 1)) & 0x3FF00000u);
    return __hiloint2double((int)hi, 0);
}

__global__ void __launch_bounds__(SC_THREADS, 1) scan_f64_kernel(const ScanParams p) {
    extern __shared__ __align__(128) uint8_t smem[];
    double* red = reinterpret_cast<double*>(smem + SC_STAGES * SC_STAGE_BYTES);  // [4][SC_BM]

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int wm = warp & 1, wn = warp >> 1;
    const int lr = lane >> 2, lc = lane & 3;
    const int64_t total_steps = (int64_t)p.NP * p.KT;
    const int pa = (int)(p.n / SC_BN);            // panel holding column n (= v)
    const int ca = (int)(p.n - (int64_t)pa * SC_BN);

    for (int64_t mb = blockIdx.x; mb < p.num_blocks; mb += gridDim.x) {
        const int64_t j0 = mb * SC_BM;

        auto load_step = [&](int64_t s) {
            const int pnl = (int)(s / p.KT);
            const int kt = (int)(s - (int64_t)pnl * p.KT);
            uint8_t* st = smem + (s % SC_STAGES) * SC_STAGE_BYTES;
            {   // W panel slice: 128 columns x 32 rows of doubles; thread -> (column tid/2, 128-byte half)
                const int c = tid >> 1, h = tid & 1;
                const double* src = p.Wp + ((int64_t)pnl * SC_BN + c) * p.Kpad + (int64_t)kt * SC_BK + h * 16;
                uint8_t* dst = st + c * (SC_B_STRIDE * 8) + h * 128;
#pragma unroll
                for (int i = 0; i < 8; i++) ptx::cp_async_16(dst + i * 16, reinterpret_cast<const uint8_t*>(src) + i * 16);
            }
            {   // Mt slice: 128 markers x 32 individuals (bytes); thread -> (row tid/2, 16-byte half)
                const int r = tid >> 1, h = tid & 1;
                int64_t j = j0 + r;
                if (j >= p.L) j = p.L - 1;  // tail block: duplicate a valid row, result discarded
                const int8_t* src = p.Mt + j * p.pitch + (int64_t)kt * SC_BK + h * 16;
                ptx::cp_async_16(st + SC_B_BYTES + r * SC_A_STRIDE + h * 16, src);
            }
        };

        double acc[8][4][2];
#pragma unroll
        for (int mt = 0; mt < 8; mt++)
#pragma unroll
            for (int t = 0; t < 4; t++) acc[mt][t][0] = acc[mt][t][1] = 0.0;
        double vara_run = 0.0;  // threads 0..127: running vara of marker row tid

        __syncthreads();  // previous marker block fully done with shared memory
        for (int64_t s = 0; s < SC_STAGES - 1; s++) {
            if (s < total_steps) load_step(s);
            ptx::cp_async_commit();
        }

        for (int64_t s = 0; s < total_steps; s++) {
            ptx::cp_async_wait<SC_STAGES - 2>();
            __syncthreads();
            if (s + SC_STAGES - 1 < total_steps) load_step(s + SC_STAGES - 1);
            ptx::cp_async_commit();

            const uint8_t* st = smem + (s % SC_STAGES) * SC_STAGE_BYTES;
            const double* sB = reinterpret_cast<const double*>(st);
            const int8_t* sA = reinterpret_cast<const int8_t*>(st + SC_B_BYTES);
#pragma unroll
            for (int ks = 0; ks < SC_BK / 4; ks++) {
                double b[4];
#pragma unroll
                for (int t = 0; t < 4; t++) b[t] = sB[(wn * 32 + t * 8 + lr) * SC_B_STRIDE + ks * 4 + lc];
#pragma unroll
                for (int mt = 0; mt < 8; mt++) {
                    const double a = s8_to_f64((int)sA[(wm * 64 + mt * 8 + lr) * SC_A_STRIDE + ks * 4 + lc]);
#pragma unroll
                    for (int t = 0; t < 4; t++) ptx::dmma_884(acc[mt][t][0], acc[mt][t][1], a, b[t]);
                }
            }

            const int pnl = (int)(s / p.KT);
            if (s - (int64_t)pnl * p.KT == p.KT - 1) {
                // ---------------- panel epilogue: fused row-dot with the marker's own genotypes
                const int64_t cbase = (int64_t)pnl * SC_BN + wn * 32 + lc * 2;
#pragma unroll
                for (int mt = 0; mt < 8; mt++) {
                    const int row = wm * 64 + mt * 8 + lr;
                    int64_t j = j0 + row;
                    if (j >= p.L) j = p.L - 1;
                    const int8_t* mrow = p.Mt + j * p.pitch + cbase;
                    double sum = 0.0;
#pragma unroll
                    for (int t = 0; t < 4; t++) {
                        const short two = *reinterpret_cast<const short*>(mrow + t * 8);
                        const int m0 = (int)(int8_t)(two & 0xFF), m1 = (int)(int8_t)((two >> 8) & 0xFF);
                        sum += acc[mt][t][0] * s8_to_f64(m0);
                        sum += acc[mt][t][1] * s8_to_f64(m1);
                    }
                    sum += __shfl_xor_sync(0xffffffffu, sum, 1);
                    sum += __shfl_xor_sync(0xffffffffu, sum, 2);
                    if (lc == 0) red[wn * SC_BM + row] = sum;
                    if (pnl == pa) {  // a_j = T[j][n]
#pragma unroll
                        for (int t = 0; t < 4; t++)
#pragma unroll
                            for (int e = 0; e < 2; e++)
                                if (wn * 32 + t * 8 + lc * 2 + e == ca && j0 + row < p.L)
                                    p.out_a[j0 + row] = acc[mt][t][e];
                    }
#pragma unroll
                    for (int t = 0; t < 4; t++) acc[mt][t][0] = acc[mt][t][1] = 0.0;
                }
                __syncthreads();
                if (tid < SC_BM)
                    vara_run += (red[tid] + red[SC_BM + tid]) + (red[2 * SC_BM + tid] + red[3 * SC_BM + tid]);
                // red[] is rewritten only after the next panel's k loop (many barriers later)
            }
        }
        ptx::cp_async_wait<0>();
        __syncthreads();
        if (tid < SC_BM && j0 + tid < p.L) {
            const int64_t j = j0 + tid;
            bool zero = false;
            for (int z = 0; z < p.n_zero; z++) zero |= (p.zero_rows[z] == j);
            p.out_vara[j] = zero ? 0.0 : vara_run;
            if (zero) p.out_a[j] = 0.0;  // ordered after the panel epilogue's write by the barriers above
        }
    }
}

}  // namespace eg

extern "C" int64_t eg_scan_wp_elems(int64_t n) {
    return eg::round_up(n + 1, eg::SC_BN) * eg::round_up(n, eg::SC_BK);
}

namespace eg {
int launch_scan(const int8_t* d_Mt, int64_t L, int64_t n, int64_t pitch, const double* d_Wp,
                const int64_t* d_zero_rows, int n_zero, double* d_a, double* d_vara, cudaStream_t st) {
    ScanParams p;
    p.Mt = d_Mt; p.L = L; p.n = n; p.pitch = pitch; p.Wp = d_Wp;
    p.Kpad = round_up(n, SC_BK);
    p.NP = (int32_t)(round_up(n + 1, SC_BN) / SC_BN);
    p.KT = (int32_t)(p.Kpad / SC_BK);
    p.zero_rows = d_zero_rows; p.n_zero = n_zero;
    p.out_a = d_a; p.out_vara = d_vara;
    p.num_blocks = (L + SC_BM - 1) / SC_BM;
    EG_CUDA(cudaFuncSetAttribute(scan_f64_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SC_SMEM_BYTES));
    const int64_t grid = p.num_blocks < num_sms() ? p.num_blocks : num_sms();
    scan_f64_kernel<<<(unsigned)grid, SC_THREADS, SC_SMEM_BYTES, st>>>(p);
    return check_launch("scan_f64_kernel");
}
}  // namespace eg

extern "C" int eg_dev_scan(const int8_t* d_Mt, int64_t L, int64_t n, int64_t pitch, const double* d_Wp,
                           const int64_t* h_zero_rows, int64_t n_zero, double* d_a, double* d_vara, void* stream) {
    using namespace eg;
    if (!d_Mt || !d_Wp || !d_a || !d_vara || L <= 0 || n <= 0 || (pitch & 15) || pitch < round_up(n + 1, SC_BN) ||
        ((uintptr_t)d_Mt & 15) || n_zero < 0 || (n_zero > 0 && !h_zero_rows))
        return set_error(EG_ERR_ARG, "eg_dev_scan: bad argument (Mt pitch must be >= round_up(n+1,128))");
    cudaStream_t st = (cudaStream_t)stream;
    int64_t* d_zero = nullptr;
    if (n_zero > 0) {
        EG_CUDA(cudaMalloc(&d_zero, n_zero * sizeof(int64_t)));
        int rc = check_cuda(cudaMemcpyAsync(d_zero, h_zero_rows, n_zero * sizeof(int64_t), cudaMemcpyHostToDevice, st),
                            "zero rows H2D");
        if (rc) { cudaFree(d_zero); return rc; }
    }
    int rc = launch_scan(d_Mt, L, n, pitch, d_Wp, d_zero, (int)n_zero, d_a, d_vara, st);
    if (d_zero) {
        cudaStreamSynchronize(st);
        cudaFree(d_zero);
    }
    return rc;
}
