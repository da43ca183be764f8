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
// same order on every CTA, so the panels are shared through L2.  Inside: 16 warps (4 x 4), warp
// tile 32 x 32 (4 warps per SM sub-partition keep the DMMA pipe fed: ptxas must space dependent
// DMMAs of one warp), DMMA.8x8x4 (mma.sync.m8n8k4.f64), int8 -> f64 conversion with integer ops
// only, done one stage ahead into a swizzled FP64 shared-memory tile, 4-stage cp.async ring.  The row-dot is fused into the panel epilogue with
// a fixed reduction order (registers -> 4-lane shuffle -> 4 warps through shared memory ->
// panel order), so a marker's a/vara depend only on its genotypes, never on its position,
// tile or GPU: identical marker rows give bit-identical results (tie rule of find_qtl.R:76-80).
#include "common.cuh"
#include "ptx.cuh"

namespace eg {

constexpr int SC_BM = 128;
constexpr int SC_BN = 128;
constexpr int SC_BK = 32;                 // k-tile: two 16-deep halves, each row/column half = one 128-byte segment
constexpr int SC_THREADS = 512;           // 16 warps = 4 (markers) x 4 (columns), warp tile 32 x 32
constexpr int SC_BSTAGES = 4;             // ring of W tiles and raw Mt tiles (cp.async)
constexpr int SC_HALF_BYTES = 128 * 128;  // 128 rows (or columns) x 16 doubles
constexpr int SC_B_BYTES = 2 * SC_HALF_BYTES;   // 32 KB per W k-tile
constexpr int SC_A64_BYTES = 2 * SC_HALF_BYTES; // 32 KB per converted Mt k-tile
constexpr int SC_RAW_BYTES = SC_BM * SC_BK;     // 4 KB per raw int8 Mt k-tile
constexpr int SC_OFF_A64 = SC_BSTAGES * SC_B_BYTES;
constexpr int SC_OFF_RAW = SC_OFF_A64 + 2 * SC_A64_BYTES;
constexpr int SC_OFF_RED = SC_OFF_RAW + SC_BSTAGES * SC_RAW_BYTES;
constexpr int SC_SMEM_BYTES = SC_OFF_RED + 4 * SC_BM * 8;

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

// genotype byte b (0..3) of w in {-1,0,1} as an IEEE double: integer pipe only
template <int B>
__device__ __forceinline__ double s8_byte_to_f64(uint32_t w) {
    constexpr uint32_t sel = (uint32_t)B | ((8u | B) << 4) | ((8u | B) << 8) | ((8u | B) << 12);
    const uint32_t g = __byte_perm(w, 0u, sel);  // sign-extended byte
    const uint32_t hi = (g & 0x80000000u) | ((g & 1u) * 0x3FF00000u);
    return __hiloint2double((int)hi, 0);
}
__device__ __forceinline__ double s8_to_f64(int g) {  // g in {-1,0,1}
    const uint32_t hi = ((uint32_t)g & 0x80000000u) | (((uint32_t)g & 1u) * 0x3FF00000u);
    return __hiloint2double((int)hi, 0);
}

// Shared-memory tiles.  Both operands are staged as doubles in [half][row-or-column][16] blocks whose
// 16-byte chunks are XOR-swizzled with (row & 7) (the SWIZZLE_128B pattern):
//   W tile   : cp.async from the column-major Wp, chunk c of column j at  j*128 + ((c ^ (j&7)) * 16)
//   Mt tile  : raw int8 via cp.async, then converted ONE STAGE AHEAD by all threads into the same layout
// so the DMMA loop contains only LDS.128 and DMMA (measured: ALU work inside the DMMA stream costs
// ~25 % of the FP64 tensor rate, independent ALU work next to it is free).
// K permutation: within a 16-deep half, lane lc (= lane & 3) supplies k = 4*lc + 2*p + e for the MMA
// steps (p,e) in {0,1}^2 -- the same bijection for A and B, so the product is unchanged; it makes each
// lane's operands two 16-byte chunks (2*lc, 2*lc+1), and chunk ^ (row&7) spreads a quarter-warp over
// all 8 bank groups (conflict-free LDS.128).
__global__ void __launch_bounds__(SC_THREADS, 1) scan_f64_kernel(const ScanParams p) {
    extern __shared__ __align__(128) uint8_t smem[];
    double* red = reinterpret_cast<double*>(smem + SC_OFF_RED);  // [4][SC_BM]

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int wm = warp & 3, wn = warp >> 2;
    const int lr = lane >> 2, lc = lane & 3;
    const int total_steps = p.NP * p.KT;
    const int pa = (int)(p.n / SC_BN);            // panel holding column n (= v)
    const int ca = (int)(p.n - (int64_t)pa * SC_BN);

    // fragment offsets inside a half block (bytes): row/col = base + 8*tile + lr  ->  (row & 7) == lr
    const int frag0 = (wm * 32 + lr) * 128 + (((2 * lc) ^ lr) * 16);      // A, p = 0 ; +mt*1024
    const int frag1 = (wm * 32 + lr) * 128 + (((2 * lc + 1) ^ lr) * 16);  // A, p = 1
    const int fragb0 = (wn * 32 + lr) * 128 + (((2 * lc) ^ lr) * 16);     // B, p = 0 ; +t*1024
    const int fragb1 = (wn * 32 + lr) * 128 + (((2 * lc + 1) ^ lr) * 16); // B, p = 1

    for (int64_t mb = blockIdx.x; mb < p.num_blocks; mb += gridDim.x) {
        const int64_t j0 = mb * SC_BM;

        // ---- asynchronous loads of k-tile s: W tile (32 KB) + raw Mt tile (4 KB)
        auto load_step = [&](int s) {
            const int pnl = s / p.KT;
            const int kt = s - pnl * p.KT;
            const int slot = (int)(s % SC_BSTAGES);
            {   // thread -> (column tid/4, 64-byte quarter q of the column's 256 bytes = half q/2, chunks 4*(q&1)..+3)
                const int c = tid >> 2, q = tid & 3;
                const uint8_t* src = reinterpret_cast<const uint8_t*>(
                    p.Wp + ((int64_t)pnl * SC_BN + c) * p.Kpad + (int64_t)kt * SC_BK) + q * 64;
                uint8_t* dst = smem + slot * SC_B_BYTES + (q >> 1) * SC_HALF_BYTES + c * 128;
#pragma unroll
                for (int i = 0; i < 4; i++) {
                    const int chunk = (q & 1) * 4 + i;
                    ptx::cp_async_16(dst + ((chunk ^ (c & 7)) * 16), src + i * 16);
                }
            }
            if (tid < 2 * SC_BM) {  // raw Mt: row tid/2, 16-byte half
                const int r = tid >> 1, h = tid & 1;
                int64_t j = j0 + r;
                if (j >= p.L) j = p.L - 1;  // tail block: duplicate a valid row, result discarded
                ptx::cp_async_16(smem + SC_OFF_RAW + slot * SC_RAW_BYTES + r * SC_BK + h * 16,
                                 p.Mt + j * p.pitch + (int64_t)kt * SC_BK + h * 16);
            }
        };
        // ---- int8 -> f64 conversion of k-tile s into A64[s & 1]; thread -> (row tid/4, 8 genotypes)
        auto convert_step = [&](int s) {
            const int r = tid >> 2, q = tid & 3;  // q: half q/2, k = 8*(q&1) .. +7 inside the half
            const uint2 w = *reinterpret_cast<const uint2*>(smem + SC_OFF_RAW + (s % SC_BSTAGES) * SC_RAW_BYTES +
                                                            r * SC_BK + q * 8);
            uint8_t* dst = smem + SC_OFF_A64 + (int)(s & 1) * SC_A64_BYTES + (q >> 1) * SC_HALF_BYTES + r * 128;
            const int c0 = (q & 1) * 4;  // first 16-byte chunk (2 doubles each)
            *reinterpret_cast<double2*>(dst + (((c0 + 0) ^ (r & 7)) * 16)) = make_double2(s8_byte_to_f64<0>(w.x), s8_byte_to_f64<1>(w.x));
            *reinterpret_cast<double2*>(dst + (((c0 + 1) ^ (r & 7)) * 16)) = make_double2(s8_byte_to_f64<2>(w.x), s8_byte_to_f64<3>(w.x));
            *reinterpret_cast<double2*>(dst + (((c0 + 2) ^ (r & 7)) * 16)) = make_double2(s8_byte_to_f64<0>(w.y), s8_byte_to_f64<1>(w.y));
            *reinterpret_cast<double2*>(dst + (((c0 + 3) ^ (r & 7)) * 16)) = make_double2(s8_byte_to_f64<2>(w.y), s8_byte_to_f64<3>(w.y));
        };

        double acc[4][4][2];
#pragma unroll
        for (int mt = 0; mt < 4; mt++)
#pragma unroll
            for (int t = 0; t < 4; t++) acc[mt][t][0] = acc[mt][t][1] = 0.0;
        double vara_run = 0.0;  // threads 0..127: running vara of marker row tid

        __syncthreads();  // previous marker block fully done with shared memory
        for (int s = 0; s < SC_BSTAGES - 1; s++) {  // groups 0,1,2
            if (s < total_steps) load_step(s);
            ptx::cp_async_commit();
        }
        ptx::cp_async_wait<SC_BSTAGES - 2>();  // group 0 landed
        __syncthreads();
        convert_step(0);

        for (int s = 0; s < total_steps; s++) {
            ptx::cp_async_wait<SC_BSTAGES - 3>();  // groups <= s+1 landed (W(s), raw Mt(s+1))
            __syncthreads();                       // ... for everybody; A64(s) written; tile s-1 fully consumed
            if (s + SC_BSTAGES - 1 < total_steps) load_step(s + SC_BSTAGES - 1);
            ptx::cp_async_commit();
            if (s + 1 < total_steps) convert_step(s + 1);  // independent of this tile's DMMAs

            const uint8_t* sB = smem + (s % SC_BSTAGES) * SC_B_BYTES;
            const uint8_t* sA = smem + SC_OFF_A64 + (int)(s & 1) * SC_A64_BYTES;
#pragma unroll
            for (int h = 0; h < 2; h++) {
#pragma unroll
                for (int pp = 0; pp < 2; pp++) {
                    double2 a[4], b[4];
#pragma unroll
                    for (int mt = 0; mt < 4; mt++)
                        a[mt] = *reinterpret_cast<const double2*>(sA + h * SC_HALF_BYTES + (pp ? frag1 : frag0) + mt * 1024);
#pragma unroll
                    for (int t = 0; t < 4; t++)
                        b[t] = *reinterpret_cast<const double2*>(sB + h * SC_HALF_BYTES + (pp ? fragb1 : fragb0) + t * 1024);
#pragma unroll
                    for (int mt = 0; mt < 4; mt++)
#pragma unroll
                        for (int t = 0; t < 4; t++) ptx::dmma_884(acc[mt][t][0], acc[mt][t][1], a[mt].x, b[t].x);
#pragma unroll
                    for (int mt = 0; mt < 4; mt++)
#pragma unroll
                        for (int t = 0; t < 4; t++) ptx::dmma_884(acc[mt][t][0], acc[mt][t][1], a[mt].y, b[t].y);
                }
            }

            const int pnl = s / p.KT;
            if (s - pnl * p.KT == p.KT - 1) {
                // ---------------- panel epilogue: fused row-dot with the marker's own genotypes
                const int64_t cbase = (int64_t)pnl * SC_BN + wn * 32 + lc * 2;
#pragma unroll
                for (int mt = 0; mt < 4; mt++) {
                    const int row = wm * 32 + mt * 8 + lr;
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
                // red[] is rewritten only after the next panel's k loop (at least one barrier later)
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
