// K3p -- the scan's n^3 pre-products  X = V * S,  W = S * X  (reference: src/calculate_a_and_vara_rcpp.cpp:97-98)
// evaluated on the int8 tensor cores instead of two cuBLAS DGEMMs.
//
// Same idea as scan_i8.cu, now with BOTH operands FP64 (an Ozaki-style splitting):
//   every column c of a matrix P is written as  P_kc = 2^(e_c-55) * sum_{s=0..6} q_s(k,c) 256^(6-s) + r,
//   q_s balanced base-256 digits (int8), |r| <= 2^(e_c-56): the full significand of the column's largest entry.
//   With left rows  A_i. = 2^(ea_i-55) sum_p a_p 256^(6-p)  and right columns  B_.j = 2^(eb_j-55) sum_q b_q 256^(6-q):
//       (A B)_ij = 2^(ea_i+eb_j-110) * 256^6 * sum_{d=0..12} 256^(6-d) * D_d(i,j),   D_d = sum_{p+q=d} sum_k a_p(i,k) b_q(k,j).
//   D_d is an EXACT int32 (one int8 GEMM per (p,q), accumulated in TMEM; terms are grouped so that
//   #terms * n * 2^14 < 2^31).  The anti-diagonals d = 0..6 are kept (28 int8 GEMMs); d >= 7 is below
//   6 * 2^-56 of rowmax * colmax * n, i.e. under the rounding error bound of an FP64 GEMM.  The levels are
//   combined in FP64 from the smallest weight to the largest (7 roundings per entry, fixed order).
// The left operand is needed by ROWS; both S and V are symmetric on this path (checked on the device by the
// caller, eg_dev_inputs_symmetric; otherwise the cuBLAS path is taken), so rows are read as columns.
//
// Kernel: the warp-specialised tcgen05 pipeline of syrk_i8.cu (TMA producer, one-thread UMMA issuer M128 N256
// K32, double-buffered TMEM, 4 epilogue warps).  Unit = (128 x 256 output tile, level chunk); a CTA owns a tile
// for all its chunks (the running FP64 sum is read-modify-written by the same thread, L2 resident), and the CTAs
// of a wave work on the same chunk of neighbouring tiles so that one slice of the operand rows of the wave
// (~45 MB at n = 10k) is shared through L2.
#include <cstdlib>
#include <vector>

#include "common.cuh"
#include "ptx.cuh"

namespace eg {

constexpr int PI_BM = 128;
constexpr int PI_BN = 256;
constexpr int PI_BK = 128;
constexpr int PI_STAGES = 4;
constexpr int PI_A_BYTES = PI_BM * PI_BK;
constexpr int PI_B_BYTES = PI_BN * PI_BK;
constexpr int PI_STAGE_BYTES = PI_A_BYTES + PI_B_BYTES;
constexpr int PI_THREADS = 192;
constexpr int PI_TMEM_COLS = 512;
constexpr int PI_SLICES = 7;
constexpr int PI_LEVELS = 7;                               // anti-diagonals d = p + q kept
constexpr int PI_TERMS = PI_LEVELS * (PI_LEVELS + 1) / 2;  // 28
constexpr int PI_SMEM_BYTES = PI_STAGES * PI_STAGE_BYTES + 1024 + 256;
// K lock-step (same soft barrier as syrk_i8.cu, finer grain): one (p, q) term of a wave touches ~50 MB of operand
// rows, the running FP64 tiles another ~38 MB, and without the throttle the CTAs drift apart until neither stays
// in L2 (measured at n = 10k: 196 GB of DRAM reads, L2 hit 41 %, DRAM-bound at 43 % tensor pipe).
constexpr int PI_SROWS = 12;  // tile rows per super-row of the tile order (a wave = PI_SROWS x ~12 tiles)     (EAGLE_PREP_SROWS)
constexpr int PI_PHASE = 8;   // k-blocks per phase                                                   (EAGLE_PREP_PHASE)
constexpr int PI_LAG = 4;     // phases a producer may run ahead of the slowest CTA of its wave, <= 8  (EAGLE_PREP_LAG)

struct PrepParams {
    int64_t M, N, ld;      // output block (rows of the left operand x columns of the right), leading dimension
    double* out;
    const double* sL;      // 2^(e-55) per left row of the block
    const double* sR;      // 2^(e-55) per right column of the block
    int32_t lrow0, rrow0;  // first row of the block inside the left / right slice arrays
    int32_t KB;
    int64_t diag_shift;    // upper_only: entry (i, j) of the block is needed iff j + diag_shift >= i
    int32_t upper_only;
    const int2* tiles;
    int32_t ntiles, nchunks, ngroups;
    uint8_t tp[PI_TERMS], tq[PI_TERMS];
    uint8_t cbeg[PI_TERMS + 1];
    double cw[PI_TERMS];   // 256^(6-d) of the chunk
    uint32_t* phase_ctr;   // [ngroups * phases_per_tile], zeroed before the launch; null = no flow control
    int32_t phases_per_tile, phase_len, lag;
};

__device__ __forceinline__ void pi_red_add(uint32_t* p, uint32_t v) {
    asm volatile("red.relaxed.gpu.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t pi_ld(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

__global__ void __launch_bounds__(PI_THREADS, 1)
prep_i8_kernel(const __grid_constant__ CUtensorMap tmapL, const __grid_constant__ CUtensorMap tmapR, const PrepParams p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + PI_STAGES * PI_STAGE_BYTES);
    uint64_t* full = bars;
    uint64_t* empty = bars + PI_STAGES;
    uint64_t* tmem_full = bars + 2 * PI_STAGES;
    uint64_t* tmem_empty = tmem_full + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp == 0 && lane == 0) {
        ptx::prefetch_tmap(&tmapL);
        ptx::prefetch_tmap(&tmapR);
        for (int s = 0; s < PI_STAGES; s++) {
            ptx::mbar_init(&full[s], 1);
            ptx::mbar_init(&empty[s], 1);
        }
        for (int a = 0; a < 2; a++) {
            ptx::mbar_init(&tmem_full[a], 1);
            ptx::mbar_init(&tmem_empty[a], 4);
        }
        ptx::fence_mbar_init();
    }
    if (warp == 1) ptx::tmem_alloc<PI_TMEM_COLS>(tmem_slot);
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    // unit (group gi, chunk c) of this CTA: tile gi * gridDim.x + blockIdx.x, chunks in order
    const int nun = p.ngroups * p.nchunks;

    if (warp == 0) {
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            int known = -1, r = 0;  // every CTA of the wave has landed all global phases <= known; r = k-block in the tile
            auto need_of = [&](int gp) -> uint32_t {
                const int left = p.ntiles - (gp / p.phases_per_tile) * (int)gridDim.x;
                return (uint32_t)(left < (int)gridDim.x ? left : (int)gridDim.x);
            };
            for (int u = 0; u < nun; u++) {
                const int gi = u / p.nchunks, c = u - gi * p.nchunks;
                const int tile = gi * (int)gridDim.x + (int)blockIdx.x;
                if (tile >= p.ntiles) break;
                const int2 t = p.tiles[tile];
                if (c == 0) r = 0;
                for (int term = p.cbeg[c]; term < p.cbeg[c + 1]; term++) {
                    const int sp = p.tp[term], sq = p.tq[term];
                    for (int kb = 0; kb < p.KB; kb++, r++) {
                        if (p.phase_ctr && (r % p.phase_len) == 0) {
                            const int gp = gi * p.phases_per_tile + r / p.phase_len;
                            if (gp - p.lag > known) {
                                uint32_t v[8];
#pragma unroll
                                for (int q = 0; q < 8; q++) v[q] = q < p.lag ? pi_ld(p.phase_ctr + max(gp - 1 - q, 0)) : 0u;
#pragma unroll
                                for (int q = 7; q >= 0; q--)
                                    if (q < p.lag && gp - 1 - q >= 0 && gp - 1 - q > known && v[q] >= need_of(gp - 1 - q)) known = gp - 1 - q;
                                uint32_t spins = 0;
                                while (gp - p.lag > known) {
                                    if (pi_ld(p.phase_ctr + gp - p.lag) >= need_of(gp - p.lag)) known = gp - p.lag;
                                    else if (++spins > (1u << 24)) {
                                        printf("eagle: prep_i8 flow control timed out (block %d phase %d)\n", (int)blockIdx.x, gp);
                                        __trap();
                                    }
                                }
                            }
                        }
                        ptx::mbar_wait(&empty[stage], phase ^ 1);
                        uint8_t* sA = smem + stage * PI_STAGE_BYTES;
                        uint8_t* sB = sA + PI_A_BYTES;
                        ptx::mbar_expect_tx(&full[stage], PI_STAGE_BYTES);
                        ptx::tma_load_3d(sA, &tmapL, kb * PI_BK, p.lrow0 + t.x * PI_BM, sp, &full[stage]);
                        ptx::tma_load_3d(sB, &tmapR, kb * PI_BK, p.rrow0 + t.y * PI_BN, sq, &full[stage]);
                        ptx::tma_load_3d(sB + PI_B_BYTES / 2, &tmapR, kb * PI_BK, p.rrow0 + t.y * PI_BN + 128, sq, &full[stage]);
                        if (++stage == PI_STAGES) { stage = 0; phase ^= 1; }
                    }
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            constexpr uint32_t idesc = ptx::make_idesc_i8(PI_BM, PI_BN);
            int stage = 0;
            uint32_t phase = 0;
            int acc = 0;
            uint32_t acc_phase = 0;
            int r = 0;
            const int rtot = (int)p.cbeg[p.nchunks] * p.KB;
            for (int u = 0; u < nun; u++) {
                const int gi = u / p.nchunks, c = u - gi * p.nchunks;
                const int tile = gi * (int)gridDim.x + (int)blockIdx.x;
                if (tile >= p.ntiles) break;
                if (c == 0) r = 0;
                ptx::mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
                ptx::tc_fence_after();
                const uint32_t d_tmem = tmem_base + (uint32_t)(acc * PI_BN);
                bool first = true;
                for (int term = p.cbeg[c]; term < p.cbeg[c + 1]; term++) {
                    for (int kb = 0; kb < p.KB; kb++, r++) {
                        ptx::mbar_wait(&full[stage], phase);
                        if (p.phase_ctr && ((r % p.phase_len) == p.phase_len - 1 || r == rtot - 1))
                            pi_red_add(p.phase_ctr + (int64_t)gi * p.phases_per_tile + r / p.phase_len, 1u);
                        ptx::tc_fence_after();
                        const uint32_t a_addr = ptx::smem_u32(smem + stage * PI_STAGE_BYTES);
                        const uint64_t a_desc = ptx::make_desc_k_sw128(a_addr);
                        const uint64_t b_desc = ptx::make_desc_k_sw128(a_addr + PI_A_BYTES);
#pragma unroll
                        for (int k = 0; k < PI_BK / 32; k++) {
                            ptx::umma_i8(d_tmem, a_desc + (uint64_t)(2 * k), b_desc + (uint64_t)(2 * k), idesc, first ? 0u : 1u);
                            first = false;
                        }
                        ptx::umma_commit(&empty[stage]);
                        if (++stage == PI_STAGES) { stage = 0; phase ^= 1; }
                    }
                }
                ptx::umma_commit(&tmem_full[acc]);
                if (++acc == 2) { acc = 0; acc_phase ^= 1; }
            }
        }
    } else {
        // epilogue: thread = one output row; running FP64 sum kept in the output itself
        const int q = warp & 3;
        int acc = 0;
        uint32_t acc_phase = 0;
        for (int u = 0; u < nun; u++) {
            const int gi = u / p.nchunks, c = u - gi * p.nchunks;
            const int tile = gi * (int)gridDim.x + (int)blockIdx.x;
            if (tile >= p.ntiles) break;
            const int2 t = p.tiles[tile];
            const bool first = c == 0, last = c == p.nchunks - 1;
            const double w = p.cw[c];
            const int64_t row0 = (int64_t)t.x * PI_BM + q * 32;
            const int64_t row = row0 + lane;
            const double srow = (last && row < p.M) ? p.sL[row] * 281474976710656.0 /* 256^6 */ : 0.0;
            ptx::mbar_wait(&tmem_full[acc], acc_phase);
            ptx::tc_fence_after();
#pragma unroll 1
            for (int cc = 0; cc < PI_BN / 32; cc++) {
                const int64_t col0 = (int64_t)t.y * PI_BN + cc * 32;
                if (col0 >= p.N) continue;
                if (p.upper_only && col0 + 31 + p.diag_shift < row0) continue;
                uint32_t v[32];
                ptx::tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * PI_BN + cc * 32), v);
                // the 32 running sums of this thread are fetched together, under the TMEM load: read one by one between
                // the stores (which the compiler must assume to alias) they cost 256 serial L2 round trips per chunk
                double* o = p.out + row + col0 * p.ld;
                const bool live = row < p.M;
                double run[32];
#pragma unroll
                for (int j = 0; j < 32; j++) run[j] = (live && !first && col0 + j < p.N) ? o[(int64_t)j * p.ld] : 0.0;
                ptx::tmem_ld_wait();
                if (live) {
#pragma unroll
                    for (int j = 0; j < 32; j++) {
                        if (col0 + j < p.N) {
                            double x = w * (double)(int)v[j];          // exact: |D| < 2^31, w a power of two
                            if (!first) x += run[j];                   // one rounding per level
                            if (last) x *= srow * __ldg(p.sR + col0 + j);
                            o[(int64_t)j * p.ld] = x;
                        }
                    }
                }
            }
            ptx::tc_fence_before();
            __syncwarp();
            if (lane == 0) ptx::mbar_arrive(&tmem_empty[acc]);
            if (++acc == 2) { acc = 0; acc_phase ^= 1; }
        }
    }

    ptx::tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        ptx::tc_fence_after();
        ptx::tmem_dealloc<PI_TMEM_COLS>(tmem_base);
    }
}

// ------------------------------------------------------------------ the same product on CTA pairs (tcgen05 cta_group::2)
// Output tile 256 x 256 over a cluster of two CTAs: each CTA stages its own 128 left rows and HALF of the 256 right rows
// (32 KB per k-block and SM instead of 48 KB, six ring stages instead of four) -- the kernel above is bound by the bytes it
// can keep in flight from L2, not by power (ncu: 52 % tensor pipe at 1.79 GHz, DRAM 19 %).  Roles as in scan_i8_pair_kernel:
// both CTAs run a TMA producer whose loads complete on the LEADER's `full` barrier, the leader's thread issues the M = 256
// MMAs and multicasts its commits to both CTAs, both CTAs' epilogue warps hand the accumulator back on the leader's barrier.
constexpr int PP_STAGES = 6;
constexpr int PP_B_BYTES = 128 * PI_BK;                       // this CTA's half of the right operand
constexpr int PP_STAGE_BYTES = PI_A_BYTES + PP_B_BYTES;       // 32 KB
constexpr int PP_SMEM_BYTES = PP_STAGES * PP_STAGE_BYTES + 1024 + 256;
constexpr int PP_BM = 2 * PI_BM;                              // tile rows of the pair

__global__ void __launch_bounds__(PI_THREADS, 1)
prep_i8_pair_kernel(const __grid_constant__ CUtensorMap tmapL, const __grid_constant__ CUtensorMap tmapR, const PrepParams p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + PP_STAGES * PP_STAGE_BYTES);
    uint64_t* full = bars;
    uint64_t* empty = bars + PP_STAGES;
    uint64_t* tmem_full = bars + 2 * PP_STAGES;
    uint64_t* tmem_empty = tmem_full + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = ptx::cluster_ctarank();
    const int cid = (int)ptx::cluster_id_x();
    const int ncl = (int)gridDim.x / 2;
    if (warp == 0 && lane == 0) {
        ptx::prefetch_tmap(&tmapL);
        ptx::prefetch_tmap(&tmapR);
        for (int s = 0; s < PP_STAGES; s++) {
            ptx::mbar_init(&full[s], 1);
            ptx::mbar_init(&empty[s], 1);
        }
        for (int a = 0; a < 2; a++) {
            ptx::mbar_init(&tmem_full[a], 1);
            ptx::mbar_init(&tmem_empty[a], 8);   // 4 epilogue warps in each CTA
        }
        ptx::fence_mbar_init();
    }
    if (warp == 1) ptx::tmem_alloc_pair<PI_TMEM_COLS>(tmem_slot);
    ptx::tc_fence_before();
    ptx::cluster_sync();
    ptx::tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    // unit (group gi, chunk c) of this cluster: tile gi * ncl + cid, chunks in order
    const int nun = p.ngroups * p.nchunks;

    if (warp == 0) {
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            int known = -1, r = 0;
            auto need_of = [&](int gp) -> uint32_t {
                const int left = p.ntiles - (gp / p.phases_per_tile) * ncl;
                return (uint32_t)(left < ncl ? left : ncl);
            };
            for (int u = 0; u < nun; u++) {
                const int gi = u / p.nchunks, c = u - gi * p.nchunks;
                const int tile = gi * ncl + cid;
                if (tile >= p.ntiles) break;
                const int2 t = p.tiles[tile];
                if (c == 0) r = 0;
                for (int term = p.cbeg[c]; term < p.cbeg[c + 1]; term++) {
                    const int sp = p.tp[term], sq = p.tq[term];
                    for (int kb = 0; kb < p.KB; kb++, r++) {
                        if (rank == 0 && p.phase_ctr && (r % p.phase_len) == 0) {
                            const int gp = gi * p.phases_per_tile + r / p.phase_len;
                            if (gp - p.lag > known) {
                                uint32_t v[8];
#pragma unroll
                                for (int q = 0; q < 8; q++) v[q] = q < p.lag ? pi_ld(p.phase_ctr + max(gp - 1 - q, 0)) : 0u;
#pragma unroll
                                for (int q = 7; q >= 0; q--)
                                    if (q < p.lag && gp - 1 - q >= 0 && gp - 1 - q > known && v[q] >= need_of(gp - 1 - q)) known = gp - 1 - q;
                                uint32_t spins = 0;
                                while (gp - p.lag > known) {
                                    if (pi_ld(p.phase_ctr + gp - p.lag) >= need_of(gp - p.lag)) known = gp - p.lag;
                                    else if (++spins > (1u << 24)) {
                                        printf("eagle: prep_i8 (pair) flow control timed out (cluster %d phase %d)\n", cid, gp);
                                        __trap();
                                    }
                                }
                            }
                        }
                        ptx::mbar_wait(&empty[stage], phase ^ 1);
                        uint8_t* sA = smem + stage * PP_STAGE_BYTES;
                        uint8_t* sB = sA + PI_A_BYTES;
                        const uint32_t lead_full = ptx::mapa(ptx::smem_u32(&full[stage]), 0);
                        if (rank == 0) ptx::mbar_expect_tx(&full[stage], 2 * PP_STAGE_BYTES);
                        ptx::tma_load_3d_pair(sA, &tmapL, kb * PI_BK, p.lrow0 + t.x * PP_BM + (int)rank * PI_BM, sp, lead_full);
                        ptx::tma_load_3d_pair(sB, &tmapR, kb * PI_BK, p.rrow0 + t.y * PI_BN + (int)rank * 128, sq, lead_full);
                        if (++stage == PP_STAGES) { stage = 0; phase ^= 1; }
                    }
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0 && rank == 0) {
            constexpr uint32_t idesc = ptx::make_idesc_i8(PP_BM, PI_BN);
            int stage = 0;
            uint32_t phase = 0;
            int acc = 0;
            uint32_t acc_phase = 0;
            int r = 0;
            const int rtot = (int)p.cbeg[p.nchunks] * p.KB;
            for (int u = 0; u < nun; u++) {
                const int gi = u / p.nchunks, c = u - gi * p.nchunks;
                const int tile = gi * ncl + cid;
                if (tile >= p.ntiles) break;
                if (c == 0) r = 0;
                ptx::mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
                ptx::tc_fence_after();
                const uint32_t d_tmem = tmem_base + (uint32_t)(acc * PI_BN);
                bool first = true;
                for (int term = p.cbeg[c]; term < p.cbeg[c + 1]; term++) {
                    for (int kb = 0; kb < p.KB; kb++, r++) {
                        ptx::mbar_wait(&full[stage], phase);
                        if (p.phase_ctr && ((r % p.phase_len) == p.phase_len - 1 || r == rtot - 1))
                            pi_red_add(p.phase_ctr + (int64_t)gi * p.phases_per_tile + r / p.phase_len, 1u);
                        ptx::tc_fence_after();
                        const uint32_t a_addr = ptx::smem_u32(smem + stage * PP_STAGE_BYTES);
                        const uint64_t a_desc = ptx::make_desc_k_sw128(a_addr);
                        const uint64_t b_desc = ptx::make_desc_k_sw128(a_addr + PI_A_BYTES);
#pragma unroll
                        for (int k = 0; k < PI_BK / 32; k++) {
                            ptx::umma_i8_pair(d_tmem, a_desc + (uint64_t)(2 * k), b_desc + (uint64_t)(2 * k), idesc, first ? 0u : 1u);
                            first = false;
                        }
                        ptx::umma_commit_pair(&empty[stage], 3);
                        if (++stage == PP_STAGES) { stage = 0; phase ^= 1; }
                    }
                }
                ptx::umma_commit_pair(&tmem_full[acc], 3);
                if (++acc == 2) { acc = 0; acc_phase ^= 1; }
            }
        }
    } else {
        // epilogue: thread = one output row of this CTA's half of the tile; running FP64 sum kept in the output itself
        const int q = warp & 3;
        int acc = 0;
        uint32_t acc_phase = 0;
        for (int u = 0; u < nun; u++) {
            const int gi = u / p.nchunks, c = u - gi * p.nchunks;
            const int tile = gi * ncl + cid;
            if (tile >= p.ntiles) break;
            const int2 t = p.tiles[tile];
            const bool first = c == 0, last = c == p.nchunks - 1;
            const double w = p.cw[c];
            const int64_t row0 = (int64_t)t.x * PP_BM + (int64_t)rank * PI_BM + q * 32;
            const int64_t row = row0 + lane;
            const double srow = (last && row < p.M) ? p.sL[row] * 281474976710656.0 /* 256^6 */ : 0.0;
            ptx::mbar_wait(&tmem_full[acc], acc_phase);
            ptx::tc_fence_after();
#pragma unroll 1
            for (int cc = 0; cc < PI_BN / 32; cc++) {
                const int64_t col0 = (int64_t)t.y * PI_BN + cc * 32;
                if (col0 >= p.N) continue;
                if (p.upper_only && col0 + 31 + p.diag_shift < row0) continue;
                uint32_t v[32];
                ptx::tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * PI_BN + cc * 32), v);
                // the 32 running sums of this thread are fetched together, under the TMEM load: read one by one between
                // the stores (which the compiler must assume to alias) they cost 256 serial L2 round trips per chunk
                double* o = p.out + row + col0 * p.ld;
                const bool live = row < p.M;
                double run[32];
#pragma unroll
                for (int j = 0; j < 32; j++) run[j] = (live && !first && col0 + j < p.N) ? o[(int64_t)j * p.ld] : 0.0;
                ptx::tmem_ld_wait();
                if (live) {
#pragma unroll
                    for (int j = 0; j < 32; j++) {
                        if (col0 + j < p.N) {
                            double x = w * (double)(int)v[j];          // exact: |D| < 2^31, w a power of two
                            if (!first) x += run[j];                   // one rounding per level
                            if (last) x *= srow * __ldg(p.sR + col0 + j);
                            o[(int64_t)j * p.ld] = x;
                        }
                    }
                }
            }
            ptx::tc_fence_before();
            __syncwarp();
            if (lane == 0) ptx::mbar_arrive_cluster(ptx::mapa(ptx::smem_u32(&tmem_empty[acc]), 0));
            if (++acc == 2) { acc = 0; acc_phase ^= 1; }
        }
    }

    ptx::tc_fence_before();
    ptx::cluster_sync();  // no CTA leaves while its peer can still signal it or read its shared memory
    if (warp == 1) {
        ptx::tc_fence_after();
        ptx::tmem_dealloc_pair<PI_TMEM_COLS>(tmem_base);
    }
}

// ------------------------------------------------------------------ slicing of the columns of a column-major matrix
// |x| bit patterns order like unsigned integers, and NaN > Inf > finite: a plain integer max finds the column's
// largest magnitude and lets a NaN / Inf poison the column.
// rowscale (optional): the matrix sliced is diag(rowscale) * A  (one rounding per entry)
__global__ void __launch_bounds__(256) pi_colscale_kernel(const double* __restrict__ A, int64_t rows, int64_t ld,
                                                          int32_t* __restrict__ expo, double* __restrict__ scale,
                                                          const double* __restrict__ rowscale) {
    const int64_t c = blockIdx.x;
    unsigned long long m = 0;
    for (int64_t i = threadIdx.x; i < rows; i += 256) {
        const double x = rowscale ? A[i + c * ld] * rowscale[i] : A[i + c * ld];
        const unsigned long long b = (unsigned long long)__double_as_longlong(x) & 0x7FFFFFFFFFFFFFFFull;
        m = b > m ? b : m;
    }
    __shared__ unsigned long long sh[8];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const unsigned long long other = __shfl_xor_sync(0xffffffffu, m, o);
        m = other > m ? other : m;
    }
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = m;
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < 8; w++) m = sh[w] > m ? sh[w] : m;
        const double amax = __longlong_as_double((long long)m);
        int e = 0;
        double s = 0.0;
        if (m >= 0x7FF0000000000000ull) {
            s = __longlong_as_double(0x7FF8000000000000LL);  // NaN / Inf: as in the FP64 product, the column is lost
        } else if (amax > 0.0) {
            if (frexp(amax, &e) >= 0.9921875) e++;           // keeps the top digit, carry included, <= 127
            s = ldexp(1.0, e - 55);
        }
        expo[c] = e;
        scale[c] = s;
    }
}
// Q[s][c][k], k contiguous (Kp bytes per row, zero beyond `rows`); thread -> 4 consecutive k of one column
__global__ void __launch_bounds__(256) pi_slice_kernel(const double* __restrict__ A, int64_t rows, int64_t ld,
                                                       const int32_t* __restrict__ expo, int8_t* __restrict__ Q, int64_t Kp,
                                                       int64_t slice_stride, const double* __restrict__ rowscale) {
    const int64_t c = blockIdx.y;
    const int64_t i0 = ((int64_t)blockIdx.x * 256 + threadIdx.x) * 4;
    if (i0 >= Kp) return;
    uint32_t out[PI_SLICES];
#pragma unroll
    for (int s = 0; s < PI_SLICES; s++) out[s] = 0;
    const int e = expo[c];
#pragma unroll
    for (int d = 0; d < 4; d++) {
        const int64_t i = i0 + d;
        if (i < rows) {
            const double xs = ldexp(rowscale ? A[i + c * ld] * rowscale[i] : A[i + c * ld], 55 - e);
            long long X = fabs(xs) < 3.6e16 ? __double2ll_rn(xs) : 0;  // |xs| <= 127 * 2^48 unless the column is poisoned
#pragma unroll
            for (int s = PI_SLICES - 1; s > 0; s--) {
                const int q = (int)(int8_t)(X & 0xFF);
                out[s] |= ((uint32_t)q & 0xFFu) << (8 * d);
                X = (X - q) >> 8;
            }
            out[0] |= ((uint32_t)(int)X & 0xFFu) << (8 * d);
        }
    }
    int8_t* base = Q + c * Kp + i0;
#pragma unroll
    for (int s = 0; s < PI_SLICES; s++) *reinterpret_cast<uint32_t*>(base + (int64_t)s * slice_stride) = out[s];
}

// ------------------------------------------------------------------ host side
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static PFN_encodeTiled pi_encode_fn() {
    static PFN_encodeTiled fn = nullptr;
    if (!fn) {
        void* f = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<PFN_encodeTiled>(f);
    }
    return fn;
}
// slices [7][rows][Kp] -> 3-D map {Kp, rows, 7}, boxes of 128 B x 128 rows of one slice
static int pi_make_map(CUtensorMap* m, const void* base, int64_t Kp, int64_t rows) {
    PFN_encodeTiled enc = pi_encode_fn();
    if (!enc) return set_error(EG_ERR_CUDA, "cuTensorMapEncodeTiled is not available");
    const cuuint64_t gdim[3] = {(cuuint64_t)Kp, (cuuint64_t)rows, (cuuint64_t)PI_SLICES};
    const cuuint64_t gstride[2] = {(cuuint64_t)Kp, (cuuint64_t)(Kp * rows)};
    const cuuint32_t box[3] = {128, 128, 1};
    const cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, const_cast<void*>(base), gdim, gstride, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return set_error(EG_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d)", (int)r);
    return EG_OK;
}

// EAGLE_PREP_PAIR=1: the CTA-pair kernel (read at every call)
static bool prep_i8_pair() {
    const char* e = getenv("EAGLE_PREP_PAIR");
    return e && e[0] == '1';
}

struct PrepWorkspace {
    int device = -1;
    int8_t* qS = nullptr;  size_t qS_cap = 0;
    int8_t* qV = nullptr;  size_t qV_cap = 0;
    int8_t* qX = nullptr;  size_t qX_cap = 0;
    double* sc = nullptr;  size_t sc_cap = 0;     // scales: S | V | X
    int32_t* ex = nullptr; size_t ex_cap = 0;
    int2* tiles = nullptr; size_t tiles_cap = 0;
    uint32_t* phase = nullptr; size_t phase_cap = 0;
};
static thread_local PrepWorkspace g_pi;

// CUDA events around the two prep_i8_kernel launches of the last launch_prepare_i8 (roofline reporting)
static thread_local cudaEvent_t g_pi_ev[4] = {nullptr, nullptr, nullptr, nullptr};
static thread_local double g_pi_ops = 0.0;
static thread_local int g_pi_marks = 0;
void prep_kernel_times(double* ms, double* ops) {
    *ms = 0.0;
    *ops = g_pi_ops;
    for (int i = 0; i + 1 < g_pi_marks; i += 2) {
        float f = 0.f;
        if (cudaEventSynchronize(g_pi_ev[i + 1]) == cudaSuccess && cudaEventElapsedTime(&f, g_pi_ev[i], g_pi_ev[i + 1]) == cudaSuccess)
            *ms += f;
    }
}

void prep_i8_release() {
    cudaFree(g_pi.qS); cudaFree(g_pi.qV); cudaFree(g_pi.qX); cudaFree(g_pi.sc); cudaFree(g_pi.ex); cudaFree(g_pi.tiles); cudaFree(g_pi.phase);
    g_pi = PrepWorkspace();
}
template <class T>
static bool pi_grow(T** ptr, size_t* cap, size_t need) {
    if (need <= *cap && *ptr) return true;
    if (*ptr) cudaFree(*ptr);
    *ptr = nullptr;
    *cap = 0;
    if (malloc_retry((void**)ptr, need * sizeof(T)) != cudaSuccess) {
        cudaGetLastError();
        return false;
    }
    *cap = need;
    return true;
}

static int pi_slice(const double* d_A, int64_t rows, int64_t ld, int64_t ncols, int8_t* Q, int64_t Kp, int32_t* expo,
                    double* scale, cudaStream_t st, const double* d_rowscale = nullptr) {
    pi_colscale_kernel<<<(unsigned)ncols, 256, 0, st>>>(d_A, rows, ld, expo, scale, d_rowscale);
    EG_TRY(check_launch("pi_colscale_kernel"));
    pi_slice_kernel<<<dim3((unsigned)((Kp / 4 + 255) / 256), (unsigned)ncols), 256, 0, st>>>(d_A, rows, ld, expo, Q, Kp,
                                                                                          Kp * ncols, d_rowscale);
    return check_launch("pi_slice_kernel");
}

// out (M x N block, column-major, ld) = rows [lrow0, lrow0+M) of the left slices  x  rows [rrow0, rrow0+N) of the right
// slices (both K-major, K = n).  upper_only: entries with  j + diag_shift < i  are not needed (left untouched).
static int pi_product(const int8_t* qL, int64_t lrows, const double* sL, int64_t lrow0, int64_t M, const int8_t* qR,
                      int64_t rrows, const double* sR, int64_t rrow0, int64_t N, int64_t n, int64_t Kp, double* out,
                      int64_t ld, bool upper_only, int64_t diag_shift, cudaStream_t st) {
    // tiles ordered in compact blocks of 12 tile rows so that a wave of CTAs shares its operand rows through L2
    const bool pair = prep_i8_pair();
    const int BMt = pair ? PP_BM : PI_BM;                                          // tile rows: 256 over a CTA pair
    const int TM = (int)((M + BMt - 1) / BMt), TN = (int)((N + PI_BN - 1) / PI_BN);
    std::vector<int2> h;
    int srows = PI_SROWS;
    if (const char* e = getenv("EAGLE_PREP_SROWS")) srows = atoi(e) > 0 ? atoi(e) : srows;
    if (pair) srows = (srows + 1) / 2;                                             // the same rows per super-row
    for (int sr = 0; sr < TM; sr += srows)
        for (int tj = 0; tj < TN; tj++)
            for (int ti = sr; ti < TM && ti < sr + srows; ti++) {
                if (upper_only && (int64_t)(tj + 1) * PI_BN - 1 + diag_shift < (int64_t)ti * BMt) continue;
                h.push_back(make_int2(ti, tj));
            }
    if (h.empty()) return EG_OK;
    if (!pi_grow(&g_pi.tiles, &g_pi.tiles_cap, h.size())) return set_error(EG_ERR_ALLOC, "prep_i8: tile table");
    EG_CUDA(cudaMemcpyAsync(g_pi.tiles, h.data(), h.size() * sizeof(int2), cudaMemcpyHostToDevice, st));
    EG_CUDA(cudaStreamSynchronize(st));  // h goes out of scope

    PrepParams p;
    p.M = M; p.N = N; p.ld = ld; p.out = out;
    p.sL = sL + lrow0; p.sR = sR + rrow0;
    p.lrow0 = (int32_t)lrow0; p.rrow0 = (int32_t)rrow0;
    p.KB = (int32_t)(Kp / PI_BK);
    p.diag_shift = diag_shift;
    p.upper_only = upper_only ? 1 : 0;
    p.tiles = g_pi.tiles;
    p.ntiles = (int32_t)h.size();
    // level chunks, smallest weight first; #terms per int32 accumulation bounded by 2^31 / (n * 2^14)
    const int64_t tmax64 = (((int64_t)1 << 31) - 1) / (n * 16384);
    const int tmax = (int)(tmax64 > PI_LEVELS ? PI_LEVELS : tmax64);
    if (tmax < 1) return set_error(EG_ERR_ARG, "prep_i8: n = %lld too large for int32 accumulation", (long long)n);
    int nt = 0, nc = 0;
    for (int d = PI_LEVELS - 1; d >= 0; d--) {
        for (int p0 = 0; p0 <= d; p0 += tmax) {
            p.cbeg[nc] = (uint8_t)nt;
            double w = 1.0;
            for (int e = 0; e < PI_LEVELS - 1 - d; e++) w *= 256.0;
            p.cw[nc] = w;
            for (int pp = p0; pp <= d && pp < p0 + tmax; pp++) {
                p.tp[nt] = (uint8_t)pp;
                p.tq[nt] = (uint8_t)(d - pp);
                nt++;
            }
            nc++;
        }
    }
    p.cbeg[nc] = (uint8_t)nt;
    p.nchunks = nc;
    const int sms = num_sms();
    const int workers_max = pair ? sms / 2 : sms;                                  // CTAs, or CTA pairs
    const int workers = p.ntiles < workers_max ? p.ntiles : workers_max;
    const int grid = pair ? 2 * workers : workers;
    p.ngroups = (p.ntiles + workers - 1) / workers;

    CUtensorMap tL, tR;
    EG_TRY(pi_make_map(&tL, qL, Kp, lrows));
    EG_TRY(pi_make_map(&tR, qR, Kp, rrows));
    if (pair) EG_CUDA(cudaFuncSetAttribute(prep_i8_pair_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, PP_SMEM_BYTES));
    else EG_CUDA(cudaFuncSetAttribute(prep_i8_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, PI_SMEM_BYTES));
    p.phase_len = PI_PHASE;
    p.lag = PI_LAG;
    if (const char* e = getenv("EAGLE_PREP_PHASE")) p.phase_len = atoi(e) > 0 ? atoi(e) : p.phase_len;
    if (const char* e = getenv("EAGLE_PREP_LAG")) p.lag = atoi(e) > 0 && atoi(e) <= 8 ? atoi(e) : p.lag;
    p.phases_per_tile = (nt * p.KB + p.phase_len - 1) / p.phase_len;
    const size_t nctr = (size_t)p.ngroups * p.phases_per_tile;
    const char* env_fc = getenv("EAGLE_PREP_FLOWCTL");
    const bool flow = !(env_fc && env_fc[0] == '0') && workers > 1;
    p.phase_ctr = nullptr;
    if (flow) {
        if (!pi_grow(&g_pi.phase, &g_pi.phase_cap, nctr)) return set_error(EG_ERR_ALLOC, "prep_i8: flow-control counters");
        EG_CUDA(cudaMemsetAsync(g_pi.phase, 0, nctr * sizeof(uint32_t), st));
        p.phase_ctr = g_pi.phase;
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)grid);
    cfg.blockDim = dim3(PI_THREADS);
    cfg.dynamicSmemBytes = pair ? PP_SMEM_BYTES : PI_SMEM_BYTES;
    cfg.stream = st;
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeCooperative;  // all CTAs co-resident: the soft barrier cannot deadlock
    attr[0].val.cooperative = flow ? 1 : 0;
    attr[1].id = cudaLaunchAttributeClusterDimension;
    attr[1].val.clusterDim.x = 2;
    attr[1].val.clusterDim.y = 1;
    attr[1].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pair ? 2 : 1;
    if (!g_pi_ev[0])
        for (int i = 0; i < 4; i++) cudaEventCreate(&g_pi_ev[i]);
    if (g_pi_marks <= 2) {
        cudaEventRecord(g_pi_ev[g_pi_marks], st);
        g_pi_ops += (double)p.ntiles * nt * p.KB * 2.0 * BMt * PI_BN * PI_BK;  // executed int8 ops
    }
    if (pair) EG_CUDA(cudaLaunchKernelEx(&cfg, prep_i8_pair_kernel, tL, tR, p));
    else EG_CUDA(cudaLaunchKernelEx(&cfg, prep_i8_kernel, tL, tR, p));
    if (g_pi_marks <= 2) {
        cudaEventRecord(g_pi_ev[g_pi_marks + 1], st);
        g_pi_marks += 2;
    }
    return check_launch("prep_i8_kernel");
}

// Columns [col0, col1) of W = S (V S), rows 0 .. col1-1 (W symmetric), into Wp (ld = Kpad); X = V S[:, col0:col1]
// goes through d_tmp (n x nc, ld n).  *done = false (and nothing written) when the slices do not fit in memory:
// the caller then takes the cuBLAS path.
int launch_prepare_i8(const double* d_S, const double* d_V, int64_t n, int64_t col0, int64_t col1, double* d_tmp,
                      double* d_Wp, int64_t Kpad, cudaStream_t st, bool* done) {
    *done = false;
    g_pi_marks = 0;
    g_pi_ops = 0.0;
    int dev = 0;
    EG_CUDA(cudaGetDevice(&dev));
    if (g_pi.device != dev) {
        if (g_pi.device >= 0) prep_i8_release();
        g_pi.device = dev;
    }
    if (n * 16384 >= ((int64_t)1 << 31)) return EG_OK;  // a single term would not fit int32
    const int64_t Kp = round_up(n, 128), nc = col1 - col0;
    size_t free_b = 0, total_b = 0;
    EG_CUDA(cudaMemGetInfo(&free_b, &total_b));
    const size_t need = (size_t)PI_SLICES * Kp * (size_t)(2 * n + nc);
    const size_t have = g_pi.qS_cap + g_pi.qV_cap + g_pi.qX_cap;
    if (need > have && need - have + ((size_t)2 << 30) > free_b) return EG_OK;
    if (!pi_grow(&g_pi.qS, &g_pi.qS_cap, (size_t)PI_SLICES * Kp * n) || !pi_grow(&g_pi.qV, &g_pi.qV_cap, (size_t)PI_SLICES * Kp * n) ||
        !pi_grow(&g_pi.qX, &g_pi.qX_cap, (size_t)PI_SLICES * Kp * nc) || !pi_grow(&g_pi.sc, &g_pi.sc_cap, (size_t)(2 * n + nc)) ||
        !pi_grow(&g_pi.ex, &g_pi.ex_cap, (size_t)(2 * n + nc)))
        return EG_OK;
    double *sS = g_pi.sc, *sV = g_pi.sc + n, *sX = g_pi.sc + 2 * n;
    int32_t *eS = g_pi.ex, *eV = g_pi.ex + n, *eX = g_pi.ex + 2 * n;
    EG_TRY(pi_slice(d_S, n, n, n, g_pi.qS, Kp, eS, sS, st));
    EG_TRY(pi_slice(d_V, n, n, n, g_pi.qV, Kp, eV, sV, st));
    // calculate_a_and_vara_rcpp.cpp:97   X = V * S[:, col0:col1]   (rows of V read as its columns: V symmetric)
    EG_TRY(pi_product(g_pi.qV, n, sV, 0, n, g_pi.qS, n, sS, col0, nc, n, Kp, d_tmp, n, false, 0, st));
    EG_TRY(pi_slice(d_tmp, n, n, nc, g_pi.qX, Kp, eX, sX, st));
    // :98   W[0:col1, col0:col1] = S[0:col1, :] * X   (rows of S read as its columns: S symmetric)
    EG_TRY(pi_product(g_pi.qS, n, sS, 0, col1, g_pi.qX, nc, sX, 0, nc, n, Kp, d_Wp + col0 * Kpad, Kpad, true, col0, st));
    *done = true;
    return EG_OK;
}

// W0 = A A^T (upper triangle, into Wp with ld = Kpad) for A = U diag(rs), given Ut = U^T column-major (so that column i
// of diag(rs) Ut is row i of A) -- the ONE n^3 product the scan's right-hand side needs per forward iteration once the
// algebra runs in the basis of eigen(K) (csrc/eigbasis.cu): both operands are the same set of digit slices.
// *done = false when the slices do not fit: the caller takes the cuBLAS path.
int launch_prepare_eig_i8(const double* d_Ut, const double* d_rs, int64_t n, double* d_Wp, int64_t Kpad, cudaStream_t st,
                          bool* done) {
    *done = false;
    g_pi_marks = 0;
    g_pi_ops = 0.0;
    int dev = 0;
    EG_CUDA(cudaGetDevice(&dev));
    if (g_pi.device != dev) {
        if (g_pi.device >= 0) prep_i8_release();
        g_pi.device = dev;
    }
    if (n * 16384 >= ((int64_t)1 << 31)) return EG_OK;
    const int64_t Kp = round_up(n, 128);
    size_t free_b = 0, total_b = 0;
    EG_CUDA(cudaMemGetInfo(&free_b, &total_b));
    const size_t need = (size_t)PI_SLICES * Kp * (size_t)n;
    if (need > g_pi.qS_cap && need - g_pi.qS_cap + ((size_t)2 << 30) > free_b) return EG_OK;
    if (!pi_grow(&g_pi.qS, &g_pi.qS_cap, need) || !pi_grow(&g_pi.sc, &g_pi.sc_cap, (size_t)n) ||
        !pi_grow(&g_pi.ex, &g_pi.ex_cap, (size_t)n))
        return EG_OK;
    EG_TRY(pi_slice(d_Ut, n, n, n, g_pi.qS, Kp, g_pi.ex, g_pi.sc, st, d_rs));
    EG_TRY(pi_product(g_pi.qS, n, g_pi.sc, 0, n, g_pi.qS, n, g_pi.sc, 0, n, n, Kp, d_Wp, Kpad, true, 0, st));
    *done = true;
    return EG_OK;
}

}  // namespace eg
