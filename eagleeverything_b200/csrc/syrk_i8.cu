// K2 -- genomic relationship product  C = M * M^T  on the int8 tcgen05 tensor cores.
//
// Replaces the FP64 Eigen product `MMt.noalias() = genoMat * genoMat.transpose()`
// (reference: src/calculateMMt_rcpp.cpp:95, blocked variants :138, :161).  Genotypes are
// {-1,0,1}, so s8 x s8 -> s32 accumulation is EXACT (|entry| <= L < 2^31) and any summation
// order gives the same integers as the reference's double arithmetic.
//
// Layout: M is the resident int8 store, n rows x K columns, row-major (markers contiguous =
// K-major for both operands of M*M^T), row pitch a multiple of 128 B, zero padded.
// One TMA tensor map over it (128 B x 128 row boxes, SWIZZLE_128B) feeds both operands.
//
// Kernel: persistent, warp specialised, 192 threads:
//   warp 0   TMA producer  (3 box loads / stage: A 128 rows, B 2 x 128 rows), 4-stage smem ring
//   warp 1   UMMA issuer   tcgen05.mma.cta_group::1.kind::i8, M=128 N=256 K=32, 4 per stage,
//                          accumulators in TMEM, double buffered (2 x 256 columns)
//   warps 2-5 epilogue     tcgen05.ld 32x32b -> registers -> red.global.add.s32 into C
// Work unit = (output tile 128 x 256 that touches the upper triangle, K chunk).  K is split only
// when there are too few tiles to fill the SMs; integer atomics keep the result exact and
// order independent.  Tiles are ordered in compact super-rows so that one wave of CTAs shares
// operand rows in L2 while all of them stream along K together.
//
// K-phase flow control.  The L2 reuse above only happens if the CTAs of a wave stay within the L2
// window along K (126 MB / (rows of the wave x 128 B) ~ 200 k-blocks); measured without it at
// n=10k, L=1M: 372 GB of DRAM reads for a 10 GB operand (L2 hit 38 %), DRAM-bound at 47 % of the
// tensor roofline.  So K is cut into phases of SY_PHASE k-blocks; a CTA publishes "phase p landed
// in my shared memory" with one relaxed red.global.add, and no producer starts phase p before every
// active CTA has published phase p - SY_LAG (probed optimistically, a few phases per round trip).  The grid is launched cooperatively (all CTAs
// co-resident), so the soft barrier cannot deadlock.
#include <cstdlib>
#include <vector>

#include "common.cuh"
#include "ptx.cuh"

namespace eg {

constexpr int SY_BM = 128;
constexpr int SY_BN = 256;
constexpr int SY_BK = 128;  // bytes == int8 elements per stage along K (one 128-B swizzle atom)
constexpr int SY_STAGES = 4;
constexpr int SY_A_BYTES = SY_BM * SY_BK;
constexpr int SY_B_BYTES = SY_BN * SY_BK;
constexpr int SY_STAGE_BYTES = SY_A_BYTES + SY_B_BYTES;
constexpr int SY_THREADS = 192;
constexpr int SY_TMEM_COLS = 512;
constexpr int SY_PHASE = 16;  // k-blocks per flow-control phase (2 KB of K)
constexpr int SY_LAG = 4;     // a producer may run at most this many phases ahead of the slowest CTA
constexpr int SY_SMEM_BYTES = SY_STAGES * SY_STAGE_BYTES + 1024 /*align*/ + 256 /*barriers*/;

struct SyrkParams {
    int64_t n;
    int64_t ldc;
    int32_t* C;
    const int2* tiles;   // (ti, tj) per tile
    int32_t ntiles;
    int32_t kblocks_total;
    int32_t kblocks_per_chunk;
    int32_t nunits;
    uint32_t* phase_ctr;       // [rounds * phases_per_unit], zeroed before the launch; null = no flow control
    int32_t phases_per_unit;
};

// The counters are a throttle only (no data is passed through them): relaxed, GPU-scope accesses.
__device__ __forceinline__ void red_relaxed_add(uint32_t* p, uint32_t v) {
    asm volatile("red.relaxed.gpu.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_relaxed(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

// KB == false: M row-major, 2-D tensor map {K, rows};  KB == true: M K-blocked [K/128][rows][128],
// 3-D tensor map {128, rows, K/128} -- one K-block of all rows is a contiguous range of memory.
template <bool KB>
__global__ void __launch_bounds__(SY_THREADS, 1)
syrk_i8_kernel(const __grid_constant__ CUtensorMap tmap, const SyrkParams p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + SY_STAGES * SY_STAGE_BYTES);
    uint64_t* full = bars;                       // [SY_STAGES]
    uint64_t* empty = bars + SY_STAGES;          // [SY_STAGES]
    uint64_t* tmem_full = bars + 2 * SY_STAGES;  // [2]
    uint64_t* tmem_empty = tmem_full + 2;        // [2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (warp == 0 && lane == 0) {
        ptx::prefetch_tmap(&tmap);
        for (int s = 0; s < SY_STAGES; s++) {
            ptx::mbar_init(&full[s], 1);
            ptx::mbar_init(&empty[s], 1);
        }
        for (int a = 0; a < 2; a++) {
            ptx::mbar_init(&tmem_full[a], 1);
            ptx::mbar_init(&tmem_empty[a], 4);
        }
        ptx::fence_mbar_init();
    }
    if (warp == 1) ptx::tmem_alloc<SY_TMEM_COLS>(tmem_slot);
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ------------------------------------------------------------ TMA producer
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            int round = 0;
            int known = -1;  // every active CTA is known to have landed all global phases <= known
            auto need_of = [&](int gp) -> uint32_t {
                const int left = p.nunits - (gp / p.phases_per_unit) * (int)gridDim.x;
                return (uint32_t)(left < (int)gridDim.x ? left : (int)gridDim.x);
            };
            for (int u = blockIdx.x; u < p.nunits; u += gridDim.x, round++) {
                const int kc = u / p.ntiles;
                const int2 t = p.tiles[u - kc * p.ntiles];
                const int kb0 = kc * p.kblocks_per_chunk;
                const int kb1 = min(kb0 + p.kblocks_per_chunk, p.kblocks_total);
                for (int kb = kb0; kb < kb1; kb++) {
                    if (p.phase_ctr && ((kb - kb0) % SY_PHASE) == 0) {
                        // do not start global phase gp before everybody has landed phase gp - SY_LAG
                        const int gp = round * p.phases_per_unit + (kb - kb0) / SY_PHASE;
                        if (gp - SY_LAG > known) {
                            // optimistic probes (independent loads, one round trip): newest first
                            uint32_t v[SY_LAG];
#pragma unroll
                            for (int q = 0; q < SY_LAG; q++) v[q] = ld_relaxed(p.phase_ctr + max(gp - 1 - q, 0));
#pragma unroll
                            for (int q = SY_LAG - 1; q >= 0; q--)
                                if (gp - 1 - q >= 0 && gp - 1 - q > known && v[q] >= need_of(gp - 1 - q)) known = gp - 1 - q;
                            uint32_t spins = 0;
                            while (gp - SY_LAG > known) {
                                if (ld_relaxed(p.phase_ctr + gp - SY_LAG) >= need_of(gp - SY_LAG)) known = gp - SY_LAG;
                                else if (++spins > (1u << 24)) {
                                    printf("eagle: syrk flow control timed out (block %d phase %d)\n", (int)blockIdx.x, gp);
                                    __trap();
                                }
                            }
                        }
                    }
                    ptx::mbar_wait(&empty[stage], phase ^ 1);
                    uint8_t* sA = smem + stage * SY_STAGE_BYTES;
                    uint8_t* sB = sA + SY_A_BYTES;
                    ptx::mbar_expect_tx(&full[stage], SY_STAGE_BYTES);
                    if (KB) {
                        ptx::tma_load_3d(sA, &tmap, 0, t.x * SY_BM, kb, &full[stage]);
                        ptx::tma_load_3d(sB, &tmap, 0, t.y * SY_BN, kb, &full[stage]);
                        ptx::tma_load_3d(sB + SY_B_BYTES / 2, &tmap, 0, t.y * SY_BN + 128, kb, &full[stage]);
                    } else {
                        ptx::tma_load_2d(sA, &tmap, kb * SY_BK, t.x * SY_BM, &full[stage]);
                        ptx::tma_load_2d(sB, &tmap, kb * SY_BK, t.y * SY_BN, &full[stage]);
                        ptx::tma_load_2d(sB + SY_B_BYTES / 2, &tmap, kb * SY_BK, t.y * SY_BN + 128, &full[stage]);
                    }
                    if (++stage == SY_STAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ------------------------------------------------------------ UMMA issuer
        if (lane == 0) {
            constexpr uint32_t idesc = ptx::make_idesc_i8(SY_BM, SY_BN);
            int stage = 0;
            uint32_t phase = 0;
            int acc = 0;
            uint32_t acc_phase = 0;
            int round = 0;
            for (int u = blockIdx.x; u < p.nunits; u += gridDim.x, round++) {
                const int kc = u / p.ntiles;
                const int kb0 = kc * p.kblocks_per_chunk;
                const int kb1 = min(kb0 + p.kblocks_per_chunk, p.kblocks_total);
                ptx::mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
                ptx::tc_fence_after();
                const uint32_t d_tmem = tmem_base + (uint32_t)(acc * SY_BN);
                for (int kb = kb0; kb < kb1; kb++) {
                    ptx::mbar_wait(&full[stage], phase);
                    if (p.phase_ctr) {
                        // the operands of this k-block are in shared memory: this CTA no longer needs them in L2
                        const int rel = kb - kb0;
                        if ((rel % SY_PHASE) == SY_PHASE - 1 || kb == kb1 - 1) {
                            const int ph = rel / SY_PHASE;
                            uint32_t* c = p.phase_ctr + (int64_t)round * p.phases_per_unit;
                            red_relaxed_add(c + ph, 1u);
                            if (kb == kb1 - 1)  // short last chunk: publish the phases this unit does not have
                                for (int q = ph + 1; q < p.phases_per_unit; q++) red_relaxed_add(c + q, 1u);
                        }
                    }
                    ptx::tc_fence_after();
                    const uint32_t a_addr = ptx::smem_u32(smem + stage * SY_STAGE_BYTES);
                    const uint64_t a_desc = ptx::make_desc_k_sw128(a_addr);
                    const uint64_t b_desc = ptx::make_desc_k_sw128(a_addr + SY_A_BYTES);
#pragma unroll
                    for (int k = 0; k < SY_BK / 32; k++) {
                        // advance 32 bytes along K inside the swizzle atom: +2 in 16-byte units
                        ptx::umma_i8(d_tmem, a_desc + (uint64_t)(2 * k), b_desc + (uint64_t)(2 * k), idesc,
                                     (kb > kb0 || k > 0) ? 1u : 0u);
                    }
                    ptx::umma_commit(&empty[stage]);  // frees the smem stage once these MMAs retire
                    if (++stage == SY_STAGES) { stage = 0; phase ^= 1; }
                }
                ptx::umma_commit(&tmem_full[acc]);    // accumulator ready for the epilogue
                if (++acc == 2) { acc = 0; acc_phase ^= 1; }
            }
        }
    } else {
        // ------------------------------------------------------------ epilogue (warps 2..5)
        const int q = warp & 3;  // TMEM lane quarter this warp may access
        int acc = 0;
        uint32_t acc_phase = 0;
        for (int u = blockIdx.x; u < p.nunits; u += gridDim.x) {
            const int kc = u / p.ntiles;
            const int2 t = p.tiles[u - kc * p.ntiles];
            ptx::mbar_wait(&tmem_full[acc], acc_phase);
            ptx::tc_fence_after();
            const int64_t row0 = (int64_t)t.x * SY_BM + q * 32;
            const int64_t row = row0 + lane;
            int32_t* crow = p.C + row * p.ldc;
#pragma unroll 1
            for (int c = 0; c < SY_BN / 32; c++) {
                const int64_t col0 = (int64_t)t.y * SY_BN + c * 32;
                if (col0 + 31 < row0 || col0 >= p.n) continue;  // strictly lower triangle / out of range
                uint32_t v[32];
                ptx::tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * SY_BN + c * 32), v);
                ptx::tmem_ld_wait();
                if (row < p.n) {
#pragma unroll
                    for (int j = 0; j < 32; j++) {
                        if (col0 + j < p.n) atomicAdd(crow + col0 + j, (int32_t)v[j]);
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
        ptx::tmem_dealloc<SY_TMEM_COLS>(tmem_base);
    }
}

// ------------------------------------------------------------------ host side
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static PFN_encodeTiled get_encode_fn() {
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

struct TileTable {
    int64_t n = -1;
    int device = -1;
    int2* d_tiles = nullptr;
    int ntiles = 0;
    uint32_t* d_phase = nullptr;  // flow-control counters
    size_t phase_cap = 0;
};
static thread_local TileTable g_tiles;

// tiles (ti, tj) with tj >= ti/2 (128-row x 256-column tiles touching the upper triangle), ordered
// by super-rows of 16 tile rows, then tile column, then tile row.
static int build_tiles(int64_t n, cudaStream_t st) {
    int dev = 0;
    EG_CUDA(cudaGetDevice(&dev));
    if (g_tiles.n == n && g_tiles.device == dev && g_tiles.d_tiles) return EG_OK;
    if (g_tiles.device != dev && g_tiles.d_phase) {
        cudaFree(g_tiles.d_phase);
        g_tiles.d_phase = nullptr;
        g_tiles.phase_cap = 0;
    }
    const int TM = (int)((n + SY_BM - 1) / SY_BM), TN = (int)((n + SY_BN - 1) / SY_BN);
    std::vector<int2> h;
    for (int sr = 0; sr < TM; sr += 16)
        for (int tj = sr >> 1; tj < TN; tj++)
            for (int ti = sr; ti < TM && ti < sr + 16; ti++)
                if (tj >= (ti >> 1)) h.push_back(make_int2(ti, tj));
    if (g_tiles.d_tiles) cudaFree(g_tiles.d_tiles);
    g_tiles.d_tiles = nullptr;
    EG_CUDA(cudaMalloc(&g_tiles.d_tiles, h.size() * sizeof(int2)));
    EG_CUDA(cudaMemcpyAsync(g_tiles.d_tiles, h.data(), h.size() * sizeof(int2), cudaMemcpyHostToDevice, st));
    EG_CUDA(cudaStreamSynchronize(st));
    g_tiles.n = n;
    g_tiles.device = dev;
    g_tiles.ntiles = (int)h.size();
    return EG_OK;
}

void syrk_release_cache() {
    if (g_tiles.d_tiles) cudaFree(g_tiles.d_tiles);
    if (g_tiles.d_phase) cudaFree(g_tiles.d_phase);
    g_tiles = TileTable();
}

}  // namespace eg

static int syrk_launch(const int8_t* d_M, int64_t n, int64_t kcols, int64_t pitch, bool kblocked, int32_t* d_C,
                       int64_t ldc, void* stream) {
    using namespace eg;
    if (!d_M || !d_C || n <= 0 || kcols < 0 || ldc < n || ((uintptr_t)d_M & 127) ||
        (!kblocked && ((pitch & 127) || pitch < round_up(kcols, 128))))
        return set_error(EG_ERR_ARG, "eg_dev_syrk_i8: bad argument");
    if (kcols == 0) return EG_OK;
    cudaStream_t st = (cudaStream_t)stream;
    PFN_encodeTiled enc = get_encode_fn();
    if (!enc) return set_error(EG_ERR_CUDA, "cuTensorMapEncodeTiled is not available (no CUDA driver?)");
    EG_TRY(build_tiles(n, st));

    CUtensorMap tmap;
    const int64_t kblocks = (kcols + SY_BK - 1) / SY_BK;
    CUresult r;
    if (kblocked) {
        const cuuint64_t gdim[3] = {128, (cuuint64_t)n, (cuuint64_t)kblocks};
        const cuuint64_t gstride[2] = {128, (cuuint64_t)n * 128};
        const cuuint32_t box[3] = {128, 128, 1};
        const cuuint32_t estr[3] = {1, 1, 1};
        r = enc(&tmap, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, const_cast<int8_t*>(d_M), gdim, gstride, box, estr,
                CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    } else {
        const cuuint64_t gdim[2] = {(cuuint64_t)round_up(kcols, 128), (cuuint64_t)n};
        const cuuint64_t gstride[1] = {(cuuint64_t)pitch};
        const cuuint32_t box[2] = {128, 128};
        const cuuint32_t estr[2] = {1, 1};
        r = enc(&tmap, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, const_cast<int8_t*>(d_M), gdim, gstride, box, estr,
                CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    }
    if (r != CUDA_SUCCESS) return set_error(EG_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d)", (int)r);

    SyrkParams p;
    p.n = n;
    p.ldc = ldc;
    p.C = d_C;
    p.tiles = g_tiles.d_tiles;
    p.ntiles = g_tiles.ntiles;
    p.kblocks_total = (int32_t)kblocks;
    const int sms = num_sms();
    int kchunks = 1;
    const int target = 4 * sms;
    if (p.ntiles < target) {
        kchunks = (target + p.ntiles - 1) / p.ntiles;
        int maxc = p.kblocks_total / 32;  // at least 4096 markers per chunk
        if (maxc < 1) maxc = 1;
        if (kchunks > maxc) kchunks = maxc;
    }
    p.kblocks_per_chunk = (p.kblocks_total + kchunks - 1) / kchunks;
    kchunks = (p.kblocks_total + p.kblocks_per_chunk - 1) / p.kblocks_per_chunk;
    p.nunits = p.ntiles * kchunks;

    auto kern = kblocked ? syrk_i8_kernel<true> : syrk_i8_kernel<false>;
    EG_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SY_SMEM_BYTES));
    const int grid = p.nunits < sms ? p.nunits : sms;

    // flow-control counters: one per (round, phase); zeroed on the launch stream
    p.phases_per_unit = (p.kblocks_per_chunk + SY_PHASE - 1) / SY_PHASE;
    const int rounds = (p.nunits + grid - 1) / grid;
    const size_t nctr = (size_t)rounds * p.phases_per_unit;
    const char* env_fc = getenv("EAGLE_SYRK_FLOWCTL");
    const bool flow = !(env_fc && env_fc[0] == '0') && grid > 1;
    p.phase_ctr = nullptr;
    if (flow) {
        if (nctr > g_tiles.phase_cap) {
            if (g_tiles.d_phase) cudaFree(g_tiles.d_phase);
            g_tiles.d_phase = nullptr;
            EG_CUDA(cudaMalloc(&g_tiles.d_phase, nctr * sizeof(uint32_t)));
            g_tiles.phase_cap = nctr;
        }
        EG_CUDA(cudaMemsetAsync(g_tiles.d_phase, 0, nctr * sizeof(uint32_t), st));
        p.phase_ctr = g_tiles.d_phase;
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)grid);
    cfg.blockDim = dim3(SY_THREADS);
    cfg.dynamicSmemBytes = SY_SMEM_BYTES;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeCooperative;  // all CTAs co-resident: the soft barrier cannot deadlock
    attr[0].val.cooperative = flow ? 1 : 0;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    EG_CUDA(cudaLaunchKernelEx(&cfg, kern, tmap, p));
    return check_launch("syrk_i8_kernel");
}

extern "C" int eg_dev_syrk_i8(const int8_t* d_M, int64_t n, int64_t kcols, int64_t pitch, int32_t* d_C, int64_t ldc,
                              void* stream) {
    return syrk_launch(d_M, n, kcols, pitch, false, d_C, ldc, stream);
}
// M in the K-blocked layout [ceil(kcols/128)][n][128] (eg_dev_decode_kb): every K step of the contraction reads
// one contiguous n*128-byte range instead of n rows a whole row pitch apart.
extern "C" int eg_dev_syrk_i8_kb(const int8_t* d_Mkb, int64_t n, int64_t kcols, int32_t* d_C, int64_t ldc, void* stream) {
    return syrk_launch(d_Mkb, n, kcols, 0, true, d_C, ldc, stream);
}
