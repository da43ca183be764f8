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
};

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
            for (int u = blockIdx.x; u < p.nunits; u += gridDim.x) {
                const int kc = u / p.ntiles;
                const int2 t = p.tiles[u - kc * p.ntiles];
                const int kb0 = kc * p.kblocks_per_chunk;
                const int kb1 = min(kb0 + p.kblocks_per_chunk, p.kblocks_total);
                for (int kb = kb0; kb < kb1; kb++) {
                    ptx::mbar_wait(&empty[stage], phase ^ 1);
                    uint8_t* sA = smem + stage * SY_STAGE_BYTES;
                    uint8_t* sB = sA + SY_A_BYTES;
                    ptx::mbar_expect_tx(&full[stage], SY_STAGE_BYTES);
                    ptx::tma_load_2d(sA, &tmap, kb * SY_BK, t.x * SY_BM, &full[stage]);
                    ptx::tma_load_2d(sB, &tmap, kb * SY_BK, t.y * SY_BN, &full[stage]);
                    ptx::tma_load_2d(sB + SY_B_BYTES / 2, &tmap, kb * SY_BK, t.y * SY_BN + 128, &full[stage]);
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
            for (int u = blockIdx.x; u < p.nunits; u += gridDim.x) {
                const int kc = u / p.ntiles;
                const int kb0 = kc * p.kblocks_per_chunk;
                const int kb1 = min(kb0 + p.kblocks_per_chunk, p.kblocks_total);
                ptx::mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
                ptx::tc_fence_after();
                const uint32_t d_tmem = tmem_base + (uint32_t)(acc * SY_BN);
                for (int kb = kb0; kb < kb1; kb++) {
                    ptx::mbar_wait(&full[stage], phase);
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
};
static thread_local TileTable g_tiles;

// tiles (ti, tj) with tj >= ti/2 (128-row x 256-column tiles touching the upper triangle), ordered
// by super-rows of 16 tile rows, then tile column, then tile row.
static int build_tiles(int64_t n, cudaStream_t st) {
    int dev = 0;
    EG_CUDA(cudaGetDevice(&dev));
    if (g_tiles.n == n && g_tiles.device == dev && g_tiles.d_tiles) return EG_OK;
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
    g_tiles = TileTable();
}

}  // namespace eg

extern "C" int eg_dev_syrk_i8(const int8_t* d_M, int64_t n, int64_t kcols, int64_t pitch, int32_t* d_C, int64_t ldc,
                              void* stream) {
    using namespace eg;
    if (!d_M || !d_C || n <= 0 || kcols < 0 || (pitch & 127) || pitch < round_up(kcols, 128) || ldc < n ||
        ((uintptr_t)d_M & 127))
        return set_error(EG_ERR_ARG, "eg_dev_syrk_i8: bad argument");
    if (kcols == 0) return EG_OK;
    cudaStream_t st = (cudaStream_t)stream;
    PFN_encodeTiled enc = get_encode_fn();
    if (!enc) return set_error(EG_ERR_CUDA, "cuTensorMapEncodeTiled is not available (no CUDA driver?)");
    EG_TRY(build_tiles(n, st));

    CUtensorMap tmap;
    const cuuint64_t gdim[2] = {(cuuint64_t)round_up(kcols, 128), (cuuint64_t)n};
    const cuuint64_t gstride[1] = {(cuuint64_t)pitch};
    const cuuint32_t box[2] = {128, 128};
    const cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(&tmap, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, const_cast<int8_t*>(d_M), gdim, gstride, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return set_error(EG_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d)", (int)r);

    SyrkParams p;
    p.n = n;
    p.ldc = ldc;
    p.C = d_C;
    p.tiles = g_tiles.d_tiles;
    p.ntiles = g_tiles.ntiles;
    p.kblocks_total = (int32_t)((kcols + SY_BK - 1) / SY_BK);
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

    static thread_local bool attr_set = false;
    if (!attr_set) {
        EG_CUDA(cudaFuncSetAttribute(syrk_i8_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SY_SMEM_BYTES));
        attr_set = true;
    }
    const int grid = p.nunits < sms ? p.nunits : sms;
    syrk_i8_kernel<<<grid, SY_THREADS, SY_SMEM_BYTES, st>>>(tmap, p);
    return check_launch("syrk_i8_kernel");
}
