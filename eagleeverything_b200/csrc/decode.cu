// K1 -- packed no-space ASCII genotypes -> int8 store.
//
// STORE ENCODING.  The reference decodes a character c to  c - '0' - 1  (AA = -1, AB = 0, BB = +1).  Every quantity
// on the path is invariant under a global sign flip of M (M M^T, m^T W m) or flips with it (a = Mt v), so the stores
// hold the NEGATED value  1 - (c - '0')  (AA = +1 = 0x01, AB = 0, BB = -1 = 0xFF) and the few consumers that see the
// sign (GEMV, ReadBlock, extract_geno) flip it back.  Reason: power.  AA is the most frequent class in genotype data,
// and -1 = 0xFF in most operand bytes costs the int8 tensor cores ~15 % of their throughput under the 1 kW cap
// (scripts/microbench/int8_encoding_power.py: cuBLASLt int8 sustains 2.25 POP/s with AA = 0xFF, 2.59 with AA = 0x01).
//
// Replaces the character loop of ReadBlock (reference: src/ReadBlock.cpp:47-58,
// value = line[ii] - '0' - 1) for the file format written by CreateASCIInospace.cpp:119-122:
// row r of the image starts at byte r*(cols_total+1), one byte per genotype, '\n' after each row.
//
// HBM-bound: algorithmic traffic = 1 byte read + 1 byte written per genotype.
// Source rows start at arbitrary byte alignment (pitch = L+1), so each work unit (one row chunk,
// or a few whole short rows) is staged into shared memory with ONE 16-byte-aligned bulk-async
// copy (cp.async.bulk, completion on an mbarrier, 4-stage ring of 16 KB per CTA, 3 CTAs per SM); warps then read aligned
// 16-byte vectors, realign them with funnel shifts (the misalignment is warp-uniform), subtract
// '1' bytewise, validate and emit aligned 16-byte stores into the padded int8 matrix.
#include <cstdlib>

#include "common.cuh"
#include "ptx.cuh"

namespace eg {

constexpr int DEC_THREADS = 256;
constexpr int DEC_WARPS = DEC_THREADS / 32;
constexpr int DEC_STAGES = 4;
constexpr int DEC_SPAN = 16384;                // max bytes of source per unit (before alignment slack)
constexpr int DEC_STAGE_BYTES = DEC_SPAN + 64; // 16 B head slack + 32 B tail over-read slack, 16-B multiple
constexpr int DEC_SMEM_BYTES = DEC_STAGES * DEC_STAGE_BYTES + 1024;

struct DecodeParams {
    const uint8_t* src;
    int64_t src_pitch;
    int64_t rows, cols;
    int8_t* dst;
    int64_t dst_pitch;
    int64_t kb_rows;          // 0: row-major dst (pitch dst_pitch); else K-blocked dst [col/128][kb_rows][128]
    int32_t* err;
    int32_t rows_per_unit;    // R
    int32_t chunk_bytes;      // CW (multiple of 512)
    int32_t chunks_per_row;
    int64_t num_units;
};

// geometry of one unit, computed once by the producer thread and published through shared memory
struct __align__(16) UnitGeom {
    int64_t r0;
    int64_t c0;
    int32_t nrows;
    int32_t out_bytes;   // output bytes per row in this chunk (multiple of 16, includes the zero pad)
    uint32_t o0;         // misalignment of (row r0, col c0) inside the staged span
    uint32_t bytes;      // span length (multiple of 16), 0 when the chunk is pure padding
};

__device__ __forceinline__ UnitGeom unit_geom(const DecodeParams& p, int64_t u, const uint8_t** a0_out) {
    UnitGeom g;
    int64_t ru;
    int32_t ch;
    if (p.kb_rows) {
        // K-blocked destination: row units fastest, so that CTAs running at the same time write adjacent rows of
        // the same 128-marker blocks (contiguous memory) instead of 128-byte pieces a whole block apart
        const int64_t row_units = p.num_units / p.chunks_per_row;
        ch = (int32_t)(u / row_units);
        ru = u - (int64_t)ch * row_units;
    } else {
        ru = u / p.chunks_per_row;
        ch = (int32_t)(u - ru * p.chunks_per_row);
    }
    g.r0 = ru * p.rows_per_unit;
    int64_t left = p.rows - g.r0;
    g.nrows = (int32_t)(left < p.rows_per_unit ? left : p.rows_per_unit);
    g.c0 = (int64_t)ch * p.chunk_bytes;
    int64_t ob = p.dst_pitch - g.c0;
    g.out_bytes = (int32_t)(ob < p.chunk_bytes ? ob : p.chunk_bytes);
    int64_t cend = g.c0 + p.chunk_bytes;
    if (cend > p.cols) cend = p.cols;
    if (cend <= g.c0) {
        *a0_out = nullptr; g.o0 = 0; g.bytes = 0;
        return g;
    }
    const uint8_t* first = p.src + g.r0 * p.src_pitch + g.c0;
    const uint8_t* last = p.src + (g.r0 + g.nrows - 1) * p.src_pitch + cend;  // exclusive
    uintptr_t fa = (uintptr_t)first;
    uintptr_t a0 = fa & ~(uintptr_t)15;
    uintptr_t a1 = ((uintptr_t)last + 15) & ~(uintptr_t)15;
    *a0_out = (const uint8_t*)a0;
    g.o0 = (uint32_t)(fa - a0);
    g.bytes = (uint32_t)(a1 - a0);
    return g;
}

// 16 source bytes starting at staged offset `soff` (any alignment; soff & 15 is warp-uniform) ->
// 16 genotypes in {-1,0,1}.  `bad` collects bits of bytes outside {'0','1','2'} among the first
// `nvalid` bytes; bytes beyond nvalid are forced to 0 (row padding).
__device__ __forceinline__ uint4 decode16(const uint8_t* sb, uint32_t soff, int nvalid, uint32_t& bad) {
    const uint32_t al = soff & ~15u, o = soff & 15u;
    const uint4 q0 = *reinterpret_cast<const uint4*>(sb + al);
    const uint4 q1 = *reinterpret_cast<const uint4*>(sb + al + 16);
    const uint32_t sh = (o & 3u) * 8u;
    uint32_t x[4];
    switch (o >> 2) {
        case 0:
            x[0] = __funnelshift_r(q0.x, q0.y, sh); x[1] = __funnelshift_r(q0.y, q0.z, sh);
            x[2] = __funnelshift_r(q0.z, q0.w, sh); x[3] = __funnelshift_r(q0.w, q1.x, sh);
            break;
        case 1:
            x[0] = __funnelshift_r(q0.y, q0.z, sh); x[1] = __funnelshift_r(q0.z, q0.w, sh);
            x[2] = __funnelshift_r(q0.w, q1.x, sh); x[3] = __funnelshift_r(q1.x, q1.y, sh);
            break;
        case 2:
            x[0] = __funnelshift_r(q0.z, q0.w, sh); x[1] = __funnelshift_r(q0.w, q1.x, sh);
            x[2] = __funnelshift_r(q1.x, q1.y, sh); x[3] = __funnelshift_r(q1.y, q1.z, sh);
            break;
        default:
            x[0] = __funnelshift_r(q0.w, q1.x, sh); x[1] = __funnelshift_r(q1.x, q1.y, sh);
            x[2] = __funnelshift_r(q1.y, q1.z, sh); x[3] = __funnelshift_r(q1.z, q1.w, sh);
            break;
    }
    uint32_t out[4];
#pragma unroll
    for (int k = 0; k < 4; k++) {
        const uint32_t d = x[k] ^ 0x30303030u;                         // '0','1','2' -> 0,1,2
        uint32_t b = (d & 0xFCFCFCFCu) | ((d & (d >> 1)) & 0x01010101u);  // any byte > 2
        uint32_t v = (0x81818181u - d) ^ 0x80808080u;                  // bytewise 1 - d (store encoding), no inter-byte borrow
        if (nvalid < 16) {
            const int nvk = nvalid - 4 * k;
            const uint32_t m = nvk >= 4 ? 0xFFFFFFFFu : (nvk > 0 ? ((1u << (8 * nvk)) - 1u) : 0u);
            b &= m;
            v &= m;
        }
        bad |= b;
        out[k] = v;
    }
    return make_uint4(out[0], out[1], out[2], out[3]);
}

__global__ void __launch_bounds__(DEC_THREADS) decode_ascii_kernel(const DecodeParams p) {
    extern __shared__ __align__(128) uint8_t dsm[];
    uint8_t* stage0 = dsm;
    UnitGeom* geom = reinterpret_cast<UnitGeom*>(dsm + DEC_STAGES * DEC_STAGE_BYTES);      // [DEC_STAGES]
    uint64_t* full = reinterpret_cast<uint64_t*>(dsm + DEC_STAGES * DEC_STAGE_BYTES + 512); // [DEC_STAGES]

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) {
        for (int s = 0; s < DEC_STAGES; s++) ptx::mbar_init(&full[s], 1);
        ptx::fence_mbar_init();
    }
    __syncthreads();

    const int64_t first_unit = blockIdx.x;
    const int64_t stride = gridDim.x;
    const int64_t my_units = p.num_units > first_unit ? (p.num_units - first_unit + stride - 1) / stride : 0;

    auto issue = [&](int64_t i) {  // thread 0 only
        const int s = (int)(i % DEC_STAGES);
        const uint8_t* a0;
        UnitGeom g = unit_geom(p, first_unit + i * stride, &a0);
        geom[s] = g;  // published by the release of the arrive below
        if (g.bytes) {
            ptx::mbar_expect_tx(&full[s], g.bytes);
            ptx::bulk_g2s(stage0 + s * DEC_STAGE_BYTES, a0, g.bytes, &full[s]);
        } else {
            ptx::mbar_arrive(&full[s]);
        }
    };
    if (tid == 0)
        for (int64_t i = 0; i < DEC_STAGES - 1 && i < my_units; i++) issue(i);

    uint32_t bad = 0;
    const uint32_t src_pitch32 = (uint32_t)p.src_pitch;
    for (int64_t i = 0; i < my_units; i++) {
        __syncthreads();  // everyone is done with unit i-1, whose stage (and geometry slot) is refilled next
        if (tid == 0 && i + DEC_STAGES - 1 < my_units) issue(i + DEC_STAGES - 1);
        const int s = (int)(i % DEC_STAGES);
        ptx::mbar_wait(&full[s], (uint32_t)((i / DEC_STAGES) & 1));

        const UnitGeom g = geom[s];
        const uint8_t* sb = stage0 + s * DEC_STAGE_BYTES;
        const int nvec = g.out_bytes >> 4;                       // 16-byte vectors per row in this chunk
        int64_t ncol = p.cols - g.c0;                            // valid source bytes per row in this chunk
        const int nvalid_row = ncol >= g.out_bytes ? g.out_bytes : (ncol > 0 ? (int)ncol : 0);
        uint32_t bad_here = 0;
        // destination of the 16-byte vector at byte vb of this chunk (chunks start at multiples of 128 columns):
        // row-major  dst + row*pitch + c0 + vb;   K-blocked  dst + (((c0+vb)>>7)*rows + row)*128 + (vb&127)
        const int64_t kbr = p.kb_rows;
        if (g.nrows == 1) {
            int8_t* drow = kbr ? p.dst + ((g.c0 >> 7) * kbr + g.r0) * 128 : p.dst + g.r0 * p.dst_pitch + g.c0;
            for (int v = warp * 32 + lane; v < nvec; v += DEC_THREADS) {
                const int vb = v << 4;
                int nv = nvalid_row - vb;
                nv = nv > 16 ? 16 : nv;
                uint4 out = make_uint4(0, 0, 0, 0);
                if (nv > 0) out = decode16(sb, g.o0 + (uint32_t)vb, nv, bad_here);
                int8_t* d = kbr ? drow + (int64_t)(vb >> 7) * kbr * 128 + (vb & 127) : drow + vb;
                *reinterpret_cast<uint4*>(d) = out;
            }
        } else {
            for (int rr = warp; rr < g.nrows; rr += DEC_WARPS) {   // one warp per (short) row
                int8_t* drow = kbr ? p.dst + ((g.c0 >> 7) * kbr + g.r0 + rr) * 128 : p.dst + (g.r0 + rr) * p.dst_pitch + g.c0;
                const uint32_t rbase = g.o0 + (uint32_t)rr * src_pitch32;
                for (int v = lane; v < nvec; v += 32) {
                    const int vb = v << 4;
                    int nv = nvalid_row - vb;
                    nv = nv > 16 ? 16 : nv;
                    uint4 out = make_uint4(0, 0, 0, 0);
                    if (nv > 0) out = decode16(sb, rbase + (uint32_t)vb, nv, bad_here);
                    int8_t* d = kbr ? drow + (int64_t)(vb >> 7) * kbr * 128 + (vb & 127) : drow + vb;
                    *reinterpret_cast<uint4*>(d) = out;
                }
            }
        }
        if (bad_here && !bad) {
            bad = 1;
            if (atomicExch(&p.err[0], 1) == 0) {  // first reporter records the unit (row of the chunk, chunk column)
                p.err[1] = (int32_t)(g.r0 & 0x7FFFFFFF);
                p.err[2] = (int32_t)(g.c0 & 0x7FFFFFFF);
                p.err[3] = (int32_t)(g.r0 >> 31);
            }
        }
    }
}

// ------------------------------------------------------------------ K-blocked destination, long rows
// The K-blocked M store [col/128][rows][128] wants, per 128-marker block, the 128-byte pieces of ADJACENT rows next
// to each other.  The row-chunk kernel above writes one row's 16 KB as 128 pieces 128*rows bytes apart (measured at
// n = 10k: 71 % of the copy bandwidth, against 79 % at n = 2k).  Here a unit is 16 rows x 1 KB of columns: sixteen
// ~1 KB bulk copies in (one per row, each 16-byte aligned on its own), eight 2 KB contiguous ranges out.  A
// producer warp (lanes = rows) and 8 consumer warps are decoupled by full/empty mbarriers: no block-wide barrier.
constexpr int DK_ROWS = 16;
constexpr int DK_CW = 1024;                          // source bytes per row per unit (8 marker blocks)
constexpr int DK_SUB = DK_CW + 64;                   // 16 B head slack + 32 B over-read slack, 64-B multiple
constexpr int DK_STAGE_BYTES = DK_ROWS * DK_SUB;     // 17,408
constexpr int DK_STAGES = 4;
constexpr int DK_THREADS = 288;                      // 8 consumer warps + 1 producer warp
constexpr int DK_SMEM_BYTES = DK_STAGES * DK_STAGE_BYTES + 1024;

struct __align__(16) DkGeom {
    int64_t r0, c0;
    int32_t nrows, out_bytes;
    uint32_t o0[DK_ROWS];
};

__global__ void __launch_bounds__(DK_THREADS) decode_kb_kernel(const DecodeParams p) {
    extern __shared__ __align__(128) uint8_t dsm[];
    uint8_t* stage0 = dsm;
    DkGeom* geom = reinterpret_cast<DkGeom*>(dsm + DK_STAGES * DK_STAGE_BYTES);                 // [DK_STAGES] x 96 B
    uint64_t* full = reinterpret_cast<uint64_t*>(dsm + DK_STAGES * DK_STAGE_BYTES + 512);        // [DK_STAGES]
    uint64_t* empty = full + DK_STAGES;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) {
        for (int s = 0; s < DK_STAGES; s++) {
            ptx::mbar_init(&full[s], 1);
            ptx::mbar_init(&empty[s], 8);
        }
        ptx::fence_mbar_init();
    }
    __syncthreads();

    const int64_t row_units = (p.rows + DK_ROWS - 1) / DK_ROWS;
    const int64_t first_unit = blockIdx.x, stride = gridDim.x;
    const int64_t my_units = p.num_units > first_unit ? (p.num_units - first_unit + stride - 1) / stride : 0;

    if (warp == 8) {
        // ------------------------------------------------------------ producer: lane r stages row r of the unit
        for (int64_t i = 0; i < my_units; i++) {
            const int s = (int)(i % DK_STAGES);
            ptx::mbar_wait(&empty[s], (uint32_t)(((i / DK_STAGES) & 1) ^ 1));
            const int64_t u = first_unit + i * stride;
            const int64_t ch = u / row_units, ru = u - ch * row_units;   // row groups fastest: neighbours write neighbours
            const int64_t r0 = ru * DK_ROWS, c0 = ch * DK_CW;
            const int64_t left = p.rows - r0;
            const int nrows = (int)(left < DK_ROWS ? left : DK_ROWS);
            int64_t cend = c0 + DK_CW;
            if (cend > p.cols) cend = p.cols;
            uint32_t bytes = 0, o0 = 0;
            const uint8_t* a0 = nullptr;
            if (lane < nrows && cend > c0) {
                const uintptr_t fa = (uintptr_t)(p.src + (r0 + lane) * p.src_pitch + c0);
                const uintptr_t la = (uintptr_t)(p.src + (r0 + lane) * p.src_pitch + cend);
                const uintptr_t b0 = fa & ~(uintptr_t)15, b1 = (la + 15) & ~(uintptr_t)15;
                a0 = (const uint8_t*)b0;
                o0 = (uint32_t)(fa - b0);
                bytes = (uint32_t)(b1 - b0);
            }
            if (lane < DK_ROWS) geom[s].o0[lane] = o0;
            if (lane == 0) {
                geom[s].r0 = r0;
                geom[s].c0 = c0;
                geom[s].nrows = nrows;
                const int64_t ob = p.dst_pitch - c0;
                geom[s].out_bytes = (int32_t)(ob < DK_CW ? ob : DK_CW);
            }
            uint32_t total = bytes;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) total += __shfl_xor_sync(0xffffffffu, total, o);
            __syncwarp();
            if (lane == 0) {
                if (total) ptx::mbar_expect_tx(&full[s], total);
                else ptx::mbar_arrive(&full[s]);
            }
            __syncwarp();
            if (bytes) ptx::bulk_g2s(stage0 + s * DK_STAGE_BYTES + lane * DK_SUB, a0, bytes, &full[s]);
        }
        return;
    }

    // ---------------------------------------------------------------- consumers: warp w decodes rows w and w + 8
    uint32_t bad = 0;
    for (int64_t i = 0; i < my_units; i++) {
        const int s = (int)(i % DK_STAGES);
        ptx::mbar_wait(&full[s], (uint32_t)((i / DK_STAGES) & 1));
        const DkGeom& g = geom[s];
        const int64_t r0 = g.r0, c0 = g.c0;
        const int nrows = g.nrows, nvec = g.out_bytes >> 4;
        const int64_t ncol = p.cols - c0;
        const int nvalid_row = ncol >= g.out_bytes ? g.out_bytes : (ncol > 0 ? (int)ncol : 0);
#pragma unroll
        for (int h = 0; h < 2; h++) {
            const int rr = warp + 8 * h;
            if (rr < nrows) {
                const uint8_t* sb = stage0 + s * DK_STAGE_BYTES + rr * DK_SUB;
                const uint32_t o0 = g.o0[rr];
                int8_t* drow = p.dst + ((c0 >> 7) * p.kb_rows + r0 + rr) * 128;
                uint32_t bad_here = 0;
                for (int v = lane; v < nvec; v += 32) {
                    const int vb = v << 4;
                    int nv = nvalid_row - vb;
                    nv = nv > 16 ? 16 : nv;
                    uint4 out = make_uint4(0, 0, 0, 0);
                    if (nv > 0) out = decode16(sb, o0 + (uint32_t)vb, nv, bad_here);
                    *reinterpret_cast<uint4*>(drow + (int64_t)(vb >> 7) * p.kb_rows * 128 + (vb & 127)) = out;
                }
                if (__any_sync(0xffffffffu, bad_here != 0) && !bad) {
                    bad = 1;
                    if (lane == 0 && atomicExch(&p.err[0], 1) == 0) {  // first reporter records (row, chunk column)
                        p.err[1] = (int32_t)((r0 + rr) & 0x7FFFFFFF);
                        p.err[2] = (int32_t)(c0 & 0x7FFFFFFF);
                        p.err[3] = (int32_t)((r0 + rr) >> 31);
                    }
                }
            }
        }
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive(&empty[s]);
    }
}

}  // namespace eg

static int decode_launch(const uint8_t* d_src, int64_t src_pitch, int64_t src_bytes_avail, int64_t rows, int64_t cols,
                         int8_t* d_dst, int64_t dst_pitch, int64_t kb_rows, int32_t* d_err, void* stream) {
    using namespace eg;
    if (!d_src || !d_dst || !d_err || rows < 0 || cols < 0 || src_pitch < cols || dst_pitch < cols ||
        (dst_pitch & 127) || ((uintptr_t)d_dst & 15))
        return set_error(EG_ERR_ARG, "eg_dev_decode: bad argument");
    if (rows == 0 || dst_pitch == 0) return EG_OK;
    // the staged span of the last unit is rounded up to 16 bytes relative to the absolute address
    {
        uintptr_t end = (uintptr_t)d_src + (rows - 1) * src_pitch + cols;
        uintptr_t end16 = (end + 15) & ~(uintptr_t)15;
        if (end16 > (uintptr_t)d_src + src_bytes_avail)
            return set_error(EG_ERR_ARG, "eg_dev_decode: source buffer needs 16 bytes of readable slack");
    }
    DecodeParams p;
    p.src = d_src; p.src_pitch = src_pitch; p.rows = rows; p.cols = cols;
    p.dst = d_dst; p.dst_pitch = dst_pitch; p.kb_rows = kb_rows; p.err = d_err;
    const char* env_dk = getenv("EAGLE_DECODE_KB_TILES");
    if (kb_rows && dst_pitch >= 4096 && !(env_dk && env_dk[0] == '0')) {
        // K-blocked destination, long rows: 16-row x 1 KB units (decode_kb_kernel)
        p.rows_per_unit = DK_ROWS;
        p.chunk_bytes = DK_CW;
        p.chunks_per_row = (int32_t)((dst_pitch + DK_CW - 1) / DK_CW);
        p.num_units = ((rows + DK_ROWS - 1) / DK_ROWS) * p.chunks_per_row;
        EG_CUDA(cudaFuncSetAttribute(decode_kb_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, DK_SMEM_BYTES));
        const int64_t cap = (int64_t)num_sms() * 3;
        decode_kb_kernel<<<(unsigned)(p.num_units < cap ? p.num_units : cap), DK_THREADS, DK_SMEM_BYTES, (cudaStream_t)stream>>>(p);
        return check_launch("decode_kb_kernel");
    }
    if (dst_pitch >= 4096 || src_pitch > DEC_SPAN / 2) {
        p.rows_per_unit = 1;
        p.chunk_bytes = DEC_SPAN;
    } else {
        // short rows: a unit is R whole rows (contiguous in the image)
        p.chunk_bytes = (int32_t)((dst_pitch + 511) & ~511LL);
        int64_t R = DEC_SPAN / src_pitch;
        p.rows_per_unit = (int32_t)(R < 1 ? 1 : (R > 64 ? 64 : R));
    }
    p.chunks_per_row = (int32_t)((dst_pitch + p.chunk_bytes - 1) / p.chunk_bytes);
    p.num_units = ((rows + p.rows_per_unit - 1) / p.rows_per_unit) * p.chunks_per_row;
    EG_CUDA(cudaFuncSetAttribute(decode_ascii_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, DEC_SMEM_BYTES));
    int64_t grid = p.num_units < (int64_t)num_sms() * 3 ? p.num_units : (int64_t)num_sms() * 3;
    decode_ascii_kernel<<<(unsigned)grid, DEC_THREADS, DEC_SMEM_BYTES, (cudaStream_t)stream>>>(p);
    return check_launch("decode_ascii_kernel");
}

extern "C" int eg_dev_decode(const uint8_t* d_src, int64_t src_pitch, int64_t src_bytes_avail, int64_t rows,
                             int64_t cols, int8_t* d_dst, int64_t dst_pitch, int32_t* d_err, void* stream) {
    return decode_launch(d_src, src_pitch, src_bytes_avail, rows, cols, d_dst, dst_pitch, 0, d_err, stream);
}

// K-blocked destination: [ceil(cols/128)][kb_rows][128] bytes; this call fills rows [row0, row0+rows) of it
// (d_dst points at the start of the whole store).  The tail of the last block is zero-filled.
extern "C" int eg_dev_decode_kb(const uint8_t* d_src, int64_t src_pitch, int64_t src_bytes_avail, int64_t rows,
                                int64_t cols, int8_t* d_dst, int64_t kb_rows, int64_t row0, int32_t* d_err,
                                void* stream) {
    if (kb_rows <= 0 || row0 < 0 || row0 + rows > kb_rows) return eg::set_error(EG_ERR_ARG, "eg_dev_decode_kb: bad rows");
    return decode_launch(d_src, src_pitch, src_bytes_avail, rows, cols, d_dst + row0 * 128, eg::round_up(cols, 128),
                         kb_rows, d_err, stream);
}
