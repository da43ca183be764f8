// K1 -- packed no-space ASCII genotypes -> int8 in {-1,0,1}.
//
// Replaces the character loop of ReadBlock (reference: src/ReadBlock.cpp:47-58,
// value = line[ii] - '0' - 1) for the file format written by CreateASCIInospace.cpp:119-122:
// row r of the image starts at byte r*(cols_total+1), one byte per genotype, '\n' after each row.
//
// HBM-bound: algorithmic traffic = 1 byte read + 1 byte written per genotype.
// Source rows start at arbitrary byte alignment (pitch = L+1), so each work unit (one row chunk,
// or a few whole short rows) is staged into shared memory with ONE 16-byte-aligned bulk-async
// copy (cp.async.bulk, completion on an mbarrier, 4-stage ring per CTA); warps then read aligned
// 16-byte vectors, realign them with funnel shifts (the misalignment is warp-uniform), subtract
// '1' bytewise, validate and emit aligned 16-byte stores into the padded int8 matrix.
#include "common.cuh"
#include "ptx.cuh"

namespace eg {

constexpr int DEC_THREADS = 256;
constexpr int DEC_STAGES = 4;
constexpr int DEC_SPAN = 8192;                 // max bytes of source per unit (before alignment slack)
constexpr int DEC_STAGE_BYTES = DEC_SPAN + 64; // 16 B head slack + 32 B tail over-read slack, 16-B multiple

struct DecodeParams {
    const uint8_t* src;
    int64_t src_pitch;
    int64_t rows, cols;
    int8_t* dst;
    int64_t dst_pitch;
    int32_t* err;
    int32_t rows_per_unit;    // R
    int32_t chunk_bytes;      // CW (multiple of 512)
    int32_t chunks_per_row;
    int64_t num_units;
};

struct UnitGeom {
    int64_t r0;
    int32_t nrows;
    int64_t c0;
    int32_t out_bytes;   // output bytes per row in this chunk (multiple of 16, includes zero pad)
    const uint8_t* a0;   // 16-B aligned start of the staged span
    uint32_t o0;         // misalignment of (row r0, col c0) inside the span
    uint32_t bytes;      // span length (multiple of 16), 0 when the chunk is pure padding
};

__device__ __forceinline__ UnitGeom unit_geom(const DecodeParams& p, int64_t u) {
    UnitGeom g;
    int64_t ru = u / p.chunks_per_row;
    int32_t ch = (int32_t)(u - ru * p.chunks_per_row);
    g.r0 = ru * p.rows_per_unit;
    int64_t left = p.rows - g.r0;
    g.nrows = (int32_t)(left < p.rows_per_unit ? left : p.rows_per_unit);
    g.c0 = (int64_t)ch * p.chunk_bytes;
    int64_t ob = p.dst_pitch - g.c0;
    g.out_bytes = (int32_t)(ob < p.chunk_bytes ? ob : p.chunk_bytes);
    int64_t cend = g.c0 + p.chunk_bytes;
    if (cend > p.cols) cend = p.cols;
    if (cend <= g.c0) {
        g.a0 = nullptr; g.o0 = 0; g.bytes = 0;
        return g;
    }
    const uint8_t* first = p.src + g.r0 * p.src_pitch + g.c0;
    const uint8_t* last = p.src + (g.r0 + g.nrows - 1) * p.src_pitch + cend;  // exclusive
    uintptr_t fa = (uintptr_t)first;
    uintptr_t a0 = fa & ~(uintptr_t)15;
    uintptr_t a1 = ((uintptr_t)last + 15) & ~(uintptr_t)15;
    g.a0 = (const uint8_t*)a0;
    g.o0 = (uint32_t)(fa - a0);
    g.bytes = (uint32_t)(a1 - a0);
    return g;
}

__global__ void __launch_bounds__(DEC_THREADS) decode_ascii_kernel(const DecodeParams p) {
    __shared__ __align__(128) uint8_t stage[DEC_STAGES][DEC_STAGE_BYTES];
    __shared__ __align__(8) uint64_t full[DEC_STAGES];

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) {
        for (int s = 0; s < DEC_STAGES; s++) ptx::mbar_init(&full[s], 1);
        ptx::fence_mbar_init();
    }
    __syncthreads();

    const int64_t first_unit = blockIdx.x;
    const int64_t stride = gridDim.x;
    int64_t my_units = p.num_units > first_unit ? (p.num_units - first_unit + stride - 1) / stride : 0;

    auto issue = [&](int64_t i) {  // thread 0 only
        UnitGeom g = unit_geom(p, first_unit + i * stride);
        uint64_t* bar = &full[i % DEC_STAGES];
        if (g.bytes) {
            ptx::mbar_expect_tx(bar, g.bytes);
            ptx::bulk_g2s(stage[i % DEC_STAGES], g.a0, g.bytes, bar);
        } else {
            ptx::mbar_arrive(bar);
        }
    };
    if (tid == 0)
        for (int64_t i = 0; i < DEC_STAGES - 1 && i < my_units; i++) issue(i);

    uint32_t bad_any = 0;
    int64_t bad_row = 0, bad_col = 0;
    for (int64_t i = 0; i < my_units; i++) {
        __syncthreads();  // everyone is done with unit i-1, whose stage is refilled next
        if (tid == 0 && i + DEC_STAGES - 1 < my_units) issue(i + DEC_STAGES - 1);
        const int s = (int)(i % DEC_STAGES);
        ptx::mbar_wait(&full[s], (uint32_t)((i / DEC_STAGES) & 1));

        const UnitGeom g = unit_geom(p, first_unit + i * stride);
        const uint8_t* sb = stage[s];
        const int segs_per_row = (g.out_bytes + 511) >> 9;
        const int nseg = g.nrows * segs_per_row;
        for (int sidx = warp; sidx < nseg; sidx += DEC_THREADS / 32) {
            const int rr = sidx / segs_per_row;
            const int sg = sidx - rr * segs_per_row;
            const int vbyte = (sg * 32 + lane) * 16;  // byte offset inside the chunk
            if (vbyte >= g.out_bytes) continue;
            const int64_t col = g.c0 + vbyte;
            int64_t nv64 = p.cols - col;
            const int nvalid = nv64 >= 16 ? 16 : (nv64 > 0 ? (int)nv64 : 0);
            uint4 out = make_uint4(0, 0, 0, 0);
            if (nvalid > 0) {
                const uint32_t soff = g.o0 + (uint32_t)rr * (uint32_t)p.src_pitch + (uint32_t)vbyte;
                const uint32_t al = soff & ~15u, o = soff & 15u;  // o is warp-uniform
                const uint4 q0 = *reinterpret_cast<const uint4*>(sb + al);
                const uint4 q1 = *reinterpret_cast<const uint4*>(sb + al + 16);
                const uint32_t sh = (o & 3u) * 8u;
                uint32_t x0, x1, x2, x3;
                switch (o >> 2) {
                    case 0:
                        x0 = __funnelshift_r(q0.x, q0.y, sh); x1 = __funnelshift_r(q0.y, q0.z, sh);
                        x2 = __funnelshift_r(q0.z, q0.w, sh); x3 = __funnelshift_r(q0.w, q1.x, sh);
                        break;
                    case 1:
                        x0 = __funnelshift_r(q0.y, q0.z, sh); x1 = __funnelshift_r(q0.z, q0.w, sh);
                        x2 = __funnelshift_r(q0.w, q1.x, sh); x3 = __funnelshift_r(q1.x, q1.y, sh);
                        break;
                    case 2:
                        x0 = __funnelshift_r(q0.z, q0.w, sh); x1 = __funnelshift_r(q0.w, q1.x, sh);
                        x2 = __funnelshift_r(q1.x, q1.y, sh); x3 = __funnelshift_r(q1.y, q1.z, sh);
                        break;
                    default:
                        x0 = __funnelshift_r(q0.w, q1.x, sh); x1 = __funnelshift_r(q1.x, q1.y, sh);
                        x2 = __funnelshift_r(q1.y, q1.z, sh); x3 = __funnelshift_r(q1.z, q1.w, sh);
                        break;
                }
                uint32_t xs[4] = {x0, x1, x2, x3};
                uint32_t os[4];
                uint32_t bad = 0;
#pragma unroll
                for (int k = 0; k < 4; k++) {
                    const int nvk = nvalid - 4 * k;
                    const uint32_t m = nvk >= 4 ? 0xFFFFFFFFu : (nvk > 0 ? ((1u << (8 * nvk)) - 1u) : 0u);
                    const uint32_t d = __vsub4(xs[k], 0x30303030u);           // byte - '0'
                    bad |= __vcmpgtu4(d, 0x02020202u) & m;                     // not in {0,1,2}
                    os[k] = __vsub4(d, 0x01010101u) & m;                       // -> {-1,0,1}, pad = 0
                }
                out = make_uint4(os[0], os[1], os[2], os[3]);
                if (bad) {
                    bad_any = 1;
                    bad_row = g.r0 + rr;
                    bad_col = col;
                }
            }
            *reinterpret_cast<uint4*>(p.dst + (g.r0 + rr) * p.dst_pitch + col) = out;
        }
    }
    if (bad_any) {
        if (atomicExch(&p.err[0], 1) == 0) {
            p.err[1] = (int32_t)(bad_row & 0x7FFFFFFF);
            p.err[2] = (int32_t)(bad_col & 0x7FFFFFFF);
            p.err[3] = (int32_t)(bad_row >> 31);
        }
    }
}

}  // namespace eg

extern "C" int eg_dev_decode(const uint8_t* d_src, int64_t src_pitch, int64_t src_bytes_avail, int64_t rows,
                             int64_t cols, int8_t* d_dst, int64_t dst_pitch, int32_t* d_err, void* stream) {
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
    p.dst = d_dst; p.dst_pitch = dst_pitch; p.err = d_err;
    if (dst_pitch >= 4096 || src_pitch > DEC_SPAN) {
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
    int64_t grid = p.num_units < (int64_t)num_sms() * 6 ? p.num_units : (int64_t)num_sms() * 6;
    decode_ascii_kernel<<<(unsigned)grid, DEC_THREADS, 0, (cudaStream_t)stream>>>(p);
    return check_launch("decode_ascii_kernel");
}
