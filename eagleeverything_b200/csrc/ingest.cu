// SURVEY.md section 8(f), rank 3 -- the ingest side of the genotype files on the device.
//
//   tokenise   whitespace-separated genotype text -> no-space ASCII rows      src/CreateASCIInospace.cpp:67-125
//   encode     int8 store -> no-space ASCII rows (the writer of Mt.ascii)     src/createMt_ASCII_rcpp.cpp:99-118
//
// TOKENISER.  The reference reads the text file line by line (getline), splits each line at whitespace
// (operator>> : space, \t, \v, \f, \r), compares every token with BB, AB, AA and the missing-value string IN THAT
// ORDER and writes '2', '1', '0', '1'; it stops at the first token that matches none of them, or at the first row
// whose token count differs from dims[1] (checked after the row's tokens).  On the device the text is one flat byte
// array; token k of the file belongs to row k / L and column k % L as long as every earlier row holds exactly L
// tokens, so two prefix sums over the bytes -- token starts and newlines -- place every token:
//   tok_count_kernel   token starts and newlines per 32 KB chunk
//   tok_scan_kernel    exclusive prefix over the chunks (one CTA; 64-bit)
//   tok_emit_kernel    per chunk: the same masks again and a CTA-wide scan per 4 KB step; every token start classifies
//                      its token, every newline checks "tokens so far == (row+1)*L", and both emit one byte.  The output
//                      offset row*(L+1)+col equals (tokens before) + (newlines before): the output is a stream
//                      compaction of the events, staged in shared memory and written as aligned 16-byte vectors.
// Errors are events with a byte position (start of the offending token / the newline that ends the short or long
// row); the smallest position is kept with one atomicMin, which is the event the reference's sequential loop meets
// first.  The host turns it into the reference's message (row number, token text or column count).
// HBM-bound: algorithmic bytes = text + (L+1) per row out; the text is read twice (count, emit).
//
// ENCODER.  The output image is treated as a flat byte stream: one thread = one aligned 16-byte vector of the file;
// vectors inside one row take two aligned 16-byte loads of the store and a funnel shift, the 1-in-(cols+1)/16 vectors
// that straddle a '\n' go byte by byte.  Algorithmic bytes = 2 per genotype.
#include <cstring>

#include "common.cuh"

namespace eg {

constexpr int TK_THREADS = 256;
constexpr int TK_STEP = TK_THREADS * 16;      // bytes per CTA step (one 16-byte vector per thread)
constexpr int TK_STEPS = 8;
constexpr int TK_CHUNK = TK_STEP * TK_STEPS;  // 32 KB per unit
constexpr int TK_MAXLEN = 15;                 // longest genotype code accepted

struct TokCodes {  // order of comparison in the reference: BB, AB, AA, missing (CreateASCIInospace.cpp:84-93)
    uint8_t s[4][16];
    int32_t len[4];
    uint8_t out[4];
};

__device__ __forceinline__ bool is_ws(uint32_t b) { return b == 0x20u || (b - 9u) <= 4u; }  // isspace in the C locale

// 0xFF / 0x00 per byte -> one bit per byte
__device__ __forceinline__ uint32_t byte_flags(uint32_t m) { return ((m & 0x08040201u) * 0x01010101u) >> 24; }
// masks over the 16 bytes at p0: whitespace (bytes at or beyond nbytes count as whitespace) and '\n'
__device__ __forceinline__ void byte_masks(const uint4& v, int64_t p0, int64_t nbytes, uint32_t& ws, uint32_t& nl) {
    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
    ws = 0;
    nl = 0;
#pragma unroll
    for (int k = 0; k < 4; k++) {
        const uint32_t sp = __vcmpeq4(w[k], 0x20202020u) | (__vcmpgeu4(w[k], 0x09090909u) & __vcmpleu4(w[k], 0x0D0D0D0Du));
        ws |= byte_flags(sp) << (4 * k);
        nl |= byte_flags(__vcmpeq4(w[k], 0x0A0A0A0Au)) << (4 * k);
    }
    const int64_t left = nbytes - p0;
    if (left < 16) {
        const uint32_t valid = left <= 0 ? 0u : ((1u << left) - 1u);
        ws |= ~valid & 0xFFFFu;
        nl &= valid;
    }
}
// token starts: a non-whitespace byte whose predecessor is whitespace (or the start of the text)
__device__ __forceinline__ uint32_t token_starts(uint32_t ws, const uint8_t* text, int64_t p0) {
    const uint32_t prev_ws = (p0 == 0) ? 1u : (is_ws(text[p0 - 1]) ? 1u : 0u);
    return ~ws & ((ws << 1) | prev_ws) & 0xFFFFu;
}

// the text buffer is readable (not necessarily meaningful) for 32 bytes beyond nbytes
__global__ void __launch_bounds__(TK_THREADS) tok_count_kernel(const uint8_t* __restrict__ text, int64_t nbytes, int64_t nchunks,
                                                               uint32_t* __restrict__ counts) {
    __shared__ uint32_t red[2][TK_THREADS / 32];
    for (int64_t ch = blockIdx.x; ch < nchunks; ch += gridDim.x) {
        uint32_t ntok = 0, nnl = 0;
#pragma unroll 2
        for (int s = 0; s < TK_STEPS; s++) {
            const int64_t p0 = ch * TK_CHUNK + (int64_t)s * TK_STEP + threadIdx.x * 16;
            if (p0 < nbytes) {
                const uint4 v = *reinterpret_cast<const uint4*>(text + p0);
                uint32_t ws, nl;
                byte_masks(v, p0, nbytes, ws, nl);
                ntok += __popc(token_starts(ws, text, p0));
                nnl += __popc(nl);
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            ntok += __shfl_xor_sync(0xffffffffu, ntok, o);
            nnl += __shfl_xor_sync(0xffffffffu, nnl, o);
        }
        __syncthreads();
        if ((threadIdx.x & 31) == 0) {
            red[0][threadIdx.x >> 5] = ntok;
            red[1][threadIdx.x >> 5] = nnl;
        }
        __syncthreads();
        if (threadIdx.x < 2) {
            uint32_t t = 0;
            for (int w = 0; w < TK_THREADS / 32; w++) t += red[threadIdx.x][w];
            counts[2 * ch + threadIdx.x] = t;
        }
    }
}

// exclusive prefix of the per-chunk counts; prefix[2*nchunks], prefix[2*nchunks+1] = totals.  One CTA of 1024.
__global__ void __launch_bounds__(1024) tok_scan_kernel(const uint32_t* __restrict__ counts, int64_t nchunks,
                                                        int64_t* __restrict__ prefix) {
    __shared__ uint32_t wsum[2][32];
    __shared__ int64_t carry[2];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x < 2) carry[threadIdx.x] = 0;
    __syncthreads();
    for (int64_t base = 0; base < nchunks; base += 1024) {
        const int64_t ch = base + threadIdx.x;
        uint32_t a = ch < nchunks ? counts[2 * ch] : 0u, b = ch < nchunks ? counts[2 * ch + 1] : 0u;
        uint32_t ia = a, ib = b;  // inclusive scans inside the warp (a tile sums to < 2^32: 1024 * 32768)
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t ua = __shfl_up_sync(0xffffffffu, ia, o), ub = __shfl_up_sync(0xffffffffu, ib, o);
            if (lane >= o) { ia += ua; ib += ub; }
        }
        if (lane == 31) { wsum[0][warp] = ia; wsum[1][warp] = ib; }
        __syncthreads();
        if (warp == 0) {
            uint32_t sa = wsum[0][lane], sb = wsum[1][lane];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const uint32_t ua = __shfl_up_sync(0xffffffffu, sa, o), ub = __shfl_up_sync(0xffffffffu, sb, o);
                if (lane >= o) { sa += ua; sb += ub; }
            }
            wsum[0][lane] = sa;  // inclusive over warps
            wsum[1][lane] = sb;
        }
        __syncthreads();
        const uint32_t wa = warp ? wsum[0][warp - 1] : 0u, wb = warp ? wsum[1][warp - 1] : 0u;
        if (ch < nchunks) {
            prefix[2 * ch] = carry[0] + (int64_t)(wa + ia - a);
            prefix[2 * ch + 1] = carry[1] + (int64_t)(wb + ib - b);
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            carry[0] += (int64_t)wsum[0][31];
            carry[1] += (int64_t)wsum[1][31];
        }
        __syncthreads();
    }
    if (threadIdx.x < 2) prefix[2 * nchunks + threadIdx.x] = carry[threadIdx.x];
}

// which code the token starting at p is: index into TokCodes (a copy in shared memory), or -1.  Deliberately not inlined:
// sixteen unrolled copies of it made the kernel 37,000 instructions long and instruction-fetch bound.
__device__ __noinline__ int classify_token(const uint8_t* __restrict__ text, int64_t p, int64_t nbytes, const TokCodes& c) {
    for (int k = 0; k < 4; k++) {
        const int len = c.len[k];
        if (len == 0 || p + len > nbytes) continue;
        bool same = true;
        for (int j = 0; j < len; j++) same &= (text[p + j] == c.s[k][j]);
        if (same && (p + len == nbytes || is_ws(text[p + len]))) return k;
    }
    return -1;
}

// The output is a stream compaction of the events: a token start writes its code, a '\n' writes '\n', and the byte goes
// to offset (tokens before) + (newlines before) -- as long as every earlier row holds exactly L tokens that IS
// row * (L + 1) + column.  Each 4 KB step stages its bytes in shared memory and writes them out as aligned 16-byte vectors.
// One-character tokens (the usual 0/1/2 or A/H/B files) are classified by a 256-entry table in shared memory;
// longer tokens (e.g. the missing-value string "NA") go through classify_token.
__global__ void __launch_bounds__(TK_THREADS) tok_emit_kernel(const uint8_t* __restrict__ text, int64_t nbytes, int64_t nchunks,
                                                              const int64_t* __restrict__ prefix, int64_t L, const TokCodes codes_param,
                                                              uint8_t* __restrict__ out, int64_t out_bytes,
                                                              unsigned long long* __restrict__ err_pos) {
    __shared__ uint32_t wsum[TK_THREADS / 32];
    __shared__ TokCodes codes;  // dynamic indexing of a kernel parameter would put it on every thread's stack
    if (threadIdx.x < sizeof(TokCodes) / 4) reinterpret_cast<uint32_t*>(&codes)[threadIdx.x] = reinterpret_cast<const uint32_t*>(&codes_param)[threadIdx.x];
    __syncthreads();
    __shared__ __align__(16) uint8_t stage[TK_STEP + 32];  // at most one event per text byte
    __shared__ uint8_t lut1[256];  // one-character token -> output byte, 0 = none of the codes
    lut1[threadIdx.x] = 0;
    __syncthreads();
    if (threadIdx.x == 0)
        for (int q = 3; q >= 0; q--)  // the first match in BB, AB, AA, missing order wins: written last
            if (codes.len[q] == 1) lut1[codes.s[q][0]] = codes.out[q];
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int64_t ch = blockIdx.x; ch < nchunks; ch += gridDim.x) {
        int64_t base_tok = prefix[2 * ch], base_nl = prefix[2 * ch + 1];
        const uint4 blank = make_uint4(0x20202020u, 0x20202020u, 0x20202020u, 0x20202020u);
        uint4 vnext = blank;  // the vector of the NEXT step is in flight while this one is processed
        {
            const int64_t q0 = ch * TK_CHUNK + threadIdx.x * 16;
            if (q0 < nbytes) vnext = *reinterpret_cast<const uint4*>(text + q0);
        }
        for (int s = 0; s < TK_STEPS; s++) {
            const int64_t p0 = ch * TK_CHUNK + (int64_t)s * TK_STEP + threadIdx.x * 16;
            if (ch * TK_CHUNK + (int64_t)s * TK_STEP >= nbytes) break;  // CTA-uniform
            const uint4 v = vnext;
            vnext = blank;
            if (s + 1 < TK_STEPS && p0 + TK_STEP < nbytes) vnext = *reinterpret_cast<const uint4*>(text + p0 + TK_STEP);
            uint32_t ws = 0xFFFFu, nl = 0, ts = 0;
            if (p0 < nbytes) {
                byte_masks(v, p0, nbytes, ws, nl);
                ts = token_starts(ws, text, p0);
            }
            // CTA-wide exclusive scan of (token starts << 16 | newlines): at most 2048 / 4096 per step
            const uint32_t mine = ((uint32_t)__popc(ts) << 16) | (uint32_t)__popc(nl);
            uint32_t inc = mine;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const uint32_t u = __shfl_up_sync(0xffffffffu, inc, o);
                if (lane >= o) inc += u;
            }
            __syncthreads();  // wsum and stage of the previous step have been read
            if (lane == 31) wsum[warp] = inc;
            __syncthreads();
            uint32_t before = inc - mine, total = 0;
#pragma unroll
            for (int w = 0; w < TK_THREADS / 32; w++) {
                const uint32_t x = wsum[w];
                if (w < warp) before += x;
                total += x;
            }
            uint32_t tl = before >> 16, rl = before & 0xFFFFu;  // tokens / newlines of this step before this thread's bytes
            uint32_t e = tl + rl;                               // = its first slot in the stage
            // does the byte after this vector end a token that starts at byte 15?
            const bool single15 = (ts & 0x8000u) && (p0 + 16 >= nbytes || is_ws(text[p0 + 16]));
            const uint64_t lo = (uint64_t)v.x | ((uint64_t)v.y << 32), hi = (uint64_t)v.z | ((uint64_t)v.w << 32);
            uint32_t ev = ts | nl;
            while (ev) {  // events in byte order; every lane of a dense file has the same number of them
                const int i = __ffs(ev) - 1;
                ev &= ev - 1;
                if ((nl >> i) & 1u) {
                    if (base_tok + tl != (base_nl + rl + 1) * L) atomicMin(err_pos, (unsigned long long)(p0 + i));
                    stage[e++] = '\n';
                    rl++;
                } else {
                    const bool single = (i < 15) ? ((ws >> (i + 1)) & 1u) : single15;
                    uint32_t ob;
                    if (single) {
                        ob = lut1[(uint32_t)((i < 8 ? lo : hi) >> (8 * (i & 7))) & 0xFFu];
                    } else {
                        const int k = classify_token(text, p0 + i, nbytes, codes);
                        ob = k >= 0 ? codes.out[k] : 0u;
                    }
                    if (!ob) atomicMin(err_pos, (unsigned long long)(p0 + i));
                    stage[e++] = ob ? (uint8_t)ob : (uint8_t)'?';
                    tl++;
                }
            }
            // a last line without '\n' is a row too (getline returns it): the thread that holds the last byte closes it
            const bool closes = p0 < nbytes && nbytes <= p0 + 16 && text[nbytes - 1] != '\n';
            if (closes) {
                if (base_tok + tl != (base_nl + rl + 1) * L) atomicMin(err_pos, (unsigned long long)nbytes);
                stage[e++] = '\n';  // e == all events of the step here: nothing follows the last byte
            }
            __syncthreads();
            // copy-out: stage[0, nev) -> out[o0, o0 + nev), as 16-byte vectors aligned in `out`
            const int64_t o0 = base_tok + base_nl;
            const uint32_t tot_ev = (total >> 16) + (total & 0xFFFFu);
            const int64_t last_pos = ch * TK_CHUNK + (int64_t)s * TK_STEP + TK_STEP;  // the closing thread is in this step?
            const uint32_t nev = tot_ev + ((nbytes <= last_pos && text[nbytes - 1] != '\n') ? 1u : 0u);
            const uint32_t head = (uint32_t)((16 - (o0 & 15)) & 15);  // bytes before the first aligned vector
            for (uint32_t j = threadIdx.x; j < head && j < nev; j += TK_THREADS)
                if (o0 + j < out_bytes) out[o0 + j] = stage[j];
            if (nev > head) {
                const uint32_t nvec = (nev - head) >> 4, tail0 = head + (nvec << 4);
                const uint32_t sh = (head & 3u) * 8u;
                const uint32_t* st32 = reinterpret_cast<const uint32_t*>(stage);
                for (uint32_t j = threadIdx.x; j < nvec; j += TK_THREADS) {
                    const uint32_t so = head + (j << 4);  // stage offset of this vector: so & 3 == head & 3
                    const uint32_t* q = st32 + (so >> 2);
                    const uint32_t a0 = q[0], a1 = q[1], a2 = q[2], a3 = q[3], a4 = q[4];
                    const uint4 x = make_uint4(__funnelshift_r(a0, a1, sh), __funnelshift_r(a1, a2, sh), __funnelshift_r(a2, a3, sh),
                                               __funnelshift_r(a3, a4, sh));
                    const int64_t oo = o0 + so;
                    if (oo + 16 <= out_bytes) *reinterpret_cast<uint4*>(out + oo) = x;
                    else
                        for (int b = 0; b < 16; b++)
                            if (oo + b < out_bytes) out[oo + b] = stage[so + b];
                }
                for (uint32_t j = tail0 + threadIdx.x; j < nev; j += TK_THREADS)
                    if (o0 + j < out_bytes) out[o0 + j] = stage[j];
            }
            base_tok += (total >> 16);
            base_nl += (total & 0xFFFFu);
        }
    }
}

// ------------------------------------------------------------------ PLINK ped files (src/CreateASCIInospace_PLINK.cpp:16-248)
// A ped line is 6 leading fields and two allele characters per SNP.  The reference checks the token count of the line
// against dims[1] (:64-76), skips six tokens and then reads dims[1] - 6 single non-blank CHARACTERS (:88-93); with
// one-character allele tokens -- the only kind it reads correctly -- character k of the line is token 6 + k.  Stage 1
// (ped_emit_kernel, the tokeniser's scans again) therefore gathers token 6 + k of every row into a dense rows x 2*nsnp
// byte matrix, offset (tokens before) - 6 * (newlines before + 1), and reports the first line with a wrong token count
// or a multi-character allele token.  Stage 2 (ped_genotype_kernel) is the reference's per-SNP state machine (:99-196):
// the two alleles seen so far are per-column state carried down the rows, so one thread owns one SNP and walks the rows
// (reads and writes coalesced across the threads of a warp); columns are independent, and the sequential loop's first
// error is the smallest (row, snp) key.
__global__ void __launch_bounds__(TK_THREADS) ped_emit_kernel(const uint8_t* __restrict__ text, int64_t nbytes, int64_t nchunks,
                                                              const int64_t* __restrict__ prefix, int64_t ncols, uint8_t* __restrict__ out,
                                                              int64_t out_rows, unsigned long long* __restrict__ err_pos) {
    __shared__ uint32_t wsum[TK_THREADS / 32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t width = ncols - 6;
    for (int64_t ch = blockIdx.x; ch < nchunks; ch += gridDim.x) {
        int64_t base_tok = prefix[2 * ch], base_nl = prefix[2 * ch + 1];
        for (int s = 0; s < TK_STEPS; s++) {
            const int64_t p0 = ch * TK_CHUNK + (int64_t)s * TK_STEP + threadIdx.x * 16;
            if (ch * TK_CHUNK + (int64_t)s * TK_STEP >= nbytes) break;  // CTA-uniform
            uint4 v = make_uint4(0x20202020u, 0x20202020u, 0x20202020u, 0x20202020u);
            uint32_t ws = 0xFFFFu, nl = 0, ts = 0;
            if (p0 < nbytes) {
                v = *reinterpret_cast<const uint4*>(text + p0);
                byte_masks(v, p0, nbytes, ws, nl);
                ts = token_starts(ws, text, p0);
            }
            const uint32_t mine = ((uint32_t)__popc(ts) << 16) | (uint32_t)__popc(nl);
            uint32_t inc = mine;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const uint32_t u = __shfl_up_sync(0xffffffffu, inc, o);
                if (lane >= o) inc += u;
            }
            __syncthreads();
            if (lane == 31) wsum[warp] = inc;
            __syncthreads();
            uint32_t before = inc - mine, total = 0;
#pragma unroll
            for (int w = 0; w < TK_THREADS / 32; w++) {
                const uint32_t x = wsum[w];
                if (w < warp) before += x;
                total += x;
            }
            int64_t t = base_tok + (before >> 16), r = base_nl + (before & 0xFFFFu);
            const bool single15 = (ts & 0x8000u) && (p0 + 16 >= nbytes || is_ws(text[p0 + 16]));
            const uint64_t lo = (uint64_t)v.x | ((uint64_t)v.y << 32), hi = (uint64_t)v.z | ((uint64_t)v.w << 32);
            uint32_t ev = ts | nl;
            while (ev) {
                const int i = __ffs(ev) - 1;
                ev &= ev - 1;
                if ((nl >> i) & 1u) {
                    if (t != (r + 1) * ncols) atomicMin(err_pos, (unsigned long long)(p0 + i));
                    r++;
                } else {
                    const int64_t col = t - r * ncols;
                    if (col >= 6 && col < ncols && r < out_rows) {
                        const bool single = (i < 15) ? ((ws >> (i + 1)) & 1u) : single15;
                        if (!single) atomicMin(err_pos, (unsigned long long)(p0 + i));
                        out[r * width + (col - 6)] = (uint8_t)(((i < 8 ? lo : hi) >> (8 * (i & 7))) & 0xFFu);
                    }
                    t++;
                }
            }
            if (p0 < nbytes && nbytes <= p0 + 16 && text[nbytes - 1] != '\n') {
                if (t != (r + 1) * ncols) atomicMin(err_pos, (unsigned long long)nbytes);
            }
            base_tok += (total >> 16);
            base_nl += (total & 0xFFFFu);
        }
    }
}

// alleles: rows x 2*nsnp characters.  state: 2*nsnp bytes (alleles0, alleles1 per SNP), carried between pieces of a file.
// keys[0]: smallest (row_base + row) * nsnp + snp with a third allele; keys[1]: the same for the first missing allele.
__global__ void __launch_bounds__(256) ped_genotype_kernel(const uint8_t* __restrict__ alleles, int64_t rows, int64_t nsnp,
                                                           int first_piece, int64_t row_base, uint8_t* __restrict__ state,
                                                           uint8_t* __restrict__ out, unsigned long long* __restrict__ keys) {
    for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < nsnp; i += (int64_t)gridDim.x * 256) {
        uint8_t a0 = 'I', a1 = 'I';
        if (!first_piece) { a0 = state[2 * i]; a1 = state[2 * i + 1]; }
        bool warned = false;
        for (int64_t r = 0; r < rows; r++) {
            const uchar2 ab = *reinterpret_cast<const uchar2*>(alleles + (r * nsnp + i) * 2);
            uint8_t a = ab.x, b = ab.y;
            const bool miss = a == '0' || b == '0' || a == '-' || b == '-';
            if (first_piece && r == 0) {  // :99-111
                a0 = miss ? (uint8_t)'I' : a;
                a1 = miss ? (uint8_t)'I' : b;
            }
            if (miss) {  // :118-133
                if (!warned) atomicMin(&keys[1], (unsigned long long)((row_base + r) * nsnp + i));
                warned = true;
                a = b = 'I';
            }
            bool bad = false;
#pragma unroll
            for (int j = 1; j >= 0; --j) {  // :137-172, second allele first
                const uint8_t x = j ? b : a;
                if (x != a0 && x != a1 && x != 'I') {
                    if (a0 == 'I') a0 = x;
                    else if (a1 == 'I') a1 = x;
                    else if (a0 == a1) a1 = x;
                    else bad = true;
                }
                if (bad) break;
            }
            if (bad) {
                atomicMin(&keys[0], (unsigned long long)((row_base + r) * nsnp + i));
                break;
            }
            uint8_t g = '1';  // :177-190
            if (a != 'I' && b != 'I' && a == b) g = (a == a0) ? '0' : '2';
            out[r * (nsnp + 1) + i] = g;
            if (i == 0) out[r * (nsnp + 1) + nsnp] = '\n';
        }
        state[2 * i] = a0;
        state[2 * i + 1] = a1;
    }
}

// ------------------------------------------------------------------ store -> no-space ASCII rows
// rows [row0, row0 + nrows) of a row-major store; out: nrows * (cols + 1) bytes, 16-byte aligned base.
__global__ void __launch_bounds__(256) encode_ascii_kernel(const int8_t* __restrict__ store, int64_t pitch, int64_t cols, int64_t row0,
                                                           int64_t nrows, uint8_t* __restrict__ out) {
    const int64_t line = cols + 1, total = nrows * line, nvec = (total + 15) / 16;
    for (int64_t vi = (int64_t)blockIdx.x * 256 + threadIdx.x; vi < nvec; vi += (int64_t)gridDim.x * 256) {
        const int64_t o = vi * 16;
        const int64_t r = o / line, c = o - r * line;
        if (c + 16 <= cols && o + 16 <= total) {
            const int8_t* src = store + (row0 + r) * pitch + c;
            const uintptr_t a = (uintptr_t)src;
            const uint32_t mis = (uint32_t)(a & 15u), sh = (mis & 3u) * 8u;
            const uint4 q0 = *reinterpret_cast<const uint4*>(a - mis);
            // the store's pitch leaves at least one pad byte after the row, the allocation 256 more: readable
            const uint4 q1 = mis ? *reinterpret_cast<const uint4*>(a - mis + 16) : q0;
            const uint32_t w8[8] = {q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, q1.z, q1.w};
            uint32_t x[4];
#pragma unroll
            for (int k = 0; k < 4; k++) {
                uint32_t lo = w8[k], hi = w8[k + 1];
                if ((mis >> 2) == 1) { lo = w8[k + 1]; hi = w8[k + 2]; }
                if ((mis >> 2) == 2) { lo = w8[k + 2]; hi = w8[k + 3]; }
                if ((mis >> 2) == 3) { lo = w8[k + 3]; hi = w8[(k + 4) & 7]; }
                const uint32_t g = __funnelshift_r(lo, hi, sh);  // four store bytes 1 - code in {0x01, 0x00, 0xFF}
                x[k] = (0xB1B1B1B1u - (g & 0x7F7F7F7Fu)) ^ (~g & 0x80808080u);  // bytewise '1' - g = '0' + code, borrow-free
            }
            *reinterpret_cast<uint4*>(out + o) = make_uint4(x[0], x[1], x[2], x[3]);
        } else {
            for (int j = 0; j < 16 && o + j < total; j++) {
                const int64_t oo = o + j, rr = oo / line, cc = oo - rr * line;
                out[oo] = (cc == cols) ? (uint8_t)'\n' : (uint8_t)('1' - store[(row0 + rr) * pitch + cc]);
            }
        }
    }
}

static int fill_codes(TokCodes& c, const char* AA, const char* AB, const char* BB, const char* missing) {
    const char* s[4] = {BB, AB, AA, missing};
    const uint8_t o[4] = {'2', '1', '0', '1'};
    memset(&c, 0, sizeof(c));
    for (int k = 0; k < 4; k++) {
        const size_t len = s[k] ? strlen(s[k]) : 0;
        if (len > TK_MAXLEN) return set_error(EG_ERR_ARG, "genotype code \"%s\" is longer than %d characters", s[k], TK_MAXLEN);
        for (size_t j = 0; j < len; j++)
            if (s[k][j] == ' ' || (s[k][j] >= 9 && s[k][j] <= 13)) return set_error(EG_ERR_ARG, "genotype code \"%s\" contains whitespace", s[k]);
        memcpy(c.s[k], s[k] ? s[k] : "", len);
        c.len[k] = (int32_t)len;
        c.out[k] = o[k];
    }
    return EG_OK;
}

}  // namespace eg

using namespace eg;

extern "C" int64_t eg_tokenise_chunks(int64_t nbytes) { return nbytes <= 0 ? 0 : (nbytes + TK_CHUNK - 1) / TK_CHUNK; }
extern "C" int64_t eg_tokenise_chunk_bytes(void) { return TK_CHUNK; }

// d_text: nbytes of text, 16-byte aligned, readable for 32 bytes beyond nbytes.  d_counts: 2 * chunks uint32 scratch.
// d_prefix: 2 * (chunks + 1) int64; on return d_prefix[2*chunks] = tokens, d_prefix[2*chunks + 1] = '\n' bytes in the text.
extern "C" int eg_dev_tokenise_scan(const uint8_t* d_text, int64_t nbytes, uint32_t* d_counts, int64_t* d_prefix, void* stream) {
    if (!d_text || !d_counts || !d_prefix || nbytes <= 0 || ((uintptr_t)d_text & 15))
        return set_error(EG_ERR_ARG, "eg_dev_tokenise_scan: bad argument");
    const int64_t nch = eg_tokenise_chunks(nbytes), cap = (int64_t)num_sms() * 8;
    tok_count_kernel<<<(unsigned)(nch < cap ? nch : cap), TK_THREADS, 0, (cudaStream_t)stream>>>(d_text, nbytes, nch, d_counts);
    EG_TRY(check_launch("tok_count_kernel"));
    tok_scan_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(d_counts, nch, d_prefix);
    return check_launch("tok_scan_kernel");
}
// d_out: out_rows * (cols + 1) bytes.  d_err_pos: one uint64 set to UINT64_MAX by the caller; afterwards the byte position
// of the first error event (start of a token that is none of the codes; the '\n' -- or nbytes for an unterminated last
// line -- that ends a row with a token count other than cols), or still UINT64_MAX.
extern "C" int eg_dev_tokenise_emit(const uint8_t* d_text, int64_t nbytes, const int64_t* d_prefix, int64_t cols, const char* AA,
                                    const char* AB, const char* BB, const char* missing, uint8_t* d_out, int64_t out_rows,
                                    uint64_t* d_err_pos, void* stream) {
    if (!d_text || !d_prefix || !d_out || !d_err_pos || nbytes <= 0 || cols <= 0 || out_rows < 0 || ((uintptr_t)d_text & 15))
        return set_error(EG_ERR_ARG, "eg_dev_tokenise_emit: bad argument");
    TokCodes codes;
    EG_TRY(fill_codes(codes, AA, AB, BB, missing));
    const int64_t nch = eg_tokenise_chunks(nbytes), cap = (int64_t)num_sms() * 8;
    tok_emit_kernel<<<(unsigned)(nch < cap ? nch : cap), TK_THREADS, 0, (cudaStream_t)stream>>>(
        d_text, nbytes, nch, d_prefix, cols, codes, d_out, out_rows * (cols + 1), reinterpret_cast<unsigned long long*>(d_err_pos));
    return check_launch("tok_emit_kernel");
}

// rows [row0, row0 + nrows) of a row-major int8 store as no-space ASCII: nrows * (cols + 1) bytes at d_out (16-byte aligned)
extern "C" int eg_dev_encode_ascii(const int8_t* d_store, int64_t pitch, int64_t cols, int64_t row0, int64_t nrows, uint8_t* d_out,
                                   void* stream) {
    if (!d_store || !d_out || pitch <= cols || (pitch & 15) || cols <= 0 || row0 < 0 || nrows < 0 || ((uintptr_t)d_out & 15))
        return set_error(EG_ERR_ARG, "eg_dev_encode_ascii: bad argument");
    if (nrows == 0) return EG_OK;
    const int64_t nvec = (nrows * (cols + 1) + 15) / 16, nb = (nvec + 255) / 256, cap = (int64_t)num_sms() * 16;
    encode_ascii_kernel<<<(unsigned)(nb < cap ? nb : cap), 256, 0, (cudaStream_t)stream>>>(d_store, pitch, cols, row0, nrows, d_out);
    return check_launch("encode_ascii_kernel");
}

// PLINK ped, stage 1: token 6 + k of every line -> d_alleles[row * (ncols - 6) + k] (ncols = 6 + 2 * nsnp tokens per line).
// d_prefix from eg_dev_tokenise_scan.  *d_err_pos as in eg_dev_tokenise_emit: the '\n' closing a line whose token count is
// not ncols, or the start of an allele token longer than one character.
extern "C" int eg_dev_ped_alleles(const uint8_t* d_text, int64_t nbytes, const int64_t* d_prefix, int64_t ncols, uint8_t* d_alleles,
                                  int64_t out_rows, uint64_t* d_err_pos, void* stream) {
    if (!d_text || !d_prefix || !d_alleles || !d_err_pos || nbytes <= 0 || ncols <= 6 || ((ncols - 6) & 1) || out_rows < 0 ||
        ((uintptr_t)d_text & 15))
        return set_error(EG_ERR_ARG, "eg_dev_ped_alleles: bad argument");
    const int64_t nch = eg_tokenise_chunks(nbytes), cap = (int64_t)num_sms() * 8;
    ped_emit_kernel<<<(unsigned)(nch < cap ? nch : cap), TK_THREADS, 0, (cudaStream_t)stream>>>(
        d_text, nbytes, nch, d_prefix, ncols, d_alleles, out_rows, reinterpret_cast<unsigned long long*>(d_err_pos));
    return check_launch("ped_emit_kernel");
}
// PLINK ped, stage 2: rows x 2*nsnp allele characters -> rows x (nsnp + 1) no-space ASCII.  d_state: 2*nsnp bytes, written
// when first_piece != 0 and carried to the next piece otherwise.  d_keys: 2 x uint64, set to UINT64_MAX by the caller:
// [0] smallest (row_base + row) * nsnp + snp where a third allele appears, [1] the same for the first missing allele.
extern "C" int eg_dev_ped_genotypes(const uint8_t* d_alleles, int64_t rows, int64_t nsnp, int first_piece, int64_t row_base,
                                    uint8_t* d_state, uint8_t* d_out, uint64_t* d_keys, void* stream) {
    if (!d_alleles || !d_state || !d_out || !d_keys || rows < 0 || nsnp <= 0 || ((uintptr_t)d_alleles & 1))
        return set_error(EG_ERR_ARG, "eg_dev_ped_genotypes: bad argument");
    if (rows == 0) return EG_OK;
    const int64_t nb = (nsnp + 255) / 256, cap = (int64_t)num_sms() * 8;
    ped_genotype_kernel<<<(unsigned)(nb < cap ? nb : cap), 256, 0, (cudaStream_t)stream>>>(
        d_alleles, rows, nsnp, first_piece, row_base, d_state, d_out, reinterpret_cast<unsigned long long*>(d_keys));
    return check_launch("ped_genotype_kernel");
}
