// Host <-> device transfers of PAGEABLE memory at close to the PCIe rate.
//
// Everything the reference hands to its C++ entry points is pageable: R owns S = inv_MMt_sqrt, V = dim_reduced_vara and
// the result matrices (Eigen::Map views of R memory, src/RcppExports.cpp:61-62), and the genotypes are files
// (M.ascii / Mt.ascii, R/create_ascii.R:15-16) that sit in the page cache.  cudaMemcpy from pageable memory is staged by
// the driver through one thread and ran at 10 - 11 GB/s on this box (1.6 GB of S and V: 140 ms; the 10 GB of an ASCII
// file: 1.6 s) against 55 GB/s from page-locked memory.  Page-locking the caller's buffers per call (cudaHostRegister)
// costs more than it saves for buffers that are used once.  Here instead: a ring of page-locked slots per GPU, a pool of
// copier threads shared by all GPUs that fill (or drain) the slots in parallel -- memcpy for memory, pread for files
// (no page-table traffic for a file that is only read once) -- and the DMA of slot k overlapped with the filling of
// slots k+1 .. k+R-1.
#include <pthread.h>
#include <unistd.h>

#include <algorithm>
#include <atomic>
#include <condition_variable>
#include <cstring>
#include <deque>
#include <functional>
#include <mutex>
#include <thread>
#include <vector>

#include "common.cuh"
#include "hostio.cuh"

namespace eg {

// ------------------------------------------------------------------ copier pool (process-wide)
namespace {
struct Latch {
    std::mutex mu;
    std::condition_variable cv;
    int pending = 0;
    bool failed = false;
    void add(int k) { std::lock_guard<std::mutex> lk(mu); pending += k; }
    void done(bool ok) {
        std::lock_guard<std::mutex> lk(mu);
        if (!ok) failed = true;
        if (--pending == 0) cv.notify_all();
    }
    bool wait() {
        std::unique_lock<std::mutex> lk(mu);
        cv.wait(lk, [&] { return pending == 0; });
        const bool ok = !failed;
        failed = false;
        return ok;
    }
};
struct Job {
    std::function<bool()> fn;
    Latch* latch;
};
struct CopierPool {
    std::vector<std::thread> th;
    std::mutex mu;
    std::condition_variable cv;
    std::deque<Job> q;
    bool stop = false;
    void start(int n) {
        for (int i = 0; i < n; i++)
            th.emplace_back([this] {
                for (;;) {
                    Job j;
                    {
                        std::unique_lock<std::mutex> lk(mu);
                        cv.wait(lk, [&] { return stop || !q.empty(); });
                        if (stop && q.empty()) return;
                        j = std::move(q.front());
                        q.pop_front();
                    }
                    j.latch->done(j.fn());
                }
            });
    }
    void submit(std::function<bool()> fn, Latch* l) {
        l->add(1);
        {
            std::lock_guard<std::mutex> lk(mu);
            q.push_back(Job{std::move(fn), l});
        }
        cv.notify_one();
    }
    ~CopierPool() {
        {
            std::lock_guard<std::mutex> lk(mu);
            stop = true;
        }
        cv.notify_all();
        for (auto& t : th) t.join();
    }
};
CopierPool* pool() {
    static CopierPool* P = [] {
        auto* p = new CopierPool();   // lives until process exit (never destroyed: threads may be mid-copy at exit otherwise)
        const char* e = getenv("EAGLE_HOST_THREADS");
        int n = e ? atoi(e) : (int)std::min<unsigned>(16u, std::max(2u, std::thread::hardware_concurrency()));
        cpu_set_t set;   // respect the affinity mask of a container
        if (!e && pthread_getaffinity_np(pthread_self(), sizeof(set), &set) == 0) n = std::min(n, std::max(2, CPU_COUNT(&set)));
        p->start(std::max(1, n));
        return p;
    }();
    return P;
}

// ------------------------------------------------------------------ page-locked ring (per GPU thread)
constexpr int RING = 8;
struct Ring {
    uint8_t* slot[RING] = {};
    cudaEvent_t ev[RING] = {};
    bool ev_used[RING] = {};
    Latch latch[RING];
    size_t bytes = 0;
    int device = -1;
};
thread_local Ring t_ring;

size_t slot_bytes() {
    static size_t b = [] {
        const char* e = getenv("EAGLE_STAGE_SLOT_MB");
        const long mb = e ? atol(e) : 8;
        return (size_t)std::max(1L, mb) << 20;
    }();
    return b;
}
size_t sub_bytes() { return (size_t)1 << 20; }   // one copier job

int ring_ready() {
    int dev = 0;
    EG_CUDA(cudaGetDevice(&dev));
    if (t_ring.slot[0] && t_ring.device == dev) return EG_OK;
    hostio_release();
    t_ring.bytes = slot_bytes();
    for (int i = 0; i < RING; i++) {
        if (cudaHostAlloc((void**)&t_ring.slot[i], t_ring.bytes, cudaHostAllocDefault) != cudaSuccess) {
            cudaGetLastError();
            hostio_release();
            return set_error(EG_ERR_ALLOC, "out of page-locked host memory for the staging ring");
        }
        EG_CUDA(cudaEventCreateWithFlags(&t_ring.ev[i], cudaEventDisableTiming));
        t_ring.ev_used[i] = false;
    }
    t_ring.device = dev;
    return EG_OK;
}
int slot_free_for_host_write(int s) {   // the DMA that last read this slot has finished
    if (t_ring.ev_used[s]) EG_CUDA(cudaEventSynchronize(t_ring.ev[s]));
    return EG_OK;
}

bool read_range(const HostSrc& src, size_t off, uint8_t* dst, size_t len) {
    if (src.fd >= 0) {
        while (len) {
            const ssize_t got = pread(src.fd, dst, len, (off_t)off);
            if (got <= 0) return false;
            dst += got;
            off += (size_t)got;
            len -= (size_t)got;
        }
        return true;
    }
    memcpy(dst, src.p + off, len);
    return true;
}
}  // namespace

void hostio_release() {
    for (int i = 0; i < RING; i++) {
        if (t_ring.slot[i]) cudaFreeHost(t_ring.slot[i]);
        if (t_ring.ev[i]) cudaEventDestroy(t_ring.ev[i]);
        t_ring.slot[i] = nullptr;
        t_ring.ev[i] = nullptr;
        t_ring.ev_used[i] = false;
    }
    t_ring.device = -1;
}

bool host_is_pinned(const void* p) {
    cudaPointerAttributes at;
    const bool pinned = cudaPointerGetAttributes(&at, p) == cudaSuccess && at.type == cudaMemoryTypeHost;
    cudaGetLastError();
    return pinned;
}

// rows x width bytes from src (row r at src_off + r * src_pitch) -> device (row r at d_dst + r * d_pitch).  Returns when
// the source has been read completely and every DMA is enqueued on `st` (the data is on the device in stream order).
int h2d_staged_2d(void* d_dst, size_t d_pitch, const HostSrc& src, size_t src_off, size_t src_pitch, size_t width, size_t rows,
                  cudaStream_t st) {
    if (!rows || !width) return EG_OK;
    if (src.fd < 0 && host_is_pinned(src.p + src_off)) {   // page-locked already: plain DMA
        if (rows == 1 || (src_pitch == width && d_pitch == width))
            return check_cuda(cudaMemcpyAsync(d_dst, src.p + src_off, rows * width, cudaMemcpyHostToDevice, st), "H2D");
        return check_cuda(cudaMemcpy2DAsync(d_dst, d_pitch, src.p + src_off, src_pitch, width, rows, cudaMemcpyHostToDevice, st), "H2D");
    }
    EG_TRY(ring_ready());
    const size_t SB = t_ring.bytes;
    // a chunk is a slot-full: whole rows when rows fit a slot, else pieces of one row
    const bool by_rows = width <= SB;
    const size_t rows_per_chunk = by_rows ? std::max<size_t>(1, SB / width) : 1;
    const size_t pieces_per_row = by_rows ? 1 : (width + SB - 1) / SB;
    const size_t nchunks = by_rows ? (rows + rows_per_chunk - 1) / rows_per_chunk : rows * pieces_per_row;
    auto chunk_geom = [&](size_t k, size_t* r0, size_t* nr, size_t* c0, size_t* w) {
        if (by_rows) {
            *r0 = k * rows_per_chunk;
            *nr = std::min(rows_per_chunk, rows - *r0);
            *c0 = 0;
            *w = width;
        } else {
            *r0 = k / pieces_per_row;
            *nr = 1;
            *c0 = (k % pieces_per_row) * SB;
            *w = std::min(SB, width - *c0);
        }
    };
    size_t next_submit = 0, next_issue = 0;
    bool ok = true;
    while (next_issue < nchunks) {
        while (next_submit < nchunks && next_submit - next_issue < (size_t)RING) {
            const int s = (int)(next_submit % RING);
            EG_TRY(slot_free_for_host_write(s));
            size_t r0, nr, c0, w;
            chunk_geom(next_submit, &r0, &nr, &c0, &w);
            uint8_t* slot = t_ring.slot[s];
            if (nr == 1) {   // pieces of a long row: split the piece over several copier jobs
                for (size_t o = 0; o < w; o += sub_bytes()) {
                    const size_t len = std::min(sub_bytes(), w - o);
                    const size_t so = src_off + r0 * src_pitch + c0 + o;
                    pool()->submit([src, so, slot, o, len] { return read_range(src, so, slot + o, len); }, &t_ring.latch[s]);
                }
            } else {         // a block of rows: jobs of whole rows
                const size_t rows_per_job = std::max<size_t>(1, sub_bytes() / w);
                for (size_t j0 = 0; j0 < nr; j0 += rows_per_job) {
                    const size_t j1 = std::min(nr, j0 + rows_per_job);
                    pool()->submit([src, src_off, src_pitch, r0, j0, j1, w, slot] {
                        for (size_t j = j0; j < j1; j++)
                            if (!read_range(src, src_off + (r0 + j) * src_pitch, slot + j * w, w)) return false;
                        return true;
                    }, &t_ring.latch[s]);
                }
            }
            next_submit++;
        }
        const int s = (int)(next_issue % RING);
        ok = t_ring.latch[s].wait() && ok;
        size_t r0, nr, c0, w;
        chunk_geom(next_issue, &r0, &nr, &c0, &w);
        uint8_t* dst = (uint8_t*)d_dst + r0 * d_pitch + c0;
        cudaError_t e = (nr == 1 || d_pitch == w) ? cudaMemcpyAsync(dst, t_ring.slot[s], nr * w, cudaMemcpyHostToDevice, st)
                                                  : cudaMemcpy2DAsync(dst, d_pitch, t_ring.slot[s], w, w, nr, cudaMemcpyHostToDevice, st);
        if (e != cudaSuccess) {
            for (size_t k = next_issue + 1; k < next_submit; k++) t_ring.latch[k % RING].wait();   // jobs hold references
            return check_cuda(e, "H2D from the staging ring");
        }
        cudaEventRecord(t_ring.ev[s], st);
        t_ring.ev_used[s] = true;
        next_issue++;
    }
    if (!ok) return set_error(EG_ERR_OPEN, "short read while staging a file to the device");
    return EG_OK;
}

int h2d_staged(void* d_dst, const void* h_src, size_t bytes, cudaStream_t st) {
    HostSrc s;
    s.p = (const uint8_t*)h_src;
    return h2d_staged_2d(d_dst, bytes, s, 0, bytes, bytes, 1, st);
}

// device -> pageable host memory; returns when h_dst is complete.  (`st` must already hold the work that produces d_src.)
int d2h_staged(void* h_dst, const void* d_src, size_t bytes, cudaStream_t st) {
    if (!bytes) return EG_OK;
    if (host_is_pinned(h_dst)) {
        EG_CUDA(cudaMemcpyAsync(h_dst, d_src, bytes, cudaMemcpyDeviceToHost, st));
        return check_cuda(cudaStreamSynchronize(st), "D2H");
    }
    EG_TRY(ring_ready());
    const size_t SB = t_ring.bytes, nchunks = (bytes + SB - 1) / SB;
    size_t next_dma = 0, next_copy = 0;
    while (next_copy < nchunks) {
        while (next_dma < nchunks && next_dma - next_copy < (size_t)RING) {
            const int s = (int)(next_dma % RING);
            t_ring.latch[s].wait();                       // the previous occupant has been copied out
            EG_TRY(slot_free_for_host_write(s));          // ... and an earlier H2D out of this slot has finished
            const size_t off = next_dma * SB, len = std::min(SB, bytes - off);
            EG_CUDA(cudaMemcpyAsync(t_ring.slot[s], (const uint8_t*)d_src + off, len, cudaMemcpyDeviceToHost, st));
            cudaEventRecord(t_ring.ev[s], st);
            t_ring.ev_used[s] = true;
            next_dma++;
        }
        const int s = (int)(next_copy % RING);
        EG_CUDA(cudaEventSynchronize(t_ring.ev[s]));
        const size_t off = next_copy * SB, len = std::min(SB, bytes - off);
        const uint8_t* slot = t_ring.slot[s];
        for (size_t o = 0; o < len; o += sub_bytes()) {
            const size_t l = std::min(sub_bytes(), len - o);
            uint8_t* dst = (uint8_t*)h_dst + off + o;
            pool()->submit([dst, slot, o, l] { memcpy(dst, slot + o, l); return true; }, &t_ring.latch[s]);
        }
        next_copy++;
    }
    for (int s = 0; s < RING; s++) t_ring.latch[s].wait();
    return EG_OK;
}

}  // namespace eg
