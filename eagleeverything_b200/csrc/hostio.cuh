// Staged transfers between pageable host memory / files and the device (csrc/hostio.cu).
#pragma once
#include <cstddef>
#include <cstdint>
#include <cuda_runtime.h>

namespace eg {

// Where host bytes come from: memory (p) or, when fd >= 0, a file read with pread (p may then hold a read-only mapping of
// the same file for the callers that want to look at a few bytes).
struct HostSrc {
    const uint8_t* p = nullptr;
    int fd = -1;
};

bool host_is_pinned(const void* p);
int h2d_staged(void* d_dst, const void* h_src, size_t bytes, cudaStream_t st);
int h2d_staged_2d(void* d_dst, size_t d_pitch, const HostSrc& src, size_t src_off, size_t src_pitch, size_t width, size_t rows,
                  cudaStream_t st);
int d2h_staged(void* h_dst, const void* d_src, size_t bytes, cudaStream_t st);
void hostio_release();   // frees the calling thread's page-locked ring

}  // namespace eg
