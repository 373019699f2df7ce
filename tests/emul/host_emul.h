// host_emul.h -- TEST FIXTURE ONLY.  Lets g++ compile adfwi_b200/csrc/*.cu as plain C++ so the
// kernel LOGIC (gather-form adjoints, region masks, workspace plans, time-loop orchestration)
// can be checked against the oracle in the GPU-less CI container.  Every "kernel launch" becomes
// a serial loop over the grid.  The product library libadfwi_b200.so is never built from this
// header and the package adfwi_b200 never loads the emulation library (tests/emul/libadfwi_emul.so).
//
// Race probe: ADFWI_EMUL_REVERSE=1 walks blocks/threads in reverse order; a kernel whose result
// depends on the thread order (an in-place neighbour hazard) gives different answers.
#pragma once
#include <stdint.h>
#include <stddef.h>
#include <stdlib.h>
#include <string.h>
#include <algorithm>
#include <cmath>

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __restrict__ __restrict
#define __launch_bounds__(...)

struct uint3_e { unsigned x, y, z; };
struct dim3 {
    unsigned x, y, z;
    dim3(unsigned a = 1, unsigned b = 1, unsigned c = 1) : x(a), y(b), z(c) {}
};
static thread_local uint3_e threadIdx, blockIdx;
static thread_local dim3 blockDim, gridDim;

typedef int cudaError_t;
typedef void* cudaStream_t;
enum { cudaSuccess = 0 };
enum cudaMemcpyKind { cudaMemcpyDeviceToDevice = 3 };
static inline cudaError_t cudaPeekAtLastError() { return cudaSuccess; }
static inline const char* cudaGetErrorString(cudaError_t) { return "emulated cuda error"; }
static inline cudaError_t cudaMemsetAsync(void* p, int v, size_t n, cudaStream_t) { memset(p, v, n); return cudaSuccess; }
static inline cudaError_t cudaMemcpyAsync(void* d, const void* s, size_t n, cudaMemcpyKind, cudaStream_t) { memmove(d, s, n); return cudaSuccess; }

using std::min;
using std::max;
template <typename T> static inline void __stcs(T* p, T v) { *p = v; }
template <typename T> static inline T __ldcs(const T* p) { return *p; }
template <typename T> static inline T __ldg(const T* p) { return *p; }
static inline float atomicAdd(float* p, float v) { float o = *p; *p = o + v; return o; }
static inline void __syncthreads_unsupported();

static inline bool emul_reverse() { static int r = getenv("ADFWI_EMUL_REVERSE") ? atoi(getenv("ADFWI_EMUL_REVERSE")) : 0; return r != 0; }

#define ADFWI_KERNEL(...) __VA_ARGS__
#define ADFWI_LAUNCH(kern, grd, blk, strm, ...)                                              \
    do {                                                                                     \
        const dim3 g__ = (grd), b__ = (blk);                                                 \
        (void)(strm);                                                                        \
        gridDim = g__; blockDim = b__;                                                       \
        const size_t nb__ = (size_t)g__.x * g__.y * g__.z, nt__ = (size_t)b__.x * b__.y * b__.z; \
        const bool rev__ = emul_reverse();                                                   \
        for (size_t bi__ = 0; bi__ < nb__; ++bi__) {                                         \
            const size_t bb__ = rev__ ? nb__ - 1 - bi__ : bi__;                              \
            blockIdx.x = bb__ % g__.x; blockIdx.y = (bb__ / g__.x) % g__.y; blockIdx.z = bb__ / ((size_t)g__.x * g__.y); \
            for (size_t ti__ = 0; ti__ < nt__; ++ti__) {                                     \
                const size_t tt__ = rev__ ? nt__ - 1 - ti__ : ti__;                          \
                threadIdx.x = tt__ % b__.x; threadIdx.y = (tt__ / b__.x) % b__.y; threadIdx.z = tt__ / ((size_t)b__.x * b__.y); \
                kern(__VA_ARGS__);                                                           \
            }                                                                                \
        }                                                                                    \
    } while (0)
