// common.cuh -- shared host/device helpers of libadfwi_b200 (sm_100a only).
#pragma once
#ifdef ADFWI_HOST_EMUL
#include "host_emul.h"      // tests/emul: serial host emulation of the launches (test fixture only)
#else
#include <cuda_runtime.h>
#define ADFWI_KERNEL(...) __VA_ARGS__
#define ADFWI_LAUNCH(kern, grd, blk, strm, ...) kern<<<(grd), (blk), 0, (strm)>>>(__VA_ARGS__)
#endif
#include <stdint.h>
#include <stddef.h>
#include <atomic>
#include "../../include/adfwi_b200.h"

#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ < 1000)
#error "libadfwi_b200 is written for sm_100a (B200) only"
#endif

namespace adfwi {

extern std::atomic<uint64_t> g_launches;   // diagnostic counter behind adfwi_launch_count()

inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }
inline int cdiv(int a, int b) { return (a + b - 1) / b; }

// bump allocator over the caller's workspace (units: bytes, 256-B aligned regions)
struct Carver {
    char* base; size_t off;
    explicit Carver(void* p) : base((char*)p), off(0) {}
    template <typename T> T* take(size_t n) {
        T* r = base ? (T*)(base + off) : nullptr;
        off = align_up(off + n * sizeof(T), 256);
        return r;
    }
};

#define ADFWI_LAUNCH_CHECK()                                   \
    do {                                                       \
        ::adfwi::g_launches.fetch_add(1, std::memory_order_relaxed); \
        cudaError_t e__ = cudaPeekAtLastError();               \
        if (e__ != cudaSuccess) return (int)e__;               \
    } while (0)

#define ADFWI_CUDA(call)                                       \
    do {                                                       \
        cudaError_t e__ = (call);                              \
        if (e__ != cudaSuccess) return (int)e__;               \
    } while (0)

}  // namespace adfwi
