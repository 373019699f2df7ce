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

constexpr int kMaxDevices = 64;            // per-device caches (SM count, kernel attributes) are indexed by device ordinal

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

// ---- optional sampled per-kernel-class device timing (diagnostics for bench.py's roofline) ----
// When enabled through adfwi_timing_enable(every_n), every n-th launch of each class is bracketed
// by CUDA events recorded on the launching stream; adfwi_timing_collect() reads them back.
enum KernelClass {
    KC_AC_FWD_P = 0, KC_AC_FWD_UW, KC_AC_RECORD, KC_AC_ADJ_INJECT, KC_AC_ADJ_A, KC_AC_ADJ_B,
    KC_EL_FWD_STRESS, KC_EL_FWD_VEL, KC_EL_RECORD, KC_EL_ADJ_INJECT, KC_EL_ADJ_VEL, KC_EL_ADJ_STRESS,
    KC_AC_FWD_FUSED, KC_AC_ADJ_FUSED, KC_OTHER, KC_EL_FWD_FUSED, KC_EL_ADJ_FUSED, KC_AC_FWD_PERSIST, KC_AC_ADJ_PERSIST, KC_COUNT
};
#ifdef ADFWI_HOST_EMUL
struct TimedLaunch { TimedLaunch(int, cudaStream_t) {} };
#else
struct TimedLaunch {      // RAII: start event in the constructor, stop event in the destructor
    int slot; cudaStream_t st;
    TimedLaunch(int cls, cudaStream_t s);
    ~TimedLaunch();
};
#endif

// NVTX range around every compute entry point of the C ABI (nvtx3 is header-only: a no-op function-pointer test unless a tool such
// as ncu --nvtx / nsys injected itself); `ncu --nvtx --nvtx-include "adfwi_acoustic_backward/"` then filters by entry point.
#ifdef ADFWI_HOST_EMUL
struct NvtxRange { explicit NvtxRange(const char*) {} };
#else
struct NvtxRange { explicit NvtxRange(const char* name); ~NvtxRange(); };
#endif
#define ADFWI_NVTX(name) ::adfwi::NvtxRange nvtx_range__(name)

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
