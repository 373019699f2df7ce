// api.cu -- error strings, ABI version and the launch counter of libadfwi_b200.
#include "common.cuh"

#include <mutex>
#include <vector>
#ifndef ADFWI_HOST_EMUL
#include <nvtx3/nvToolsExt.h>
#endif

namespace adfwi {
std::atomic<uint64_t> g_launches{0};

#ifndef ADFWI_HOST_EMUL
// sampled event timing -- process-wide diagnostic state, off by default
static std::mutex g_tm;
static int g_every = 0;
static uint64_t g_seen[KC_COUNT];
struct Sample { int cls; cudaEvent_t a, b; };
static std::vector<Sample> g_samples;
static const size_t kMaxSamples = 8192;

TimedLaunch::TimedLaunch(int cls, cudaStream_t s) : slot(-1), st(s)
{
    if (g_every <= 0) return;
    std::lock_guard<std::mutex> lk(g_tm);
    // the cluster-persistent kernels run a whole sweep per launch: sample every one of them
    const uint64_t every = (cls == KC_AC_FWD_PERSIST || cls == KC_AC_ADJ_PERSIST) ? 1 : (uint64_t)g_every;
    if ((g_seen[cls]++ % every) != 0 || g_samples.size() >= kMaxSamples) return;
    Sample sm; sm.cls = cls;
    if (cudaEventCreate(&sm.a) != cudaSuccess || cudaEventCreate(&sm.b) != cudaSuccess) return;
    cudaEventRecord(sm.a, st);
    slot = (int)g_samples.size();
    g_samples.push_back(sm);
}
TimedLaunch::~TimedLaunch()
{
    if (slot < 0) return;
    std::lock_guard<std::mutex> lk(g_tm);
    cudaEventRecord(g_samples[slot].b, st);
}

NvtxRange::NvtxRange(const char* name) { nvtxRangePushA(name); }
NvtxRange::~NvtxRange() { nvtxRangePop(); }
#endif
}  // namespace adfwi

#ifndef ADFWI_HOST_EMUL
extern "C" void adfwi_timing_enable(int every_n)
{
    std::lock_guard<std::mutex> lk(adfwi::g_tm);
    adfwi::g_every = every_n;
    for (int i = 0; i < adfwi::KC_COUNT; ++i) adfwi::g_seen[i] = 0;
}
extern "C" int adfwi_timing_collect(float* ms_sum, int* count, int n)
{
    std::lock_guard<std::mutex> lk(adfwi::g_tm);
    for (int i = 0; i < n; ++i) { ms_sum[i] = 0.f; count[i] = 0; }
    for (auto& s : adfwi::g_samples) {
        float ms = 0.f;
        if (cudaEventSynchronize(s.b) == cudaSuccess && cudaEventElapsedTime(&ms, s.a, s.b) == cudaSuccess && s.cls < n) {
            ms_sum[s.cls] += ms; count[s.cls] += 1;
        }
        cudaEventDestroy(s.a); cudaEventDestroy(s.b);
    }
    adfwi::g_samples.clear();
    return adfwi::KC_COUNT;
}
#else
extern "C" void adfwi_timing_enable(int) {}
extern "C" int adfwi_timing_collect(float*, int*, int) { return adfwi::KC_COUNT; }
#endif

extern "C" const char* adfwi_strerror(int code)
{
    switch (code) {
        case ADFWI_OK: return "ok";
        case ADFWI_E_NULL: return "adfwi: required pointer is NULL";
        case ADFWI_E_DIMS: return "adfwi: grid dimensions / counts out of the supported range";
        case ADFWI_E_WORKSPACE: return "adfwi: workspace smaller than adfwi_*_workspace_bytes()";
        case ADFWI_E_ORDER: return "adfwi: fd_order must be 4 or 6";
        case ADFWI_E_MODE: return "adfwi: descriptor mode does not allow this call (save_history=0?)";
        default: break;
    }
    if (code > 0) return cudaGetErrorString((cudaError_t)code);
    return "adfwi: unknown error";
}

extern "C" int adfwi_abi_version(void) { return ADFWI_ABI_VERSION; }
extern "C" uint64_t adfwi_launch_count(void) { return adfwi::g_launches.load(); }
