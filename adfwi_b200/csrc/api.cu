// api.cu -- error strings, ABI version and the launch counter of libadfwi_b200.
#include "common.cuh"

namespace adfwi { std::atomic<uint64_t> g_launches{0}; }

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
