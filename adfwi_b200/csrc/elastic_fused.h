// elastic_fused.h -- entry points of the TMA-staged split-PML elastic pipeline (elastic_fused.cu)
#pragma once
#include "common.cuh"
#ifndef ADFWI_HOST_EMUL
namespace adfwi {
bool elf_supported(const adfwi_elastic_desc* d);
size_t elf_workspace_bytes(const adfwi_elastic_desc* d);
// coef = C11, C13, C33, C55, bx, bz (dense caller planes [nzp][nxp]); bcx, bcz dense PML profiles
int elf_forward(const adfwi_elastic_desc* d, const float* const* coef, const float* bcx, const float* bcz, const float* mt,
                const float* src_v, const int64_t* sx, const int64_t* sz, const int64_t* rx, const int64_t* rz,
                float* const* rcv, float* const* illum, void* ws, cudaStream_t st);
int elf_backward(const adfwi_elastic_desc* d, const float* const* coef, const float* bcx, const float* bcz, const float* mt,
                 const float* src_v, const int64_t* sx, const int64_t* sz, const int64_t* rx, const int64_t* rz,
                 const float* const* g_rcv, float* const* g_coef, float* g_src, void* ws, cudaStream_t st);
// sponge (ABL) boundary: ela_f / ela_b (elastic_abl_fused.inl); damp = dense sponge plane [nzp][nxp]
bool ela_supported(const adfwi_elastic_desc* d);
size_t ela_workspace_bytes(const adfwi_elastic_desc* d);
int ela_forward(const adfwi_elastic_desc* d, const float* const* coef, const float* damp, const float* mt,
                const float* src_v, const int64_t* sx, const int64_t* sz, const int64_t* rx, const int64_t* rz,
                float* const* rcv, float* const* illum, void* ws, cudaStream_t st);
int ela_backward(const adfwi_elastic_desc* d, const float* mt, const float* src_v, const int64_t* sx, const int64_t* sz,
                 const float* const* g_rcv, float* const* g_coef, float* g_src, void* ws, cudaStream_t st);
}
#endif
