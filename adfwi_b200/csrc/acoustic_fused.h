// acoustic_fused.h -- entry points of the fused TMA-staged acoustic pipeline (acoustic_fused.cu)
#pragma once
#include "common.cuh"
#ifndef ADFWI_HOST_EMUL
namespace adfwi {
size_t acf_workspace_bytes(const adfwi_acoustic_desc* d);
int acf_group_size(const adfwi_acoustic_desc* d);
// coef = alpha1, kappa1, alpha2, kappa2, kappa3 (dense caller planes)
int acf_forward(const adfwi_acoustic_desc* d, const float* const* coef, const float* src_v, const int64_t* sx, const int64_t* sz,
                const int64_t* rx, const int64_t* rz, float* rcv_p, float* rcv_u, float* rcv_w,
                float* illum_p, float* illum_u, float* illum_w, void* ws, cudaStream_t st);
int acf_backward(const adfwi_acoustic_desc* d, const float* const* coef, const float* src_v, const int64_t* sx, const int64_t* sz,
                 const int64_t* rx, const int64_t* rz, const float* gp, const float* gu, const float* gw,
                 float* g_alpha1, float* g_alpha2, float* g_src, void* ws, cudaStream_t st);
}
#endif
