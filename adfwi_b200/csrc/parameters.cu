// parameters.cu -- the model-side producers of the elastic coefficient planes, one kernel each way:
//
//   adfwi_elastic_moduli_*   Thomsen / velocity parameters -> the six planes the P-SV kernels consume, in their own (ragged) shapes:
//                            thomsen_to_elastic_moduli (ADFWI/model/parameters.py:71-107), the TI fill-in C55 = C44 (and the HTI
//                            swap, :156-181), b = 1/rho (:47-69) and parameter_staggered_grid (:184-213):
//                              C11, C13, C33 (nz,nx);  C55 (nz-2,nx-2) = 0.2*(C44[1:-1,1:-1] + C44[2:,1:-1] + C44[1:-1,2:] + C44[2:,1:-1] + C44[2:,2:]);
//                              bx (nz,nx-1) = 0.5*(b[:,:-1] + b[:,1:]);  bz (nz-1,nx) = 0.5*(b[:-1] + b[1:])
//                            ~25 eager elementwise / slicing ops upstream (and their autograd mirrors); same association and roundings here
//                            (-fmad=false, IEEE division and square root), so the planes are bit-identical to the eager chain.
//   adfwi_elastic_pad_*      the six replicate paddings of forward_kernel (ADFWI/propagator/elastic_kernels.py:176-216, :935-946): each
//                            plane padded FROM ITS OWN SHAPE by `pml` columns left / right, `pml` rows below and `top` rows above, then
//                            zero-extended to the full (nzp,nxp) grid (what the region slices of :303-310 see); and the
//                            transpose (sum of the replicated cells) for the backward pass.
#include "common.cuh"
#ifndef ADFWI_HOST_EMUL
#include <math.h>

namespace adfwi {
namespace {

struct MdGeom { int nz, nx, hti; };
struct MdIn { const float *vp, *vs, *rho, *eps, *delta; };
struct MdOut { float *c11, *c13, *c33, *c55, *bx, *bz; };

__device__ __forceinline__ float md_c44(const MdIn& in, size_t c) { const float vs = in.vs[c]; return (vs * vs) * in.rho[c]; }

// one thread per cell of the (nz,nx) grid; the ragged planes are written by the threads whose cell is their top-left corner
__global__ void __launch_bounds__(256)
md_forward(const MdGeom g, const MdIn in, const MdOut out)
{
    const int j = blockIdx.x * blockDim.x + threadIdx.x, i = blockIdx.y;
    if (j >= g.nx) return;
    const int nx = g.nx, nz = g.nz;
    const size_t c = (size_t)i * nx + j;
    const float vp = in.vp[c], rho = in.rho[c];
    const float C33 = (vp * vp) * rho;                                  // alpha**2*rho (:102)
    const float C44 = md_c44(in, c);                                    // beta**2*rho (:103)
    const float C11 = C33 * (1.0f + 2.0f * in.eps[c]);                  // C33*(1+2*eps) (:104)
    const float A = C33 - C44;
    const float C13 = sqrtf(((2.0f * C33) * A) * in.delta[c] + A * A) - C44;   // (:106)
    out.c11[c] = g.hti ? C33 : C11;                                     // HTI: the rotated tensor swaps C11 and C33 (:178-179)
    out.c13[c] = C13;
    out.c33[c] = g.hti ? C11 : C33;
    const float b = 1.0f / rho;
    if (j < nx - 1) out.bx[(size_t)i * (nx - 1) + j] = 0.5f * (b + 1.0f / in.rho[c + 1]);
    if (i < nz - 1) out.bz[c] = 0.5f * (b + 1.0f / in.rho[c + nx]);
    if (i < nz - 2 && j < nx - 2) {                                     // C55[i,j] of the (nz-2,nx-2) plane (:208-209)
        const float s = (((md_c44(in, c + nx + 1) + md_c44(in, c + 2 * (size_t)nx + 1)) + md_c44(in, c + nx + 2)) + md_c44(in, c + 2 * (size_t)nx + 1)) +
                        md_c44(in, c + 2 * (size_t)nx + 2);
        out.c55[(size_t)i * (nx - 2) + j] = 0.2f * s;
    }
}

struct MdGrad { const float *c11, *c13, *c33, *c55, *bx, *bz; };          // cotangents of the six planes, own shapes
struct MdGout { float *vp, *vs, *rho, *eps, *delta; };                   // nullable each

__device__ __forceinline__ float md_at(const float* p, int i, int j, int h, int w) { return (i >= 0 && i < h && j >= 0 && j < w) ? p[(size_t)i * w + j] : 0.f; }

__global__ void __launch_bounds__(256)
md_backward(const MdGeom g, const MdIn in, const MdGrad gr, const MdGout go)
{
    const int j = blockIdx.x * blockDim.x + threadIdx.x, i = blockIdx.y;
    if (j >= g.nx) return;
    const int nx = g.nx, nz = g.nz;
    const size_t c = (size_t)i * nx + j;
    const float vp = in.vp[c], vs = in.vs[c], rho = in.rho[c], eps = in.eps[c], delta = in.delta[c];
    const float C33 = vp * vp * rho, C44 = vs * vs * rho, A = C33 - C44;
    const float R = sqrtf(2.0f * C33 * A * delta + A * A);
    const float g11 = g.hti ? gr.c33[c] : gr.c11[c], g33 = g.hti ? gr.c11[c] : gr.c33[c], g13 = gr.c13[c];
    // transposes of the staggered averages
    const float gb = 0.5f * (md_at(gr.bx, i, j, nz, nx - 1) + md_at(gr.bx, i, j - 1, nz, nx - 1)) +
                     0.5f * (md_at(gr.bz, i, j, nz - 1, nx) + md_at(gr.bz, i - 1, j, nz - 1, nx));
    const float g44s = 0.2f * (md_at(gr.c55, i - 1, j - 1, nz - 2, nx - 2) + 2.0f * md_at(gr.c55, i - 2, j - 1, nz - 2, nx - 2) +
                               md_at(gr.c55, i - 1, j - 2, nz - 2, nx - 2) + md_at(gr.c55, i - 2, j - 2, nz - 2, nx - 2));
    // C13 = R - C44
    const float d13_33 = (delta * (A + C33) + A) / R, d13_44 = -(C33 * delta + A) / R - 1.0f, d13_dl = C33 * A / R;
    const float t33 = g33 + g11 * (1.0f + 2.0f * eps) + g13 * d13_33;
    const float t44 = g44s + g13 * d13_44;
    if (go.vp) go.vp[c] = t33 * 2.0f * vp * rho;
    if (go.vs) go.vs[c] = t44 * 2.0f * vs * rho;
    if (go.rho) go.rho[c] = t33 * vp * vp + t44 * vs * vs - gb / (rho * rho);
    if (go.eps) go.eps[c] = g11 * 2.0f * C33;
    if (go.delta) go.delta[c] = g13 * d13_dl;
}

// ---- replicate padding from each plane's own shape + zero extension to the full grid ----------------------------------------------
struct PdGeom { int nzp, nxp, pml, top; int h[6], w[6]; };
struct Pd6 { const float* p[6]; };
struct Pd6o { float* p[6]; };

__global__ void __launch_bounds__(256)
pd_forward(const PdGeom g, const Pd6 in, const Pd6o out)
{
    const int j = blockIdx.x * blockDim.x + threadIdx.x, i = blockIdx.y;
    if (j >= g.nxp) return;
    const size_t o = (size_t)i * g.nxp + j;
#pragma unroll
    for (int k = 0; k < 6; ++k) {
        const int h = g.h[k], w = g.w[k];
        float v = 0.f;
        if (i < h + g.top + g.pml && j < w + 2 * g.pml) {
            const int ii = min(max(i - g.top, 0), h - 1), jj = min(max(j - g.pml, 0), w - 1);
            v = in.p[k][(size_t)ii * w + jj];
        }
        out.p[k][o] = v;
    }
}
// transpose: own cell (ii,jj) collects every padded cell that replicates it (a rectangle for edge and corner cells)
__global__ void __launch_bounds__(256)
pd_backward(const PdGeom g, int k, const float* __restrict__ gfull, float* __restrict__ gown)
{
    const int jj = blockIdx.x * blockDim.x + threadIdx.x, ii = blockIdx.y;
    const int h = g.h[k], w = g.w[k];
    if (jj >= w || ii >= h) return;
    const int i0 = ii == 0 ? 0 : ii + g.top, i1 = ii == h - 1 ? h + g.top + g.pml : ii + g.top + 1;
    const int j0 = jj == 0 ? 0 : jj + g.pml, j1 = jj == w - 1 ? w + 2 * g.pml : jj + g.pml + 1;
    double acc = 0.0;
    for (int i = i0; i < min(i1, g.nzp); ++i)
        for (int j = j0; j < min(j1, g.nxp); ++j) acc += (double)gfull[(size_t)i * g.nxp + j];
    gown[(size_t)ii * w + jj] = (float)acc;
}

int md_geom(const adfwi_elastic_moduli_desc* d, MdGeom* g)
{
    if (!d) return ADFWI_E_NULL;
    if (d->nz < 4 || d->nx < 4) return ADFWI_E_DIMS;
    g->nz = d->nz; g->nx = d->nx; g->hti = d->hti ? 1 : 0;
    return ADFWI_OK;
}
int pd_geom(const adfwi_elastic_pad_desc* d, PdGeom* g)
{
    if (!d) return ADFWI_E_NULL;
    if (d->nz < 4 || d->nx < 4 || d->pml < 0 || d->top < 0) return ADFWI_E_DIMS;
    if (d->nzp < d->nz + d->top + d->pml || d->nxp < d->nx + 2 * d->pml) return ADFWI_E_DIMS;
    g->nzp = d->nzp; g->nxp = d->nxp; g->pml = d->pml; g->top = d->top;
    const int h[6] = {d->nz, d->nz, d->nz, d->nz - 2, d->nz, d->nz - 1}, w[6] = {d->nx, d->nx, d->nx, d->nx - 2, d->nx - 1, d->nx};
    for (int k = 0; k < 6; ++k) { g->h[k] = h[k]; g->w[k] = w[k]; }
    return ADFWI_OK;
}

}  // namespace
}  // namespace adfwi

using namespace adfwi;

extern "C" int adfwi_elastic_moduli_forward(const adfwi_elastic_moduli_desc* desc, const float* vp, const float* vs, const float* rho,
                                            const float* eps, const float* delta, float* const* planes, void* stream)
{
    ADFWI_NVTX("adfwi_elastic_moduli_forward");
    MdGeom g;
    int rc = md_geom(desc, &g);
    if (rc) return rc;
    if (!vp || !vs || !rho || !eps || !delta || !planes) return ADFWI_E_NULL;
    for (int k = 0; k < 6; ++k) if (!planes[k]) return ADFWI_E_NULL;
    MdIn in{vp, vs, rho, eps, delta};
    MdOut out{planes[0], planes[1], planes[2], planes[3], planes[4], planes[5]};
    md_forward<<<dim3(cdiv(g.nx, 256), g.nz), 256, 0, (cudaStream_t)stream>>>(g, in, out);
    ADFWI_LAUNCH_CHECK();
    return ADFWI_OK;
}

extern "C" int adfwi_elastic_moduli_backward(const adfwi_elastic_moduli_desc* desc, const float* vp, const float* vs, const float* rho,
                                             const float* eps, const float* delta, const float* const* g_planes,
                                             float* g_vp, float* g_vs, float* g_rho, float* g_eps, float* g_delta, void* stream)
{
    ADFWI_NVTX("adfwi_elastic_moduli_backward");
    MdGeom g;
    int rc = md_geom(desc, &g);
    if (rc) return rc;
    if (!vp || !vs || !rho || !eps || !delta || !g_planes) return ADFWI_E_NULL;
    for (int k = 0; k < 6; ++k) if (!g_planes[k]) return ADFWI_E_NULL;
    MdIn in{vp, vs, rho, eps, delta};
    MdGrad gr{g_planes[0], g_planes[1], g_planes[2], g_planes[3], g_planes[4], g_planes[5]};
    MdGout go{g_vp, g_vs, g_rho, g_eps, g_delta};
    md_backward<<<dim3(cdiv(g.nx, 256), g.nz), 256, 0, (cudaStream_t)stream>>>(g, in, gr, go);
    ADFWI_LAUNCH_CHECK();
    return ADFWI_OK;
}

extern "C" int adfwi_elastic_pad_forward(const adfwi_elastic_pad_desc* desc, const float* const* planes, float* const* full, void* stream)
{
    ADFWI_NVTX("adfwi_elastic_pad_forward");
    PdGeom g;
    int rc = pd_geom(desc, &g);
    if (rc) return rc;
    if (!planes || !full) return ADFWI_E_NULL;
    Pd6 in; Pd6o out;
    for (int k = 0; k < 6; ++k) { if (!planes[k] || !full[k]) return ADFWI_E_NULL; in.p[k] = planes[k]; out.p[k] = full[k]; }
    pd_forward<<<dim3(cdiv(g.nxp, 256), g.nzp), 256, 0, (cudaStream_t)stream>>>(g, in, out);
    ADFWI_LAUNCH_CHECK();
    return ADFWI_OK;
}

extern "C" int adfwi_elastic_pad_backward(const adfwi_elastic_pad_desc* desc, const float* const* g_full, float* const* g_planes, void* stream)
{
    ADFWI_NVTX("adfwi_elastic_pad_backward");
    PdGeom g;
    int rc = pd_geom(desc, &g);
    if (rc) return rc;
    if (!g_full || !g_planes) return ADFWI_E_NULL;
    for (int k = 0; k < 6; ++k) {
        if (!g_planes[k]) continue;                 // gradient of this plane not wanted
        if (!g_full[k]) return ADFWI_E_NULL;
        pd_backward<<<dim3(cdiv(g.w[k], 256), g.h[k]), 256, 0, (cudaStream_t)stream>>>(g, k, g_full[k], g_planes[k]);
        ADFWI_LAUNCH_CHECK();
    }
    return ADFWI_OK;
}
#endif  // !ADFWI_HOST_EMUL
