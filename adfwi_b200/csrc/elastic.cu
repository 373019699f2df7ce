// elastic.cu -- 2-D P-SV velocity-stress staggered-grid solver (Cij planes: iso / VTI / HTI),
// O(2,4) and O(2,6), split-field PML or multiplicative sponge (ABL), free surface on/off:
// forward time loop and hand-written adjoint for sm_100a.
// Semantics: ADFWI/propagator/elastic_kernels.py:339-418 / :495-575 (split PML), :709-774 /
// :844-908 (ABL), operators :66-108; adjoint = reverse-mode derivative of those loops
// (SURVEY.md Appendix A.2).  C15 = C35 = 0 (ADFWI/model/parameters.py:38-44) are dropped: x+0*y
// is exact.  Compiled with -fmad=false: forward records are bit-identical to eager PyTorch on CPU.
//
// Adjoint structure (all gathers, no atomics on fields): per reverse step
//   el_adj_pre    : record cotangents + transpose of the free-surface velocity edits (atomicAdd)
//   el_adj_vel    : own-cell transpose of the velocity update -> g_bx,g_bz, scratch planes m1..m4
//   el_adj_vel_g  : gathers D^T(m) into the stress(-sum) cotangents
//   el_adj_stress : own-cell transpose of the stress update -> g_C11..g_C55, scratch planes nA..nD
//   el_adj_stress_g : gathers D^T(n) into lambda_vx, lambda_vz (pre-step)
#include "common.cuh"
#include "elastic_fused.h"

namespace adfwi {

struct ElGeom {
    int nzp, nxp, nt, NNv, h, fs;     // h = NN+1; fs = free surface flag
    int nz, nx, nabc, zoff;           // physical grid, illumination crop offset
    size_t plane;
    float dt, dx, dz, dt_dx, dt_dz, half_dt;
    float c[3];
};

struct ElShots { int begin, end; };

// field indices inside the state block of one shot
enum { F_TXX_X = 0, F_TXX_Z, F_TZZ_X, F_TZZ_Z, F_TXZ_X, F_TXZ_Z, F_VX_X, F_VX_Z, F_VZ_X, F_VZ_Z,
       F_VX, F_VZ, F_TXX, F_TZZ, F_TXZ, F_COUNT };

// ---- one-sided staggered operators (elastic_kernels.py:66-108), a points at cell (i,j) ----------
template <int NN> __device__ __forceinline__ float d_xf(const float* a, const float* c)
{ float s = c[0] * (a[1] - a[0]);
#pragma unroll
  for (int k = 1; k < NN; ++k) s = s + c[k] * (a[k + 1] - a[-k]);
  return s; }
template <int NN> __device__ __forceinline__ float d_zf(const float* a, int n, const float* c)
{ float s = c[0] * (a[n] - a[0]);
#pragma unroll
  for (int k = 1; k < NN; ++k) s = s + c[k] * (a[(k + 1) * n] - a[-k * n]);
  return s; }
template <int NN> __device__ __forceinline__ float d_xb(const float* a, const float* c)
{ float s = c[0] * (a[0] - a[-1]);
#pragma unroll
  for (int k = 1; k < NN; ++k) s = s + c[k] * (a[k] - a[-k - 1]);
  return s; }
template <int NN> __device__ __forceinline__ float d_zb(const float* a, int n, const float* c)
{ float s = c[0] * (a[0] - a[-n]);
#pragma unroll
  for (int k = 1; k < NN; ++k) s = s + c[k] * (a[k * n] - a[(-k - 1) * n]);
  return s; }

// ---- gathers of the operator transposes; m is zero outside the update region -------------------
__device__ __forceinline__ float at0(const float* m, int i, int j, int nzp, int nxp)
{ return (i >= 0 && i < nzp && j >= 0 && j < nxp) ? m[(size_t)i * nxp + j] : 0.f; }
template <int NN> __device__ __forceinline__ float t_xf(const float* m, int i, int j, int nzp, int nxp, const float* c)
{ float s = 0.f;
#pragma unroll
  for (int k = 0; k < NN; ++k) s += c[k] * (at0(m, i, j - k - 1, nzp, nxp) - at0(m, i, j + k, nzp, nxp));
  return s; }
template <int NN> __device__ __forceinline__ float t_zf(const float* m, int i, int j, int nzp, int nxp, const float* c)
{ float s = 0.f;
#pragma unroll
  for (int k = 0; k < NN; ++k) s += c[k] * (at0(m, i - k - 1, j, nzp, nxp) - at0(m, i + k, j, nzp, nxp));
  return s; }
template <int NN> __device__ __forceinline__ float t_xb(const float* m, int i, int j, int nzp, int nxp, const float* c)
{ float s = 0.f;
#pragma unroll
  for (int k = 0; k < NN; ++k) s += c[k] * (at0(m, i, j - k, nzp, nxp) - at0(m, i, j + k + 1, nzp, nxp));
  return s; }
template <int NN> __device__ __forceinline__ float t_zb(const float* m, int i, int j, int nzp, int nxp, const float* c)
{ float s = 0.f;
#pragma unroll
  for (int k = 0; k < NN; ++k) s += c[k] * (at0(m, i - k, j, nzp, nxp) - at0(m, i + k + 1, j, nzp, nxp));
  return s; }

#define EL_CELL()                                                    \
    const int j = blockIdx.x * blockDim.x + threadIdx.x;             \
    const int i = blockIdx.y * blockDim.y + threadIdx.y;             \
    const int s = sh.begin + blockIdx.z;                             \
    if (j >= g.nxp || i >= g.nzp || s >= sh.end) return;             \
    const int nxp = g.nxp, nzp = g.nzp;                              \
    const size_t cc = (size_t)i * nxp + j;                           \
    const bool inR = (i >= NN) && (i < nzp - NN) && (j >= NN) && (j < nxp - NN);

// ------------------------------------------------------------------------------------------
// forward: stress update + source + sums + free-surface mirrors      (:341-384 / :720-742)
// st = state block of all shots: field f of shot s at st + (s*F_COUNT + f)*plane
// hist slot of (shot s, local step tl): hist + ((s*hist_len + tl)*5)*plane : vx, vz, txx, tzz, txz
// ------------------------------------------------------------------------------------------
template <int NN, bool PML, bool FS, bool SAVE>
__global__ void __launch_bounds__(256)
el_fwd_stress(const ElGeom g, const ElShots sh, const float* __restrict__ C11, const float* __restrict__ C13,
              const float* __restrict__ C33, const float* __restrict__ C55, const float* __restrict__ bcx,
              const float* __restrict__ bcz, float* __restrict__ st, const float* __restrict__ mt,
              const float* __restrict__ src_v, const int64_t* __restrict__ sx, const int64_t* __restrict__ sz,
              float* __restrict__ hist, int hist_len, int tl, int it)
{
    EL_CELL();
    float* S = st + (size_t)s * F_COUNT * g.plane;
    const float* vx = S + F_VX * g.plane + cc;
    const float* vz = S + F_VZ * g.plane + cc;
    float* H = SAVE ? hist + ((size_t)s * hist_len + tl) * 5 * g.plane : nullptr;
    if (SAVE) { __stcs(H + cc, vx[0]); __stcs(H + g.plane + cc, vz[0]); }
    if (FS && i < NN) {           // rows above the region only ever receive the mirrored sums
        if (SAVE) {               // history cells nobody else writes must read as zero in the adjoint
            __stcs(H + 2 * g.plane + cc, 0.f);
            if (i != g.h - 2) __stcs(H + 3 * g.plane + cc, 0.f);
            if (i != g.h - 2 && i != g.h - 3) __stcs(H + 4 * g.plane + cc, 0.f);
        }
        return;
    }
    const bool is_src = ((int64_t)i == sz[s]) && ((int64_t)j == sx[s]);
    if (!inR && !is_src) {        // outside the region the split fields and their sums stay zero
        if (SAVE) { __stcs(H + 2 * g.plane + cc, 0.f); __stcs(H + 3 * g.plane + cc, 0.f); __stcs(H + 4 * g.plane + cc, 0.f); }
        return;
    }
    float txx, tzz, txz;
    float sxx = 0.f, szz = 0.f, sxz = 0.f;
    if (is_src) {
        const float* M = mt + (size_t)s * 9;
        const float v = src_v[(size_t)s * g.nt + it];
        if (PML) { sxx = (-(M[0] / 2.0f)) * v; szz = (-(M[8] / 2.0f)) * v; sxz = (-(M[2] / 2.0f)) * v; }
        else { const float sc = (float)(-1.0 / 3.0); sxx = (sc * M[0]) * v; szz = (sc * M[8]) * v; sxz = (sc * M[2]) * v; }
    }
    if (PML) {
        float* p0 = S + F_TXX_X * g.plane + cc; float* p1 = S + F_TXX_Z * g.plane + cc;
        float* p2 = S + F_TZZ_X * g.plane + cc; float* p3 = S + F_TZZ_Z * g.plane + cc;
        float* p4 = S + F_TXZ_X * g.plane + cc; float* p5 = S + F_TXZ_Z * g.plane + cc;
        float a0 = *p0, a1 = *p1, a2 = *p2, a3 = *p3, a4 = *p4, a5 = *p5;
        if (inR) {
            const float bx_ = bcx[cc], bz_ = bcz[cc];
            const float pxd = 1.0f + g.half_dt * bx_, pxn = 1.0f - g.half_dt * bx_;
            const float pzd = 1.0f + g.half_dt * bz_, pzn = 1.0f - g.half_dt * bz_;
            const float pxi = 1.0f / pxd, pzi = 1.0f / pzd;
            const float dxb_vx = d_xb<NN>(vx, g.c), dzb_vz = d_zb<NN>(vz, nxp, g.c);
            const float dxf_vz = d_xf<NN>(vz, g.c), dzf_vx = d_zf<NN>(vx, nxp, g.c);
            const float c11 = C11[cc], c13 = C13[cc], c33 = C33[cc], c55 = C55[cc];
            a0 = (pxn * a0 + g.dt_dx * (c11 * dxb_vx)) * pxi;
            a1 = (pzn * a1 + g.dt_dz * (c13 * dzb_vz)) * pzi;
            a2 = (pxn * a2 + g.dt_dx * (c13 * dxb_vx)) * pxi;
            a3 = (pzn * a3 + g.dt_dz * (c33 * dzb_vz)) * pzi;
            a4 = (pxn * a4 + g.dt_dx * (c55 * dxf_vz)) * pxi;
            a5 = (pzn * a5 + g.dt_dz * (c55 * dzf_vx)) * pzi;
        }
        if (is_src) { a0 += sxx; a1 += sxx; a2 += szz; a3 += szz; a4 += sxz; a5 += sxz; }
        *p0 = a0; *p1 = a1; *p2 = a2; *p3 = a3; *p4 = a4; *p5 = a5;
        txx = a0 + a1; tzz = a2 + a3; txz = a4 + a5;
    } else {
        txx = S[F_TXX * g.plane + cc]; tzz = S[F_TZZ * g.plane + cc]; txz = S[F_TXZ * g.plane + cc];
        if (inR) {
            const float dxb_vx = d_xb<NN>(vx, g.c), dzb_vz = d_zb<NN>(vz, nxp, g.c);
            const float dxf_vz = d_xf<NN>(vz, g.c), dzf_vx = d_zf<NN>(vx, nxp, g.c);
            const float c11 = C11[cc], c13 = C13[cc], c33 = C33[cc], c55 = C55[cc];
            txx = txx + g.dt * ((c11 * dxb_vx) / g.dx + (c13 * dzb_vz) / g.dz);
            tzz = tzz + g.dt * ((c13 * dxb_vx) / g.dx + (c33 * dzb_vz) / g.dz);
            txz = txz + g.dt * ((c55 * dxf_vz) / g.dx + (c55 * dzf_vx) / g.dz);
        }
        if (is_src) { txx += sxx; tzz += szz; txz += sxz; }
    }
    float* Txx = S + F_TXX * g.plane; float* Tzz = S + F_TZZ * g.plane; float* Txz = S + F_TXZ * g.plane;
    if (FS) {            // :380-384 on the sums (PML) / on the state (ABL); whole rows
        const int h = g.h;
        if (i == h - 1) {
            tzz = 0.f;
            Txz[cc - nxp] = -txz;                          // txz[h-2] = -txz[h-1]
            if (SAVE) __stcs(H + 4 * g.plane + cc - nxp, -txz);
        } else if (i == h) {
            Tzz[cc - 2 * (size_t)nxp] = -tzz;              // tzz[h-2] = -tzz[h]
            Txz[cc - 3 * (size_t)nxp] = -txz;              // txz[h-3] = -txz[h]
            if (SAVE) { __stcs(H + 3 * g.plane + cc - 2 * (size_t)nxp, -tzz); __stcs(H + 4 * g.plane + cc - 3 * (size_t)nxp, -txz); }
        }
    }
    Txx[cc] = txx; Tzz[cc] = tzz; Txz[cc] = txz;
    if (SAVE) { __stcs(H + 2 * g.plane + cc, txx); __stcs(H + 3 * g.plane + cc, tzz); __stcs(H + 4 * g.plane + cc, txz); }
}

// forward: velocity update + sums (+ sponge)      (:387-396 / :744-760)
// With a free surface the rows h-1, h are left undamped here; el_fwd_post applies the edits of
// :399-402 from the undamped values and then damps rows h-3..h.
template <int NN, bool PML, bool FS>
__global__ void __launch_bounds__(256)
el_fwd_vel(const ElGeom g, const ElShots sh, const float* __restrict__ bx, const float* __restrict__ bz,
           const float* __restrict__ bcx, const float* __restrict__ bcz, float* __restrict__ st)
{
    EL_CELL();
    if (!inR) return;
    float* S = st + (size_t)s * F_COUNT * g.plane;
    const float* txx = S + F_TXX * g.plane + cc;
    const float* tzz = S + F_TZZ * g.plane + cc;
    const float* txz = S + F_TXZ * g.plane + cc;
    const float dxf_txx = d_xf<NN>(txx, g.c), dzb_txz = d_zb<NN>(txz, nxp, g.c);
    const float dxb_txz = d_xb<NN>(txz, g.c), dzf_tzz = d_zf<NN>(tzz, nxp, g.c);
    const float bx_ = bx[cc], bz_ = bz[cc];
    if (PML) {
        const float cx = bcx[cc], cz = bcz[cc];
        const float pxd = 1.0f + g.half_dt * cx, pxn = 1.0f - g.half_dt * cx;
        const float pzd = 1.0f + g.half_dt * cz, pzn = 1.0f - g.half_dt * cz;
        float* q0 = S + F_VX_X * g.plane + cc; float* q1 = S + F_VX_Z * g.plane + cc;
        float* q2 = S + F_VZ_X * g.plane + cc; float* q3 = S + F_VZ_Z * g.plane + cc;
        const float a0 = (pxn * *q0 + ((g.dt * bx_) * dxf_txx) / g.dx) / pxd;
        const float a1 = (pzn * *q1 + ((g.dt * bx_) * dzb_txz) / g.dz) / pzd;
        const float a2 = (pxn * *q2 + ((g.dt * bz_) * dxb_txz) / g.dx) / pxd;
        const float a3 = (pzn * *q3 + ((g.dt * bz_) * dzf_tzz) / g.dz) / pzd;
        *q0 = a0; *q1 = a1; *q2 = a2; *q3 = a3;
        S[F_VX * g.plane + cc] = a0 + a1;
        S[F_VZ * g.plane + cc] = a2 + a3;
    } else {
        float vx = S[F_VX * g.plane + cc], vz = S[F_VZ * g.plane + cc];
        vx += (g.dt * bx_) * (dxf_txx / g.dx + dzb_txz / g.dz);
        vz += (g.dt * bz_) * (dxb_txz / g.dx + dzf_tzz / g.dz);
        if (FS && (i == g.h - 1 || i == g.h)) {   // undamped copies for el_fwd_post (split slots are unused by ABL)
            S[F_VX_X * g.plane + cc] = vx; S[F_VZ_X * g.plane + cc] = vz;
        }
        const float dm = bcx[cc];
        vx *= dm; vz *= dm;
        S[F_VX * g.plane + cc] = vx; S[F_VZ * g.plane + cc] = vz;
    }
}

// forward: free-surface velocity edits (:399-402), sponge on the top rows, receiver sampling (:405-409)
// grid: x over max(nxp, nr) ; blockIdx.y = 0 -> free-surface columns, 1 -> receivers; z = shot
template <int NN, bool PML, bool FS>
__global__ void el_fwd_post(const ElGeom g, const ElShots sh, const float* __restrict__ damp, float* __restrict__ st,
                            int nr, const int64_t* __restrict__ rx, const int64_t* __restrict__ rz,
                            float* r0, float* r1, float* r2, float* r3, float* r4, int it)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    const int s = sh.begin + blockIdx.z;
    if (s >= sh.end) return;
    float* S = st + (size_t)s * F_COUNT * g.plane;
    const int nxp = g.nxp;
    if (blockIdx.y == 0) {
        if (!FS) return;
        const int j = t, h = g.h;
        if (j >= nxp) return;
        float* vx = S + F_VX * g.plane; float* vz = S + F_VZ * g.plane;
        // undamped values of rows h-1, h: the state itself (PML) or the side copies (ABL)
        const float* uvx = PML ? vx : S + F_VX_X * g.plane;
        const float* uvz = PML ? vz : S + F_VZ_X * g.plane;
        if (j < NN || j >= nxp - NN) return;
        const float vz1 = uvz[(size_t)(h - 1) * nxp + j];                    // vz[h-2,j] <- vz[h-1,j]
        // vz[h-2,j+1] is the freshly assigned copy when j+1 is in J; at j+1 = nxp-NN both that cell
        // and vz[h-1,j+1] are cells nobody ever writes (outside the region): zero
        const bool nin = (j + 1 < nxp - NN);
        const float vz1n = nin ? uvz[(size_t)(h - 1) * nxp + j + 1] : 0.f;
        const float vz2n = vz1n;
        const float vxh = uvx[(size_t)h * nxp + j];
        const float nvx = (((vz2n - vz1) + vz1n) - vz1) + vxh;
        float o2 = vz1, o3 = vz1, ovx = nvx;
        if (!PML) { o2 *= damp[(size_t)(h - 2) * nxp + j]; ovx *= damp[(size_t)(h - 2) * nxp + j]; o3 *= damp[(size_t)(h - 3) * nxp + j]; }
        vz[(size_t)(h - 2) * nxp + j] = o2;
        vx[(size_t)(h - 2) * nxp + j] = ovx;
        vz[(size_t)(h - 3) * nxp + j] = o3;
    } else {
        const int r = t;
        if (r >= nr) return;
        const int64_t z = rz[r], x = rx[r];
        if (z < 0 || z >= g.nzp || x < 0 || x >= nxp) return;
        const size_t c = (size_t)z * nxp + x, o = ((size_t)s * g.nt + it) * nr + r;
        r0[o] = S[F_TXX * g.plane + c]; r1[o] = S[F_TZZ * g.plane + c]; r2[o] = S[F_TXZ * g.plane + c];
        r3[o] = S[F_VX * g.plane + c];  r4[o] = S[F_VZ * g.plane + c];
    }
}

// squares of the five sum fields of the current state, summed over shots (:414-418)
__global__ void el_illum_acc(const ElGeom g, const ElShots sh, const float* __restrict__ st, float* __restrict__ ill)
{
    const size_t q = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= g.plane) return;
    const int order[5] = {F_TXX, F_TZZ, F_TXZ, F_VX, F_VZ};
    for (int k = 0; k < 5; ++k) {
        float a = 0.f;
        for (int s = sh.begin; s < sh.end; ++s) { const float v = st[((size_t)s * F_COUNT + order[k]) * g.plane + q]; a += v * v; }
        ill[k * g.plane + q] += a;
    }
}

__global__ void el_illum_crop(const ElGeom g, const float* __restrict__ ill, float* o0, float* o1, float* o2, float* o3, float* o4)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x, z = blockIdx.y;
    if (x >= g.nx || z >= g.nz) return;
    const size_t c = (size_t)(z + g.zoff) * g.nxp + (x + g.nabc), o = (size_t)z * g.nx + x;
    float* out[5] = {o0, o1, o2, o3, o4};
    for (int k = 0; k < 5; ++k) if (out[k]) out[k][o] = ill[k * g.plane + c];
}

// ------------------------------------------------------------------------------------------
// adjoint.  L = cotangent block, same field slots as the state (F_TXX..F_TXZ = mu (PML) /
// lambda of the stresses (ABL)); scr = 6 scratch planes per shot.
// ------------------------------------------------------------------------------------------
// 10T + 9T: record cotangents and transpose of the free-surface velocity edits (pure adds into
// rows h-1 and h; rows h-2,h-3 are only read -- they are overwritten by el_adj_stress_g).
// ABL: the sponge (8T) multiplies lambda_v before the free-surface transpose; that product is
// formed on the fly in el_adj_vel, so here only the record scatter is done for ABL.
template <int NN, bool PML, bool FS>
__global__ void el_adj_pre(const ElGeom g, const ElShots sh, float* __restrict__ Lb, int nr,
                           const int64_t* __restrict__ rx, const int64_t* __restrict__ rz,
                           const float* g0, const float* g1, const float* g2, const float* g3, const float* g4, int it)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    const int s = sh.begin + blockIdx.z;
    if (s >= sh.end) return;
    float* L = Lb + (size_t)s * F_COUNT * g.plane;
    const int nxp = g.nxp;
    if (blockIdx.y == 0) {
        if (!FS || !PML) return;
        const int j = t, h = g.h;
        if (j < NN || j >= nxp - NN) return;
        const float* lvx = L + F_VX * g.plane; const float* lvz = L + F_VZ * g.plane;
        const float qj = lvx[(size_t)(h - 2) * nxp + j];
        const float qm = (j - 1 >= NN) ? lvx[(size_t)(h - 2) * nxp + j - 1] : 0.f;
        const float add_vz = lvz[(size_t)(h - 2) * nxp + j] + lvz[(size_t)(h - 3) * nxp + j] + 2.0f * (qm - qj);
        atomicAdd(L + F_VZ * g.plane + (size_t)(h - 1) * nxp + j, add_vz);
        atomicAdd(L + F_VX * g.plane + (size_t)h * nxp + j, qj);
    } else {
        const int r = t;
        if (r >= nr) return;
        const int64_t z = rz[r], x = rx[r];
        if (z < 0 || z >= g.nzp || x < 0 || x >= nxp) return;
        const size_t c = (size_t)z * nxp + x, o = ((size_t)s * g.nt + it) * nr + r;
        if (g0) atomicAdd(L + F_TXX * g.plane + c, g0[o]);
        if (g1) atomicAdd(L + F_TZZ * g.plane + c, g1[o]);
        if (g2) atomicAdd(L + F_TXZ * g.plane + c, g2[o]);
        if (g3) atomicAdd(L + F_VX * g.plane + c, g3[o]);
        if (g4) atomicAdd(L + F_VZ * g.plane + c, g4[o]);
    }
}

// 8T,7T,6T own-cell part: velocity-update transpose on the region.
// hist slot: vx, vz (pre-step), txx, tzz, txz (post-free-surface) of this step.
// scr planes 0..3 <- m1 (Dxf^T -> txx), m2 (Dzb^T -> txz), m3 (Dxb^T -> txz), m4 (Dzf^T -> tzz);
// ABL also scr 4,5 <- q_vx, q_vz (the damped, free-surface-transposed lambda_v of the region).
template <int NN, bool PML, bool FS>
__global__ void __launch_bounds__(256)
el_adj_vel(const ElGeom g, const ElShots sh, const float* __restrict__ bx, const float* __restrict__ bz,
           const float* __restrict__ bcx, const float* __restrict__ bcz, float* __restrict__ Lb,
           float* __restrict__ scrb, const float* __restrict__ hist, int hist_len, int tl,
           float* __restrict__ gpart)
{
    EL_CELL();
    if (!inR) return;
    float* L = Lb + (size_t)s * F_COUNT * g.plane;
    float* scr = scrb + (size_t)s * 6 * g.plane;
    const float* H = hist + ((size_t)s * hist_len + tl) * 5 * g.plane;
    const float* txx = H + 2 * g.plane + cc; const float* tzz = H + 3 * g.plane + cc; const float* txz = H + 4 * g.plane + cc;
    const float e1 = d_xf<NN>(txx, g.c), e2 = d_zb<NN>(txz, nxp, g.c), e3 = d_xb<NN>(txz, g.c), e4 = d_zf<NN>(tzz, nxp, g.c);
    const float bx_ = bx[cc], bz_ = bz[cc];
    float* gp = gpart + (size_t)blockIdx.z * 6 * g.plane;
    if (PML) {
        const float cx = bcx[cc], cz = bcz[cc];
        const float pxd = 1.0f + g.half_dt * cx, pxn = 1.0f - g.half_dt * cx;
        const float pzd = 1.0f + g.half_dt * cz, pzn = 1.0f - g.half_dt * cz;
        const float lvx = L[F_VX * g.plane + cc], lvz = L[F_VZ * g.plane + cc];
        const float q1 = L[F_VX_X * g.plane + cc] + lvx, q2 = L[F_VX_Z * g.plane + cc] + lvx;
        const float q3 = L[F_VZ_X * g.plane + cc] + lvz, q4 = L[F_VZ_Z * g.plane + cc] + lvz;
        const float w1 = q1 * g.dt / g.dx / pxd, w2 = q2 * g.dt / g.dz / pzd;
        const float w3 = q3 * g.dt / g.dx / pxd, w4 = q4 * g.dt / g.dz / pzd;
        gp[4 * g.plane + cc] += w1 * e1 + w2 * e2;
        gp[5 * g.plane + cc] += w3 * e3 + w4 * e4;
        scr[0 * g.plane + cc] = w1 * bx_; scr[1 * g.plane + cc] = w2 * bx_;
        scr[2 * g.plane + cc] = w3 * bz_; scr[3 * g.plane + cc] = w4 * bz_;
        L[F_VX_X * g.plane + cc] = pxn * q1 / pxd; L[F_VX_Z * g.plane + cc] = pzn * q2 / pzd;
        L[F_VZ_X * g.plane + cc] = pxn * q3 / pxd; L[F_VZ_Z * g.plane + cc] = pzn * q4 / pzd;
    } else {
        const float* lvx = L + F_VX * g.plane; const float* lvz = L + F_VZ * g.plane;
        const float* dm = bcx;
        float qx = dm[cc] * lvx[cc], qz = dm[cc] * lvz[cc];          // 8T
        if (FS) {                                                     // 7T (gather form)
            const int h = g.h;
            if (i == h) qx += dm[(size_t)(h - 2) * nxp + j] * lvx[(size_t)(h - 2) * nxp + j];
            if (i == h - 1) {
                const size_t a2 = (size_t)(h - 2) * nxp + j, a3 = (size_t)(h - 3) * nxp + j;
                const float qj = dm[a2] * lvx[a2];
                const float qm = (j - 1 >= NN) ? dm[a2 - 1] * lvx[a2 - 1] : 0.f;
                qz += dm[a2] * lvz[a2] + dm[a3] * lvz[a3] + 2.0f * (qm - qj);
            }
        }
        const float wx = qx * g.dt, wz = qz * g.dt;
        gp[4 * g.plane + cc] += wx * (e1 / g.dx + e2 / g.dz);
        gp[5 * g.plane + cc] += wz * (e3 / g.dx + e4 / g.dz);
        scr[0 * g.plane + cc] = wx * bx_ / g.dx; scr[1 * g.plane + cc] = wx * bx_ / g.dz;
        scr[2 * g.plane + cc] = wz * bz_ / g.dx; scr[3 * g.plane + cc] = wz * bz_ / g.dz;
        scr[4 * g.plane + cc] = qx; scr[5 * g.plane + cc] = qz;
    }
}

// 6T gather part: stress(-sum) cotangents += D^T(m)   (every cell of the grid)
template <int NN>
__global__ void __launch_bounds__(256)
el_adj_vel_g(const ElGeom g, const ElShots sh, float* __restrict__ Lb, const float* __restrict__ scrb)
{
    EL_CELL();
    (void)inR;
    float* L = Lb + (size_t)s * F_COUNT * g.plane;
    const float* scr = scrb + (size_t)s * 6 * g.plane;
    L[F_TXX * g.plane + cc] += t_xf<NN>(scr + 0 * g.plane, i, j, nzp, nxp, g.c);
    L[F_TXZ * g.plane + cc] += t_zb<NN>(scr + 1 * g.plane, i, j, nzp, nxp, g.c) + t_xb<NN>(scr + 2 * g.plane, i, j, nzp, nxp, g.c);
    L[F_TZZ * g.plane + cc] += t_zf<NN>(scr + 3 * g.plane, i, j, nzp, nxp, g.c);
}

// 5T,4T,3T,2T,1T own-cell part: stress-update transpose on the region.
// scr planes 0..3 <- nA (Dxb^T -> vx), nB (Dzb^T -> vz), nC (Dxf^T -> vz), nD (Dzf^T -> vx)
template <int NN, bool PML, bool FS>
__global__ void __launch_bounds__(256)
el_adj_stress(const ElGeom g, const ElShots sh, const float* __restrict__ C11, const float* __restrict__ C13,
              const float* __restrict__ C33, const float* __restrict__ C55, const float* __restrict__ bcx,
              const float* __restrict__ bcz, float* __restrict__ Lb, float* __restrict__ scrb,
              const float* __restrict__ hist, int hist_len, int tl, float* __restrict__ gpart,
              const float* __restrict__ mt, const int64_t* __restrict__ sx, const int64_t* __restrict__ sz,
              float* __restrict__ g_src, int it)
{
    EL_CELL();
    if (!inR) return;
    float* L = Lb + (size_t)s * F_COUNT * g.plane;
    float* scr = scrb + (size_t)s * 6 * g.plane;
    const float* H = hist + ((size_t)s * hist_len + tl) * 5 * g.plane;
    const float* vx = H + cc; const float* vz = H + g.plane + cc;
    const float dxb_vx = d_xb<NN>(vx, g.c), dzb_vz = d_zb<NN>(vz, nxp, g.c);
    const float dxf_vz = d_xf<NN>(vz, g.c), dzf_vx = d_zf<NN>(vx, nxp, g.c);
    const float c11 = C11[cc], c13 = C13[cc], c33 = C33[cc], c55 = C55[cc];
    float mxx = L[F_TXX * g.plane + cc], mzz = L[F_TZZ * g.plane + cc], mxz = L[F_TXZ * g.plane + cc];
    if (FS) {                                                  // 5T / 4T(ABL) in gather form
        const int h = g.h;
        if (i == h) {
            mxz -= L[F_TXZ * g.plane + cc - 3 * (size_t)nxp];
            mzz -= L[F_TZZ * g.plane + cc - 2 * (size_t)nxp];
        } else if (i == h - 1) {
            mxz -= L[F_TXZ * g.plane + cc - nxp];
            mzz = 0.f;
        }
    }
    float* gp = gpart + (size_t)blockIdx.z * 6 * g.plane;
    const bool is_src = g_src && ((int64_t)i == sz[s]) && ((int64_t)j == sx[s]);
    if (PML) {
        const float cx = bcx[cc], cz = bcz[cc];
        const float pxd = 1.0f + g.half_dt * cx, pxn = 1.0f - g.half_dt * cx;
        const float pzd = 1.0f + g.half_dt * cz, pzn = 1.0f - g.half_dt * cz;
        const float pxi = 1.0f / pxd, pzi = 1.0f / pzd;
        const float l1 = L[F_TXX_X * g.plane + cc] + mxx, l2 = L[F_TXX_Z * g.plane + cc] + mxx;   // 4T
        const float l3 = L[F_TZZ_X * g.plane + cc] + mzz, l4 = L[F_TZZ_Z * g.plane + cc] + mzz;
        const float l5 = L[F_TXZ_X * g.plane + cc] + mxz, l6 = L[F_TXZ_Z * g.plane + cc] + mxz;
        if (is_src) {                                                                             // 3T
            const float* M = mt + (size_t)s * 9;
            g_src[(size_t)s * g.nt + it] = -(M[0] / 2.0f) * (l1 + l2) - (M[8] / 2.0f) * (l3 + l4) - (M[2] / 2.0f) * (l5 + l6);
        }
        const float q1 = l1 * pxi * g.dt_dx, q2 = l2 * pzi * g.dt_dz, q3 = l3 * pxi * g.dt_dx;
        const float q4 = l4 * pzi * g.dt_dz, q5 = l5 * pxi * g.dt_dx, q6 = l6 * pzi * g.dt_dz;
        gp[0 * g.plane + cc] += q1 * dxb_vx;
        gp[1 * g.plane + cc] += q2 * dzb_vz + q3 * dxb_vx;
        gp[2 * g.plane + cc] += q4 * dzb_vz;
        gp[3 * g.plane + cc] += q5 * dxf_vz + q6 * dzf_vx;
        scr[0 * g.plane + cc] = q1 * c11 + q3 * c13;
        scr[1 * g.plane + cc] = q2 * c13 + q4 * c33;
        scr[2 * g.plane + cc] = q5 * c55;
        scr[3 * g.plane + cc] = q6 * c55;
        L[F_TXX_X * g.plane + cc] = pxn * (l1 * pxi); L[F_TXX_Z * g.plane + cc] = pzn * (l2 * pzi);
        L[F_TZZ_X * g.plane + cc] = pxn * (l3 * pxi); L[F_TZZ_Z * g.plane + cc] = pzn * (l4 * pzi);
        L[F_TXZ_X * g.plane + cc] = pxn * (l5 * pxi); L[F_TXZ_Z * g.plane + cc] = pzn * (l6 * pzi);
    } else {
        if (is_src) {
            const float* M = mt + (size_t)s * 9;
            const float sc = (float)(-1.0 / 3.0);
            g_src[(size_t)s * g.nt + it] = sc * (M[0] * mxx + M[8] * mzz + M[2] * mxz);
        }
        const float qx = mxx * g.dt, qz = mzz * g.dt, qs = mxz * g.dt;
        gp[0 * g.plane + cc] += qx * dxb_vx / g.dx;
        gp[1 * g.plane + cc] += qx * dzb_vz / g.dz + qz * dxb_vx / g.dx;
        gp[2 * g.plane + cc] += qz * dzb_vz / g.dz;
        gp[3 * g.plane + cc] += qs * (dxf_vz / g.dx + dzf_vx / g.dz);
        scr[0 * g.plane + cc] = (qx * c11 + qz * c13) / g.dx;
        scr[1 * g.plane + cc] = (qx * c13 + qz * c33) / g.dz;
        scr[2 * g.plane + cc] = qs * c55 / g.dx;
        scr[3 * g.plane + cc] = qs * c55 / g.dz;
        // the free-surface transposed values become the state (rows h-2,h-3 are zeroed by el_adj_stress_g)
        if (FS && (i == g.h || i == g.h - 1)) { L[F_TZZ * g.plane + cc] = mzz; L[F_TXZ * g.plane + cc] = mxz; }
    }
}

// 2T/1T gather part: new lambda_vx, lambda_vz (pre-step) for every cell; clears what must be zero
// before the next (earlier) step: mu planes (PML), mirrored stress rows (ABL).
template <int NN, bool PML, bool FS>
__global__ void __launch_bounds__(256)
el_adj_stress_g(const ElGeom g, const ElShots sh, float* __restrict__ Lb, const float* __restrict__ scrb)
{
    EL_CELL();
    float* L = Lb + (size_t)s * F_COUNT * g.plane;
    const float* scr = scrb + (size_t)s * 6 * g.plane;
    float nvx = t_xb<NN>(scr + 0 * g.plane, i, j, nzp, nxp, g.c) + t_zf<NN>(scr + 3 * g.plane, i, j, nzp, nxp, g.c);
    float nvz = t_zb<NN>(scr + 1 * g.plane, i, j, nzp, nxp, g.c) + t_xf<NN>(scr + 2 * g.plane, i, j, nzp, nxp, g.c);
    if (PML) {
        L[F_TXX * g.plane + cc] = 0.f; L[F_TZZ * g.plane + cc] = 0.f; L[F_TXZ * g.plane + cc] = 0.f;
    } else {
        if (inR) { nvx += scr[4 * g.plane + cc]; nvz += scr[5 * g.plane + cc]; }
        if (FS) {
            const int h = g.h;
            if (i == h - 2) { L[F_TZZ * g.plane + cc] = 0.f; L[F_TXZ * g.plane + cc] = 0.f; }
            if (i == h - 3) L[F_TXZ * g.plane + cc] = 0.f;
        }
    }
    L[F_VX * g.plane + cc] = nvx; L[F_VZ * g.plane + cc] = nvz;
}

__global__ void el_reduce_parts(size_t plane, int nparts, int k, const float* __restrict__ part, float* __restrict__ out)
{
    const size_t q = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= plane) return;
    float a = 0.f;
    for (int p = 0; p < nparts; ++p) a += part[((size_t)p * 6 + k) * plane + q];
    out[q] = a;
}

// ------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------
struct ElPlan {
    ElGeom g;
    int ns, nr, NN, PML, FS, save, n_segments;
    int K, nseg, nckpt, G;
    int nstate;               // planes checkpointed per shot: 12 (PML) / 5 (ABL)
    float *st, *L, *scr, *hist, *ckpt, *gpart, *ill;
    size_t bytes;
};

static int el_make_plan(const adfwi_elastic_desc* d, void* ws, ElPlan* P)
{
    if (!d) return ADFWI_E_NULL;
    if (d->fd_order != 4 && d->fd_order != 6) return ADFWI_E_ORDER;
    const int NN = d->fd_order / 2;
    if (d->nzp < 4 * NN + 4 || d->nxp < 4 * NN + 4 || d->ns < 1 || d->nt < 1 || d->nr < 0) return ADFWI_E_DIMS;
    if (d->nz < 1 || d->nx < 1 || d->nabc < 0) return ADFWI_E_DIMS;
    const int zoff = d->free_surface ? NN : NN + d->nabc;
    if (zoff + d->nz > d->nzp || d->nabc + d->nx > d->nxp) return ADFWI_E_DIMS;
    ElGeom& g = P->g;
    g.nzp = d->nzp; g.nxp = d->nxp; g.nt = d->nt; g.NNv = NN; g.h = NN + 1; g.fs = d->free_surface ? 1 : 0;
    g.nz = d->nz; g.nx = d->nx; g.nabc = d->nabc; g.zoff = zoff;
    g.plane = (size_t)d->nzp * d->nxp;
    g.dt = d->dt; g.dx = d->dx; g.dz = d->dz; g.dt_dx = d->dt_dx; g.dt_dz = d->dt_dz; g.half_dt = d->half_dt;
    for (int k = 0; k < 3; ++k) g.c[k] = d->fdc[k];
    P->ns = d->ns; P->nr = d->nr; P->NN = NN; P->PML = d->abc_pml ? 1 : 0; P->FS = g.fs;
    P->save = d->save_history ? 1 : 0;
    P->n_segments = d->n_segments > 0 ? d->n_segments : 1;
    int K = d->ckpt_interval;
    if (K <= 0 || K >= d->nt) K = d->nt;
    P->K = K; P->nseg = cdiv(d->nt, K); P->nckpt = P->nseg > 2 ? P->nseg - 2 : 0;
    P->nstate = P->PML ? 12 : 5;
    int G = d->shots_per_group;
    if (G <= 0) {
        const size_t per_shot = g.plane * sizeof(float) * (P->PML ? 15 : 8);
        G = (int)((size_t)(64u << 20) / per_shot);
        if (G < 1) G = 1;
    }
    if (G > d->ns) G = d->ns;
    P->G = G;
    Carver cv(ws);
    const size_t sp = (size_t)d->ns * g.plane;
    P->st = cv.take<float>(sp * F_COUNT);
    P->ill = cv.take<float>(5 * g.plane);
    P->L = P->scr = P->hist = P->ckpt = P->gpart = nullptr;
    if (P->save) {
        P->L = cv.take<float>(sp * F_COUNT);
        P->scr = cv.take<float>(sp * 6);
        P->gpart = cv.take<float>((size_t)G * 6 * g.plane);
        if (P->nckpt) P->ckpt = cv.take<float>((size_t)P->nckpt * P->nstate * sp);
        P->hist = cv.take<float>((size_t)K * 5 * sp);
    }
    P->bytes = cv.off;
    return ADFWI_OK;
}

// state <-> checkpoint copies (ABL keeps its 5 fields in slots F_VX,F_VZ,F_TXX,F_TZZ,F_TXZ = 10..14)
static int el_copy_state(const ElPlan& P, cudaStream_t st, int sb, int se, float* ck, bool to_ckpt)
{
    const int first = P.PML ? 0 : F_VX, cnt = P.nstate;
    for (int s = sb; s < se; ++s) {
        float* a = P.st + ((size_t)s * F_COUNT + first) * P.g.plane;
        float* b = ck + (size_t)s * cnt * P.g.plane;
        const size_t n = (size_t)cnt * P.g.plane * sizeof(float);
        ADFWI_CUDA(cudaMemcpyAsync(to_ckpt ? b : a, to_ckpt ? a : b, n, cudaMemcpyDeviceToDevice, st));
    }
    return ADFWI_OK;
}

struct ElArgs {
    const float *C11, *C13, *C33, *C55, *bx, *bz, *bcx, *bcz, *mt, *src_v;
    const int64_t *sx, *sz, *rx, *rz;
};

template <int NN, bool PML, bool FS>
static int el_forward_step_t(const ElPlan& P, cudaStream_t st, int sb, int se, int it, bool save, int tl,
                             const ElArgs& a, float* const* rcv)
{
    const dim3 blk(64, 4);
    const ElShots sh{sb, se};
    const dim3 grd(cdiv(P.g.nxp, blk.x), cdiv(P.g.nzp, blk.y), se - sb);
    {
        TimedLaunch tl_(KC_EL_FWD_STRESS, st);
        if (save) ADFWI_LAUNCH(ADFWI_KERNEL(el_fwd_stress<NN, PML, FS, true>), grd, blk, st, P.g, sh, a.C11, a.C13, a.C33, a.C55, a.bcx, a.bcz,
                               P.st, a.mt, a.src_v, a.sx, a.sz, P.hist, P.K, tl, it);
        else      ADFWI_LAUNCH(ADFWI_KERNEL(el_fwd_stress<NN, PML, FS, false>), grd, blk, st, P.g, sh, a.C11, a.C13, a.C33, a.C55, a.bcx, a.bcz,
                               P.st, a.mt, a.src_v, a.sx, a.sz, P.hist, P.K, tl, it);
    }
    ADFWI_LAUNCH_CHECK();
    {
        TimedLaunch tl_(KC_EL_FWD_VEL, st);
        ADFWI_LAUNCH(ADFWI_KERNEL(el_fwd_vel<NN, PML, FS>), grd, blk, st, P.g, sh, a.bx, a.bz, a.bcx, a.bcz, P.st);
    }
    ADFWI_LAUNCH_CHECK();
    if (FS || (rcv && P.nr > 0)) {
        const int nr = rcv ? P.nr : 0;
        const int width = P.g.nxp > nr ? P.g.nxp : nr;
        TimedLaunch tl_(KC_EL_RECORD, st);
        ADFWI_LAUNCH(ADFWI_KERNEL(el_fwd_post<NN, PML, FS>), dim3(cdiv(width, 128), 2, se - sb), 128, st, P.g, sh, a.bcx, P.st, nr, a.rx, a.rz,
                     rcv ? rcv[0] : nullptr, rcv ? rcv[1] : nullptr, rcv ? rcv[2] : nullptr, rcv ? rcv[3] : nullptr,
                     rcv ? rcv[4] : nullptr, it);
        ADFWI_LAUNCH_CHECK();
    }
    return ADFWI_OK;
}

template <int NN, bool PML, bool FS>
static int el_forward_t(const ElPlan& P, cudaStream_t st, const ElArgs& a, float* const* rcv, float* const* illum)
{
    const ElGeom& g = P.g;
    const int nt = g.nt;
    const int csz = cdiv(nt, P.n_segments);
    if (illum) ADFWI_CUDA(cudaMemsetAsync(P.ill, 0, sizeof(float) * 5 * g.plane, st));
    for (int sb = 0; sb < P.ns; sb += P.G) {
        const int se = sb + P.G < P.ns ? sb + P.G : P.ns;
        ADFWI_CUDA(cudaMemsetAsync(P.st + (size_t)sb * F_COUNT * g.plane, 0, (size_t)(se - sb) * F_COUNT * g.plane * sizeof(float), st));
        for (int it = 0; it < nt; ++it) {
            const int seg = it / P.K, tl = it - seg * P.K;
            if (P.save && tl == 0 && seg >= 1 && seg <= P.nseg - 2) {
                int rc = el_copy_state(P, st, sb, se, P.ckpt + (size_t)(seg - 1) * P.nstate * P.ns * g.plane, true);
                if (rc) return rc;
            }
            const bool save = P.save && seg == P.nseg - 1;
            int rc = el_forward_step_t<NN, PML, FS>(P, st, sb, se, it, save, tl, a, rcv);
            if (rc) return rc;
            if (illum && ((it + 1) % csz == 0 || it == nt - 1)) {
                ADFWI_LAUNCH(ADFWI_KERNEL(el_illum_acc), cdiv((int)g.plane, 256), 256, st, g, ElShots{sb, se}, P.st, P.ill);
                ADFWI_LAUNCH_CHECK();
            }
        }
    }
    if (illum) {
        ADFWI_LAUNCH(ADFWI_KERNEL(el_illum_crop), dim3(cdiv(g.nx, 128), g.nz), 128, st, g, P.ill, illum[0], illum[1], illum[2], illum[3], illum[4]);
        ADFWI_LAUNCH_CHECK();
    }
    return ADFWI_OK;
}

template <int NN, bool PML, bool FS>
static int el_backward_t(const ElPlan& P, cudaStream_t st, const ElArgs& a, const float* const* g_rcv,
                         float* const* g_coef, float* g_src)
{
    const ElGeom& g = P.g;
    const int nt = g.nt;
    const dim3 blk(64, 4);
    const bool have_g = P.nr > 0 && (g_rcv[0] || g_rcv[1] || g_rcv[2] || g_rcv[3] || g_rcv[4]);
    ADFWI_CUDA(cudaMemsetAsync(P.gpart, 0, sizeof(float) * (size_t)P.G * 6 * g.plane, st));
    for (int sb = 0; sb < P.ns; sb += P.G) {
        const int se = sb + P.G < P.ns ? sb + P.G : P.ns;
        const ElShots sh{sb, se};
        const dim3 grd(cdiv(g.nxp, blk.x), cdiv(g.nzp, blk.y), se - sb);
        ADFWI_CUDA(cudaMemsetAsync(P.L + (size_t)sb * F_COUNT * g.plane, 0, (size_t)(se - sb) * F_COUNT * g.plane * sizeof(float), st));
        ADFWI_CUDA(cudaMemsetAsync(P.scr + (size_t)sb * 6 * g.plane, 0, (size_t)(se - sb) * 6 * g.plane * sizeof(float), st));
        for (int seg = P.nseg - 1; seg >= 0; --seg) {
            const int t0 = seg * P.K, t1 = t0 + P.K < nt ? t0 + P.K : nt;
            if (seg != P.nseg - 1) {
                if (seg == 0) {
                    ADFWI_CUDA(cudaMemsetAsync(P.st + (size_t)sb * F_COUNT * g.plane, 0, (size_t)(se - sb) * F_COUNT * g.plane * sizeof(float), st));
                } else {
                    int rc = el_copy_state(P, st, sb, se, P.ckpt + (size_t)(seg - 1) * P.nstate * P.ns * g.plane, false);
                    if (rc) return rc;
                }
                for (int it = t0; it < t1; ++it) {
                    int rc = el_forward_step_t<NN, PML, FS>(P, st, sb, se, it, true, it - t0, a, nullptr);
                    if (rc) return rc;
                }
            }
            for (int it = t1 - 1; it >= t0; --it) {
                const int tl = it - t0;
                if (have_g || (FS && PML)) {
                    const int nr = have_g ? P.nr : 0;
                    const int width = g.nxp > nr ? g.nxp : nr;
                    TimedLaunch tl_(KC_EL_ADJ_INJECT, st);
                    ADFWI_LAUNCH(ADFWI_KERNEL(el_adj_pre<NN, PML, FS>), dim3(cdiv(width, 128), 2, se - sb), 128, st, g, sh, P.L, nr, a.rx, a.rz,
                                 g_rcv[0], g_rcv[1], g_rcv[2], g_rcv[3], g_rcv[4], it);
                    ADFWI_LAUNCH_CHECK();
                }
                {
                    TimedLaunch tl_(KC_EL_ADJ_VEL, st);
                    ADFWI_LAUNCH(ADFWI_KERNEL(el_adj_vel<NN, PML, FS>), grd, blk, st, g, sh, a.bx, a.bz, a.bcx, a.bcz, P.L, P.scr, P.hist, P.K, tl, P.gpart);
                    ADFWI_LAUNCH_CHECK();
                    ADFWI_LAUNCH(ADFWI_KERNEL(el_adj_vel_g<NN>), grd, blk, st, g, sh, P.L, P.scr);
                    ADFWI_LAUNCH_CHECK();
                }
                {
                    TimedLaunch tl_(KC_EL_ADJ_STRESS, st);
                    ADFWI_LAUNCH(ADFWI_KERNEL(el_adj_stress<NN, PML, FS>), grd, blk, st, g, sh, a.C11, a.C13, a.C33, a.C55, a.bcx, a.bcz, P.L, P.scr,
                                 P.hist, P.K, tl, P.gpart, a.mt, a.sx, a.sz, g_src, it);
                    ADFWI_LAUNCH_CHECK();
                    ADFWI_LAUNCH(ADFWI_KERNEL(el_adj_stress_g<NN, PML, FS>), grd, blk, st, g, sh, P.L, P.scr);
                    ADFWI_LAUNCH_CHECK();
                }
            }
        }
    }
    for (int k = 0; k < 6; ++k) {
        ADFWI_LAUNCH(ADFWI_KERNEL(el_reduce_parts), cdiv((int)g.plane, 256), 256, st, g.plane, P.G, k, P.gpart, g_coef[k]);
        ADFWI_LAUNCH_CHECK();
    }
    return ADFWI_OK;
}

#define EL_DISPATCH(fn, ...)                                                            \
    (P.NN == 2 ? (P.PML ? (P.FS ? fn<2, true, true>(__VA_ARGS__) : fn<2, true, false>(__VA_ARGS__))   \
                        : (P.FS ? fn<2, false, true>(__VA_ARGS__) : fn<2, false, false>(__VA_ARGS__))) \
               : (P.PML ? (P.FS ? fn<3, true, true>(__VA_ARGS__) : fn<3, true, false>(__VA_ARGS__))   \
                        : (P.FS ? fn<3, false, true>(__VA_ARGS__) : fn<3, false, false>(__VA_ARGS__))))

}  // namespace adfwi

using namespace adfwi;

// The TMA-staged pipelines (elastic_fused.cu: split-field PML; elastic_abl_fused.inl: sponge / ABL) are the
// default; the generic kernels of this file run when the caller sets bit 0 of desc->reserved[0] (used by
// the tests to cross-check the pipelines) or for grids beyond the fused paths' index packing.
static bool el_use_fused(const adfwi_elastic_desc* d)
{
#ifdef ADFWI_HOST_EMUL
    (void)d; return false;
#else
    return elf_supported(d) || ela_supported(d);
#endif
}
#ifndef ADFWI_HOST_EMUL
static size_t el_fused_bytes(const adfwi_elastic_desc* d) { return d->abc_pml ? elf_workspace_bytes(d) : ela_workspace_bytes(d); }
#endif

extern "C" size_t adfwi_elastic_workspace_bytes(const adfwi_elastic_desc* desc)
{
    ElPlan P;
    if (el_make_plan(desc, nullptr, &P) != ADFWI_OK) return 0;
#ifndef ADFWI_HOST_EMUL
    if (el_use_fused(desc)) return el_fused_bytes(desc);
#endif
    return P.bytes;
}

static int el_check_args(const ElPlan& P, const float* const* coef, const float* bcx, const float* bcz, const float* mt,
                         const float* src_v, const int64_t* sx, const int64_t* sz, const int64_t* rx, const int64_t* rz,
                         void* ws, size_t wsb, ElArgs* a)
{
    if (!coef || !bcx || !mt || !src_v || !sx || !sz || !ws) return ADFWI_E_NULL;
    for (int k = 0; k < 6; ++k) if (!coef[k]) return ADFWI_E_NULL;
    if (P.PML && !bcz) return ADFWI_E_NULL;
    if (P.nr > 0 && (!rx || !rz)) return ADFWI_E_NULL;
    if (wsb < P.bytes) return ADFWI_E_WORKSPACE;
    a->C11 = coef[0]; a->C13 = coef[1]; a->C33 = coef[2]; a->C55 = coef[3]; a->bx = coef[4]; a->bz = coef[5];
    a->bcx = bcx; a->bcz = bcz; a->mt = mt; a->src_v = src_v; a->sx = sx; a->sz = sz; a->rx = rx; a->rz = rz;
    return ADFWI_OK;
}

extern "C" int adfwi_elastic_forward(const adfwi_elastic_desc* desc, const float* const* coef,
                                     const float* bcx, const float* bcz, const float* mt,
                                     const float* src_v, const int64_t* src_x, const int64_t* src_z,
                                     const int64_t* rcv_x, const int64_t* rcv_z,
                                     float* const* rcv, float* const* illum,
                                     void* workspace, size_t workspace_bytes, void* stream)
{
    ADFWI_NVTX("adfwi_elastic_forward");
    ElPlan P;
    int rc = el_make_plan(desc, workspace, &P);
    if (rc) return rc;
    ElArgs a;
    rc = el_check_args(P, coef, bcx, bcz, mt, src_v, src_x, src_z, rcv_x, rcv_z, workspace, el_use_fused(desc) ? (size_t)-1 : workspace_bytes, &a);
    if (rc) return rc;
    if (P.nr > 0) { if (!rcv) return ADFWI_E_NULL; for (int k = 0; k < 5; ++k) if (!rcv[k]) return ADFWI_E_NULL; }
    cudaStream_t st = (cudaStream_t)stream;
#ifndef ADFWI_HOST_EMUL
    if (el_use_fused(desc)) {
        if (workspace_bytes < el_fused_bytes(desc)) return ADFWI_E_WORKSPACE;
        if (!desc->abc_pml) return ela_forward(desc, coef, bcx, mt, src_v, src_x, src_z, rcv_x, rcv_z, rcv, illum, workspace, st);
        return elf_forward(desc, coef, bcx, bcz, mt, src_v, src_x, src_z, rcv_x, rcv_z, rcv, illum, workspace, st);
    }
#endif
    return EL_DISPATCH(el_forward_t, P, st, a, rcv, illum);
}

extern "C" int adfwi_elastic_backward(const adfwi_elastic_desc* desc, const float* const* coef,
                                      const float* bcx, const float* bcz, const float* mt,
                                      const float* src_v, const int64_t* src_x, const int64_t* src_z,
                                      const int64_t* rcv_x, const int64_t* rcv_z,
                                      const float* const* g_rcv, float* const* g_coef, float* g_src_v,
                                      void* workspace, size_t workspace_bytes, void* stream)
{
    ADFWI_NVTX("adfwi_elastic_backward");
    ElPlan P;
    int rc = el_make_plan(desc, workspace, &P);
    if (rc) return rc;
    if (!P.save) return ADFWI_E_MODE;
    ElArgs a;
    rc = el_check_args(P, coef, bcx, bcz, mt, src_v, src_x, src_z, rcv_x, rcv_z, workspace, el_use_fused(desc) ? (size_t)-1 : workspace_bytes, &a);
    if (rc) return rc;
    if (!g_rcv || !g_coef) return ADFWI_E_NULL;
    for (int k = 0; k < 6; ++k) if (!g_coef[k]) return ADFWI_E_NULL;
    cudaStream_t st = (cudaStream_t)stream;
#ifndef ADFWI_HOST_EMUL
    if (el_use_fused(desc)) {
        if (workspace_bytes < el_fused_bytes(desc)) return ADFWI_E_WORKSPACE;
        if (!desc->abc_pml) return ela_backward(desc, mt, src_v, src_x, src_z, g_rcv, g_coef, g_src_v, workspace, st);
        return elf_backward(desc, coef, bcx, bcz, mt, src_v, src_x, src_z, rcv_x, rcv_z, g_rcv, g_coef, g_src_v, workspace, st);
    }
#endif
    return EL_DISPATCH(el_backward_t, P, st, a, g_rcv, g_coef, g_src_v);
}
