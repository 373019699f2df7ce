// acoustic_fused.cu -- fused, TMA-staged iso-acoustic time step for sm_100a (the fast path).
//
// One launch per time step per shot group:
//   ac_fwd_fused : P update + source + free-surface mirror + U/W update + free surface + receiver
//                  sampling + illumination + stencil-history store   (acoustic_kernels.py:115-174)
//   ac_adj_fused : receiver-cotangent injection + the whole reverse step 7T..1T of SURVEY.md
//                  Appendix A.1 + g_alpha1 accumulation + g_src
// Each CTA owns 64x32 tiles; the old p,u,w (or lambda) tiles with their halos (4 cells in x, 3 in z)
// are brought into shared memory by TMA (cp.async.bulk.tensor.3d, hardware zero fill outside the
// grid), the intermediate pressure (forward) / pressure cotangent (adjoint) lives only in shared
// memory (halo recompute), and the new fields go to the other buffer of a ping-pong pair.
// Same arithmetic and association as the generic kernels in acoustic.cu (-fmad=false): forward
// records stay bit-identical to the CPU reference.
//
// Used when the density gradient is not requested; otherwise acoustic.cu's generic kernels run.
#include "common.cuh"
#ifndef ADFWI_HOST_EMUL
#include "tma.cuh"
#include "acoustic_fused.h"

namespace adfwi {

namespace {

constexpr int TX = 64, TZ = 32;             // tile interior
constexpr int HX = 4, HZ = 3;               // halo of the staged rectangle
constexpr int RX = TX + 2 * HX;             // 72 floats per staged row (16-B multiple)
constexpr int RZ = TZ + 2 * HZ;             // 38 staged rows
constexpr int QZ = TZ + 3, QX = TX + 3;     // region of the intermediate field: rows/cols [-1, T+2)
constexpr int QS = QX + 1;                  // its shared-memory row stride (68)
constexpr int NTHREADS = 256;
constexpr int RECT_BYTES = ((RZ * RX * 4 + 127) / 128) * 128;   // 11008
constexpr int REG_BYTES = ((QZ * QS * 4 + 127) / 128) * 128;    // 9600

struct FGeom {
    int nzp, nxp, ld, fs, zlo, nabc, nt;
    int ntx, ntz;
    size_t plane;            // nzp*ld
    float c1, c2, dt;
};

struct FwdArgs {
    const float *a1, *k1, *a2, *k2, *k3;       // pitched coefficient planes
    float *p_out, *u_out, *w_out;
    const float* src_v; const int64_t *sx, *sz;
    float* hist; int hist_len, tl, it;
    int nr; const int64_t *rx, *rz; const int* rzrange;
    float *rcv_p, *rcv_u, *rcv_w;
    float *ill_p, *ill_u; int acc_u;
    int s_begin, s_end;
};

struct AdjArgs {
    const float *a1, *k1, *k2, *k3;
    float *lp_out, *lu_out, *lw_out;
    const int64_t *sx, *sz;
    const float* hist; int hist_len, tl, it;
    int nr; const int64_t *rx, *rz; const int* rzrange;
    const float *gp, *gu, *gw;
    float* g1part; float* g_src;
    int s_begin, s_end;
};

__device__ __forceinline__ bool mbar_try(uint64_t* bar, uint32_t parity)
{
    uint32_t ok;
    asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
                 : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return ok != 0;
}

// ------------------------------------------------------------------------------------------
template <bool FS, bool SAVE, bool ILLUM>
__global__ void __launch_bounds__(NTHREADS, 3)
ac_fwd_fused(const __grid_constant__ CUtensorMap tm_p, const __grid_constant__ CUtensorMap tm_u,
             const __grid_constant__ CUtensorMap tm_w, const FGeom g, const FwdArgs a)
{
    extern __shared__ __align__(128) unsigned char smem_raw[];
    float* ps = (float*)smem_raw;
    float* us = (float*)(smem_raw + RECT_BYTES);
    float* ws = (float*)(smem_raw + 2 * RECT_BYTES);
    float* pn = (float*)(smem_raw + 3 * RECT_BYTES);
    uint64_t* bar = (uint64_t*)(smem_raw + 3 * RECT_BYTES + REG_BYTES);
    const int tid = threadIdx.x;
    if (tid == 0) mbar_init(bar, 1);
    __syncthreads();
    const int nsh = a.s_end - a.s_begin;
    const int nitems = g.ntx * g.ntz * nsh;
    const int rzmin = a.nr > 0 ? a.rzrange[0] : 1 << 30, rzmax = a.nr > 0 ? a.rzrange[1] : -1;
    uint32_t parity = 0;
    const int ld = g.ld;
    for (int item = blockIdx.x; item < nitems; item += gridDim.x) {
        const int sl = item % nsh, tile = item / nsh;
        const int s = a.s_begin + sl;
        const int tzi = tile / g.ntx, txi = tile - tzi * g.ntx;
        const int X0 = txi * TX, Z0 = g.zlo + tzi * TZ;
        if (tid == 0) {
            fence_proxy_async();
            mbar_expect_tx(bar, 3 * RZ * RX * 4);
            tma_load_3d(ps, &tm_p, X0 - HX, Z0 - HZ, s, bar);
            tma_load_3d(us, &tm_u, X0 - HX, Z0 - HZ, s, bar);
            tma_load_3d(ws, &tm_w, X0 - HX, Z0 - HZ, s, bar);
        }
        const int szs = (int)a.sz[s], sxs = (int)a.sx[s];
        const float srcval = g.dt * a.src_v[(size_t)s * g.nt + a.it];
        float* Hs = SAVE ? a.hist + ((size_t)s * a.hist_len + a.tl) * g.plane : nullptr;
        while (!mbar_try(bar, parity)) {}
        parity ^= 1;
        // ---- phase 1: new pressure on the region rows [-1,TZ+2) x cols [-1,TX+2) -----------------
        for (int idx = tid; idx < QZ * QX; idx += NTHREADS) {
            const int r = idx / QX, c = idx - r * QX;
            const int gz = Z0 - 1 + r, gx = X0 - 1 + c;
            const int sr = r + 2, sc = c + 3;
            float pv = ps[sr * RX + sc];
            const bool inreg = (gz >= g.fs + 1) && (gz < g.nzp - 2) && (gx >= 2) && (gx < g.nxp - 2);
            if (inreg) {
                const size_t cg = (size_t)gz * ld + gx;
                const float t1 = 1.0f - a.k1[cg], al = a.a1[cg];
                const float* uu = us + sr * RX + sc;
                const float* ww = ws + sr * RX + sc;
                const float S1 = ((uu[0] - uu[-1]) + ww[0]) - ww[-RX];
                const float S2 = ((uu[1] - uu[-2]) + ww[RX]) - ww[-2 * RX];
                const float S = g.c1 * S1 + g.c2 * S2;
                if (SAVE && r >= 1 && r <= TZ && c >= 1 && c <= TX) __stcs(Hs + cg, S);
                pv = t1 * pv - al * S;
            }
            if (gz == szs && gx == sxs) pv = pv + srcval;
            pn[r * QS + c] = pv;
        }
        __syncthreads();
        if (FS && tzi == 0) {        // p[fs-1] = -p[fs+1] : region rows 1 and 3 of the first tile row
            if (tid < QX) pn[1 * QS + tid] = -pn[3 * QS + tid];
            __syncthreads();
        }
        // ---- phase 2: velocities on the interior, stores, illumination -----------------------------
        const int tx = tid & (TX - 1), ty = tid >> 6;
        const int gx = X0 + tx;
#pragma unroll 2
        for (int jj = 0; jj < TZ / 4; ++jj) {
            const int rr = ty + 4 * jj;
            const int gz = Z0 + rr;
            if (gz < g.nzp && gx < g.nxp) {
                const int r = rr + 1, c = tx + 1, sr = rr + HZ, sc = tx + HX;
                const size_t cg = (size_t)gz * ld + gx;
                const size_t o = (size_t)s * g.plane + cg;
                const float p0 = pn[r * QS + c];
                float uv = us[sr * RX + sc], wv = ws[sr * RX + sc];
                const bool inU = (gz >= g.fs) && (gz < g.nzp - 1) && (gx >= 1) && (gx < g.nxp - 2);
                const bool inW = (gz >= g.fs) && (gz < g.nzp - 2) && (gx >= 1) && (gx < g.nxp - 1);
                if (inU || inW) {
                    const float al = a.a2[cg];
                    if (inU) {
                        const float t2 = 1.0f - a.k2[cg];
                        uv = t2 * uv - al * (g.c1 * (pn[r * QS + c + 1] - p0) + g.c2 * (pn[r * QS + c + 2] - pn[r * QS + c - 1]));
                    }
                    if (inW) {
                        const float t3 = 1.0f - a.k3[cg];
                        wv = t3 * wv - al * (g.c1 * (pn[(r + 1) * QS + c] - p0) + g.c2 * (pn[(r + 2) * QS + c] - pn[(r - 1) * QS + c]));
                    }
                }
                a.p_out[o] = p0;
                a.u_out[o] = uv;
                if (!(FS && gz == g.fs - 1)) a.w_out[o] = wv;
                if (FS && gz == g.fs) a.w_out[o - ld] = wv;                 // w[fs-1] = w[fs]
                if (a.nr > 0) { us[sr * RX + sc] = uv; ws[sr * RX + sc] = (FS && gz == g.fs - 1) ? 0.f : wv; }
                if (ILLUM && gz >= g.nabc && gz < g.nzp - g.nabc && gx >= g.nabc && gx < g.nxp - g.nabc) {
                    atomicAdd(a.ill_p + cg, p0 * p0);
                    if (a.acc_u) atomicAdd(a.ill_u + cg, uv * uv);
                }
            }
        }
        __syncthreads();
        // ---- receivers inside this tile's interior (acoustic_kernels.py:167-169) -------------------
        if (a.nr > 0 && rzmax >= Z0 && rzmin < Z0 + TZ) {
            for (int r = tid; r < a.nr; r += NTHREADS) {
                const int z = (int)a.rz[r] - Z0, x = (int)a.rx[r] - X0;
                if (z >= 0 && z < TZ && x >= 0 && x < TX) {
                    const size_t o = ((size_t)s * g.nt + a.it) * a.nr + r;
                    a.rcv_p[o] = pn[(z + 1) * QS + x + 1];
                    if (a.rcv_u) a.rcv_u[o] = us[(z + HZ) * RX + x + HX];
                    if (a.rcv_w) a.rcv_w[o] = ws[(z + HZ) * RX + x + HX];
                }
            }
            __syncthreads();
        }
    }
}

// ------------------------------------------------------------------------------------------
template <bool FS>
__global__ void __launch_bounds__(NTHREADS, 3)
ac_adj_fused(const __grid_constant__ CUtensorMap tm_lp, const __grid_constant__ CUtensorMap tm_lu,
             const __grid_constant__ CUtensorMap tm_lw, const __grid_constant__ CUtensorMap tm_a2,
             const FGeom g, const AdjArgs a)
{
    extern __shared__ __align__(128) unsigned char smem_raw[];
    float* lps = (float*)smem_raw;
    float* lus = (float*)(smem_raw + RECT_BYTES);
    float* lws = (float*)(smem_raw + 2 * RECT_BYTES);
    float* a2s = (float*)(smem_raw + 3 * RECT_BYTES);
    float* lp2 = (float*)(smem_raw + 4 * RECT_BYTES);
    float* mps = (float*)(smem_raw + 4 * RECT_BYTES + REG_BYTES);
    uint64_t* bar = (uint64_t*)(smem_raw + 4 * RECT_BYTES + 2 * REG_BYTES);
    const int tid = threadIdx.x;
    if (tid == 0) mbar_init(bar, 1);
    __syncthreads();
    const int nsh = a.s_end - a.s_begin;
    const int nitems = g.ntx * g.ntz * nsh;
    const bool have_g = a.nr > 0 && (a.gp || a.gu || a.gw);
    const int rzmin = have_g ? a.rzrange[0] : 1 << 30, rzmax = have_g ? a.rzrange[1] : -1;
    uint32_t parity = 0;
    const int ld = g.ld, fs = g.fs, nzp = g.nzp, nxp = g.nxp;
    for (int item = blockIdx.x; item < nitems; item += gridDim.x) {
        const int sl = item % nsh, tile = item / nsh;
        const int s = a.s_begin + sl;
        const int tzi = tile / g.ntx, txi = tile - tzi * g.ntx;
        const int X0 = txi * TX, Z0 = g.zlo + tzi * TZ;
        if (tid == 0) {
            fence_proxy_async();
            mbar_expect_tx(bar, 4 * RZ * RX * 4);
            tma_load_3d(lps, &tm_lp, X0 - HX, Z0 - HZ, s, bar);
            tma_load_3d(lus, &tm_lu, X0 - HX, Z0 - HZ, s, bar);
            tma_load_3d(lws, &tm_lw, X0 - HX, Z0 - HZ, s, bar);
            tma_load_2d(a2s, &tm_a2, X0 - HX, Z0 - HZ, bar);
        }
        const int szs = (int)a.sz[s], sxs = (int)a.sx[s];
        const float* Hs = a.hist + ((size_t)s * a.hist_len + a.tl) * g.plane;
        while (!mbar_try(bar, parity)) {}
        parity ^= 1;
        // ---- 7T: receiver cotangents into the staged rectangle (duplicates legal -> shared atomics)
        if (have_g && rzmax >= Z0 - HZ && rzmin < Z0 + TZ + HZ) {
            for (int r = tid; r < a.nr; r += NTHREADS) {
                const int z = (int)a.rz[r] - (Z0 - HZ), x = (int)a.rx[r] - (X0 - HX);
                if (z >= 0 && z < RZ && x >= 0 && x < RX) {
                    const size_t o = ((size_t)s * g.nt + a.it) * a.nr + r;
                    if (a.gp) atomicAdd(lps + z * RX + x, a.gp[o]);
                    if (a.gu) atomicAdd(lus + z * RX + x, a.gu[o]);
                    if (a.gw) atomicAdd(lws + z * RX + x, a.gw[o]);
                }
            }
            __syncthreads();
        }
        if (FS && tzi == 0) {        // 6T: lambda_w[fs] += lambda_w[fs-1]; lambda_w[fs-1] = 0  (staged rows 4 and 3)
            if (tid < RX) { lws[(HZ + 1) * RX + tid] += lws[HZ * RX + tid]; lws[HZ * RX + tid] = 0.f; }
            __syncthreads();
        }
        // ---- phase 1: lambda_p after undoing W and U (5T, 4T) on the region -------------------------
        for (int idx = tid; idx < QZ * QX; idx += NTHREADS) {
            const int r = idx / QX, c = idx - r * QX;
            const int gz = Z0 - 1 + r, gx = X0 - 1 + c;
            const int sr = r + 2, sc = c + 3;
            float acc = lps[sr * RX + sc];
            if (gz >= 0 && gz < nzp && gx >= 0 && gx < nxp) {
                float qw[4], qu[4];
                const bool xw = (gx >= 1) && (gx < nxp - 1);
                const bool zu = (gz >= fs) && (gz < nzp - 1);
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const int zz = gz - 2 + k, xx = gx - 2 + k;
                    const int iw = (sr - 2 + k) * RX + sc, iu = sr * RX + sc - 2 + k;
                    qw[k] = (xw && zz >= fs && zz < nzp - 2) ? (-a2s[iw]) * lws[iw] : 0.f;
                    qu[k] = (zu && xx >= 1 && xx < nxp - 2) ? (-a2s[iu]) * lus[iu] : 0.f;
                }
                acc += g.c1 * qw[1] - g.c1 * qw[2] + g.c2 * qw[0] - g.c2 * qw[3];
                acc += g.c1 * qu[1] - g.c1 * qu[2] + g.c2 * qu[0] - g.c2 * qu[3];
            }
            lp2[r * QS + c] = acc;
        }
        __syncthreads();
        if (FS && tzi == 0) {        // 3T: lambda_p[fs+1] -= lambda_p[fs-1]; lambda_p[fs-1] = 0 (region rows 3 and 1)
            if (tid < QX) { lp2[3 * QS + tid] -= lp2[1 * QS + tid]; lp2[1 * QS + tid] = 0.f; }
            __syncthreads();
        }
        // ---- phase 1b: m_p = -alpha1 * lambda_p on the P cells of the region ------------------------
        for (int idx = tid; idx < QZ * QX; idx += NTHREADS) {
            const int r = idx / QX, c = idx - r * QX;
            const int gz = Z0 - 1 + r, gx = X0 - 1 + c;
            const bool inP = (gz >= fs + 1) && (gz < nzp - 2) && (gx >= 2) && (gx < nxp - 2);
            mps[r * QS + c] = inP ? (-a.a1[(size_t)gz * ld + gx]) * lp2[r * QS + c] : 0.f;
        }
        __syncthreads();
        // ---- phase 2: new lambda_u, lambda_w, lambda_p on the interior (5T,4T,1T), g_alpha1, g_src ---
        const int tx = tid & (TX - 1), ty = tid >> 6;
        const int gx = X0 + tx;
#pragma unroll 2
        for (int jj = 0; jj < TZ / 4; ++jj) {
            const int rr = ty + 4 * jj;
            const int gz = Z0 + rr;
            if (gz < nzp && gx < nxp) {
                const int r = rr + 1, c = tx + 1, sr = rr + HZ, sc = tx + HX;
                const size_t cg = (size_t)gz * ld + gx;
                const size_t o = (size_t)s * g.plane + cg;
                const bool inP = (gz >= fs + 1) && (gz < nzp - 2) && (gx >= 2) && (gx < nxp - 2);
                const bool inU = (gz >= fs) && (gz < nzp - 1) && (gx >= 1) && (gx < nxp - 2);
                const bool inW = (gz >= fs) && (gz < nzp - 2) && (gx >= 1) && (gx < nxp - 1);
                const float t1 = inP ? 1.0f - a.k1[cg] : 1.0f;
                const float t2 = inU ? 1.0f - a.k2[cg] : 1.0f;
                const float t3 = inW ? 1.0f - a.k3[cg] : 1.0f;
                const float* m = mps + r * QS + c;
                const float qp = lp2[r * QS + c];
                const float du = g.c1 * m[0] - g.c1 * m[1] + g.c2 * m[-1] - g.c2 * m[2];
                const float dw = g.c1 * m[0] - g.c1 * m[QS] + g.c2 * m[-QS] - g.c2 * m[2 * QS];
                a.lu_out[o] = t2 * lus[sr * RX + sc] + du;
                a.lw_out[o] = t3 * lws[sr * RX + sc] + dw;
                a.lp_out[o] = t1 * qp;
                if (inP) {
                    float* gp1 = a.g1part + (size_t)sl * g.plane + cg;
                    *gp1 = *gp1 - qp * __ldcs(Hs + cg);
                }
                if (a.g_src && gz == szs && gx == sxs) a.g_src[(size_t)s * g.nt + a.it] = g.dt * qp;
            }
        }
        __syncthreads();
    }
}

__global__ void acf_rz_range(int nr, const int64_t* __restrict__ rz, int* __restrict__ out)
{
    __shared__ int smin[256], smax[256];
    int mn = 1 << 30, mx = -1;
    for (int r = threadIdx.x; r < nr; r += blockDim.x) { const int z = (int)rz[r]; mn = min(mn, z); mx = max(mx, z); }
    smin[threadIdx.x] = mn; smax[threadIdx.x] = mx;
    __syncthreads();
    for (int k = 128; k > 0; k >>= 1) {
        if ((int)threadIdx.x < k) { smin[threadIdx.x] = min(smin[threadIdx.x], smin[threadIdx.x + k]); smax[threadIdx.x] = max(smax[threadIdx.x], smax[threadIdx.x + k]); }
        __syncthreads();
    }
    if (threadIdx.x == 0) { out[0] = smin[0]; out[1] = smax[0]; }
}

// pitched partial planes -> dense caller plane
__global__ void acf_reduce_parts(int nzp, int nxp, int ld, int nparts, const float* __restrict__ part, float* __restrict__ out)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x, z = blockIdx.y;
    if (x >= nxp || z >= nzp) return;
    float acc = 0.f;
    for (int k = 0; k < nparts; ++k) acc += part[((size_t)k * nzp + z) * ld + x];
    out[(size_t)z * nxp + x] = acc;
}

__global__ void acf_sumsq(size_t n, size_t plane, int s_begin, int s_end, const float* __restrict__ f, float* __restrict__ out)
{
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float acc = 0.f;
    for (int s = s_begin; s < s_end; ++s) { const float v = f[(size_t)s * plane + i]; acc += v * v; }
    out[i] += acc;
}

__global__ void acf_illum_finalize(int nzp, int nxp, int ld, int nabc, const float* __restrict__ ip, const float* __restrict__ iu,
                                   const float* __restrict__ iw, float* op, float* ou, float* ow)
{
    const int nx = nxp - 2 * nabc, nz = nzp - 2 * nabc;
    const int x = blockIdx.x * blockDim.x + threadIdx.x, z = blockIdx.y;
    if (x >= nx || z >= nz) return;
    const size_t c = (size_t)(z + nabc) * ld + (x + nabc), o = (size_t)z * nx + x;
    const float p = ip[c], u = iu[c];
    if (op) op[o] = p;
    if (ou) ou[o] = p + u;
    if (ow) ow[o] = p + (u + iw[c]);
}

struct FPlan {
    FGeom g;
    int ns, nr, FS, save, n_segments;
    int K, nseg, nckpt, G;
    float *coef[5];                 // a1,k1,a2,k2,k3 pitched
    float *st[2][3];                // p,u,w ping-pong
    float *lam[2][3];               // lambda ping-pong
    float *hist, *ckpt, *g1part, *ill_p, *ill_u, *ill_w;
    int* rzrange;
    size_t bytes;
};

int acf_make_plan(const adfwi_acoustic_desc* d, void* ws, FPlan* P)
{
    FGeom& g = P->g;
    g.nzp = d->nzp; g.nxp = d->nxp; g.ld = (d->nxp + 3) / 4 * 4; g.nabc = d->nabc; g.nt = d->nt;
    g.fs = d->free_surface ? d->nabc : 1;
    g.zlo = g.fs - 1;
    g.ntx = cdiv(g.nxp, TX); g.ntz = cdiv(g.nzp - g.zlo, TZ);
    g.plane = (size_t)g.nzp * g.ld;
    g.c1 = d->c1; g.c2 = d->c2; g.dt = d->dt;
    P->ns = d->ns; P->nr = d->nr; P->FS = d->free_surface ? 1 : 0;
    P->save = d->save_history ? 1 : 0;
    P->n_segments = d->n_segments > 0 ? d->n_segments : 1;
    int K = d->ckpt_interval;
    if (K <= 0 || K >= d->nt) K = d->nt;
    P->K = K; P->nseg = cdiv(d->nt, K); P->nckpt = P->nseg > 2 ? P->nseg - 2 : 0;
    int G = d->shots_per_group;
    if (G <= 0) {
        const size_t per_shot = g.plane * sizeof(float) * 6;     // two buffers of three fields
        G = (int)((size_t)(72u << 20) / per_shot);
        if (G < 1) G = 1;
    }
    if (G > d->ns) G = d->ns;
    P->G = G;
    Carver cv(ws);
    const size_t sp = (size_t)d->ns * g.plane;
    for (int k = 0; k < 5; ++k) P->coef[k] = cv.take<float>(g.plane);
    for (int b = 0; b < 2; ++b) for (int f = 0; f < 3; ++f) P->st[b][f] = cv.take<float>(sp);
    P->ill_p = cv.take<float>(g.plane); P->ill_u = cv.take<float>(g.plane); P->ill_w = cv.take<float>(g.plane);
    P->rzrange = cv.take<int>(64);
    P->hist = P->ckpt = P->g1part = nullptr;
    for (int b = 0; b < 2; ++b) for (int f = 0; f < 3; ++f) P->lam[b][f] = nullptr;
    if (P->save) {
        for (int b = 0; b < 2; ++b) for (int f = 0; f < 3; ++f) P->lam[b][f] = cv.take<float>(sp);
        P->g1part = cv.take<float>((size_t)G * g.plane);
        if (P->nckpt) P->ckpt = cv.take<float>((size_t)P->nckpt * 3 * sp);
        P->hist = cv.take<float>((size_t)K * sp);
    }
    P->bytes = cv.off;
    return ADFWI_OK;
}

constexpr int FWD_SMEM = 3 * RECT_BYTES + REG_BYTES + 64;
constexpr int ADJ_SMEM = 4 * RECT_BYTES + 2 * REG_BYTES + 64;

int acf_num_sms()
{
    static int n = 0;
    if (!n) { int dev = 0; cudaGetDevice(&dev); cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev); if (n <= 0) n = 148; }
    return n;
}

template <typename K> int acf_set_smem(K kern, int bytes)
{
    return (int)cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
}

int acf_copy_coefs(const FPlan& P, cudaStream_t st, const float* const* src)
{
    const FGeom& g = P.g;
    for (int k = 0; k < 5; ++k) {
        if (g.ld != g.nxp) ADFWI_CUDA(cudaMemsetAsync(P.coef[k], 0, g.plane * sizeof(float), st));
        ADFWI_CUDA(cudaMemcpy2DAsync(P.coef[k], (size_t)g.ld * 4, src[k], (size_t)g.nxp * 4, (size_t)g.nxp * 4, g.nzp, cudaMemcpyDeviceToDevice, st));
    }
    return ADFWI_OK;
}

struct StepMaps { CUtensorMap st[2][3]; CUtensorMap lam[2][3]; CUtensorMap a2; };

int acf_make_maps(const FPlan& P, StepMaps* M)
{
    const FGeom& g = P.g;
    for (int b = 0; b < 2; ++b)
        for (int f = 0; f < 3; ++f) {
            int rc = make_tmap_f32(&M->st[b][f], P.st[b][f], 3, g.nxp, g.ld, g.nzp, P.ns, RX, RZ);
            if (rc) return rc;
            if (P.save) { rc = make_tmap_f32(&M->lam[b][f], P.lam[b][f], 3, g.nxp, g.ld, g.nzp, P.ns, RX, RZ); if (rc) return rc; }
        }
    return make_tmap_f32(&M->a2, P.coef[2], 2, g.nxp, g.ld, g.nzp, 1, RX, RZ);
}

// one fused forward step of shots [sb,se): reads buffer cur, writes buffer cur^1
int acf_forward_step(const FPlan& P, const StepMaps& M, cudaStream_t st, int cur, int sb, int se, int it, bool save, int tl,
                     const float* src_v, const int64_t* sx, const int64_t* sz, const int64_t* rx, const int64_t* rz,
                     float* rcv_p, float* rcv_u, float* rcv_w, bool illum, int acc_u)
{
    const FGeom& g = P.g;
    FwdArgs a;
    a.a1 = P.coef[0]; a.k1 = P.coef[1]; a.a2 = P.coef[2]; a.k2 = P.coef[3]; a.k3 = P.coef[4];
    a.p_out = P.st[cur ^ 1][0]; a.u_out = P.st[cur ^ 1][1]; a.w_out = P.st[cur ^ 1][2];
    a.src_v = src_v; a.sx = sx; a.sz = sz;
    a.hist = P.hist; a.hist_len = P.K; a.tl = tl; a.it = it;
    a.nr = rcv_p ? P.nr : 0; a.rx = rx; a.rz = rz; a.rzrange = P.rzrange;
    a.rcv_p = rcv_p; a.rcv_u = rcv_u; a.rcv_w = rcv_w;
    a.ill_p = P.ill_p; a.ill_u = P.ill_u; a.acc_u = acc_u;
    a.s_begin = sb; a.s_end = se;
    const int nitems = g.ntx * g.ntz * (se - sb);
    const int grid = nitems < 3 * acf_num_sms() ? nitems : 3 * acf_num_sms();
    TimedLaunch tl_(KC_AC_FWD_FUSED, st);
#define LF(FSv, SVv, ILv) ac_fwd_fused<FSv, SVv, ILv><<<grid, NTHREADS, FWD_SMEM, st>>>(M.st[cur][0], M.st[cur][1], M.st[cur][2], g, a)
    if (P.FS) { if (save) { if (illum) LF(true, true, true); else LF(true, true, false); } else { if (illum) LF(true, false, true); else LF(true, false, false); } }
    else      { if (save) { if (illum) LF(false, true, true); else LF(false, true, false); } else { if (illum) LF(false, false, true); else LF(false, false, false); } }
#undef LF
    ADFWI_LAUNCH_CHECK();
    return ADFWI_OK;
}

int acf_init_kernels()
{
    static bool done = false;
    if (done) return 0;
    int rc = 0;
    rc |= acf_set_smem(ac_fwd_fused<true, true, true>, FWD_SMEM);   rc |= acf_set_smem(ac_fwd_fused<true, true, false>, FWD_SMEM);
    rc |= acf_set_smem(ac_fwd_fused<true, false, true>, FWD_SMEM);  rc |= acf_set_smem(ac_fwd_fused<true, false, false>, FWD_SMEM);
    rc |= acf_set_smem(ac_fwd_fused<false, true, true>, FWD_SMEM);  rc |= acf_set_smem(ac_fwd_fused<false, true, false>, FWD_SMEM);
    rc |= acf_set_smem(ac_fwd_fused<false, false, true>, FWD_SMEM); rc |= acf_set_smem(ac_fwd_fused<false, false, false>, FWD_SMEM);
    rc |= acf_set_smem(ac_adj_fused<true>, ADJ_SMEM);               rc |= acf_set_smem(ac_adj_fused<false>, ADJ_SMEM);
    if (!rc) done = true;
    return rc;
}

}  // namespace

size_t acf_workspace_bytes(const adfwi_acoustic_desc* d)
{
    FPlan P;
    acf_make_plan(d, nullptr, &P);
    return P.bytes;
}

int acf_group_size(const adfwi_acoustic_desc* d)
{
    FPlan P;
    acf_make_plan(d, nullptr, &P);
    return P.G;
}

int acf_forward(const adfwi_acoustic_desc* d, const float* const* coef, const float* src_v, const int64_t* sx, const int64_t* sz,
                const int64_t* rx, const int64_t* rz, float* rcv_p, float* rcv_u, float* rcv_w,
                float* illum_p, float* illum_u, float* illum_w, void* ws, cudaStream_t st)
{
    FPlan P;
    acf_make_plan(d, ws, &P);
    const FGeom& g = P.g;
    int rc = acf_init_kernels();
    if (rc) return rc;
    StepMaps M;
    rc = acf_make_maps(P, &M);
    if (rc) return rc;
    rc = acf_copy_coefs(P, st, coef);
    if (rc) return rc;
    const bool illum = illum_p || illum_u || illum_w;
    const int nt = g.nt;
    const int csz = cdiv(nt, P.n_segments);
    const int last_chunk_start = (cdiv(nt, csz) - 1) * csz;
    if (illum) {
        ADFWI_CUDA(cudaMemsetAsync(P.ill_p, 0, sizeof(float) * g.plane, st));
        ADFWI_CUDA(cudaMemsetAsync(P.ill_u, 0, sizeof(float) * g.plane, st));
        ADFWI_CUDA(cudaMemsetAsync(P.ill_w, 0, sizeof(float) * g.plane, st));
    }
    if (P.nr > 0) { acf_rz_range<<<1, 256, 0, st>>>(P.nr, rz, P.rzrange); ADFWI_LAUNCH_CHECK(); }
    for (int sb = 0; sb < P.ns; sb += P.G) {
        const int se = sb + P.G < P.ns ? sb + P.G : P.ns;
        const size_t off = (size_t)sb * g.plane, cnt = (size_t)(se - sb) * g.plane * sizeof(float);
        for (int b = 0; b < 2; ++b) for (int f = 0; f < 3; ++f) ADFWI_CUDA(cudaMemsetAsync(P.st[b][f] + off, 0, cnt, st));
        int cur = 0;
        for (int it = 0; it < nt; ++it) {
            const int seg = it / P.K, tl = it - seg * P.K;
            if (P.save && tl == 0 && seg >= 1 && seg <= P.nseg - 2) {
                float* ck = P.ckpt + (size_t)(seg - 1) * 3 * P.ns * g.plane;
                for (int f = 0; f < 3; ++f)
                    ADFWI_CUDA(cudaMemcpyAsync(ck + (size_t)f * P.ns * g.plane + off, P.st[cur][f] + off, cnt, cudaMemcpyDeviceToDevice, st));
            }
            const bool save = P.save && seg == P.nseg - 1;
            rc = acf_forward_step(P, M, st, cur, sb, se, it, save, tl, src_v, sx, sz, rx, rz,
                                  P.nr > 0 ? rcv_p : nullptr, rcv_u, rcv_w, illum, it >= last_chunk_start);
            if (rc) return rc;
            cur ^= 1;
        }
        if (illum) {
            acf_sumsq<<<cdiv((int)g.plane, 256), 256, 0, st>>>(g.plane, g.plane, sb, se, P.st[cur][2], P.ill_w);
            ADFWI_LAUNCH_CHECK();
        }
    }
    if (illum) {
        const int nx = g.nxp - 2 * g.nabc, nz = g.nzp - 2 * g.nabc;
        acf_illum_finalize<<<dim3(cdiv(nx, 128), nz), 128, 0, st>>>(g.nzp, g.nxp, g.ld, g.nabc, P.ill_p, P.ill_u, P.ill_w, illum_p, illum_u, illum_w);
        ADFWI_LAUNCH_CHECK();
    }
    return ADFWI_OK;
}

int acf_backward(const adfwi_acoustic_desc* d, const float* const* coef, const float* src_v, const int64_t* sx, const int64_t* sz,
                 const int64_t* rx, const int64_t* rz, const float* gp, const float* gu, const float* gw,
                 float* g_alpha1, float* g_src, void* ws, cudaStream_t st)
{
    FPlan P;
    acf_make_plan(d, ws, &P);
    const FGeom& g = P.g;
    int rc = acf_init_kernels();
    if (rc) return rc;
    StepMaps M;
    rc = acf_make_maps(P, &M);
    if (rc) return rc;
    // coefficient copies and the receiver z-range were left in the workspace by the forward call
    const int nt = g.nt;
    ADFWI_CUDA(cudaMemsetAsync(P.g1part, 0, sizeof(float) * (size_t)P.G * g.plane, st));
    for (int sb = 0; sb < P.ns; sb += P.G) {
        const int se = sb + P.G < P.ns ? sb + P.G : P.ns;
        const size_t off = (size_t)sb * g.plane, cnt = (size_t)(se - sb) * g.plane * sizeof(float);
        for (int b = 0; b < 2; ++b) for (int f = 0; f < 3; ++f) ADFWI_CUDA(cudaMemsetAsync(P.lam[b][f] + off, 0, cnt, st));
        int lcur = 0;
        const int nitems = g.ntx * g.ntz * (se - sb);
        const int grid = nitems < 3 * acf_num_sms() ? nitems : 3 * acf_num_sms();
        for (int seg = P.nseg - 1; seg >= 0; --seg) {
            const int t0 = seg * P.K, t1 = t0 + P.K < nt ? t0 + P.K : nt;
            if (seg != P.nseg - 1) {
                int cur = 0;
                if (seg == 0) {
                    for (int b = 0; b < 2; ++b) for (int f = 0; f < 3; ++f) ADFWI_CUDA(cudaMemsetAsync(P.st[b][f] + off, 0, cnt, st));
                } else {
                    const float* ck = P.ckpt + (size_t)(seg - 1) * 3 * P.ns * g.plane;
                    for (int f = 0; f < 3; ++f)
                        ADFWI_CUDA(cudaMemcpyAsync(P.st[0][f] + off, ck + (size_t)f * P.ns * g.plane + off, cnt, cudaMemcpyDeviceToDevice, st));
                }
                for (int it = t0; it < t1; ++it) {
                    rc = acf_forward_step(P, M, st, cur, sb, se, it, true, it - t0, src_v, sx, sz, rx, rz, nullptr, nullptr, nullptr, false, 0);
                    if (rc) return rc;
                    cur ^= 1;
                }
            }
            for (int it = t1 - 1; it >= t0; --it) {
                AdjArgs a;
                a.a1 = P.coef[0]; a.k1 = P.coef[1]; a.k2 = P.coef[3]; a.k3 = P.coef[4];
                a.lp_out = P.lam[lcur ^ 1][0]; a.lu_out = P.lam[lcur ^ 1][1]; a.lw_out = P.lam[lcur ^ 1][2];
                a.sx = sx; a.sz = sz; a.hist = P.hist; a.hist_len = P.K; a.tl = it - t0; a.it = it;
                a.nr = P.nr; a.rx = rx; a.rz = rz; a.rzrange = P.rzrange; a.gp = gp; a.gu = gu; a.gw = gw;
                a.g1part = P.g1part; a.g_src = g_src; a.s_begin = sb; a.s_end = se;
                {
                    TimedLaunch tl_(KC_AC_ADJ_FUSED, st);
                    if (P.FS) ac_adj_fused<true><<<grid, NTHREADS, ADJ_SMEM, st>>>(M.lam[lcur][0], M.lam[lcur][1], M.lam[lcur][2], M.a2, g, a);
                    else      ac_adj_fused<false><<<grid, NTHREADS, ADJ_SMEM, st>>>(M.lam[lcur][0], M.lam[lcur][1], M.lam[lcur][2], M.a2, g, a);
                }
                ADFWI_LAUNCH_CHECK();
                lcur ^= 1;
            }
        }
    }
    acf_reduce_parts<<<dim3(cdiv(g.nxp, 128), g.nzp), 128, 0, st>>>(g.nzp, g.nxp, g.ld, P.G, P.g1part, g_alpha1);
    ADFWI_LAUNCH_CHECK();
    return ADFWI_OK;
}

}  // namespace adfwi
#endif  // !ADFWI_HOST_EMUL
