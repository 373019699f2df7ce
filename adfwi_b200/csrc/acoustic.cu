// acoustic.cu -- iso-acoustic (p,u,w) staggered-grid O(2,4) solver, forward and hand-written
// adjoint, for sm_100a.  Semantics: ADFWI/propagator/acoustic_kernels.py:113-174 (forward time
// loop) and the reverse-mode derivative of that loop (what autograd produces through
// acoustic_kernels.py:268-278).  Step algebra and evaluation order: SURVEY.md Appendix A.1/A.4.
//
// Arithmetic contract: this translation unit is compiled with -fmad=false, so every fp32
// operation rounds once, in the association written here, which is the association of the eager
// PyTorch expressions.  Forward records are therefore bit-identical to the CPU reference.
//
// All kernels are gathers (no atomics on the wavefields): each thread owns one cell column
// position (z,x) and loops over the shots of its shot sub-group, so coefficient loads and the
// read-modify-write of gradient / illumination planes are amortised over shots.
#include "common.cuh"
#include "acoustic_fused.h"

namespace adfwi {

struct AcGeom {
    int nzp, nxp, fs, nabc, nt;
    size_t plane;
    float c1, c2, dt;
};

struct ShotRange { int begin, end, spt; };   // shots [begin,end), spt shots per thread (grid.z)

// ------------------------------------------------------------------------------------------
// forward, step 1-3: pressure update + source + free-surface mirror   (acoustic_kernels.py:115-136)
// ------------------------------------------------------------------------------------------
template <bool FS, bool SAVE>
__global__ void __launch_bounds__(256)
ac_fwd_p(const AcGeom g, const ShotRange sr,
         const float* __restrict__ a1, const float* __restrict__ k1,
         float* __restrict__ p, const float* __restrict__ u, const float* __restrict__ w,
         const float* __restrict__ src_v, const int64_t* __restrict__ sx,
         const int64_t* __restrict__ sz, float* __restrict__ hist, int hist_len, int tl, int it)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int z = g.fs - 1 + blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= g.nxp || z >= g.nzp) return;
    const int s0 = sr.begin + blockIdx.z * sr.spt;
    const int s1 = min(s0 + sr.spt, sr.end);
    const bool inreg = (z >= g.fs + 1) && (z < g.nzp - 2) && (x >= 2) && (x < g.nxp - 2);
    const size_t c = (size_t)z * g.nxp + x;
    float t1 = 0.f, al = 0.f;
    if (inreg) { t1 = 1.0f - k1[c]; al = a1[c]; }
    const int nxp = g.nxp;
    for (int s = s0; s < s1; ++s) {
        const size_t o = (size_t)s * g.plane + c;
        float pv = p[o];
        bool touched = false;
        if (inreg) {
            const float* us = u + o;
            const float* ws = w + o;
            float S1 = ((us[0] - us[-1]) + ws[0]) - ws[-nxp];
            float S2 = ((us[1] - us[-2]) + ws[nxp]) - ws[-2 * nxp];
            float S = g.c1 * S1 + g.c2 * S2;
            if (SAVE) __stcs(hist + ((size_t)s * hist_len + tl) * g.plane + c, S);
            pv = t1 * pv - al * S;
            touched = true;
        }
        if ((int64_t)z == sz[s] && (int64_t)x == sx[s]) {
            pv = pv + g.dt * src_v[(size_t)s * g.nt + it];
            touched = true;
        }
        if (touched) p[o] = pv;
        if (FS && z == g.fs + 1) p[o - 2 * (size_t)nxp] = -pv;   // p[fs-1] = -p[fs+1], all x
    }
}

// ------------------------------------------------------------------------------------------
// forward, step 4-6 (+8): velocity updates + free surface + illumination   (:139-164, :172-173)
// ------------------------------------------------------------------------------------------
template <bool FS, bool ILLUM>
__global__ void __launch_bounds__(256)
ac_fwd_uw(const AcGeom g, const ShotRange sr,
          const float* __restrict__ a2, const float* __restrict__ k2, const float* __restrict__ k3,
          const float* __restrict__ p, float* __restrict__ u, float* __restrict__ w,
          float* __restrict__ ill_p, float* __restrict__ ill_u, int acc_u)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int z = g.fs - 1 + blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= g.nxp || z >= g.nzp) return;
    const int s0 = sr.begin + blockIdx.z * sr.spt;
    const int s1 = min(s0 + sr.spt, sr.end);
    const int nxp = g.nxp;
    const bool inU = (z >= g.fs) && (z < g.nzp - 1) && (x >= 1) && (x < nxp - 2);
    const bool inW = (z >= g.fs) && (z < g.nzp - 2) && (x >= 1) && (x < nxp - 1);
    const bool interior = ILLUM && (z >= g.nabc) && (z < g.nzp - g.nabc) && (x >= g.nabc) && (x < nxp - g.nabc);
    const size_t c = (size_t)z * nxp + x;
    float al = 0.f, t2 = 0.f, t3 = 0.f;
    if (inU || inW) al = a2[c];
    if (inU) t2 = 1.0f - k2[c];
    if (inW) t3 = 1.0f - k3[c];
    float accp = 0.f, accu = 0.f;
    for (int s = s0; s < s1; ++s) {
        const size_t o = (size_t)s * g.plane + c;
        const float* ps = p + o;
        const float p0 = ps[0];
        float uv = 0.f;
        if (inU) {
            uv = t2 * u[o] - al * (g.c1 * (ps[1] - p0) + g.c2 * (ps[2] - ps[-1]));
            u[o] = uv;
        } else if (interior) uv = u[o];
        if (inW) {
            float wv = t3 * w[o] - al * (g.c1 * (ps[nxp] - p0) + g.c2 * (ps[2 * nxp] - ps[-nxp]));
            w[o] = wv;
            if (FS && z == g.fs) w[o - nxp] = wv;
        } else if (FS && z == g.fs) {
            w[o - nxp] = w[o];
        }
        if (interior) { accp += p0 * p0; accu += uv * uv; }
    }
    if (interior) {
        const size_t q = (size_t)blockIdx.z * g.plane + c;
        ill_p[q] += accp;
        if (acc_u) ill_u[q] += accu;
    }
}

// step 7: receiver sampling (:167-169)
__global__ void ac_record(const AcGeom g, int s_begin, int s_end, int nr,
                          const float* __restrict__ p, const float* __restrict__ u,
                          const float* __restrict__ w, const int64_t* __restrict__ rx,
                          const int64_t* __restrict__ rz, float* __restrict__ rcv_p,
                          float* __restrict__ rcv_u, float* __restrict__ rcv_w, int it)
{
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    const int s = s_begin + blockIdx.y;
    if (r >= nr || s >= s_end) return;
    const int64_t z = rz[r], x = rx[r];
    if (z < 0 || z >= g.nzp || x < 0 || x >= g.nxp) return;
    const size_t c = (size_t)s * g.plane + (size_t)z * g.nxp + x;
    const size_t o = ((size_t)s * g.nt + it) * nr + r;
    rcv_p[o] = p[c];
    if (rcv_u) rcv_u[o] = u[c];
    if (rcv_w) rcv_w[o] = w[c];
}

// sum_s w^2 of the final state (literal forward_wavefield_w, :174)
__global__ void ac_sumsq(const AcGeom g, int s_begin, int s_end, const float* __restrict__ f,
                         float* __restrict__ out)
{
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= g.plane) return;
    float a = 0.f;
    for (int s = s_begin; s < s_end; ++s) { float v = f[(size_t)s * g.plane + i]; a += v * v; }
    out[i] += a;
}

// forward_wavefield_{p,u,w} assembly (:286-288): IP = sum p^2; IU = IP + Iu; IW = IP + (Iu + w_last^2)
__global__ void ac_illum_finalize(const AcGeom g, int nsub, const float* __restrict__ ill_p,
                                  const float* __restrict__ ill_u, const float* __restrict__ ill_w,
                                  float* __restrict__ out_p, float* __restrict__ out_u,
                                  float* __restrict__ out_w)
{
    const int nx = g.nxp - 2 * g.nabc, nz = g.nzp - 2 * g.nabc;
    const int x = blockIdx.x * blockDim.x + threadIdx.x, z = blockIdx.y;
    if (x >= nx || z >= nz) return;
    const size_t c = (size_t)(z + g.nabc) * g.nxp + (x + g.nabc);
    float ip = 0.f, iu = 0.f;
    for (int k = 0; k < nsub; ++k) { ip += ill_p[k * g.plane + c]; iu += ill_u[k * g.plane + c]; }
    const size_t o = (size_t)z * nx + x;
    if (out_p) out_p[o] = ip;
    if (out_u) out_u[o] = ip + iu;
    if (out_w) out_w[o] = ip + (iu + ill_w[c]);
}

// ------------------------------------------------------------------------------------------
// adjoint
// ------------------------------------------------------------------------------------------
// 7T: scatter-add of the record cotangents (duplicate receiver cells are legal -> atomics)
__global__ void ac_adj_inject(const AcGeom g, int s_begin, int s_end, int nr,
                              float* __restrict__ lp, float* __restrict__ lu, float* __restrict__ lw,
                              const int64_t* __restrict__ rx, const int64_t* __restrict__ rz,
                              const float* __restrict__ gp, const float* __restrict__ gu,
                              const float* __restrict__ gw, int it)
{
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    const int s = s_begin + blockIdx.y;
    if (r >= nr || s >= s_end) return;
    const int64_t z = rz[r], x = rx[r];
    if (z < 0 || z >= g.nzp || x < 0 || x >= g.nxp) return;
    const size_t c = (size_t)s * g.plane + (size_t)z * g.nxp + x;
    const size_t o = ((size_t)s * g.nt + it) * nr + r;
    if (gp) atomicAdd(lp + c, gp[o]);
    if (gu) atomicAdd(lu + c, gu[o]);
    if (gw) atomicAdd(lw + c, gw[o]);
}

// 6T,5T,4T (pressure part) and 3T:  lpX = lambda_p after undoing W, U and the p free-surface
// mirror.  Also accumulates g_alpha2 when requested (needs the stored post-step pressure).
template <bool FS, bool G2>
__global__ void __launch_bounds__(256)
ac_adj_a(const AcGeom g, const ShotRange sr, const float* __restrict__ a2,
         const float* __restrict__ lpY, const float* __restrict__ lu, const float* __restrict__ lw,
         float* __restrict__ lpX, const float* __restrict__ histP, int hist_len, int tl,
         float* __restrict__ g2part)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int z = g.fs - 1 + blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= g.nxp || z >= g.nzp) return;
    const int s0 = sr.begin + blockIdx.z * sr.spt;
    const int s1 = min(s0 + sr.spt, sr.end);
    const int nxp = g.nxp, nzp = g.nzp, fs = g.fs;
    const size_t c = (size_t)z * nxp + x;
    // -alpha2 at the W cells z-2..z+1 (same x) and U cells x-2..x+1 (same z), 0 outside the regions
    float aw[4], au[4];
    const bool xw = (x >= 1) && (x < nxp - 1);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int zz = z - 2 + k;
        aw[k] = (xw && zz >= fs && zz < nzp - 2) ? -a2[(size_t)zz * nxp + x] : 0.f;
        const int xx = x - 2 + k;
        au[k] = (z >= fs && z < nzp - 1 && xx >= 1 && xx < nxp - 2) ? -a2[(size_t)z * nxp + xx] : 0.f;
    }
    const bool inU = (z >= fs) && (z < nzp - 1) && (x >= 1) && (x < nxp - 2);
    const bool inW = (z >= fs) && (z < nzp - 2) && xw;
    float gacc = 0.f;
    for (int s = s0; s < s1; ++s) {
        const size_t o = (size_t)s * g.plane + c;
        float qw[4], qu[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int zz = z - 2 + k;
            float v = 0.f;
            if (aw[k] != 0.f) {
                v = lw[o + (ptrdiff_t)(k - 2) * nxp];
                if (FS && zz == fs) v += lw[o + (ptrdiff_t)(k - 3) * nxp];   // 6T: row fs absorbs row fs-1
            }
            qw[k] = aw[k] * v;
            qu[k] = (au[k] != 0.f) ? au[k] * lu[o + (k - 2)] : 0.f;
        }
        // transposes of D+z and D+x:  +c1 m[z-1] - c1 m[z] + c2 m[z-2] - c2 m[z+1]
        float acc = lpY[o];
        acc += g.c1 * qw[1] - g.c1 * qw[2] + g.c2 * qw[0] - g.c2 * qw[3];
        acc += g.c1 * qu[1] - g.c1 * qu[2] + g.c2 * qu[0] - g.c2 * qu[3];
        if (FS) {
            if (z == fs - 1) acc = 0.f;                                  // 3T
            else if (z == fs + 1) {
                // lambda_p1[fs-1] = lpY[fs-1] - c2*m_w[fs]   (only W cell reaching row fs-1)
                float top = lpY[o - 2 * (size_t)nxp] - g.c2 * qw[1];
                acc -= top;
            }
        }
        lpX[o] = acc;
        if (G2) {
            const float* P = histP + ((size_t)s * hist_len + tl) * g.plane + c;
            if (inW) {
                float q = lw[o];
                if (FS && z == fs) q += lw[o - nxp];
                gacc -= q * (g.c1 * (P[nxp] - P[0]) + g.c2 * (P[2 * nxp] - P[-nxp]));
            }
            if (inU) gacc -= lu[o] * (g.c1 * (P[1] - P[0]) + g.c2 * (P[2] - P[-1]));
        }
    }
    if (G2 && (inU || inW)) g2part[(size_t)blockIdx.z * g.plane + c] += gacc;
}

// 5T,4T (velocity part), 2T, 1T: new lambda_u, lambda_w, lambda_p; g_alpha1 and g_src.
template <bool FS>
__global__ void __launch_bounds__(256)
ac_adj_b(const AcGeom g, const ShotRange sr, const float* __restrict__ a1,
         const float* __restrict__ k1, const float* __restrict__ k2, const float* __restrict__ k3,
         const float* __restrict__ lpX, float* __restrict__ lpY, float* __restrict__ lu,
         float* __restrict__ lw, const float* __restrict__ histS, int hist_len, int tl,
         float* __restrict__ g1part, const int64_t* __restrict__ sx, const int64_t* __restrict__ sz,
         float* __restrict__ g_src, int it)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int z = g.fs - 1 + blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= g.nxp || z >= g.nzp) return;
    const int s0 = sr.begin + blockIdx.z * sr.spt;
    const int s1 = min(s0 + sr.spt, sr.end);
    const int nxp = g.nxp, nzp = g.nzp, fs = g.fs;
    const size_t c = (size_t)z * nxp + x;
    // -alpha1 at the P cells (z, x-1..x+2) and (z-1..z+2, x); index 1 of both is the own cell
    float ax[4], az[4];
    const bool zp = (z >= fs + 1) && (z < nzp - 2), xp = (x >= 2) && (x < nxp - 2);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int xx = x - 1 + k, zz = z - 1 + k;
        ax[k] = (zp && xx >= 2 && xx < nxp - 2) ? -a1[(size_t)z * nxp + xx] : 0.f;
        az[k] = (xp && zz >= fs + 1 && zz < nzp - 2) ? -a1[(size_t)zz * nxp + x] : 0.f;
    }
    const bool inP = zp && xp;
    const bool inU = (z >= fs) && (z < nzp - 1) && (x >= 1) && (x < nxp - 2);
    const bool inW = (z >= fs) && (z < nzp - 2) && (x >= 1) && (x < nxp - 1);
    const float t1 = inP ? 1.0f - k1[c] : 1.0f;
    const float t2 = inU ? 1.0f - k2[c] : 1.0f;
    const float t3 = inW ? 1.0f - k3[c] : 1.0f;
    float gacc = 0.f;
    for (int s = s0; s < s1; ++s) {
        const size_t o = (size_t)s * g.plane + c;
        float mx[4], mz[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            mx[k] = (ax[k] != 0.f) ? ax[k] * lpX[o + (k - 1)] : 0.f;
            mz[k] = (az[k] != 0.f) ? az[k] * lpX[o + (ptrdiff_t)(k - 1) * nxp] : 0.f;
        }
        const float qp = lpX[o];
        // lambda_u: transpose of D-x:  +c1 m[x] - c1 m[x+1] + c2 m[x-1] - c2 m[x+2]
        lu[o] = t2 * lu[o] + (g.c1 * mx[1] - g.c1 * mx[2] + g.c2 * mx[0] - g.c2 * mx[3]);
        // lambda_w, with the free-surface copy w[fs-1] = w[fs] undone (6T) by the row-fs thread
        const float dw = g.c1 * mz[1] - g.c1 * mz[2] + g.c2 * mz[0] - g.c2 * mz[3];
        if (FS && z == fs) {
            lw[o] = t3 * (lw[o] + lw[o - nxp]) + dw;
            lw[o - nxp] = -g.c2 * mz[2];          // row fs-1: 0 + (only P cell fs+1 reaches it)
        } else if (!(FS && z == fs - 1)) {
            lw[o] = t3 * lw[o] + dw;
        }
        if (inP) gacc -= qp * __ldcs(histS + ((size_t)s * hist_len + tl) * g.plane + c);
        lpY[o] = t1 * qp;
        if (g_src && (int64_t)z == sz[s] && (int64_t)x == sx[s]) g_src[(size_t)s * g.nt + it] = g.dt * qp;
    }
    if (inP) g1part[(size_t)blockIdx.z * g.plane + c] += gacc;
}

__global__ void ac_reduce_parts(size_t n, size_t plane, int nsub, const float* __restrict__ part,
                                float* __restrict__ out)
{
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float a = 0.f;
    for (int k = 0; k < nsub; ++k) a += part[k * plane + i];
    out[i] = a;
}

// ------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------
struct AcPlan {
    AcGeom g;
    int ns, nr, FS;
    int K, nseg, nckpt;       // history length, number of segments, stored checkpoints
    int G, nsub, spt;         // shots per group, sub-groups (grid.z), shots per thread
    int save, need_g2, n_segments;
    // workspace regions
    float *p, *u, *w, *lpX, *lpY, *lu, *lw, *histS, *histP, *ckpt, *g1part, *g2part, *ill_p, *ill_u, *ill_w;
    size_t bytes;
};

static int ac_make_plan(const adfwi_acoustic_desc* d, void* ws, AcPlan* P)
{
    if (!d) return ADFWI_E_NULL;
    if (d->nzp < 8 || d->nxp < 8 || d->ns < 1 || d->nt < 1 || d->nr < 0 || d->nabc < 1) return ADFWI_E_DIMS;
    if (d->nzp - 2 * d->nabc < 1 || d->nxp - 2 * d->nabc < 1) return ADFWI_E_DIMS;
    AcGeom& g = P->g;
    g.nzp = d->nzp; g.nxp = d->nxp; g.nabc = d->nabc; g.nt = d->nt;
    g.fs = d->free_surface ? d->nabc : 1;
    g.plane = (size_t)d->nzp * d->nxp;
    g.c1 = d->c1; g.c2 = d->c2; g.dt = d->dt;
    P->ns = d->ns; P->nr = d->nr; P->FS = d->free_surface ? 1 : 0;
    P->save = d->save_history ? 1 : 0;
    P->need_g2 = (d->save_history && d->need_g_alpha2) ? 1 : 0;
    P->n_segments = d->n_segments > 0 ? d->n_segments : 1;
    int K = d->ckpt_interval;
    if (K <= 0 || K >= d->nt) K = d->nt;
    P->K = K; P->nseg = cdiv(d->nt, K); P->nckpt = P->nseg > 2 ? P->nseg - 2 : 0;
    // shots per group: keep one group's adjoint working set (4 planes/shot) inside ~1/2 of L2
    int G = d->shots_per_group;
    if (G <= 0) {
        const size_t per_shot = g.plane * sizeof(float) * 4;
        G = (int)((size_t)(56u << 20) / per_shot);
        if (G < 1) G = 1;
    }
    if (G > d->ns) G = d->ns;
    P->G = G;
    // enough thread parallelism per launch: ~4 resident waves of 148 SMs x 2048 threads
    int nsub = cdiv(1200000, (int)(g.plane > 1200000 ? 1200000 : g.plane));
    if (nsub > G) nsub = G;
    P->spt = cdiv(G, nsub);
    P->nsub = cdiv(G, P->spt);
    Carver cv(ws);
    const size_t sp = (size_t)d->ns * g.plane;
    P->p = cv.take<float>(sp); P->u = cv.take<float>(sp); P->w = cv.take<float>(sp);
    P->ill_p = cv.take<float>(P->nsub * g.plane);
    P->ill_u = cv.take<float>(P->nsub * g.plane);
    P->ill_w = cv.take<float>(g.plane);
    P->lpX = P->lpY = P->lu = P->lw = P->histS = P->histP = P->ckpt = P->g1part = P->g2part = nullptr;
    if (P->save) {
        P->lpX = cv.take<float>(sp); P->lpY = cv.take<float>(sp);
        P->lu = cv.take<float>(sp); P->lw = cv.take<float>(sp);
        P->g1part = cv.take<float>(P->nsub * g.plane);
        if (P->need_g2) P->g2part = cv.take<float>(P->nsub * g.plane);
        if (P->nckpt) P->ckpt = cv.take<float>((size_t)P->nckpt * 3 * sp);
        P->histS = cv.take<float>((size_t)K * sp);
        if (P->need_g2) P->histP = cv.take<float>((size_t)K * sp);
    }
    P->bytes = cv.off;
    return ADFWI_OK;
}

static inline dim3 ac_grid(const AcPlan& P, int nsub, dim3 blk)
{
    const int rows = P.g.nzp - (P.g.fs - 1);
    return dim3(cdiv(P.g.nxp, blk.x), cdiv(rows, blk.y), nsub);
}

// one forward step for shots [sb,se): P, UW kernels (+ record when rcv_p != null)
static int ac_forward_step(const AcPlan& P, cudaStream_t st, int sb, int se, int it, bool save, int tl,
                           const float* a1, const float* a2, const float* k1, const float* k2,
                           const float* k3, const float* src_v, const int64_t* sx, const int64_t* sz,
                           bool illum, int acc_u)
{
    const dim3 blk(64, 4);
    ShotRange sr{sb, se, P.spt};
    const dim3 grd = ac_grid(P, cdiv(se - sb, P.spt), blk);
#define LP(FSv, SAVEv) do { TimedLaunch tl_(KC_AC_FWD_P, st); ADFWI_LAUNCH(ADFWI_KERNEL(ac_fwd_p<FSv, SAVEv>), grd, blk, st, P.g, sr, a1, k1, P.p, P.u, P.w, src_v, sx, sz, P.histS, P.K, tl, it); } while (0)
    if (P.FS) { if (save) LP(true, true); else LP(true, false); }
    else      { if (save) LP(false, true); else LP(false, false); }
#undef LP
    ADFWI_LAUNCH_CHECK();
#define LU(FSv, ILv) do { TimedLaunch tl_(KC_AC_FWD_UW, st); ADFWI_LAUNCH(ADFWI_KERNEL(ac_fwd_uw<FSv, ILv>), grd, blk, st, P.g, sr, a2, k2, k3, P.p, P.u, P.w, P.ill_p, P.ill_u, acc_u); } while (0)
    if (P.FS) { if (illum) LU(true, true); else LU(true, false); }
    else      { if (illum) LU(false, true); else LU(false, false); }
#undef LU
    ADFWI_LAUNCH_CHECK();
    if (save && P.need_g2) {   // post-step pressure history for the density gradient
        for (int s = sb; s < se; ++s)
            ADFWI_CUDA(cudaMemcpyAsync(P.histP + ((size_t)s * P.K + tl) * P.g.plane, P.p + (size_t)s * P.g.plane,
                                       P.g.plane * sizeof(float), cudaMemcpyDeviceToDevice, st));
    }
    return ADFWI_OK;
}

}  // namespace adfwi

using namespace adfwi;

// The fused TMA pipeline (acoustic_fused.cu) is the default, with or without the density gradient; the
// generic kernels of this file run when the caller sets bit 0 of desc->reserved[0] (used by the tests to
// cross-check the two pipelines) or for grids beyond the fused path's index packing.
static bool ac_use_fused(const adfwi_acoustic_desc* d)
{
#ifdef ADFWI_HOST_EMUL
    (void)d; return false;
#else
    if (d->nzp >= 32768 || d->nxp >= 65536) return false;      // the fused path packs receiver cells as (z<<16)|x
    return !(d->reserved[0] & 1);
#endif
}

extern "C" size_t adfwi_acoustic_workspace_bytes(const adfwi_acoustic_desc* desc)
{
    AcPlan P;
    if (ac_make_plan(desc, nullptr, &P) != ADFWI_OK) return 0;
#ifndef ADFWI_HOST_EMUL
    if (ac_use_fused(desc)) return acf_workspace_bytes(desc);
#endif
    return P.bytes;
}

extern "C" int adfwi_acoustic_group_size(const adfwi_acoustic_desc* desc)
{
    AcPlan P;
    if (ac_make_plan(desc, nullptr, &P) != ADFWI_OK) return 0;
#ifndef ADFWI_HOST_EMUL
    if (ac_use_fused(desc)) return acf_group_size(desc);
#endif
    return P.G;
}

extern "C" int adfwi_acoustic_forward(const adfwi_acoustic_desc* desc,
                                      const float* alpha1, const float* alpha2, const float* kappa1,
                                      const float* kappa2, const float* kappa3,
                                      const float* src_v, const int64_t* src_x, const int64_t* src_z,
                                      const int64_t* rcv_x, const int64_t* rcv_z,
                                      float* rcv_p, float* rcv_u, float* rcv_w,
                                      float* illum_p, float* illum_u, float* illum_w,
                                      void* workspace, size_t workspace_bytes, void* stream)
{
    ADFWI_NVTX("adfwi_acoustic_forward");
    AcPlan P;
    int rc = ac_make_plan(desc, workspace, &P);
    if (rc) return rc;
    if (!alpha1 || !alpha2 || !kappa1 || !kappa2 || !kappa3 || !src_v || !src_x || !src_z || !workspace) return ADFWI_E_NULL;
    if (P.nr > 0 && (!rcv_x || !rcv_z || !rcv_p)) return ADFWI_E_NULL;
    cudaStream_t st = (cudaStream_t)stream;
#ifndef ADFWI_HOST_EMUL
    if (ac_use_fused(desc)) {
        if (workspace_bytes < acf_workspace_bytes(desc)) return ADFWI_E_WORKSPACE;
        const float* coef[5] = {alpha1, kappa1, alpha2, kappa2, kappa3};
        return acf_forward(desc, coef, src_v, src_x, src_z, rcv_x, rcv_z, rcv_p, rcv_u, rcv_w, illum_p, illum_u, illum_w, workspace, st);
    }
#endif
    if (workspace_bytes < P.bytes) return ADFWI_E_WORKSPACE;
    const AcGeom& g = P.g;
    const bool illum = illum_p || illum_u || illum_w;
    const int nt = g.nt;
    // last chunk of torch.chunk(src_v, n_segments, dim=-1): chunk size ceil(nt/n_segments)
    const int csz = cdiv(nt, P.n_segments);
    const int last_chunk_start = (cdiv(nt, csz) - 1) * csz;
    if (illum) {
        ADFWI_CUDA(cudaMemsetAsync(P.ill_p, 0, sizeof(float) * P.nsub * g.plane, st));
        ADFWI_CUDA(cudaMemsetAsync(P.ill_u, 0, sizeof(float) * P.nsub * g.plane, st));
        ADFWI_CUDA(cudaMemsetAsync(P.ill_w, 0, sizeof(float) * g.plane, st));
    }
    for (int sb = 0; sb < P.ns; sb += P.G) {
        const int se = sb + P.G < P.ns ? sb + P.G : P.ns;
        const size_t off = (size_t)sb * g.plane, cnt = (size_t)(se - sb) * g.plane * sizeof(float);
        ADFWI_CUDA(cudaMemsetAsync(P.p + off, 0, cnt, st));
        ADFWI_CUDA(cudaMemsetAsync(P.u + off, 0, cnt, st));
        ADFWI_CUDA(cudaMemsetAsync(P.w + off, 0, cnt, st));
        for (int it = 0; it < nt; ++it) {
            const int seg = it / P.K, tl = it - seg * P.K;
            if (P.save && tl == 0 && seg >= 1 && seg <= P.nseg - 2) {   // state checkpoint before step it
                float* ck = P.ckpt + (size_t)(seg - 1) * 3 * P.ns * g.plane;
                ADFWI_CUDA(cudaMemcpyAsync(ck + 0 * P.ns * g.plane + off, P.p + off, cnt, cudaMemcpyDeviceToDevice, st));
                ADFWI_CUDA(cudaMemcpyAsync(ck + 1 * P.ns * g.plane + off, P.u + off, cnt, cudaMemcpyDeviceToDevice, st));
                ADFWI_CUDA(cudaMemcpyAsync(ck + 2 * P.ns * g.plane + off, P.w + off, cnt, cudaMemcpyDeviceToDevice, st));
            }
            const bool save = P.save && seg == P.nseg - 1;
            rc = ac_forward_step(P, st, sb, se, it, save, tl, alpha1, alpha2, kappa1, kappa2, kappa3,
                                 src_v, src_x, src_z, illum, it >= last_chunk_start);
            if (rc) return rc;
            if (P.nr > 0) {
                TimedLaunch tl_(KC_AC_RECORD, st);
                ADFWI_LAUNCH(ADFWI_KERNEL(ac_record), dim3(cdiv(P.nr, 128), se - sb), 128, st, g, sb, se, P.nr, P.p, P.u, P.w, rcv_x, rcv_z,
                                                                          rcv_p, rcv_u, rcv_w, it);
                ADFWI_LAUNCH_CHECK();
            }
        }
        if (illum) {
            ADFWI_LAUNCH(ADFWI_KERNEL(ac_sumsq), cdiv((int)g.plane, 256), 256, st, g, sb, se, P.w, P.ill_w);
            ADFWI_LAUNCH_CHECK();
        }
    }
    if (illum) {
        const int nx = g.nxp - 2 * g.nabc, nz = g.nzp - 2 * g.nabc;
        ADFWI_LAUNCH(ADFWI_KERNEL(ac_illum_finalize), dim3(cdiv(nx, 128), nz), 128, st, g, P.nsub, P.ill_p, P.ill_u, P.ill_w,
                                                                 illum_p, illum_u, illum_w);
        ADFWI_LAUNCH_CHECK();
    }
    return ADFWI_OK;
}

extern "C" int adfwi_acoustic_backward(const adfwi_acoustic_desc* desc,
                                       const float* alpha1, const float* alpha2, const float* kappa1,
                                       const float* kappa2, const float* kappa3,
                                       const float* src_v, const int64_t* src_x, const int64_t* src_z,
                                       const int64_t* rcv_x, const int64_t* rcv_z,
                                       const float* g_rcv_p, const float* g_rcv_u, const float* g_rcv_w,
                                       float* g_alpha1, float* g_alpha2, float* g_src_v,
                                       void* workspace, size_t workspace_bytes, void* stream)
{
    ADFWI_NVTX("adfwi_acoustic_backward");
    AcPlan P;
    int rc = ac_make_plan(desc, workspace, &P);
    if (rc) return rc;
    if (!P.save) return ADFWI_E_MODE;
    if (!alpha1 || !alpha2 || !kappa1 || !kappa2 || !kappa3 || !src_v || !src_x || !src_z || !workspace || !g_alpha1)
        return ADFWI_E_NULL;
    if (P.need_g2 && !g_alpha2) return ADFWI_E_NULL;
    if (P.nr > 0 && (!rcv_x || !rcv_z)) return ADFWI_E_NULL;
    cudaStream_t st = (cudaStream_t)stream;
#ifndef ADFWI_HOST_EMUL
    if (ac_use_fused(desc)) {
        if (workspace_bytes < acf_workspace_bytes(desc)) return ADFWI_E_WORKSPACE;
        const float* coef[5] = {alpha1, kappa1, alpha2, kappa2, kappa3};
        return acf_backward(desc, coef, src_v, src_x, src_z, rcv_x, rcv_z, g_rcv_p, g_rcv_u, g_rcv_w, g_alpha1, g_alpha2, g_src_v, workspace, st);
    }
#endif
    if (workspace_bytes < P.bytes) return ADFWI_E_WORKSPACE;
    const AcGeom& g = P.g;
    const int nt = g.nt;
    const dim3 blk(64, 4);
    const bool have_g = (g_rcv_p || g_rcv_u || g_rcv_w) && P.nr > 0;
    ADFWI_CUDA(cudaMemsetAsync(P.g1part, 0, sizeof(float) * P.nsub * g.plane, st));
    if (P.need_g2) ADFWI_CUDA(cudaMemsetAsync(P.g2part, 0, sizeof(float) * P.nsub * g.plane, st));
    for (int sb = 0; sb < P.ns; sb += P.G) {
        const int se = sb + P.G < P.ns ? sb + P.G : P.ns;
        const size_t off = (size_t)sb * g.plane, cnt = (size_t)(se - sb) * g.plane * sizeof(float);
        ADFWI_CUDA(cudaMemsetAsync(P.lpY + off, 0, cnt, st));
        ADFWI_CUDA(cudaMemsetAsync(P.lu + off, 0, cnt, st));
        ADFWI_CUDA(cudaMemsetAsync(P.lw + off, 0, cnt, st));
        ShotRange sr{sb, se, P.spt};
        const dim3 grd = ac_grid(P, cdiv(se - sb, P.spt), blk);
        for (int seg = P.nseg - 1; seg >= 0; --seg) {
            const int t0 = seg * P.K, t1 = t0 + P.K < nt ? t0 + P.K : nt;
            if (seg != P.nseg - 1) {
                // recompute this segment's history from its checkpoint (zero state for segment 0)
                if (seg == 0) {
                    ADFWI_CUDA(cudaMemsetAsync(P.p + off, 0, cnt, st));
                    ADFWI_CUDA(cudaMemsetAsync(P.u + off, 0, cnt, st));
                    ADFWI_CUDA(cudaMemsetAsync(P.w + off, 0, cnt, st));
                } else {
                    const float* ck = P.ckpt + (size_t)(seg - 1) * 3 * P.ns * g.plane;
                    ADFWI_CUDA(cudaMemcpyAsync(P.p + off, ck + 0 * P.ns * g.plane + off, cnt, cudaMemcpyDeviceToDevice, st));
                    ADFWI_CUDA(cudaMemcpyAsync(P.u + off, ck + 1 * P.ns * g.plane + off, cnt, cudaMemcpyDeviceToDevice, st));
                    ADFWI_CUDA(cudaMemcpyAsync(P.w + off, ck + 2 * P.ns * g.plane + off, cnt, cudaMemcpyDeviceToDevice, st));
                }
                for (int it = t0; it < t1; ++it) {
                    rc = ac_forward_step(P, st, sb, se, it, true, it - t0, alpha1, alpha2, kappa1, kappa2, kappa3,
                                         src_v, src_x, src_z, false, 0);
                    if (rc) return rc;
                }
            }
            for (int it = t1 - 1; it >= t0; --it) {
                const int tl = it - t0;
                if (have_g) {
                    TimedLaunch tl_(KC_AC_ADJ_INJECT, st);
                    ADFWI_LAUNCH(ADFWI_KERNEL(ac_adj_inject), dim3(cdiv(P.nr, 128), se - sb), 128, st, g, sb, se, P.nr, P.lpY, P.lu, P.lw,
                                                                                  rcv_x, rcv_z, g_rcv_p, g_rcv_u, g_rcv_w, it);
                    ADFWI_LAUNCH_CHECK();
                }
#define LA(FSv, G2v) do { TimedLaunch tl_(KC_AC_ADJ_A, st); ADFWI_LAUNCH(ADFWI_KERNEL(ac_adj_a<FSv, G2v>), grd, blk, st, g, sr, alpha2, P.lpY, P.lu, P.lw, P.lpX, P.histP, P.K, tl, P.g2part); } while (0)
                if (P.FS) { if (P.need_g2) LA(true, true); else LA(true, false); }
                else      { if (P.need_g2) LA(false, true); else LA(false, false); }
#undef LA
                ADFWI_LAUNCH_CHECK();
                {
                TimedLaunch tl_(KC_AC_ADJ_B, st);
                if (P.FS) ADFWI_LAUNCH(ADFWI_KERNEL(ac_adj_b<true>), grd, blk, st, g, sr, alpha1, kappa1, kappa2, kappa3, P.lpX, P.lpY, P.lu, P.lw,
                                                              P.histS, P.K, tl, P.g1part, src_x, src_z, g_src_v, it);
                else      ADFWI_LAUNCH(ADFWI_KERNEL(ac_adj_b<false>), grd, blk, st, g, sr, alpha1, kappa1, kappa2, kappa3, P.lpX, P.lpY, P.lu, P.lw,
                                                               P.histS, P.K, tl, P.g1part, src_x, src_z, g_src_v, it);
                }
                ADFWI_LAUNCH_CHECK();
            }
        }
    }
    const int nb = cdiv((int)g.plane, 256);
    ADFWI_LAUNCH(ADFWI_KERNEL(ac_reduce_parts), nb, 256, st, g.plane, g.plane, P.nsub, P.g1part, g_alpha1);
    ADFWI_LAUNCH_CHECK();
    if (P.need_g2) {
        ADFWI_LAUNCH(ADFWI_KERNEL(ac_reduce_parts), nb, 256, st, g.plane, g.plane, P.nsub, P.g2part, g_alpha2);
        ADFWI_LAUNCH_CHECK();
    }
    return ADFWI_OK;
}
