/*
 * oracle/elastic_oracle.c -- CPU restatement of ADFWI's 2-D P-SV velocity-stress solvers.
 *
 * TEST INFRASTRUCTURE ONLY (see acoustic_oracle.c): the checker, never the product.
 *
 * What it restates (paths relative to the upstream reference tree):
 *   split-field PML time loop : ADFWI/propagator/elastic_kernels.py:339-418 (O4), :495-575 (O6)
 *   sponge/ABL time loop      : ADFWI/propagator/elastic_kernels.py:709-774 (O4), :844-908 (O6)
 *   FD operators              : ADFWI/propagator/elastic_kernels.py:66-108
 *   adjoints                  : reverse-mode derivative of those loops (what autograd produces
 *                               through elastic_kernels.py:975-1008), SURVEY.md Appendix A.2.
 * C15 and C35 are identically zero for every model the reference can build
 * (ADFWI/model/parameters.py:38-44); x + 0*y is exact in IEEE arithmetic, so their terms are
 * dropped.  Parity pin: tests/golden/elastic_*.npz (unmodified reference on CPU); forward records
 * must be BIT-IDENTICAL, gradients within 2e-5 relative L2 (tests/test_oracle_golden.py).
 *
 * Arithmetic: IEEE fp32, one rounding per op, association of the eager PyTorch expressions
 * (SURVEY.md Appendix A.4); build with -ffp-contract=off.
 * Layout: every array is dense [ns][nzp][nxp] / [nzp][nxp] fp32.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef struct {
    int nzp, nxp, ns, nt, nr;
    int NN;            /* fd_order/2: 2 or 3 */
    int free_surface;
    int pml;           /* 1 split-field PML, 0 sponge */
    float dt, dx, dz;  /* f32(dt), f32(dx), f32(dz) */
    float dt_dx, dt_dz;/* f32(dt/dx), f32(dt/dz) (double division) */
    float half_dt;     /* f32(0.5*dt) */
    float fdc[3];      /* DiffCoef(NN,'s') */
} el_dims;

#define AT(a, i, j) (a)[(size_t)(i) * nxp + (j)]

static inline float dxf(const float *a, int nxp, int i, int j, int NN, const float *c)
{ float s = c[0] * (AT(a, i, j + 1) - AT(a, i, j)); for (int k = 1; k < NN; ++k) s = s + c[k] * (AT(a, i, j + k + 1) - AT(a, i, j - k)); return s; }
static inline float dzf(const float *a, int nxp, int i, int j, int NN, const float *c)
{ float s = c[0] * (AT(a, i + 1, j) - AT(a, i, j)); for (int k = 1; k < NN; ++k) s = s + c[k] * (AT(a, i + k + 1, j) - AT(a, i - k, j)); return s; }
static inline float dxb(const float *a, int nxp, int i, int j, int NN, const float *c)
{ float s = c[0] * (AT(a, i, j) - AT(a, i, j - 1)); for (int k = 1; k < NN; ++k) s = s + c[k] * (AT(a, i, j + k) - AT(a, i, j - k - 1)); return s; }
static inline float dzb(const float *a, int nxp, int i, int j, int NN, const float *c)
{ float s = c[0] * (AT(a, i, j) - AT(a, i - 1, j)); for (int k = 1; k < NN; ++k) s = s + c[k] * (AT(a, i + k, j) - AT(a, i - k - 1, j)); return s; }

/* transposes (scatter of m at (i,j) into abar) */
static inline void dxf_T(float *ab, int nxp, int i, int j, int NN, const float *c, float m)
{ for (int k = 0; k < NN; ++k) { AT(ab, i, j + k + 1) += c[k] * m; AT(ab, i, j - k) -= c[k] * m; } }
static inline void dzf_T(float *ab, int nxp, int i, int j, int NN, const float *c, float m)
{ for (int k = 0; k < NN; ++k) { AT(ab, i + k + 1, j) += c[k] * m; AT(ab, i - k, j) -= c[k] * m; } }
static inline void dxb_T(float *ab, int nxp, int i, int j, int NN, const float *c, float m)
{ for (int k = 0; k < NN; ++k) { AT(ab, i, j + k) += c[k] * m; AT(ab, i, j - k - 1) -= c[k] * m; } }
static inline void dzb_T(float *ab, int nxp, int i, int j, int NN, const float *c, float m)
{ for (int k = 0; k < NN; ++k) { AT(ab, i + k, j) += c[k] * m; AT(ab, i - k - 1, j) -= c[k] * m; } }

/* state of one shot */
typedef struct {
    float *txx_x, *txx_z, *tzz_x, *tzz_z, *txz_x, *txz_z, *vx_x, *vx_z, *vz_x, *vz_z; /* PML only */
    float *txx, *tzz, *txz, *vx, *vz;
} el_state;

static void fs_stress(const el_dims *d, float *tzz, float *txz)      /* elastic_kernels.py:380-384 */
{
    const int nxp = d->nxp, h = d->NN + 1;
    for (int j = 0; j < nxp; ++j) {
        AT(tzz, h - 1, j) = 0.f;
        AT(tzz, h - 2, j) = -AT(tzz, h, j);
        AT(txz, h - 2, j) = -AT(txz, h - 1, j);
        AT(txz, h - 3, j) = -AT(txz, h, j);
    }
}

static void fs_velocity(const el_dims *d, float *vx, float *vz)      /* elastic_kernels.py:399-402 */
{
    const int nxp = d->nxp, NN = d->NN, h = NN + 1;
    for (int j = NN; j < nxp - NN; ++j) AT(vz, h - 2, j) = AT(vz, h - 1, j);
    for (int j = NN; j < nxp - NN; ++j)
        AT(vx, h - 2, j) = (((AT(vz, h - 2, j + 1) - AT(vz, h - 2, j)) + AT(vz, h - 1, j + 1)) - AT(vz, h - 1, j)) + AT(vx, h, j);
    for (int j = NN; j < nxp - NN; ++j) AT(vz, h - 3, j) = AT(vz, h - 2, j);
}

/* one forward step of one shot; coef = C11,C13,C33,C55,bx,bz ; b1,b2 = bcx,bcz (PML) or damp,NULL */
static void el_step(const el_dims *d, const float *const *coef, const float *b1, const float *b2,
                    el_state *S, float src_xx, float src_zz, float src_xz, int64_t sz, int64_t sx)
{
    const int nzp = d->nzp, nxp = d->nxp, NN = d->NN;
    const float *c = d->fdc;
    const float *C11 = coef[0], *C13 = coef[1], *C33 = coef[2], *C55 = coef[3], *bx = coef[4], *bz = coef[5];
    const size_t plane = (size_t)nzp * nxp;
    if (d->pml) {
        for (int i = NN; i < nzp - NN; ++i)
            for (int j = NN; j < nxp - NN; ++j) {
                const float pxd = 1.0f + d->half_dt * AT(b1, i, j), pxn = 1.0f - d->half_dt * AT(b1, i, j);
                const float pzd = 1.0f + d->half_dt * AT(b2, i, j), pzn = 1.0f - d->half_dt * AT(b2, i, j);
                const float pxi = 1.0f / pxd, pzi = 1.0f / pzd;
                const float dxb_vx = dxb(S->vx, nxp, i, j, NN, c), dzb_vz = dzb(S->vz, nxp, i, j, NN, c);
                const float dxf_vz = dxf(S->vz, nxp, i, j, NN, c), dzf_vx = dzf(S->vx, nxp, i, j, NN, c);
                AT(S->txx_x, i, j) = (pxn * AT(S->txx_x, i, j) + d->dt_dx * (AT(C11, i, j) * dxb_vx)) * pxi;
                AT(S->txx_z, i, j) = (pzn * AT(S->txx_z, i, j) + d->dt_dz * (AT(C13, i, j) * dzb_vz)) * pzi;
                AT(S->tzz_x, i, j) = (pxn * AT(S->tzz_x, i, j) + d->dt_dx * (AT(C13, i, j) * dxb_vx)) * pxi;
                AT(S->tzz_z, i, j) = (pzn * AT(S->tzz_z, i, j) + d->dt_dz * (AT(C33, i, j) * dzb_vz)) * pzi;
                AT(S->txz_x, i, j) = (pxn * AT(S->txz_x, i, j) + d->dt_dx * (AT(C55, i, j) * dxf_vz)) * pxi;
                AT(S->txz_z, i, j) = (pzn * AT(S->txz_z, i, j) + d->dt_dz * (AT(C55, i, j) * dzf_vx)) * pzi;
            }
        /* moment-tensor source on both halves (:366-372); src_* = (-(MT/2))*src_v */
        AT(S->txx_x, sz, sx) += src_xx; AT(S->txx_z, sz, sx) += src_xx;
        AT(S->tzz_x, sz, sx) += src_zz; AT(S->tzz_z, sz, sx) += src_zz;
        AT(S->txz_x, sz, sx) += src_xz; AT(S->txz_z, sz, sx) += src_xz;
        for (size_t q = 0; q < plane; ++q) {               /* :375-377 */
            S->txx[q] = S->txx_x[q] + S->txx_z[q];
            S->tzz[q] = S->tzz_x[q] + S->tzz_z[q];
            S->txz[q] = S->txz_x[q] + S->txz_z[q];
        }
        if (d->free_surface) fs_stress(d, S->tzz, S->txz);
        for (int i = NN; i < nzp - NN; ++i)
            for (int j = NN; j < nxp - NN; ++j) {          /* :387-394 */
                const float pxd = 1.0f + d->half_dt * AT(b1, i, j), pxn = 1.0f - d->half_dt * AT(b1, i, j);
                const float pzd = 1.0f + d->half_dt * AT(b2, i, j), pzn = 1.0f - d->half_dt * AT(b2, i, j);
                const float dxf_txx = dxf(S->txx, nxp, i, j, NN, c), dzb_txz = dzb(S->txz, nxp, i, j, NN, c);
                const float dxb_txz = dxb(S->txz, nxp, i, j, NN, c), dzf_tzz = dzf(S->tzz, nxp, i, j, NN, c);
                AT(S->vx_x, i, j) = (pxn * AT(S->vx_x, i, j) + ((d->dt * AT(bx, i, j)) * dxf_txx) / d->dx) / pxd;
                AT(S->vx_z, i, j) = (pzn * AT(S->vx_z, i, j) + ((d->dt * AT(bx, i, j)) * dzb_txz) / d->dz) / pzd;
                AT(S->vz_x, i, j) = (pxn * AT(S->vz_x, i, j) + ((d->dt * AT(bz, i, j)) * dxb_txz) / d->dx) / pxd;
                AT(S->vz_z, i, j) = (pzn * AT(S->vz_z, i, j) + ((d->dt * AT(bz, i, j)) * dzf_tzz) / d->dz) / pzd;
            }
        for (size_t q = 0; q < plane; ++q) { S->vx[q] = S->vx_x[q] + S->vx_z[q]; S->vz[q] = S->vz_x[q] + S->vz_z[q]; }
        if (d->free_surface) fs_velocity(d, S->vx, S->vz);
    } else {
        /* the stress update reads only vx,vz and writes only the stresses at (i,j): in place is safe */
        for (int i = NN; i < nzp - NN; ++i)
            for (int j = NN; j < nxp - NN; ++j) {          /* :720-725 */
                const float dxb_vx = dxb(S->vx, nxp, i, j, NN, c), dzb_vz = dzb(S->vz, nxp, i, j, NN, c);
                const float dxf_vz = dxf(S->vz, nxp, i, j, NN, c), dzf_vx = dzf(S->vx, nxp, i, j, NN, c);
                AT(S->txx, i, j) = AT(S->txx, i, j) + d->dt * ((AT(C11, i, j) * dxb_vx) / d->dx + (AT(C13, i, j) * dzb_vz) / d->dz);
                AT(S->tzz, i, j) = AT(S->tzz, i, j) + d->dt * ((AT(C13, i, j) * dxb_vx) / d->dx + (AT(C33, i, j) * dzb_vz) / d->dz);
                AT(S->txz, i, j) = AT(S->txz, i, j) + d->dt * ((AT(C55, i, j) * dxf_vz) / d->dx + (AT(C55, i, j) * dzf_vx) / d->dz);
            }
        AT(S->txx, sz, sx) += src_xx; AT(S->tzz, sz, sx) += src_zz; AT(S->txz, sz, sx) += src_xz;   /* :733-735 */
        if (d->free_surface) fs_stress(d, S->tzz, S->txz);
        /* velocity update needs the OLD stresses only at other cells -> needs no copy either */
        for (int i = NN; i < nzp - NN; ++i)
            for (int j = NN; j < nxp - NN; ++j) {          /* :749-750 */
                const float dxf_txx = dxf(S->txx, nxp, i, j, NN, c), dzb_txz = dzb(S->txz, nxp, i, j, NN, c);
                const float dxb_txz = dxb(S->txz, nxp, i, j, NN, c), dzf_tzz = dzf(S->tzz, nxp, i, j, NN, c);
                AT(S->vx, i, j) += (d->dt * AT(bx, i, j)) * (dxf_txx / d->dx + dzb_txz / d->dz);
                AT(S->vz, i, j) += (d->dt * AT(bz, i, j)) * (dxb_txz / d->dx + dzf_tzz / d->dz);
            }
        if (d->free_surface) fs_velocity(d, S->vx, S->vz);
        for (size_t q = 0; q < plane; ++q) { S->vx[q] *= b1[q]; S->vz[q] *= b1[q]; }   /* :759-760 */
    }
}

static int el_alloc(el_state *S, size_t plane, int pml)
{
    float **f = (float **)S;
    const int n = 15;
    for (int k = 0; k < n; ++k) f[k] = NULL;
    for (int k = (pml ? 0 : 10); k < n; ++k) { f[k] = calloc(plane, sizeof(float)); if (!f[k]) return -1; }
    return 0;
}
static void el_free(el_state *S) { float **f = (float **)S; for (int k = 0; k < 15; ++k) free(f[k]); }

/*
 * Forward modelling from a zero state.
 *   coef[6] = C11,C13,C33,C55,bx,bz planes [nzp][nxp]; b1,b2 = bcx,bcz (PML) or damp,NULL (ABL)
 *   mt [ns][3][3]; src_v [ns][nt]; sx,sz [ns]; rx,rz [nr]  (PADDED indices)
 *   rcv[5] = txx,tzz,txz,vx,vz records [ns][nt][nr]
 *   illum[5] (nullable) [nzp][nxp]: sum over shots of the squared fields at the LAST step of each
 *       of the n_seg chunks of torch.chunk(src_v, n_seg) (elastic_kernels.py:414-418, :1017-1021)
 *   hist (nullable): [ns][nt][5][nzp][nxp] = pre-step vx,vz and post-free-surface txx,tzz,txz
 */
int oracle_elastic_forward(const el_dims *d, const float *const *coef, const float *b1, const float *b2,
                           const float *mt, const float *src_v, const int64_t *sx, const int64_t *sz,
                           const int64_t *rx, const int64_t *rz, float *const *rcv,
                           float *const *illum, int n_seg, float *hist)
{
    const int nzp = d->nzp, nxp = d->nxp, ns = d->ns, nt = d->nt, nr = d->nr;
    const size_t plane = (size_t)nzp * nxp;
    const int csz = (nt + n_seg - 1) / n_seg;
    int err = 0;
    float *ill_part = illum ? calloc((size_t)ns * 5 * plane, sizeof(float)) : NULL;
#pragma omp parallel for schedule(dynamic)
    for (int s = 0; s < ns; ++s) {
        el_state S;
        if (el_alloc(&S, plane, d->pml)) { err = -1; continue; }
        const float *M = mt + (size_t)s * 9;
        for (int t = 0; t < nt; ++t) {
            const float sv = src_v[(size_t)s * nt + t];
            float sxx, szz, sxz;
            if (d->pml) { sxx = (-(M[0] / 2.0f)) * sv; szz = (-(M[8] / 2.0f)) * sv; sxz = (-(M[2] / 2.0f)) * sv; }
            else { const float sc = (float)(-1.0 / 3.0); sxx = (sc * M[0]) * sv; szz = (sc * M[8]) * sv; sxz = (sc * M[2]) * sv; }
            float *H = hist ? hist + ((size_t)s * nt + t) * 5 * plane : NULL;
            if (H) { memcpy(H, S.vx, plane * 4); memcpy(H + plane, S.vz, plane * 4); }
            el_step(d, coef, b1, b2, &S, sxx, szz, sxz, sz[s], sx[s]);
            if (H) {
                /* post-free-surface stresses of this step.  ABL: S.txx.. are the state itself. */
                memcpy(H + 2 * plane, S.txx, plane * 4); memcpy(H + 3 * plane, S.tzz, plane * 4); memcpy(H + 4 * plane, S.txz, plane * 4);
            }
            const float *F[5] = {S.txx, S.tzz, S.txz, S.vx, S.vz};
            for (int k = 0; k < 5; ++k)
                for (int r = 0; r < nr; ++r)
                    rcv[k][((size_t)s * nt + t) * nr + r] = F[k][(size_t)rz[r] * nxp + rx[r]];
            if (ill_part && ((t + 1) % csz == 0 || t == nt - 1))
                for (int k = 0; k < 5; ++k) {
                    float *I = ill_part + ((size_t)s * 5 + k) * plane;
                    for (size_t q = 0; q < plane; ++q) I[q] += F[k][q] * F[k][q];
                }
        }
        el_free(&S);
    }
    if (illum) {
        for (int k = 0; k < 5; ++k)
            for (size_t q = 0; q < plane; ++q) {
                float a = 0.f;
                for (int s = 0; s < ns; ++s) a += ill_part[((size_t)s * 5 + k) * plane + q];
                illum[k][q] = a;
            }
        free(ill_part);
    }
    return err;
}

static void fs_velocity_T(const el_dims *d, float *lvx, float *lvz)   /* 9T of Appendix A.2 */
{
    const int nxp = d->nxp, NN = d->NN, h = NN + 1;
    for (int j = NN; j < nxp - NN; ++j) { AT(lvz, h - 2, j) += AT(lvz, h - 3, j); AT(lvz, h - 3, j) = 0.f; }
    for (int j = NN; j < nxp - NN; ++j) {
        const float q = AT(lvx, h - 2, j);
        AT(lvx, h - 2, j) = 0.f;
        AT(lvz, h - 2, j + 1) += q; AT(lvz, h - 2, j) -= q;
        AT(lvz, h - 1, j + 1) += q; AT(lvz, h - 1, j) -= q;
        AT(lvx, h, j) += q;
    }
    for (int j = NN; j < nxp - NN; ++j) { AT(lvz, h - 1, j) += AT(lvz, h - 2, j); AT(lvz, h - 2, j) = 0.f; }
}

static void fs_stress_T(const el_dims *d, float *mtzz, float *mtxz)    /* 5T */
{
    const int nxp = d->nxp, h = d->NN + 1;
    for (int j = 0; j < nxp; ++j) {
        AT(mtxz, h, j) -= AT(mtxz, h - 3, j); AT(mtxz, h - 3, j) = 0.f;
        AT(mtxz, h - 1, j) -= AT(mtxz, h - 2, j); AT(mtxz, h - 2, j) = 0.f;
        AT(mtzz, h, j) -= AT(mtzz, h - 2, j); AT(mtzz, h - 2, j) = 0.f;
        AT(mtzz, h - 1, j) = 0.f;
    }
}

/*
 * Adjoint sweep.  g_rcv[5] (entries nullable) are the record cotangents; g_coef[6] receive the
 * gradients of C11,C13,C33,C55,bx,bz summed over shots ([nzp][nxp]); g_src (nullable) [ns][nt].
 */
int oracle_elastic_adjoint(const el_dims *d, const float *const *coef, const float *b1, const float *b2,
                           const float *mt, const int64_t *sx, const int64_t *sz,
                           const int64_t *rx, const int64_t *rz, const float *const *g_rcv,
                           const float *hist, float *const *g_coef, float *g_src)
{
    const int nzp = d->nzp, nxp = d->nxp, ns = d->ns, nt = d->nt, nr = d->nr, NN = d->NN;
    const float *c = d->fdc;
    const float *C11 = coef[0], *C13 = coef[1], *C33 = coef[2], *C55 = coef[3], *bx = coef[4], *bz = coef[5];
    const size_t plane = (size_t)nzp * nxp;
    float *gpart = calloc((size_t)ns * 6 * plane, sizeof(float));
    int err = gpart ? 0 : -1;
    if (err) return err;
#pragma omp parallel for schedule(dynamic)
    for (int s = 0; s < ns; ++s) {
        el_state L;             /* cotangents, same slots as the state */
        float *mxx = calloc(plane, 4), *mzz = calloc(plane, 4), *mxz = calloc(plane, 4);
        if (el_alloc(&L, plane, d->pml) || !mxx || !mzz || !mxz) { err = -1; continue; }
        float *G[6];
        for (int k = 0; k < 6; ++k) G[k] = gpart + ((size_t)s * 6 + k) * plane;
        const float *M = mt + (size_t)s * 9;
        for (int t = nt - 1; t >= 0; --t) {
            const float *H = hist + ((size_t)s * nt + t) * 5 * plane;
            const float *vx = H, *vz = H + plane, *txx = H + 2 * plane, *tzz = H + 3 * plane, *txz = H + 4 * plane;
            /* PML: stress sums are rebuilt each step -> scratch cotangents; ABL: they are state */
            float *Mxx = d->pml ? mxx : L.txx, *Mzz = d->pml ? mzz : L.tzz, *Mxz = d->pml ? mxz : L.txz;
            if (d->pml) { memset(mxx, 0, plane * 4); memset(mzz, 0, plane * 4); memset(mxz, 0, plane * 4); }
            float *LF[5] = {Mxx, Mzz, Mxz, L.vx, L.vz};
            for (int k = 0; k < 5; ++k)
                if (g_rcv[k])
                    for (int r = 0; r < nr; ++r)
                        LF[k][(size_t)rz[r] * nxp + rx[r]] += g_rcv[k][((size_t)s * nt + t) * nr + r];
            if (!d->pml) for (size_t q = 0; q < plane; ++q) { L.vx[q] *= b1[q]; L.vz[q] *= b1[q]; }
            if (d->free_surface) fs_velocity_T(d, L.vx, L.vz);
            if (d->pml) {
                for (size_t q = 0; q < plane; ++q) {        /* 8T */
                    L.vx_x[q] += L.vx[q]; L.vx_z[q] += L.vx[q]; L.vx[q] = 0.f;
                    L.vz_x[q] += L.vz[q]; L.vz_z[q] += L.vz[q]; L.vz[q] = 0.f;
                }
                for (int i = NN; i < nzp - NN; ++i)
                    for (int j = NN; j < nxp - NN; ++j) {   /* 7T + 6T */
                        const float pxd = 1.0f + d->half_dt * AT(b1, i, j), pxn = 1.0f - d->half_dt * AT(b1, i, j);
                        const float pzd = 1.0f + d->half_dt * AT(b2, i, j), pzn = 1.0f - d->half_dt * AT(b2, i, j);
                        float q;
                        q = AT(L.vx_x, i, j);
                        AT(G[4], i, j) += q * d->dt * dxf(txx, nxp, i, j, NN, c) / d->dx / pxd;
                        dxf_T(Mxx, nxp, i, j, NN, c, q * d->dt * AT(bx, i, j) / d->dx / pxd);
                        AT(L.vx_x, i, j) = pxn * q / pxd;
                        q = AT(L.vx_z, i, j);
                        AT(G[4], i, j) += q * d->dt * dzb(txz, nxp, i, j, NN, c) / d->dz / pzd;
                        dzb_T(Mxz, nxp, i, j, NN, c, q * d->dt * AT(bx, i, j) / d->dz / pzd);
                        AT(L.vx_z, i, j) = pzn * q / pzd;
                        q = AT(L.vz_x, i, j);
                        AT(G[5], i, j) += q * d->dt * dxb(txz, nxp, i, j, NN, c) / d->dx / pxd;
                        dxb_T(Mxz, nxp, i, j, NN, c, q * d->dt * AT(bz, i, j) / d->dx / pxd);
                        AT(L.vz_x, i, j) = pxn * q / pxd;
                        q = AT(L.vz_z, i, j);
                        AT(G[5], i, j) += q * d->dt * dzf(tzz, nxp, i, j, NN, c) / d->dz / pzd;
                        dzf_T(Mzz, nxp, i, j, NN, c, q * d->dt * AT(bz, i, j) / d->dz / pzd);
                        AT(L.vz_z, i, j) = pzn * q / pzd;
                    }
                if (d->free_surface) fs_stress_T(d, mzz, mxz);
                for (size_t q = 0; q < plane; ++q) {        /* 4T */
                    L.txx_x[q] += mxx[q]; L.txx_z[q] += mxx[q];
                    L.tzz_x[q] += mzz[q]; L.tzz_z[q] += mzz[q];
                    L.txz_x[q] += mxz[q]; L.txz_z[q] += mxz[q];
                }
                if (g_src) {                                /* 3T */
                    const size_t o = (size_t)sz[s] * nxp + sx[s];
                    g_src[(size_t)s * nt + t] = -(M[0] / 2.0f) * (L.txx_x[o] + L.txx_z[o]) - (M[8] / 2.0f) * (L.tzz_x[o] + L.tzz_z[o])
                                                - (M[2] / 2.0f) * (L.txz_x[o] + L.txz_z[o]);
                }
                for (int i = NN; i < nzp - NN; ++i)
                    for (int j = NN; j < nxp - NN; ++j) {   /* 2T + 1T */
                        const float pxd = 1.0f + d->half_dt * AT(b1, i, j), pxn = 1.0f - d->half_dt * AT(b1, i, j);
                        const float pzd = 1.0f + d->half_dt * AT(b2, i, j), pzn = 1.0f - d->half_dt * AT(b2, i, j);
                        const float pxi = 1.0f / pxd, pzi = 1.0f / pzd;
                        const float dxb_vx = dxb(vx, nxp, i, j, NN, c), dzb_vz = dzb(vz, nxp, i, j, NN, c);
                        const float dxf_vz = dxf(vz, nxp, i, j, NN, c), dzf_vx = dzf(vx, nxp, i, j, NN, c);
                        float q;
                        q = AT(L.txx_x, i, j) * pxi; AT(G[0], i, j) += q * d->dt_dx * dxb_vx;
                        dxb_T(L.vx, nxp, i, j, NN, c, q * d->dt_dx * AT(C11, i, j)); AT(L.txx_x, i, j) = pxn * q;
                        q = AT(L.txx_z, i, j) * pzi; AT(G[1], i, j) += q * d->dt_dz * dzb_vz;
                        dzb_T(L.vz, nxp, i, j, NN, c, q * d->dt_dz * AT(C13, i, j)); AT(L.txx_z, i, j) = pzn * q;
                        q = AT(L.tzz_x, i, j) * pxi; AT(G[1], i, j) += q * d->dt_dx * dxb_vx;
                        dxb_T(L.vx, nxp, i, j, NN, c, q * d->dt_dx * AT(C13, i, j)); AT(L.tzz_x, i, j) = pxn * q;
                        q = AT(L.tzz_z, i, j) * pzi; AT(G[2], i, j) += q * d->dt_dz * dzb_vz;
                        dzb_T(L.vz, nxp, i, j, NN, c, q * d->dt_dz * AT(C33, i, j)); AT(L.tzz_z, i, j) = pzn * q;
                        q = AT(L.txz_x, i, j) * pxi; AT(G[3], i, j) += q * d->dt_dx * dxf_vz;
                        dxf_T(L.vz, nxp, i, j, NN, c, q * d->dt_dx * AT(C55, i, j)); AT(L.txz_x, i, j) = pxn * q;
                        q = AT(L.txz_z, i, j) * pzi; AT(G[3], i, j) += q * d->dt_dz * dzf_vx;
                        dzf_T(L.vx, nxp, i, j, NN, c, q * d->dt_dz * AT(C55, i, j)); AT(L.txz_z, i, j) = pzn * q;
                    }
            } else {
                for (int i = NN; i < nzp - NN; ++i)
                    for (int j = NN; j < nxp - NN; ++j) {   /* 6T */
                        float q = AT(L.vx, i, j);
                        AT(G[4], i, j) += q * d->dt * (dxf(txx, nxp, i, j, NN, c) / d->dx + dzb(txz, nxp, i, j, NN, c) / d->dz);
                        float m = q * d->dt * AT(bx, i, j);
                        dxf_T(L.txx, nxp, i, j, NN, c, m / d->dx); dzb_T(L.txz, nxp, i, j, NN, c, m / d->dz);
                        q = AT(L.vz, i, j);
                        AT(G[5], i, j) += q * d->dt * (dxb(txz, nxp, i, j, NN, c) / d->dx + dzf(tzz, nxp, i, j, NN, c) / d->dz);
                        m = q * d->dt * AT(bz, i, j);
                        dxb_T(L.txz, nxp, i, j, NN, c, m / d->dx); dzf_T(L.tzz, nxp, i, j, NN, c, m / d->dz);
                    }
                if (d->free_surface) fs_stress_T(d, L.tzz, L.txz);   /* 4T */
                if (g_src) {
                    const size_t o = (size_t)sz[s] * nxp + sx[s];
                    const float sc = (float)(-1.0 / 3.0);
                    g_src[(size_t)s * nt + t] = sc * (M[0] * L.txx[o] + M[8] * L.tzz[o] + M[2] * L.txz[o]);
                }
                for (int i = NN; i < nzp - NN; ++i)
                    for (int j = NN; j < nxp - NN; ++j) {   /* 2T */
                        const float dxb_vx = dxb(vx, nxp, i, j, NN, c), dzb_vz = dzb(vz, nxp, i, j, NN, c);
                        const float dxf_vz = dxf(vz, nxp, i, j, NN, c), dzf_vx = dzf(vx, nxp, i, j, NN, c);
                        float q = AT(L.txx, i, j) * d->dt;
                        AT(G[0], i, j) += q * dxb_vx / d->dx; AT(G[1], i, j) += q * dzb_vz / d->dz;
                        dxb_T(L.vx, nxp, i, j, NN, c, q * AT(C11, i, j) / d->dx); dzb_T(L.vz, nxp, i, j, NN, c, q * AT(C13, i, j) / d->dz);
                        q = AT(L.tzz, i, j) * d->dt;
                        AT(G[1], i, j) += q * dxb_vx / d->dx; AT(G[2], i, j) += q * dzb_vz / d->dz;
                        dxb_T(L.vx, nxp, i, j, NN, c, q * AT(C13, i, j) / d->dx); dzb_T(L.vz, nxp, i, j, NN, c, q * AT(C33, i, j) / d->dz);
                        q = AT(L.txz, i, j) * d->dt;
                        AT(G[3], i, j) += q * (dxf_vz / d->dx + dzf_vx / d->dz);
                        dxf_T(L.vz, nxp, i, j, NN, c, q * AT(C55, i, j) / d->dx); dzf_T(L.vx, nxp, i, j, NN, c, q * AT(C55, i, j) / d->dz);
                    }
            }
        }
        el_free(&L); free(mxx); free(mzz); free(mxz);
    }
    for (int k = 0; k < 6; ++k)
        for (size_t q = 0; q < plane; ++q) {
            float a = 0.f;
            for (int s = 0; s < ns; ++s) a += gpart[((size_t)s * 6 + k) * plane + q];
            g_coef[k][q] = a;
        }
    free(gpart);
    return err;
}
