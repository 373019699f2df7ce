/*
 * oracle/acoustic_oracle.c -- CPU restatement of ADFWI's iso-acoustic staggered-grid solver.
 *
 * TEST INFRASTRUCTURE ONLY.  This file is the checker, never the product: only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load it.
 * The product path (adfwi_b200/) never links, imports or falls back to anything here.
 *
 * What it restates (paths relative to the upstream reference tree):
 *   forward  : ADFWI/propagator/acoustic_kernels.py:113-174  (time loop of step_forward)
 *   adjoint  : the reverse-mode derivative of that loop (what torch.autograd produces for
 *              loss.backward() through acoustic_kernels.py:268-278), hand-derived; see
 *              SURVEY.md Appendix A.1.
 * Parity pin: tests/golden/acoustic_*.npz hold inputs and outputs of the unmodified reference
 * run on CPU (tests/golden/make_golden.py); tests/test_oracle_golden.py requires the forward
 * records to be BIT-IDENTICAL to them and the gradients to agree to 2e-5 relative L2.
 *
 * Arithmetic: IEEE fp32, one rounding per operation, same association as the eager PyTorch
 * expressions (build with -ffp-contract=off; see oracle/Makefile).
 *
 * Layout: every field is a dense [ns][nzp][nxp] fp32 array (the reference's u is one column
 * narrower and w one row shorter; the extra column/row here is never touched).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define IDX(z, x) ((size_t)(z) * nxp + (x))

typedef struct {
    int nzp, nxp;      /* padded grid (nz+2*nabc, nx+2*nabc)   acoustic_kernels.py:230-231 */
    int ns, nt, nr;    /* shots in this batch, time steps, receivers */
    int free_surface;  /* acoustic_kernels.py:228 */
    int nabc;          /* absorbing layer width (free_surface_start = nabc when free_surface) */
    float dt;          /* f32(dt): the source term is f32(dt)*src_v, acoustic_kernels.py:131 */
    float c1, c2;      /* f32(9/8), f32(-1/24), acoustic_kernels.py:253-254 */
} ac_dims;

/* one forward time step for all shots; S_out/P_out (nullable) receive the stencil sum S and
 * the post-source, post-free-surface pressure of this step (what the adjoint needs). */
static void ac_step(const ac_dims *d, int it,
                    const float *a1, const float *a2, const float *k1, const float *k2,
                    const float *k3, const float *src_v, const int64_t *sx, const int64_t *sz,
                    float *p, float *u, float *w, float *S_out, float *P_out)
{
    const int nzp = d->nzp, nxp = d->nxp, ns = d->ns;
    const int fs = d->free_surface ? d->nabc : 1;
    const float c1 = d->c1, c2 = d->c2;
    const size_t plane = (size_t)nzp * nxp;

#pragma omp parallel for collapse(2) schedule(static)
    for (int s = 0; s < ns; ++s)
        for (int z = fs + 1; z < nzp - 2; ++z) {          /* acoustic_kernels.py:115-128 */
            float *ps = p + s * plane;
            const float *us = u + s * plane, *ws = w + s * plane;
            for (int x = 2; x < nxp - 2; ++x) {
                float S1 = ((us[IDX(z, x)] - us[IDX(z, x - 1)]) + ws[IDX(z, x)]) - ws[IDX(z - 1, x)];
                float S2 = ((us[IDX(z, x + 1)] - us[IDX(z, x - 2)]) + ws[IDX(z + 1, x)]) - ws[IDX(z - 2, x)];
                float S = c1 * S1 + c2 * S2;
                if (S_out) S_out[s * plane + IDX(z, x)] = S;
                ps[IDX(z, x)] = (1.0f - k1[IDX(z, x)]) * ps[IDX(z, x)] - a1[IDX(z, x)] * S;
            }
        }
    for (int s = 0; s < ns; ++s) {                         /* :131-132 source */
        float *ps = p + s * plane;
        ps[IDX(sz[s], sx[s])] = ps[IDX(sz[s], sx[s])] + d->dt * src_v[(size_t)s * d->nt + it];
    }
    if (d->free_surface)                                   /* :135-136 */
        for (int s = 0; s < ns; ++s) {
            float *ps = p + s * plane;
            for (int x = 0; x < nxp; ++x) ps[IDX(fs - 1, x)] = -ps[IDX(fs + 1, x)];
        }
    if (P_out) memcpy(P_out, p, sizeof(float) * plane * ns);

#pragma omp parallel for collapse(2) schedule(static)
    for (int s = 0; s < ns; ++s)
        for (int z = fs; z < nzp - 1; ++z) {               /* :139-160 */
            const float *ps = p + s * plane;
            float *us = u + s * plane, *ws = w + s * plane;
            for (int x = 1; x < nxp - 2; ++x)
                us[IDX(z, x)] = (1.0f - k2[IDX(z, x)]) * us[IDX(z, x)] -
                                a2[IDX(z, x)] * (c1 * (ps[IDX(z, x + 1)] - ps[IDX(z, x)]) +
                                                 c2 * (ps[IDX(z, x + 2)] - ps[IDX(z, x - 1)]));
            if (z < nzp - 2)
                for (int x = 1; x < nxp - 1; ++x)
                    ws[IDX(z, x)] = (1.0f - k3[IDX(z, x)]) * ws[IDX(z, x)] -
                                    a2[IDX(z, x)] * (c1 * (ps[IDX(z + 1, x)] - ps[IDX(z, x)]) +
                                                     c2 * (ps[IDX(z + 2, x)] - ps[IDX(z - 1, x)]));
        }
    if (d->free_surface)                                   /* :163-164 */
        for (int s = 0; s < ns; ++s) {
            float *ws = w + s * plane;
            for (int x = 0; x < nxp; ++x) ws[IDX(fs - 1, x)] = ws[IDX(fs, x)];
        }
}

/*
 * Forward modelling.  p,u,w: in/out state [ns][nzp][nxp] (zero for a fresh run).
 * rcv_*: [ns][nt][nr].  illum_p/u (nullable): [nzp][nxp] accumulators of sum_t sum_s p^2 / u^2
 * (acoustic_kernels.py:172-173; the caller crops them).  hist_S / hist_P (nullable):
 * [nt][ns][nzp][nxp] history consumed by oracle_acoustic_adjoint.
 */
int oracle_acoustic_forward(const ac_dims *d, const float *a1, const float *a2, const float *k1,
                            const float *k2, const float *k3, const float *src_v,
                            const int64_t *sx, const int64_t *sz, const int64_t *rx,
                            const int64_t *rz, float *p, float *u, float *w, float *rcv_p,
                            float *rcv_u, float *rcv_w, float *illum_p, float *illum_u,
                            float *hist_S, float *hist_P)
{
    const int nzp = d->nzp, nxp = d->nxp, ns = d->ns, nt = d->nt, nr = d->nr;
    const size_t plane = (size_t)nzp * nxp;
    for (int it = 0; it < nt; ++it) {
        ac_step(d, it, a1, a2, k1, k2, k3, src_v, sx, sz, p, u, w,
                hist_S ? hist_S + (size_t)it * ns * plane : NULL,
                hist_P ? hist_P + (size_t)it * ns * plane : NULL);
        for (int s = 0; s < ns; ++s)                      /* :167-169 */
            for (int r = 0; r < nr; ++r) {
                size_t o = ((size_t)s * nt + it) * nr + r, c = s * plane + IDX(rz[r], rx[r]);
                rcv_p[o] = p[c]; rcv_u[o] = u[c]; rcv_w[o] = w[c];
            }
        if (illum_p || illum_u) {
#pragma omp parallel for schedule(static)
            for (int z = 0; z < nzp; ++z)
                for (int x = 0; x < nxp; ++x) {
                    float sp = 0.f, su = 0.f;
                    for (int s = 0; s < ns; ++s) {
                        float pv = p[s * plane + IDX(z, x)], uv = u[s * plane + IDX(z, x)];
                        sp += pv * pv; su += uv * uv;
                    }
                    if (illum_p) illum_p[IDX(z, x)] += sp;
                    if (illum_u) illum_u[IDX(z, x)] += su;
                }
        }
    }
    return 0;
}

/*
 * Adjoint sweep: given d(loss)/d(rcv_*) produce d(loss)/d(alpha1), d(loss)/d(alpha2)
 * ([nzp][nxp], summed over shots) and optionally d(loss)/d(src_v) ([ns][nt]).
 * Scatter form of SURVEY.md Appendix A.1, steps 7T..1T.  g_a2 / hist_P / g_src may be NULL.
 */
int oracle_acoustic_adjoint(const ac_dims *d, const float *a1, const float *a2, const float *k1,
                            const float *k2, const float *k3, const int64_t *sx,
                            const int64_t *sz, const int64_t *rx, const int64_t *rz,
                            const float *g_rcv_p, const float *g_rcv_u, const float *g_rcv_w,
                            const float *hist_S, const float *hist_P, float *g_a1, float *g_a2,
                            float *g_src)
{
    const int nzp = d->nzp, nxp = d->nxp, ns = d->ns, nt = d->nt, nr = d->nr;
    const int fs = d->free_surface ? d->nabc : 1;
    const float c1 = d->c1, c2 = d->c2;
    const size_t plane = (size_t)nzp * nxp;
    float *lp = calloc(plane * ns, sizeof(float)), *lu = calloc(plane * ns, sizeof(float)),
          *lw = calloc(plane * ns, sizeof(float));
    /* per-shot gradient planes so that shots can run in parallel; reduced at the end */
    float *ga1 = calloc(plane * ns, sizeof(float));
    float *ga2 = g_a2 ? calloc(plane * ns, sizeof(float)) : NULL;
    if (!lp || !lu || !lw || !ga1 || (g_a2 && !ga2)) return -1;

#pragma omp parallel for schedule(static)
    for (int s = 0; s < ns; ++s) {
        float *Lp = lp + s * plane, *Lu = lu + s * plane, *Lw = lw + s * plane;
        float *G1 = ga1 + s * plane, *G2 = ga2 ? ga2 + s * plane : NULL;
        for (int it = nt - 1; it >= 0; --it) {
            const float *S = hist_S + ((size_t)it * ns + s) * plane;
            const float *P = hist_P ? hist_P + ((size_t)it * ns + s) * plane : NULL;
            for (int r = 0; r < nr; ++r) {                 /* 7T: gather -> scatter-add */
                size_t o = ((size_t)s * nt + it) * nr + r, c = IDX(rz[r], rx[r]);
                if (g_rcv_p) Lp[c] += g_rcv_p[o];
                if (g_rcv_u) Lu[c] += g_rcv_u[o];
                if (g_rcv_w) Lw[c] += g_rcv_w[o];
            }
            if (d->free_surface)                           /* 6T */
                for (int x = 0; x < nxp; ++x) { Lw[IDX(fs, x)] += Lw[IDX(fs - 1, x)]; Lw[IDX(fs - 1, x)] = 0.f; }
            for (int z = fs; z < nzp - 2; ++z)             /* 5T: W */
                for (int x = 1; x < nxp - 1; ++x) {
                    float q = Lw[IDX(z, x)];
                    if (G2) G2[IDX(z, x)] += -q * (c1 * (P[IDX(z + 1, x)] - P[IDX(z, x)]) + c2 * (P[IDX(z + 2, x)] - P[IDX(z - 1, x)]));
                    float m = -a2[IDX(z, x)] * q;
                    Lp[IDX(z + 1, x)] += c1 * m; Lp[IDX(z, x)] -= c1 * m;
                    Lp[IDX(z + 2, x)] += c2 * m; Lp[IDX(z - 1, x)] -= c2 * m;
                    Lw[IDX(z, x)] = (1.0f - k3[IDX(z, x)]) * q;
                }
            for (int z = fs; z < nzp - 1; ++z)             /* 4T: U */
                for (int x = 1; x < nxp - 2; ++x) {
                    float q = Lu[IDX(z, x)];
                    if (G2) G2[IDX(z, x)] += -q * (c1 * (P[IDX(z, x + 1)] - P[IDX(z, x)]) + c2 * (P[IDX(z, x + 2)] - P[IDX(z, x - 1)]));
                    float m = -a2[IDX(z, x)] * q;
                    Lp[IDX(z, x + 1)] += c1 * m; Lp[IDX(z, x)] -= c1 * m;
                    Lp[IDX(z, x + 2)] += c2 * m; Lp[IDX(z, x - 1)] -= c2 * m;
                    Lu[IDX(z, x)] = (1.0f - k2[IDX(z, x)]) * q;
                }
            if (d->free_surface)                           /* 3T */
                for (int x = 0; x < nxp; ++x) { Lp[IDX(fs + 1, x)] -= Lp[IDX(fs - 1, x)]; Lp[IDX(fs - 1, x)] = 0.f; }
            if (g_src) g_src[(size_t)s * nt + it] = d->dt * Lp[IDX(sz[s], sx[s])];   /* 2T */
            for (int z = fs + 1; z < nzp - 2; ++z)         /* 1T: P */
                for (int x = 2; x < nxp - 2; ++x) {
                    float q = Lp[IDX(z, x)];
                    G1[IDX(z, x)] += -q * S[IDX(z, x)];
                    float m = -a1[IDX(z, x)] * q;
                    Lu[IDX(z, x)] += c1 * m; Lu[IDX(z, x - 1)] -= c1 * m;
                    Lw[IDX(z, x)] += c1 * m; Lw[IDX(z - 1, x)] -= c1 * m;
                    Lu[IDX(z, x + 1)] += c2 * m; Lu[IDX(z, x - 2)] -= c2 * m;
                    Lw[IDX(z + 1, x)] += c2 * m; Lw[IDX(z - 2, x)] -= c2 * m;
                    Lp[IDX(z, x)] = (1.0f - k1[IDX(z, x)]) * q;
                }
        }
    }
    for (size_t i = 0; i < plane; ++i) {
        float s1 = 0.f, s2 = 0.f;
        for (int s = 0; s < ns; ++s) { s1 += ga1[s * plane + i]; if (ga2) s2 += ga2[s * plane + i]; }
        g_a1[i] = s1;
        if (g_a2) g_a2[i] = s2;
    }
    free(lp); free(lu); free(lw); free(ga1); free(ga2);
    return 0;
}

/* bench.py helper: torchrun exports OMP_NUM_THREADS=1; the reference arm wants all host threads */
#ifdef _OPENMP
#include <omp.h>
int oracle_set_threads(int n) { if (n > 0) omp_set_num_threads(n); return omp_get_max_threads(); }
#else
int oracle_set_threads(int n) { (void)n; return 1; }
#endif
