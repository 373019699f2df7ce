// objective.cu -- the consumers of the records and of the model planes right after the propagator:
//
//   adfwi_misfit_*          per-trace max-abs normalisation of the synthetic records (ADFWI/fwi/acoustic_fwi.py:149-150,
//                           elastic_fwi.py:224-255) fused with the misfit and its adjoint source:
//                             kind 0  Misfit_waveform_L2         (ADFWI/fwi/misfit/L2.py:22-28)
//                             kind 1  Misfit_global_correlation  (ADFWI/fwi/misfit/GlobalCorrelation.py:43-70)
//                           Records are the largest tensors on the path ([ns][nt][nr], a trace = fixed (s,r), stride nr in time).  The
//                           reference makes ~10 eager passes over them (abs, max, divide, subtract, square, scale, sum, sqrt ... and the
//                           autograd mirror of each); here: ONE pass that reduces every trace to six sums and its maximum (double
//                           accumulators), a per-trace closed form for the loss and for the coefficients of the adjoint source
//                               g[s,t,r] = c_obs * obs + c_syn * syn + c_1  (+ c_max at the trace's arg-max sample, the normalisation's own
//                               derivative),
//                           and ONE pass that writes g.  20 B per sample instead of > 100 B.
//   adfwi_regularization_*  TV / Tikhonov regularisers, 1st and 2nd order (ADFWI/fwi/regularization/{tv,tikhonov}_{1,2}order.py): the
//                           reference multiplies every model column and row with a dense difference matrix in Python loops
//                           (nx + nz small matmuls per evaluation, on the HOST when the model lives on a GPU: its work arrays are CPU
//                           tensors); here: one stencil kernel for the value and one for the gradient.
#include "common.cuh"
#ifndef ADFWI_HOST_EMUL
#include <math.h>

namespace adfwi {
namespace {

constexpr int MF_TPB = 128;

struct MfPlan {
    int ns, nt, nr, kind, normalize, tch, tlen;
    double dt;
    size_t ntr;
    // workspace
    float* pmax; int* pidx; double* psum;      // partials [tch][ntr] (psum: 5 planes of that)
    double* coef;                               // [4][ntr]: c_obs, c_syn, c_1, c_max
    int* tmax;                                  // [ntr]
    double* loss;                               // 1
    size_t bytes;
};

int mf_make_plan(const adfwi_misfit_desc* d, void* ws, MfPlan* P)
{
    if (!d) return ADFWI_E_NULL;
    if (d->ns < 1 || d->nt < 2 || d->nr < 1 || d->kind < 0 || d->kind > 1) return ADFWI_E_DIMS;
    P->ns = d->ns; P->nt = d->nt; P->nr = d->nr; P->kind = d->kind; P->normalize = d->normalize ? 1 : 0; P->dt = d->dt;
    P->ntr = (size_t)d->ns * d->nr;
    // enough threads for the streaming pass: split the time axis when there are few traces
    int tch = (int)((262144 + P->ntr - 1) / P->ntr);
    if (tch < 1) tch = 1;
    if (tch > 64) tch = 64;
    if (tch > d->nt / 16) tch = d->nt / 16 > 0 ? d->nt / 16 : 1;
    P->tlen = cdiv(d->nt, tch);
    P->tch = cdiv(d->nt, P->tlen);
    Carver cv(ws);
    P->pmax = cv.take<float>((size_t)P->tch * P->ntr);
    P->pidx = cv.take<int>((size_t)P->tch * P->ntr);
    P->psum = cv.take<double>((size_t)5 * P->tch * P->ntr);
    P->coef = cv.take<double>(4 * P->ntr);
    P->tmax = cv.take<int>(P->ntr);
    P->loss = cv.take<double>(1);
    P->bytes = cv.off;
    return ADFWI_OK;
}

// pass 1: per (shot, time chunk, receiver): max |syn| with its first arg-max, and the sums obs^2, obs*syn, syn^2, obs, syn
__global__ void __launch_bounds__(MF_TPB)
mf_stats(int nt, int nr, int tlen, size_t ntr, const float* __restrict__ syn, const float* __restrict__ obs,
         float* __restrict__ pmax, int* __restrict__ pidx, double* __restrict__ psum)
{
    const int r = blockIdx.x * MF_TPB + threadIdx.x, c = blockIdx.y, s = blockIdx.z;
    if (r >= nr) return;
    const int t0 = c * tlen, t1 = min(t0 + tlen, nt);
    const float* S = syn + ((size_t)s * nt + t0) * nr + r;
    const float* O = obs + ((size_t)s * nt + t0) * nr + r;
    float m = -1.f; int im = t0;
    double soo = 0.0, sos = 0.0, sss = 0.0, so = 0.0, ss = 0.0;
    for (int t = t0; t < t1; ++t, S += nr, O += nr) {
        const float y = __ldcs(S), o = __ldcs(O);
        const float a = fabsf(y);
        if (a > m || (a != a && m == m)) { m = a; im = t; }       // first maximum; a NaN takes over and stays (torch.max propagates NaN)
        const double yd = (double)y, od = (double)o;
        soo += od * od; sos += od * yd; sss += yd * yd; so += od; ss += yd;
    }
    const size_t q = (size_t)c * ntr + (size_t)s * nr + r;
    const size_t pl = (size_t)gridDim.y * ntr;
    pmax[q] = m; pidx[q] = im;
    psum[q] = soo; psum[pl + q] = sos; psum[2 * pl + q] = sss; psum[3 * pl + q] = so; psum[4 * pl + q] = ss;
}

// per trace: combine the chunk partials, closed-form loss and adjoint-source coefficients
__global__ void __launch_bounds__(256)
mf_finalize(int nt, int nr, int tch, size_t ntr, int kind, int normalize, double dt, const float* __restrict__ syn,
            const float* __restrict__ pmax, const int* __restrict__ pidx, const double* __restrict__ psum,
            double* __restrict__ coef, int* __restrict__ tmax, double* __restrict__ loss)
{
    __shared__ double sh[256];
    const size_t q = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    double l = 0.0;
    if (q < ntr) {
        float m = -1.f; int im = 0;
        double soo = 0.0, sos = 0.0, sss = 0.0, so = 0.0, ss = 0.0;
        const size_t pl = (size_t)tch * ntr;
        for (int c = 0; c < tch; ++c) {
            const size_t p = (size_t)c * ntr + q;
            const float a = pmax[p];
            if (a > m || (a != a && m == m)) { m = a; im = pidx[p]; }
            soo += psum[p]; sos += psum[pl + p]; sss += psum[2 * pl + p]; so += psum[3 * pl + p]; ss += psum[4 * pl + p];
        }
        double md = 1.0, sg = 0.0;
        if (normalize) {
            md = (double)m;
            const int s = (int)(q / nr), r = (int)(q - (size_t)s * nr);
            const float ym = syn[((size_t)s * nt + im) * nr + r];
            sg = ym > 0.f ? 1.0 : (ym < 0.f ? -1.0 : 0.0);
        }
        const double T = (double)nt;
        double co = 0.0, cs = 0.0, c1 = 0.0, cm = 0.0;
        if (kind == 0) {
            // y = syn/m, r = obs - y, n = sqrt(dt * sum r^2);  L = n;  dL/dy_t = -dt r_t / n
            double n2 = dt * (soo - 2.0 * sos / md + sss / (md * md));
            if (n2 < 0.0) n2 = 0.0;                    // cancellation when obs == y to round-off
            const double n = sqrt(n2);
            l = n;
            co = -dt / (n * md); cs = dt / (n * md * md);
            cm = sg * (dt / (n * md * md)) * (sos - sss / md);
        } else {
            // a = obs/|obs|, b = y/|y|, corr = mean(a b) / (sqrt(var a var b) + 1e-8) (unbiased variances), L = -corr dt
            const double No = sqrt(soo), Sy = ss / md, Ny2 = sss / (md * md), Ny = sqrt(Ny2), Soy = sos / md;
            const double Pm = Soy / (T * No * Ny);
            const double Qa = (1.0 - so * so / (T * soo)) / (T - 1.0), Qb = (1.0 - Sy * Sy / (T * Ny2)) / (T - 1.0);
            const double D = sqrt(Qa * Qb) + 1e-8;
            const double corr = Pm / D;
            if (corr == corr) {                        // the reference replaces a NaN correlation by 0 (no gradient then)
                l = -corr * dt;
                const double K = (Pm / (D * D)) * 0.5 * sqrt(Qa / Qb), f = 2.0 / ((T - 1.0) * T);
                const double cob = 1.0 / (D * T * No * Ny), cy = -Pm / (D * Ny2) - K * f * Sy * Sy / (Ny2 * Ny2), cc = K * f * Sy / Ny2;
                co = -dt * cob / md; cs = -dt * cy / (md * md); c1 = -dt * cc / md;
                cm = (sg / (md * md)) * dt * (cob * sos + cy * sss / md + cc * ss);
            }
        }
        if (!normalize) cm = 0.0;
        coef[q] = co; coef[ntr + q] = cs; coef[2 * ntr + q] = c1; coef[3 * ntr + q] = cm;
        tmax[q] = normalize ? im : -1;
    }
    sh[threadIdx.x] = l;
    __syncthreads();
    for (int k = blockDim.x / 2; k > 0; k >>= 1) {
        if ((int)threadIdx.x < k) sh[threadIdx.x] += sh[threadIdx.x + k];
        __syncthreads();
    }
    if (threadIdx.x == 0) atomicAdd(loss, sh[0]);
}

__global__ void mf_loss_out(const double* __restrict__ loss, float* __restrict__ out) { out[0] = (float)loss[0]; }

// pass 2: adjoint source g = scale * (c_obs obs + c_syn syn + c_1 [+ c_max at the arg-max sample])
__global__ void __launch_bounds__(MF_TPB)
mf_adjoint(int nt, int nr, int tlen, size_t ntr, const float* __restrict__ syn, const float* __restrict__ obs,
           const double* __restrict__ coef, const int* __restrict__ tmax, const float* __restrict__ scale, float* __restrict__ g)
{
    const int r = blockIdx.x * MF_TPB + threadIdx.x, c = blockIdx.y, s = blockIdx.z;
    if (r >= nr) return;
    const size_t q = (size_t)s * nr + r;
    const double sc = scale ? (double)scale[0] : 1.0;
    const double co = sc * coef[q], cs = sc * coef[ntr + q], c1 = sc * coef[2 * ntr + q], cm = sc * coef[3 * ntr + q];
    const int im = tmax[q];
    const int t0 = c * tlen, t1 = min(t0 + tlen, nt);
    size_t o = ((size_t)s * nt + t0) * nr + r;
    for (int t = t0; t < t1; ++t, o += nr) {
        double v = co * (double)__ldcs(obs + o) + cs * (double)__ldcs(syn + o) + c1;
        if (t == im) v += cm;
        __stcs(g + o, (float)v);
    }
}

// ---------------------------------------------------------------------------------------------------------------------------
// regularisers.  kind 0 TV 1st order, 1 Tikhonov 1st order, 2 TV 2nd order, 3 Tikhonov 2nd order.
//   1st order (tv_1order.py:29-41): dz m = (m[i,j] - m[i+1,j]) / dz for i < nz-1, 0 on the last row; dx m likewise along x
//   2nd order (tv_2order.py:29-45): dz m = (m[i-1,j] - 2 m[i,j] + m[i+1,j]) / dz for 0 < i < nz-1, 0 on the first / last row
//   (dz, dx in km: the reference divides the grid spacing by 1000; the 2nd-order operator is divided by the spacing ONCE, as upstream)
//   TV:       sum(alphax |dx m| + alphaz |dz m|)                  Tikhonov: sqrt(sum(alphax (dx m)^2 + alphaz (dz m)^2))
// ---------------------------------------------------------------------------------------------------------------------------
struct RgGeom { int nz, nx, kind; float rdx, rdz, ax, az; };

__device__ __forceinline__ void rg_derivs(const RgGeom& g, const float* __restrict__ m, int i, int j, float& dz, float& dx)
{
    const size_t c = (size_t)i * g.nx + j;
    const float m0 = m[c];
    if (g.kind < 2) {
        dz = i < g.nz - 1 ? (m0 - m[c + g.nx]) * g.rdz : 0.f;
        dx = j < g.nx - 1 ? (m0 - m[c + 1]) * g.rdx : 0.f;
    } else {
        dz = (i > 0 && i < g.nz - 1) ? ((m[c - g.nx] - 2.0f * m0) + m[c + g.nx]) * g.rdz : 0.f;
        dx = (j > 0 && j < g.nx - 1) ? ((m[c - 1] - 2.0f * m0) + m[c + 1]) * g.rdx : 0.f;
    }
}

__global__ void __launch_bounds__(256)
rg_value(const RgGeom g, const float* __restrict__ m, double* __restrict__ acc)
{
    __shared__ double sh[256];
    const int j = blockIdx.x * blockDim.x + threadIdx.x, i = blockIdx.y;
    double v = 0.0;
    if (j < g.nx) {
        float dz, dx;
        rg_derivs(g, m, i, j, dz, dx);
        v = (g.kind & 1) ? (double)(g.ax * dx * dx) + (double)(g.az * dz * dz) : (double)(g.ax * fabsf(dx)) + (double)(g.az * fabsf(dz));
    }
    sh[threadIdx.x] = v;
    __syncthreads();
    for (int k = blockDim.x / 2; k > 0; k >>= 1) {
        if ((int)threadIdx.x < k) sh[threadIdx.x] += sh[threadIdx.x + k];
        __syncthreads();
    }
    if (threadIdx.x == 0) atomicAdd(acc, sh[0]);
}
__global__ void rg_value_out(int tikhonov, const double* __restrict__ acc, float* __restrict__ out) { out[0] = (float)(tikhonov ? sqrt(acc[0]) : acc[0]); }

// gradient: transpose of the difference operators applied to w = d(value)/d(derivative)
//   TV: w = alpha * sign(d)      Tikhonov: w = alpha * d / value
__device__ __forceinline__ void rg_w(const RgGeom& g, const float* __restrict__ m, int i, int j, float inv_value, float& wz, float& wx)
{
    if (i < 0 || i >= g.nz || j < 0 || j >= g.nx) { wz = 0.f; wx = 0.f; return; }
    float dz, dx;
    rg_derivs(g, m, i, j, dz, dx);
    if (g.kind & 1) { wz = g.az * dz * inv_value; wx = g.ax * dx * inv_value; }
    else {
        wz = g.az * (dz > 0.f ? 1.f : (dz < 0.f ? -1.f : 0.f));
        wx = g.ax * (dx > 0.f ? 1.f : (dx < 0.f ? -1.f : 0.f));
    }
}
__global__ void __launch_bounds__(256)
rg_grad(const RgGeom g, const float* __restrict__ m, const double* __restrict__ acc, const float* __restrict__ scale, float* __restrict__ out)
{
    const int j = blockIdx.x * blockDim.x + threadIdx.x, i = blockIdx.y;
    if (j >= g.nx) return;
    const float inv_value = (g.kind & 1) ? (float)(1.0 / sqrt(acc[0])) : 1.f;
    float wz0, wx0, wzm, wxm, t;
    rg_w(g, m, i, j, inv_value, wz0, wx0);
    float gz, gx;
    if (g.kind < 2) {           // d[i] = (m[i] - m[i+1]) r :  dL/dm[i] = (w[i] - w[i-1]) r
        rg_w(g, m, i - 1, j, inv_value, wzm, t);
        rg_w(g, m, i, j - 1, inv_value, t, wxm);
        gz = (wz0 - wzm) * g.rdz; gx = (wx0 - wxm) * g.rdx;
    } else {                    // d[i] = (m[i-1] - 2 m[i] + m[i+1]) r :  dL/dm[i] = (w[i+1] - 2 w[i] + w[i-1]) r
        float wzp, wxp;
        rg_w(g, m, i - 1, j, inv_value, wzm, t);
        rg_w(g, m, i + 1, j, inv_value, wzp, t);
        rg_w(g, m, i, j - 1, inv_value, t, wxm);
        rg_w(g, m, i, j + 1, inv_value, t, wxp);
        gz = ((wzm - 2.0f * wz0) + wzp) * g.rdz; gx = ((wxm - 2.0f * wx0) + wxp) * g.rdx;
    }
    out[(size_t)i * g.nx + j] = (scale ? scale[0] : 1.f) * (gz + gx);
}

}  // namespace
}  // namespace adfwi

using namespace adfwi;

extern "C" size_t adfwi_misfit_workspace_bytes(const adfwi_misfit_desc* desc)
{
    MfPlan P;
    return mf_make_plan(desc, nullptr, &P) == ADFWI_OK ? P.bytes : 0;
}

extern "C" int adfwi_misfit_forward(const adfwi_misfit_desc* desc, const float* syn, const float* obs, float* loss,
                                    void* workspace, size_t workspace_bytes, void* stream)
{
    ADFWI_NVTX("adfwi_misfit_forward");
    MfPlan P;
    int rc = mf_make_plan(desc, workspace, &P);
    if (rc) return rc;
    if (!syn || !obs || !loss || !workspace) return ADFWI_E_NULL;
    if (workspace_bytes < P.bytes) return ADFWI_E_WORKSPACE;
    cudaStream_t st = (cudaStream_t)stream;
    ADFWI_CUDA(cudaMemsetAsync(P.loss, 0, sizeof(double), st));
    mf_stats<<<dim3(cdiv(P.nr, MF_TPB), P.tch, P.ns), MF_TPB, 0, st>>>(P.nt, P.nr, P.tlen, P.ntr, syn, obs, P.pmax, P.pidx, P.psum);
    ADFWI_LAUNCH_CHECK();
    mf_finalize<<<(unsigned)((P.ntr + 255) / 256), 256, 0, st>>>(P.nt, P.nr, P.tch, P.ntr, P.kind, P.normalize, P.dt, syn, P.pmax, P.pidx, P.psum,
                                                              P.coef, P.tmax, P.loss);
    ADFWI_LAUNCH_CHECK();
    mf_loss_out<<<1, 1, 0, st>>>(P.loss, loss);
    ADFWI_LAUNCH_CHECK();
    return ADFWI_OK;
}

extern "C" int adfwi_misfit_adjoint_source(const adfwi_misfit_desc* desc, const float* syn, const float* obs, const float* grad_loss,
                                           float* g_syn, void* workspace, size_t workspace_bytes, void* stream)
{
    ADFWI_NVTX("adfwi_misfit_adjoint_source");
    MfPlan P;
    int rc = mf_make_plan(desc, workspace, &P);
    if (rc) return rc;
    if (!syn || !obs || !g_syn || !workspace) return ADFWI_E_NULL;
    if (workspace_bytes < P.bytes) return ADFWI_E_WORKSPACE;
    cudaStream_t st = (cudaStream_t)stream;
    mf_adjoint<<<dim3(cdiv(P.nr, MF_TPB), P.tch, P.ns), MF_TPB, 0, st>>>(P.nt, P.nr, P.tlen, P.ntr, syn, obs, P.coef, P.tmax, grad_loss, g_syn);
    ADFWI_LAUNCH_CHECK();
    return ADFWI_OK;
}

static int rg_geom(const adfwi_regularization_desc* d, RgGeom* g)
{
    if (!d) return ADFWI_E_NULL;
    if (d->nz < 3 || d->nx < 3 || d->kind < 0 || d->kind > 3) return ADFWI_E_DIMS;
    g->nz = d->nz; g->nx = d->nx; g->kind = d->kind;
    g->rdx = (float)(1.0 / (d->dx / 1000.0)); g->rdz = (float)(1.0 / (d->dz / 1000.0));
    g->ax = (float)d->alphax; g->az = (float)d->alphaz;
    return ADFWI_OK;
}

extern "C" size_t adfwi_regularization_workspace_bytes(const adfwi_regularization_desc* desc)
{
    RgGeom g;
    return rg_geom(desc, &g) == ADFWI_OK ? 256 : 0;
}

extern "C" int adfwi_regularization_forward(const adfwi_regularization_desc* desc, const float* m, float* value,
                                            void* workspace, size_t workspace_bytes, void* stream)
{
    ADFWI_NVTX("adfwi_regularization_forward");
    RgGeom g;
    int rc = rg_geom(desc, &g);
    if (rc) return rc;
    if (!m || !value || !workspace) return ADFWI_E_NULL;
    if (workspace_bytes < 256) return ADFWI_E_WORKSPACE;
    cudaStream_t st = (cudaStream_t)stream;
    double* acc = (double*)workspace;
    ADFWI_CUDA(cudaMemsetAsync(acc, 0, sizeof(double), st));
    rg_value<<<dim3(cdiv(g.nx, 256), g.nz), 256, 0, st>>>(g, m, acc);
    ADFWI_LAUNCH_CHECK();
    rg_value_out<<<1, 1, 0, st>>>(g.kind & 1, acc, value);
    ADFWI_LAUNCH_CHECK();
    return ADFWI_OK;
}

extern "C" int adfwi_regularization_backward(const adfwi_regularization_desc* desc, const float* m, const float* grad_value, float* g_m,
                                             void* workspace, size_t workspace_bytes, void* stream)
{
    ADFWI_NVTX("adfwi_regularization_backward");
    RgGeom g;
    int rc = rg_geom(desc, &g);
    if (rc) return rc;
    if (!m || !g_m || !workspace) return ADFWI_E_NULL;
    if (workspace_bytes < 256) return ADFWI_E_WORKSPACE;
    cudaStream_t st = (cudaStream_t)stream;
    rg_grad<<<dim3(cdiv(g.nx, 256), g.nz), 256, 0, st>>>(g, m, (const double*)workspace, grad_value, g_m);
    ADFWI_LAUNCH_CHECK();
    return ADFWI_OK;
}
#endif  // !ADFWI_HOST_EMUL
