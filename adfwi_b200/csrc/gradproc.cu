// gradproc.cu -- gradient post-processing on the device.
//
// Semantics: ADFWI/propagator/gradient_process.py:30-135 (smooth2d, grad_taper, GradProcessor.forward), which the
// reference runs in float64 numpy/scipy on the host between two device<->host copies (ADFWI/fwi/acoustic_fwi.py:171-177).
// Same arithmetic in float64 here.  smooth2d's (2*span+1)^2 Gaussian has a diagonal covariance (:40), so it is applied
// as its two 1-D factors (row pass, column pass) and the normalising convolution of ones is the product of two 1-D
// partial sums; against scipy's direct 2-D sum this only changes the summation order (parity 1e-12, tests/test_gradproc_gpu.py).
// HBM-bound: a plane is read and written once per pass (the 2*span+1 taps of a pass hit L1/L2).
#include "common.cuh"
#ifndef ADFWI_HOST_EMUL
#include <math.h>

namespace adfwi {
namespace {

constexpr int GP_MAX_SPAN = 512;
constexpr int GP_MAX_BLOCKS = 256;          // partial maxima of the two-stage reductions

struct GPlan { double *g, *t0, *t1, *f, *wz, *wx, *scal; size_t bytes; };

GPlan gp_plan(int nz, int nx, void* ws)
{
    Carver cv(ws);
    GPlan P;
    const size_t n = (size_t)nz * nx;
    P.g = cv.take<double>(n); P.t0 = cv.take<double>(n); P.t1 = cv.take<double>(n);
    P.f = cv.take<double>(2 * GP_MAX_SPAN + 1);
    P.wz = cv.take<double>(nz); P.wx = cv.take<double>(nx);
    P.scal = cv.take<double>(8 + GP_MAX_BLOCKS);
    P.bytes = cv.off;
    return P;
}

// 1-D factor of the filter (:36-42): taps t = -2*span + 2k, exp(-t^2 / (2 span^2)), unit sum
__global__ void gp_factor(int span, double* __restrict__ f)
{
    if (blockIdx.x || threadIdx.x) return;
    const int m = 2 * span + 1;
    double acc = 0.0;
    for (int k = 0; k < m; ++k) {
        const double t = (-2.0 * span + 2.0 * k) / (double)span;
        const double e = exp(-0.5 * t * t);
        f[k] = e; acc += e;
    }
    for (int k = 0; k < m; ++k) f[k] /= acc;
}
// partial sums of the factor inside the plane: w[i] = sum of f[k] over 0 <= i + k - span < n  ('same', zero boundary)
__global__ void gp_partial(int n, int span, const double* __restrict__ f, double* __restrict__ w)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int k0 = max(0, span - i), k1 = min(2 * span, n - 1 - i + span);
    double acc = 0.0;
    for (int k = k0; k <= k1; ++k) acc += f[k];
    w[i] = acc;
}
__global__ void gp_conv_rows(int nz, int nx, int span, const double* __restrict__ f, const double* __restrict__ in, double* __restrict__ out)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x, z = blockIdx.y;
    if (x >= nx) return;
    const int k0 = max(0, span - x), k1 = min(2 * span, nx - 1 - x + span);
    const double* r = in + (size_t)z * nx + x - span;
    double acc = 0.0;
    for (int k = k0; k <= k1; ++k) acc += f[k] * r[k];
    out[(size_t)z * nx + x] = acc;
}
// column pass + division by the convolution of ones (:46-47)
__global__ void gp_conv_cols(int nz, int nx, int span, const double* __restrict__ f, const double* __restrict__ wz, const double* __restrict__ wx,
                             const double* __restrict__ in, double* __restrict__ out)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x, z = blockIdx.y;
    if (x >= nx) return;
    const int k0 = max(0, span - z), k1 = min(2 * span, nz - 1 - z + span);
    double acc = 0.0;
    for (int k = k0; k <= k1; ++k) acc += f[k] * in[(size_t)(z + k - span) * nx + x];
    out[(size_t)z * nx + x] = acc / (wz[z] * wx[x]);
}
// numpy's max() propagates NaN (a diverged simulation must stay visible: the reference then returns an all-NaN
// gradient); fmax() would drop it
__device__ __forceinline__ double nanmax(double m, double a) { return (a != a || m != m) ? NAN : fmax(m, a); }
// two-stage maximum.  mode 0: max(v), 1: max(|v|), 2: max(v + 1e-5), 3: max of the partial results
__global__ void gp_max(size_t n, int mode, const double* __restrict__ v, double* __restrict__ res)
{
    __shared__ double sh[256];
    double m = -INFINITY;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        double a = v[i];
        a = mode == 1 ? fabs(a) : (mode == 2 ? a + 1e-5 : a);
        m = nanmax(m, a);
    }
    sh[threadIdx.x] = m;
    __syncthreads();
    for (int s = blockDim.x / 2; s > 0; s >>= 1) {
        if ((int)threadIdx.x < s) sh[threadIdx.x] = nanmax(sh[threadIdx.x], sh[threadIdx.x + s]);
        __syncthreads();
    }
    if (threadIdx.x == 0) res[blockIdx.x] = sh[0];
}
__global__ void gp_load_f32(size_t n, const float* __restrict__ in, double* __restrict__ out)
{
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = (double)in[i];
}
// g *= w, stored in g's dtype (:98, :107: numpy's in-place product of a float32 array rounds back to float32)
__global__ void gp_mul(size_t n, int is32, const double* __restrict__ w, double* __restrict__ g)
{
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double v = g[i] * w[i];
    g[i] = is32 ? (double)(float)v : v;
}
__global__ void gp_taper_marine(int nz, int nx, int size, double* __restrict__ w)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x, z = blockIdx.y;
    if (x < nx) w[(size_t)z * nx + x] = z < size ? 0.0 : 1.0;
}
// land branch before smoothing (:61-64): falling half of hamming(2*size) on the first `size` columns of every row
__global__ void gp_taper_land_init(int nz, int nx, int size, double* __restrict__ w)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x, z = blockIdx.y;
    if (x >= nx) return;
    double v = 0.0;
    if (x < size) {
        const int M = 2 * size, n = size + x;
        v = M == 1 ? 1.0 : 0.54 - 0.46 * cos(2.0 * M_PI * n / (double)(M - 1));
    }
    w[(size_t)z * nx + x] = v;
}
// (:66-70) t /= max; t *= 1-thred; t = 1 - t; t = t*t
__global__ void gp_taper_land_finish(size_t n, double thred, const double* __restrict__ mx, double* __restrict__ w)
{
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double t = w[i] / *mx;
    t = t * (1.0 - thred);
    t = -t + 1.0;
    w[i] = t * t;
}
// (:119-122) p = p / max(p + 1e-5); p = max(p, 1e-4); g = g / p^2
__global__ void gp_precond(size_t n, const double* __restrict__ p, const double* __restrict__ mx, double* __restrict__ g)
{
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double q = p[i] / *mx;
    q = q < 0.0001 ? 0.0001 : q;
    g[i] = g[i] / (q * q);
}
// copy of the smoothed rows back into g (:128), rounded to g's dtype
__global__ void gp_store(size_t n, int is32, const double* __restrict__ in, double* __restrict__ g)
{
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) g[i] = is32 ? (double)(float)in[i] : in[i];
}
// (:135) vmax * g / max|g|, in float32 arithmetic when the plane is still float32
__global__ void gp_norm(size_t n, int is32, double vmax, const double* __restrict__ mx, const double* __restrict__ g, double* __restrict__ out)
{
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    if (is32) {
        const float t = (float)vmax * (float)g[i];
        out[i] = (double)(t / (float)*mx);
    } else {
        out[i] = vmax * g[i] / *mx;
    }
}

inline unsigned nb(size_t n) { return (unsigned)((n + 255) / 256); }

// res[0] = maximum over v[0..n) in the given mode (partials parked behind the scalars of the plan)
int gp_reduce_max(const GPlan& P, cudaStream_t st, size_t n, int mode, const double* v, double* res)
{
    unsigned blocks = nb(n);
    if (blocks > (unsigned)GP_MAX_BLOCKS) blocks = GP_MAX_BLOCKS;
    double* part = P.scal + 8;
    gp_max<<<blocks, 256, 0, st>>>(n, mode, v, part);
    ADFWI_LAUNCH_CHECK();
    gp_max<<<1, 256, 0, st>>>(blocks, 3, part, res);
    ADFWI_LAUNCH_CHECK();
    return ADFWI_OK;
}

// smooth2d (:30-49) of the nz x nx plane `in` -> `out` (may alias `in`); tmp is a scratch plane
int gp_smooth(const GPlan& P, cudaStream_t st, int nz, int nx, int span, const double* in, double* out, double* tmp)
{
    if (span < 1 || span > GP_MAX_SPAN) return ADFWI_E_DIMS;
    gp_factor<<<1, 1, 0, st>>>(span, P.f);
    ADFWI_LAUNCH_CHECK();
    gp_partial<<<cdiv(nz, 128), 128, 0, st>>>(nz, span, P.f, P.wz);
    ADFWI_LAUNCH_CHECK();
    gp_partial<<<cdiv(nx, 128), 128, 0, st>>>(nx, span, P.f, P.wx);
    ADFWI_LAUNCH_CHECK();
    const dim3 grd(cdiv(nx, 128), nz);
    gp_conv_rows<<<grd, 128, 0, st>>>(nz, nx, span, P.f, in, tmp);
    ADFWI_LAUNCH_CHECK();
    gp_conv_cols<<<grd, 128, 0, st>>>(nz, nx, span, P.f, P.wz, P.wx, tmp, out);
    ADFWI_LAUNCH_CHECK();
    return ADFWI_OK;
}

}  // namespace
}  // namespace adfwi

using namespace adfwi;

extern "C" size_t adfwi_gradproc_workspace_bytes(const adfwi_gradproc_desc* d)
{
    if (!d || d->nz <= 0 || d->nx <= 0) return 0;
    return gp_plan(d->nz, d->nx, nullptr).bytes;
}

extern "C" int adfwi_gradproc_smooth2d(int nz, int nx, int span, const double* in, double* out, void* ws, size_t ws_bytes, void* stream)
{
    ADFWI_NVTX("adfwi_gradproc_smooth2d");
    if (!in || !out || !ws) return ADFWI_E_NULL;
    if (nz <= 0 || nx <= 0) return ADFWI_E_DIMS;
    const GPlan P = gp_plan(nz, nx, ws);
    if (ws_bytes < P.bytes) return ADFWI_E_WORKSPACE;
    return gp_smooth(P, (cudaStream_t)stream, nz, nx, span, in, out, P.t0);
}

extern "C" int adfwi_gradproc_forward(const adfwi_gradproc_desc* d, const float* grad, const float* forw, const double* mask,
                                      double* out, int* out_is_f32_host, void* ws, size_t ws_bytes, void* stream)
{
    ADFWI_NVTX("adfwi_gradproc_forward");
    if (!d || !grad || !out || !ws) return ADFWI_E_NULL;
    if (d->nz <= 0 || d->nx <= 0 || d->grad_mute < 0 || d->grad_smooth < 0 || d->grad_mute > d->nz) return ADFWI_E_DIMS;
    if (d->use_illumination && !forw) return ADFWI_E_NULL;
    const int nz = d->nz, nx = d->nx;
    const size_t n = (size_t)nz * nx;
    const GPlan P = gp_plan(nz, nx, ws);
    if (ws_bytes < P.bytes) return ADFWI_E_WORKSPACE;
    cudaStream_t st = (cudaStream_t)stream;
    const dim3 grd(cdiv(nx, 128), nz);
    int is32 = 1;
    gp_load_f32<<<nb(n), 256, 0, st>>>(n, grad, P.g);
    ADFWI_LAUNCH_CHECK();
    if (d->grad_mute > 0) {                                               // :97-98
        if (d->taper_marine) {
            gp_taper_marine<<<grd, 128, 0, st>>>(nz, nx, d->grad_mute, P.t1);
            ADFWI_LAUNCH_CHECK();
        } else {
            if (d->grad_mute > nx || d->grad_mute / 2 < 1) return ADFWI_E_DIMS;
            gp_taper_land_init<<<grd, 128, 0, st>>>(nz, nx, d->grad_mute, P.t1);
            ADFWI_LAUNCH_CHECK();
            int rc = gp_smooth(P, st, nz, nx, d->grad_mute / 2, P.t1, P.t1, P.t0);
            if (rc) return rc;
            rc = gp_reduce_max(P, st, n, 0, P.t1, P.scal);
            if (rc) return rc;
            gp_taper_land_finish<<<nb(n), 256, 0, st>>>(n, d->thred, P.scal, P.t1);
            ADFWI_LAUNCH_CHECK();
        }
        gp_mul<<<nb(n), 256, 0, st>>>(n, is32, P.t1, P.g);
        ADFWI_LAUNCH_CHECK();
    }
    if (mask) {                                                           // :101-107
        gp_mul<<<nb(n), 256, 0, st>>>(n, is32, mask, P.g);
        ADFWI_LAUNCH_CHECK();
    }
    if (d->use_illumination) {                                            // :117-122
        gp_load_f32<<<nb(n), 256, 0, st>>>(n, forw, P.t1);
        ADFWI_LAUNCH_CHECK();
        int rc = gp_smooth(P, st, nz, nx, d->illum_span, P.t1, P.t1, P.t0);
        if (rc) return rc;
        rc = gp_reduce_max(P, st, n, 2, P.t1, P.scal + 1);
        if (rc) return rc;
        gp_precond<<<nb(n), 256, 0, st>>>(n, P.t1, P.scal + 1, P.g);
        ADFWI_LAUNCH_CHECK();
        is32 = 0;
    }
    if (d->grad_smooth > 0) {                                             // :125-131
        if (d->smooth_below_mute) {
            const int z0 = d->grad_mute, nzs = nz - z0;
            if (nzs > 0) {
                int rc = gp_smooth(P, st, nzs, nx, d->grad_smooth, P.g + (size_t)z0 * nx, P.t1, P.t0);
                if (rc) return rc;
                gp_store<<<nb((size_t)nzs * nx), 256, 0, st>>>((size_t)nzs * nx, is32, P.t1, P.g + (size_t)z0 * nx);
                ADFWI_LAUNCH_CHECK();
            }
        } else {
            int rc = gp_smooth(P, st, nz, nx, d->grad_smooth, P.g, P.g, P.t0);
            if (rc) return rc;
            is32 = 0;
        }
    }
    if (d->norm_grad) {                                                   // :134-135
        int rc = gp_reduce_max(P, st, n, 1, P.g, P.scal + 2);
        if (rc) return rc;
        gp_norm<<<nb(n), 256, 0, st>>>(n, is32, d->vmax, P.scal + 2, P.g, out);
        ADFWI_LAUNCH_CHECK();
    } else {
        ADFWI_CUDA(cudaMemcpyAsync(out, P.g, n * sizeof(double), cudaMemcpyDeviceToDevice, st));
    }
    if (out_is_f32_host) *out_is_f32_host = is32;
    return ADFWI_OK;
}
#endif  // !ADFWI_HOST_EMUL
