// acoustic_fused.cu -- fused, TMA-staged iso-acoustic time step for sm_100a (the fast path).
//
// One launch per time step per shot group:
//   ac_fwd_fused : P update + source + free-surface mirror + U/W update + free surface + receiver
//                  sampling + illumination + stencil-history store   (acoustic_kernels.py:115-174)
//   ac_adj_fused : receiver-cotangent injection + the whole reverse step 7T..1T of SURVEY.md
//                  Appendix A.1 + g_alpha1 accumulation + g_src
//
// Design (v2, "lean"):
//   * A CTA of 128 threads owns one 64(x) x 32(z) tile of one shot at a time (persistent loop over
//     (tile, shot) items, shot-minor so concurrently running CTAs share coefficient lines in L2).
//   * The old p,u,w (or lambda) tiles with their halos (4 cells in x, 3 in z) are brought into
//     shared memory by TMA (cp.async.bulk.tensor.3d, hardware zero fill outside the grid).
//   * Phase 1 computes the intermediate field (new pressure / pressure cotangent) on the tile plus
//     the one/two-cell ring phase 2 needs (halo recompute); every thread owns a float4 of 4 cells
//     in x and a block of 5 rows in z, so all shared-memory traffic is 64/128-bit and the z
//     neighbours are reused from registers.  Phase 2 (4 cells x 4 rows per thread) produces the
//     new fields and writes them with 128-bit stores to the other buffer of a ping-pong pair.
//   * No per-cell region predicates: the update regions of Appendix A.1 are folded into a private,
//     zero-padded copy of the coefficient planes ("coefficient pack": alpha = 0 and kappa = 0
//     outside a field's update region make the update the identity, bit for bit).
//   * Receivers are bucketed per tile once per call; a tile only touches its own receivers.
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
constexpr int NG = TX / 4;                  // float4 groups per tile row
constexpr int NGP = NG + 2;                 // + one side group left and right (ring of phase 1)
constexpr int RB = 5;                       // rows per phase-1 block
constexpr int NB1 = (TZ + 3) / RB;          // phase-1 row blocks: rows [-1, TZ+2)
constexpr int NTHREADS = NG * (TZ / 4);     // 128
static_assert((TZ + 3) % RB == 0, "phase-1 row blocks must tile rows [-1,TZ+2)");
static_assert(NGP * NB1 <= NTHREADS, "phase 1 must fit one pass");
static_assert(RX <= NTHREADS, "row fix-ups use one thread per staged column");
constexpr int RECT_BYTES = ((RZ * RX * 4 + 127) / 128) * 128;   // 11008
constexpr int CPX = 8, CPZ = 4;             // apron of the coefficient pack (cells)

struct FGeom {
    int nzp, nxp, ld, fs, zlo, nabc, nt;
    int ntx, ntz;
    int cpld;                // pitch of the coefficient pack
    size_t plane;            // nzp*ld
    float c1, c2, dt;
};

// coefficient pack: masked planes, pointers pre-offset to logical cell (0,0)
struct CoefPack { const float *a1, *k1, *a2u, *k2, *a2w, *k3; };

struct RcvBuckets { const int* start; const int* id; const int* zx; const unsigned char* nbr; };

struct FwdArgs {
    CoefPack cp;
    float *p_out, *u_out, *w_out;
    const float* src_v; const int64_t *sx, *sz;
    float* hist; int hist_len, tl, it;
    int nr; RcvBuckets rb;
    float *rcv_p, *rcv_u, *rcv_w;
    float *ill_p, *ill_u; int acc_u;
    int s_begin, s_end;
};

struct AdjArgs {
    CoefPack cp;
    float *lp_out, *lu_out, *lw_out;
    const int64_t *sx, *sz;
    const float* hist; int hist_len, tl, it;
    int nr; RcvBuckets rb;
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

__device__ __forceinline__ float4 ld4(const float* p) { return *reinterpret_cast<const float4*>(p); }
__device__ __forceinline__ float2 ld2(const float* p) { return *reinterpret_cast<const float2*>(p); }
__device__ __forceinline__ void st4(float* p, const float4& v) { *reinterpret_cast<float4*>(p) = v; }
__device__ __forceinline__ float4 ldg4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }
__device__ __forceinline__ float2 ldg2(const float* p) { return __ldg(reinterpret_cast<const float2*>(p)); }

// ------------------------------------------------------------------------------------------
template <bool FS, bool SAVE, bool ILLUM>
__global__ void __launch_bounds__(NTHREADS, 5)
ac_fwd_fused(const __grid_constant__ CUtensorMap tm_p, const __grid_constant__ CUtensorMap tm_u,
             const __grid_constant__ CUtensorMap tm_w, const FGeom g, const FwdArgs a)
{
    extern __shared__ __align__(128) unsigned char smem_raw[];
    float* ps = (float*)smem_raw;
    float* us = (float*)(smem_raw + RECT_BYTES);
    float* ws = (float*)(smem_raw + 2 * RECT_BYTES);
    float* pn = (float*)(smem_raw + 3 * RECT_BYTES);
    uint64_t* bar = (uint64_t*)(smem_raw + 4 * RECT_BYTES);
    const int tid = threadIdx.x;
    if (tid == 0) mbar_init(bar, 1);
    __syncthreads();
    const int nsh = a.s_end - a.s_begin;
    const int nitems = g.ntx * g.ntz * nsh;
    uint32_t parity = 0;
    const int ld = g.ld, cpld = g.cpld;
    const float c1 = g.c1, c2 = g.c2;
    // phase-1 role: float4 group gi in [-1, NG], row block b in [0, NB1)
    const int b1 = tid / NGP, gi1 = tid - b1 * NGP - 1;
    const bool p1_active = tid < NGP * NB1;
    const int c01 = 4 * gi1, r01 = RB * b1 - 1;
    const int so1 = (r01 + HZ) * RX + c01 + HX;
    // phase-2 role: float4 group l, row quad q
    const int q2 = tid / NG, l2 = tid - q2 * NG;
    const int c02 = 4 * l2, r02 = 4 * q2;
    const int so2 = (r02 + HZ) * RX + c02 + HX;

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
        const bool has_src = (szs >= Z0 - 1) && (szs < Z0 + TZ + 2) && (sxs >= X0 - 1) && (sxs < X0 + TX + 2);
        const int rcv_lo = a.nr > 0 ? a.rb.start[tile] : 0, rcv_hi = a.nr > 0 ? a.rb.start[tile + 1] : 0;
        const bool has_rcv = rcv_hi > rcv_lo;
        while (!mbar_try(bar, parity)) {}
        parity ^= 1;
        // ---- phase 1: new pressure on rows [-1,TZ+2) x cols [-4,TX+4) (cols [-1,TX+2) are used) ----
        if (p1_active) {
            const int gz0 = Z0 + r01, gx0 = X0 + c01;
            const size_t cpo = (size_t)gz0 * cpld + gx0;      // may be "negative": the pack has an apron
            const float* A1 = a.cp.a1 + (ptrdiff_t)cpo;
            const float* K1 = a.cp.k1 + (ptrdiff_t)cpo;
            float4 wq[RB + 3];
#pragma unroll
            for (int k = 0; k < RB + 3; ++k) wq[k] = ld4(ws + so1 + (k - 2) * RX);
            float4 pv[RB];
            const bool colok = (gi1 >= 0) && (gi1 < NG) && (gx0 < ld);
#pragma unroll
            for (int j = 0; j < RB; ++j) {
                const float* ur = us + so1 + j * RX;
                const float2 ul = ld2(ur - 2);
                const float4 um = ld4(ur);
                const float uR = ur[4];
                const float4 po = ld4(ps + so1 + j * RX);
                const float4 al = ldg4(A1 + (size_t)j * cpld);
                const float4 kk = ldg4(K1 + (size_t)j * cpld);
                const float4 w0 = wq[j + 2], wm1 = wq[j + 1], wp1 = wq[j + 3], wm2 = wq[j];
                float4 S;
                S.x = c1 * (((um.x - ul.y) + w0.x) - wm1.x) + c2 * (((um.y - ul.x) + wp1.x) - wm2.x);
                S.y = c1 * (((um.y - um.x) + w0.y) - wm1.y) + c2 * (((um.z - ul.y) + wp1.y) - wm2.y);
                S.z = c1 * (((um.z - um.y) + w0.z) - wm1.z) + c2 * (((um.w - um.x) + wp1.z) - wm2.z);
                S.w = c1 * (((um.w - um.z) + w0.w) - wm1.w) + c2 * (((uR - um.y) + wp1.w) - wm2.w);
                if (SAVE) {
                    const int r = r01 + j, gz = gz0 + j;
                    if (colok && r >= 0 && r < TZ && gz < g.nzp)
                        __stcs(reinterpret_cast<float4*>(a.hist + ((size_t)s * a.hist_len + a.tl) * g.plane + (size_t)gz * ld + gx0), S);
                }
                pv[j].x = (1.0f - kk.x) * po.x - al.x * S.x;
                pv[j].y = (1.0f - kk.y) * po.y - al.y * S.y;
                pv[j].z = (1.0f - kk.z) * po.z - al.z * S.z;
                pv[j].w = (1.0f - kk.w) * po.w - al.w * S.w;
            }
            if (has_src) {     // p[sz,sx] += dt*src   (acoustic_kernels.py:131-132)
                const int dr = szs - gz0, dc = sxs - gx0;
                if (dr >= 0 && dr < RB && dc >= 0 && dc < 4) {
                    const float srcval = g.dt * a.src_v[(size_t)s * g.nt + a.it];
#pragma unroll
                    for (int j = 0; j < RB; ++j) {
                        if (j == dr) {
                            if (dc == 0) pv[j].x = pv[j].x + srcval;
                            else if (dc == 1) pv[j].y = pv[j].y + srcval;
                            else if (dc == 2) pv[j].z = pv[j].z + srcval;
                            else pv[j].w = pv[j].w + srcval;
                        }
                    }
                }
            }
            if (FS && tzi == 0 && b1 == 0) {      // p[fs-1] = -p[fs+1]: tile rows 0 and 2 = block rows 1 and 3
                pv[1].x = -pv[3].x; pv[1].y = -pv[3].y; pv[1].z = -pv[3].z; pv[1].w = -pv[3].w;
            }
#pragma unroll
            for (int j = 0; j < RB; ++j) st4(pn + so1 + j * RX, pv[j]);
        }
        __syncthreads();
        // ---- phase 2: velocities on the interior, stores, illumination -----------------------------
        {
            const int gx0 = X0 + c02, gz0 = Z0 + r02;
            const bool colok = gx0 < ld;
            const size_t cpo = (size_t)gz0 * cpld + gx0;
            float4 P[7];
#pragma unroll
            for (int k = 0; k < 7; ++k) P[k] = ld4(pn + so2 + (k - 1) * RX);
            float4 uv[4], wv[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float* pr = pn + so2 + j * RX;
                const float pL = pr[-1];
                const float2 pR = ld2(pr + 4);
                const float4 p0 = P[j + 1], pm1 = P[j], pp1 = P[j + 2], pp2 = P[j + 3];
                const float4 uo = ld4(us + so2 + j * RX), wo = ld4(ws + so2 + j * RX);
                const float4 au = ldg4(a.cp.a2u + (ptrdiff_t)cpo + (size_t)j * cpld), k2 = ldg4(a.cp.k2 + (ptrdiff_t)cpo + (size_t)j * cpld);
                const float4 aw = ldg4(a.cp.a2w + (ptrdiff_t)cpo + (size_t)j * cpld), k3 = ldg4(a.cp.k3 + (ptrdiff_t)cpo + (size_t)j * cpld);
                uv[j].x = (1.0f - k2.x) * uo.x - au.x * (c1 * (p0.y - p0.x) + c2 * (p0.z - pL));
                uv[j].y = (1.0f - k2.y) * uo.y - au.y * (c1 * (p0.z - p0.y) + c2 * (p0.w - p0.x));
                uv[j].z = (1.0f - k2.z) * uo.z - au.z * (c1 * (p0.w - p0.z) + c2 * (pR.x - p0.y));
                uv[j].w = (1.0f - k2.w) * uo.w - au.w * (c1 * (pR.x - p0.w) + c2 * (pR.y - p0.z));
                wv[j].x = (1.0f - k3.x) * wo.x - aw.x * (c1 * (pp1.x - p0.x) + c2 * (pp2.x - pm1.x));
                wv[j].y = (1.0f - k3.y) * wo.y - aw.y * (c1 * (pp1.y - p0.y) + c2 * (pp2.y - pm1.y));
                wv[j].z = (1.0f - k3.z) * wo.z - aw.z * (c1 * (pp1.z - p0.z) + c2 * (pp2.z - pm1.z));
                wv[j].w = (1.0f - k3.w) * wo.w - aw.w * (c1 * (pp1.w - p0.w) + c2 * (pp2.w - pm1.w));
            }
            if (FS && tzi == 0 && q2 == 0) wv[0] = wv[1];        // w[fs-1] = w[fs]   (acoustic_kernels.py:163-164)
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int gz = gz0 + j;
                if (colok && gz < g.nzp) {
                    const size_t cg = (size_t)gz * ld + gx0;
                    const size_t o = (size_t)s * g.plane + cg;
                    const float4 p0 = P[j + 1];
                    st4(a.p_out + o, p0);
                    st4(a.u_out + o, uv[j]);
                    st4(a.w_out + o, wv[j]);
                    if (ILLUM) {      // per-shot-slot partial sums over the whole plane; cropped when finalised
                        float* ip = a.ill_p + (size_t)sl * g.plane + cg;
                        float4 acc = ld4(ip);
                        acc.x += p0.x * p0.x; acc.y += p0.y * p0.y; acc.z += p0.z * p0.z; acc.w += p0.w * p0.w;
                        st4(ip, acc);
                        if (a.acc_u) {
                            float* iu = a.ill_u + (size_t)sl * g.plane + cg;
                            float4 au = ld4(iu);
                            au.x += uv[j].x * uv[j].x; au.y += uv[j].y * uv[j].y; au.z += uv[j].z * uv[j].z; au.w += uv[j].w * uv[j].w;
                            st4(iu, au);
                        }
                    }
                }
                if (has_rcv) { st4(us + so2 + j * RX, uv[j]); st4(ws + so2 + j * RX, wv[j]); }
            }
        }
        // ---- receivers of this tile (acoustic_kernels.py:167-169) ----------------------------------
        if (has_rcv) {
            __syncthreads();
            for (int i = rcv_lo + tid; i < rcv_hi; i += NTHREADS) {
                const int r = a.rb.id[i], zx = a.rb.zx[i];
                const int off = ((zx >> 16) - Z0 + HZ) * RX + ((zx & 0xffff) - X0 + HX);
                const size_t o = ((size_t)s * g.nt + a.it) * a.nr + r;
                a.rcv_p[o] = pn[off];
                if (a.rcv_u) a.rcv_u[o] = us[off];
                if (a.rcv_w) a.rcv_w[o] = ws[off];
            }
        }
        __syncthreads();
    }
}

// ------------------------------------------------------------------------------------------
template <bool FS>
__global__ void __launch_bounds__(NTHREADS, 4)
ac_adj_fused(const __grid_constant__ CUtensorMap tm_lp, const __grid_constant__ CUtensorMap tm_lu,
             const __grid_constant__ CUtensorMap tm_lw, const FGeom g, const AdjArgs a)
{
    extern __shared__ __align__(128) unsigned char smem_raw[];
    float* lps = (float*)smem_raw;
    float* lus = (float*)(smem_raw + RECT_BYTES);
    float* lws = (float*)(smem_raw + 2 * RECT_BYTES);
    float* lp1 = (float*)(smem_raw + 3 * RECT_BYTES);     // lambda_p after undoing W,U (and 3T)
    float* mps = (float*)(smem_raw + 4 * RECT_BYTES);     // m = -alpha1 * lambda_p1
    uint64_t* bar = (uint64_t*)(smem_raw + 5 * RECT_BYTES);
    const int tid = threadIdx.x;
    if (tid == 0) mbar_init(bar, 1);
    __syncthreads();
    const int nsh = a.s_end - a.s_begin;
    const int nitems = g.ntx * g.ntz * nsh;
    const bool have_g = a.nr > 0 && (a.gp || a.gu || a.gw);
    uint32_t parity = 0;
    const int ld = g.ld, cpld = g.cpld;
    const float c1 = g.c1, c2 = g.c2;
    const int b1 = tid / NGP, gi1 = tid - b1 * NGP - 1;
    const bool p1_active = tid < NGP * NB1;
    const int c01 = 4 * gi1, r01 = RB * b1 - 1;
    const int so1 = (r01 + HZ) * RX + c01 + HX;
    const int q2 = tid / NG, l2 = tid - q2 * NG;
    const int c02 = 4 * l2, r02 = 4 * q2;
    const int so2 = (r02 + HZ) * RX + c02 + HX;

    for (int item = blockIdx.x; item < nitems; item += gridDim.x) {
        const int sl = item % nsh, tile = item / nsh;
        const int s = a.s_begin + sl;
        const int tzi = tile / g.ntx, txi = tile - tzi * g.ntx;
        const int X0 = txi * TX, Z0 = g.zlo + tzi * TZ;
        if (tid == 0) {
            fence_proxy_async();
            mbar_expect_tx(bar, 3 * RZ * RX * 4);
            tma_load_3d(lps, &tm_lp, X0 - HX, Z0 - HZ, s, bar);
            tma_load_3d(lus, &tm_lu, X0 - HX, Z0 - HZ, s, bar);
            tma_load_3d(lws, &tm_lw, X0 - HX, Z0 - HZ, s, bar);
        }
        const bool inject = have_g && a.rb.nbr[tile];
        while (!mbar_try(bar, parity)) {}
        parity ^= 1;
        // ---- 7T: receiver cotangents into the staged rectangle (duplicates legal -> shared atomics)
        if (inject) {
            for (int dz = -1; dz <= 1; ++dz) {
                const int tz2 = tzi + dz;
                if (tz2 < 0 || tz2 >= g.ntz) continue;
                for (int dx = -1; dx <= 1; ++dx) {
                    const int tx2 = txi + dx;
                    if (tx2 < 0 || tx2 >= g.ntx) continue;
                    const int t2 = tz2 * g.ntx + tx2;
                    const int lo = a.rb.start[t2], hi = a.rb.start[t2 + 1];
                    for (int i = lo + tid; i < hi; i += NTHREADS) {
                        const int zx = a.rb.zx[i];
                        const int z = (zx >> 16) - (Z0 - HZ), x = (zx & 0xffff) - (X0 - HX);
                        if (z >= 0 && z < RZ && x >= 0 && x < RX) {
                            const size_t o = ((size_t)s * g.nt + a.it) * a.nr + a.rb.id[i];
                            if (a.gp) atomicAdd(lps + z * RX + x, a.gp[o]);
                            if (a.gu) atomicAdd(lus + z * RX + x, a.gu[o]);
                            if (a.gw) atomicAdd(lws + z * RX + x, a.gw[o]);
                        }
                    }
                }
            }
            __syncthreads();
        }
        if (FS && tzi == 0) {        // 6T: lambda_w[fs] += lambda_w[fs-1]; lambda_w[fs-1] = 0  (staged rows HZ+1 and HZ)
            if (tid < RX) { lws[(HZ + 1) * RX + tid] += lws[HZ * RX + tid]; lws[HZ * RX + tid] = 0.f; }
            __syncthreads();
        }
        // ---- phase 1: lambda_p after undoing W and U (5T, 4T), 3T, and m = -alpha1*lambda_p ------------
        if (p1_active) {
            const int gz0 = Z0 + r01, gx0 = X0 + c01;
            const ptrdiff_t cpo = (ptrdiff_t)gz0 * cpld + gx0;
            float4 qw[RB + 3];     // (-alpha2w * lambda_w) rows r01-2 .. r01+5
#pragma unroll
            for (int k = 0; k < RB + 3; ++k) {
                const float4 lw = ld4(lws + so1 + (k - 2) * RX);
                const float4 aw = ldg4(a.cp.a2w + cpo + (ptrdiff_t)(k - 2) * cpld);
                qw[k].x = (-aw.x) * lw.x; qw[k].y = (-aw.y) * lw.y; qw[k].z = (-aw.z) * lw.z; qw[k].w = (-aw.w) * lw.w;
            }
            float4 acc[RB];
#pragma unroll
            for (int j = 0; j < RB; ++j) {
                const float* ur = lus + so1 + j * RX;
                const float* ar = a.cp.a2u + cpo + (ptrdiff_t)j * cpld;
                const float2 ul = ld2(ur - 2); const float4 um = ld4(ur); const float uR = ur[4];
                const float2 al = ldg2(ar - 2); const float4 am = ldg4(ar); const float aR = __ldg(ar + 4);
                // qu at columns c0-2 .. c0+4
                const float q0 = (-al.x) * ul.x, q1 = (-al.y) * ul.y, q2_ = (-am.x) * um.x, q3 = (-am.y) * um.y,
                            q4 = (-am.z) * um.z, q5 = (-am.w) * um.w, q6 = (-aR) * uR;
                const float4 w1 = qw[j + 1], w2 = qw[j + 2], w0 = qw[j], w3 = qw[j + 3];
                float4 v = ld4(lps + so1 + j * RX);
                // transposes of D+z and D+x:  +c1 m[z-1] - c1 m[z] + c2 m[z-2] - c2 m[z+1]
                v.x += c1 * w1.x - c1 * w2.x + c2 * w0.x - c2 * w3.x;
                v.y += c1 * w1.y - c1 * w2.y + c2 * w0.y - c2 * w3.y;
                v.z += c1 * w1.z - c1 * w2.z + c2 * w0.z - c2 * w3.z;
                v.w += c1 * w1.w - c1 * w2.w + c2 * w0.w - c2 * w3.w;
                v.x += c1 * q1 - c1 * q2_ + c2 * q0 - c2 * q3;
                v.y += c1 * q2_ - c1 * q3 + c2 * q1 - c2 * q4;
                v.z += c1 * q3 - c1 * q4 + c2 * q2_ - c2 * q5;
                v.w += c1 * q4 - c1 * q5 + c2 * q3 - c2 * q6;
                acc[j] = v;
            }
            if (FS && tzi == 0 && b1 == 0) {   // 3T: lambda_p[fs+1] -= lambda_p[fs-1]; lambda_p[fs-1] = 0 (block rows 3 and 1)
                acc[3].x -= acc[1].x; acc[3].y -= acc[1].y; acc[3].z -= acc[1].z; acc[3].w -= acc[1].w;
                acc[1] = make_float4(0.f, 0.f, 0.f, 0.f);
            }
#pragma unroll
            for (int j = 0; j < RB; ++j) {
                const float4 a1 = ldg4(a.cp.a1 + cpo + (ptrdiff_t)j * cpld);
                float4 m;
                m.x = (-a1.x) * acc[j].x; m.y = (-a1.y) * acc[j].y; m.z = (-a1.z) * acc[j].z; m.w = (-a1.w) * acc[j].w;
                st4(lp1 + so1 + j * RX, acc[j]);
                st4(mps + so1 + j * RX, m);
            }
        }
        __syncthreads();
        // ---- phase 2: new lambda_u, lambda_w, lambda_p on the interior (5T,4T,1T), g_alpha1, g_src ---
        {
            const int gx0 = X0 + c02, gz0 = Z0 + r02;
            const bool colok = gx0 < ld;
            const ptrdiff_t cpo = (ptrdiff_t)gz0 * cpld + gx0;
            const float* Hs = a.hist + ((size_t)s * a.hist_len + a.tl) * g.plane;
            float4 M[7];
#pragma unroll
            for (int k = 0; k < 7; ++k) M[k] = ld4(mps + so2 + (k - 1) * RX);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int gz = gz0 + j;
                if (colok && gz < g.nzp) {
                    const float* mr = mps + so2 + j * RX;
                    const float mL = mr[-1];
                    const float2 mR = ld2(mr + 4);
                    const float4 m0 = M[j + 1], mm1 = M[j], mp1 = M[j + 2], mp2 = M[j + 3];
                    const float4 qp = ld4(lp1 + so2 + j * RX);
                    const float4 luo = ld4(lus + so2 + j * RX), lwo = ld4(lws + so2 + j * RX);
                    const float4 k1 = ldg4(a.cp.k1 + cpo + (ptrdiff_t)j * cpld), k2 = ldg4(a.cp.k2 + cpo + (ptrdiff_t)j * cpld),
                                 k3 = ldg4(a.cp.k3 + cpo + (ptrdiff_t)j * cpld);
                    const size_t cg = (size_t)gz * ld + gx0;
                    const size_t o = (size_t)s * g.plane + cg;
                    const float4 S = __ldcs(reinterpret_cast<const float4*>(Hs + cg));
                    float4 nu, nw, np;
                    // lambda_u: transpose of D-x:  +c1 m[x] - c1 m[x+1] + c2 m[x-1] - c2 m[x+2]
                    nu.x = (1.0f - k2.x) * luo.x + (c1 * m0.x - c1 * m0.y + c2 * mL - c2 * m0.z);
                    nu.y = (1.0f - k2.y) * luo.y + (c1 * m0.y - c1 * m0.z + c2 * m0.x - c2 * m0.w);
                    nu.z = (1.0f - k2.z) * luo.z + (c1 * m0.z - c1 * m0.w + c2 * m0.y - c2 * mR.x);
                    nu.w = (1.0f - k2.w) * luo.w + (c1 * m0.w - c1 * mR.x + c2 * m0.z - c2 * mR.y);
                    // lambda_w: transpose of D-z:  +c1 m[z] - c1 m[z+1] + c2 m[z-1] - c2 m[z+2]
                    nw.x = (1.0f - k3.x) * lwo.x + (c1 * m0.x - c1 * mp1.x + c2 * mm1.x - c2 * mp2.x);
                    nw.y = (1.0f - k3.y) * lwo.y + (c1 * m0.y - c1 * mp1.y + c2 * mm1.y - c2 * mp2.y);
                    nw.z = (1.0f - k3.z) * lwo.z + (c1 * m0.z - c1 * mp1.z + c2 * mm1.z - c2 * mp2.z);
                    nw.w = (1.0f - k3.w) * lwo.w + (c1 * m0.w - c1 * mp1.w + c2 * mm1.w - c2 * mp2.w);
                    np.x = (1.0f - k1.x) * qp.x; np.y = (1.0f - k1.y) * qp.y; np.z = (1.0f - k1.z) * qp.z; np.w = (1.0f - k1.w) * qp.w;
                    st4(a.lu_out + o, nu);
                    st4(a.lw_out + o, nw);
                    st4(a.lp_out + o, np);
                    float* gp1 = a.g1part + (size_t)sl * g.plane + cg;     // masked to the P region when reduced
                    float4 ga = ld4(gp1);
                    ga.x = ga.x - qp.x * S.x; ga.y = ga.y - qp.y * S.y; ga.z = ga.z - qp.z * S.z; ga.w = ga.w - qp.w * S.w;
                    st4(gp1, ga);
                    if (a.g_src) {
                        const int dz = (int)a.sz[s] - gz, dx = (int)a.sx[s] - gx0;
                        if (dz == 0 && dx >= 0 && dx < 4)
                            a.g_src[(size_t)s * g.nt + a.it] = g.dt * (dx == 0 ? qp.x : dx == 1 ? qp.y : dx == 2 ? qp.z : qp.w);
                    }
                }
            }
        }
        __syncthreads();
    }
}

// ---- set-up kernels ------------------------------------------------------------------------------
// coefficient pack: six planes [cprows][cpld], logical cell (z,x) at [(z+CPZ)*cpld + x+CPX]; each
// plane is the caller's coefficient inside the update region of the field it drives and 0 elsewhere
// (update regions: SURVEY.md Appendix A.1 / acoustic_kernels.py:115,139,151).
__global__ void acf_pack_coefs(int nzp, int nxp, int fs, int cprows, int cpld, size_t cpplane,
                               const float* __restrict__ a1, const float* __restrict__ k1, const float* __restrict__ a2,
                               const float* __restrict__ k2, const float* __restrict__ k3, float* __restrict__ pack)
{
    const int xx = blockIdx.x * blockDim.x + threadIdx.x, zz = blockIdx.y;
    if (xx >= cpld || zz >= cprows) return;
    const int x = xx - CPX, z = zz - CPZ;
    const bool in = (z >= 0) && (z < nzp) && (x >= 0) && (x < nxp);
    const bool inP = in && (z >= fs + 1) && (z < nzp - 2) && (x >= 2) && (x < nxp - 2);
    const bool inU = in && (z >= fs) && (z < nzp - 1) && (x >= 1) && (x < nxp - 2);
    const bool inW = in && (z >= fs) && (z < nzp - 2) && (x >= 1) && (x < nxp - 1);
    const size_t c = in ? (size_t)z * nxp + x : 0;
    const size_t o = (size_t)zz * cpld + xx;
    pack[0 * cpplane + o] = inP ? a1[c] : 0.f;
    pack[1 * cpplane + o] = inP ? k1[c] : 0.f;
    pack[2 * cpplane + o] = inU ? a2[c] : 0.f;
    pack[3 * cpplane + o] = inU ? k2[c] : 0.f;
    pack[4 * cpplane + o] = inW ? a2[c] : 0.f;
    pack[5 * cpplane + o] = inW ? k3[c] : 0.f;
}

// receiver buckets: counting sort of the receivers by tile
__device__ __forceinline__ int rcv_tile(const FGeom& g, int64_t z, int64_t x)
{
    if (z < g.zlo || z >= g.nzp || x < 0 || x >= g.nxp) return -1;
    return ((int)z - g.zlo) / TZ * g.ntx + (int)x / TX;
}
__global__ void acf_rcv_count(const FGeom g, int nr, const int64_t* __restrict__ rx, const int64_t* __restrict__ rz, int* __restrict__ cnt)
{
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= nr) return;
    const int t = rcv_tile(g, rz[r], rx[r]);
    if (t >= 0) atomicAdd(cnt + t, 1);
}
__global__ void acf_rcv_scan(int ntiles, const int* __restrict__ cnt, int* __restrict__ start, int* __restrict__ cursor)
{
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        int acc = 0;
        for (int t = 0; t < ntiles; ++t) { start[t] = acc; cursor[t] = acc; acc += cnt[t]; }
        start[ntiles] = acc;
    }
}
__global__ void acf_rcv_fill(const FGeom g, int nr, const int64_t* __restrict__ rx, const int64_t* __restrict__ rz,
                             int* __restrict__ cursor, int* __restrict__ id, int* __restrict__ zx)
{
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= nr) return;
    const int t = rcv_tile(g, rz[r], rx[r]);
    if (t < 0) return;
    const int i = atomicAdd(cursor + t, 1);
    id[i] = r;
    zx[i] = ((int)rz[r] << 16) | (int)rx[r];
}
__global__ void acf_rcv_nbr(const FGeom g, const int* __restrict__ start, unsigned char* __restrict__ nbr)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= g.ntx * g.ntz) return;
    const int tz = t / g.ntx, tx = t - tz * g.ntx;
    int any = 0;
    for (int dz = -1; dz <= 1; ++dz)
        for (int dx = -1; dx <= 1; ++dx) {
            const int z2 = tz + dz, x2 = tx + dx;
            if (z2 < 0 || z2 >= g.ntz || x2 < 0 || x2 >= g.ntx) continue;
            const int t2 = z2 * g.ntx + x2;
            any |= (start[t2 + 1] > start[t2]);
        }
    nbr[t] = (unsigned char)any;
}

// pitched partial planes -> dense caller plane, masked to the P update region
__global__ void acf_reduce_parts(int nzp, int nxp, int ld, int fs, int nparts, const float* __restrict__ part, float* __restrict__ out)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x, z = blockIdx.y;
    if (x >= nxp || z >= nzp) return;
    float acc = 0.f;
    const bool inP = (z >= fs + 1) && (z < nzp - 2) && (x >= 2) && (x < nxp - 2);
    if (inP) for (int k = 0; k < nparts; ++k) acc += part[((size_t)k * nzp + z) * ld + x];
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

__global__ void acf_illum_finalize(int nzp, int nxp, int ld, int nabc, int nparts, size_t plane, const float* __restrict__ ip,
                                   const float* __restrict__ iu, const float* __restrict__ iw, float* op, float* ou, float* ow)
{
    const int nx = nxp - 2 * nabc, nz = nzp - 2 * nabc;
    const int x = blockIdx.x * blockDim.x + threadIdx.x, z = blockIdx.y;
    if (x >= nx || z >= nz) return;
    const size_t c = (size_t)(z + nabc) * ld + (x + nabc), o = (size_t)z * nx + x;
    float p = 0.f, u = 0.f;
    for (int k = 0; k < nparts; ++k) { p += ip[k * plane + c]; u += iu[k * plane + c]; }
    if (op) op[o] = p;
    if (ou) ou[o] = p + u;
    if (ow) ow[o] = p + (u + iw[c]);
}

struct FPlan {
    FGeom g;
    int ns, nr, FS, save, n_segments;
    int K, nseg, nckpt, G;
    int cprows; size_t cpplane;
    float* pack;                    // 6 masked coefficient planes
    float *st[2][3];                // p,u,w ping-pong
    float *lam[2][3];               // lambda ping-pong
    float *hist, *ckpt, *g1part, *ill_p, *ill_u, *ill_w;
    int *rcv_cnt, *rcv_start, *rcv_cursor, *rcv_id, *rcv_zx; unsigned char* rcv_nbr;
    size_t bytes;
};

int acf_make_plan(const adfwi_acoustic_desc* d, void* ws, FPlan* P)
{
    FGeom& g = P->g;
    g.nzp = d->nzp; g.nxp = d->nxp; g.ld = (d->nxp + 31) / 32 * 32; g.nabc = d->nabc; g.nt = d->nt;
    g.fs = d->free_surface ? d->nabc : 1;
    g.zlo = g.fs - 1;
    g.ntx = cdiv(g.nxp, TX); g.ntz = cdiv(g.nzp - g.zlo, TZ);
    g.plane = (size_t)g.nzp * g.ld;
    g.cpld = g.ntx * TX + 2 * CPX;
    P->cprows = g.zlo + g.ntz * TZ + 2 * CPZ;
    P->cpplane = align_up((size_t)P->cprows * g.cpld, 64);
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
    const int ntiles = g.ntx * g.ntz;
    P->pack = cv.take<float>(6 * P->cpplane);
    for (int b = 0; b < 2; ++b) for (int f = 0; f < 3; ++f) P->st[b][f] = cv.take<float>(sp);
    P->ill_p = cv.take<float>((size_t)G * g.plane); P->ill_u = cv.take<float>((size_t)G * g.plane); P->ill_w = cv.take<float>(g.plane);
    P->rcv_cnt = cv.take<int>(ntiles + 1); P->rcv_start = cv.take<int>(ntiles + 1); P->rcv_cursor = cv.take<int>(ntiles + 1);
    P->rcv_id = cv.take<int>(d->nr > 0 ? d->nr : 1); P->rcv_zx = cv.take<int>(d->nr > 0 ? d->nr : 1);
    P->rcv_nbr = cv.take<unsigned char>(ntiles);
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

constexpr int FWD_SMEM = 4 * RECT_BYTES + 64;
constexpr int ADJ_SMEM = 5 * RECT_BYTES + 64;
constexpr int FWD_CTAS_PER_SM = 5, ADJ_CTAS_PER_SM = 4;

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

CoefPack acf_pack_ptrs(const FPlan& P)
{
    const size_t o = (size_t)CPZ * P.g.cpld + CPX;
    CoefPack c;
    c.a1 = P.pack + 0 * P.cpplane + o; c.k1 = P.pack + 1 * P.cpplane + o;
    c.a2u = P.pack + 2 * P.cpplane + o; c.k2 = P.pack + 3 * P.cpplane + o;
    c.a2w = P.pack + 4 * P.cpplane + o; c.k3 = P.pack + 5 * P.cpplane + o;
    return c;
}

RcvBuckets acf_bucket_ptrs(const FPlan& P)
{
    RcvBuckets b; b.start = P.rcv_start; b.id = P.rcv_id; b.zx = P.rcv_zx; b.nbr = P.rcv_nbr;
    return b;
}

// per-call set-up left in the workspace for the matching backward call: coefficient pack + buckets
int acf_setup(const FPlan& P, cudaStream_t st, const float* const* coef, const int64_t* rx, const int64_t* rz)
{
    const FGeom& g = P.g;
    acf_pack_coefs<<<dim3(cdiv(g.cpld, 128), P.cprows), 128, 0, st>>>(g.nzp, g.nxp, g.fs, P.cprows, g.cpld, P.cpplane,
                                                                      coef[0], coef[1], coef[2], coef[3], coef[4], P.pack);
    ADFWI_LAUNCH_CHECK();
    const int ntiles = g.ntx * g.ntz;
    ADFWI_CUDA(cudaMemsetAsync(P.rcv_cnt, 0, sizeof(int) * (ntiles + 1), st));
    if (P.nr > 0) {
        acf_rcv_count<<<cdiv(P.nr, 128), 128, 0, st>>>(g, P.nr, rx, rz, P.rcv_cnt);
        ADFWI_LAUNCH_CHECK();
    }
    acf_rcv_scan<<<1, 32, 0, st>>>(ntiles, P.rcv_cnt, P.rcv_start, P.rcv_cursor);
    ADFWI_LAUNCH_CHECK();
    if (P.nr > 0) {
        acf_rcv_fill<<<cdiv(P.nr, 128), 128, 0, st>>>(g, P.nr, rx, rz, P.rcv_cursor, P.rcv_id, P.rcv_zx);
        ADFWI_LAUNCH_CHECK();
    }
    acf_rcv_nbr<<<cdiv(ntiles, 128), 128, 0, st>>>(g, P.rcv_start, P.rcv_nbr);
    ADFWI_LAUNCH_CHECK();
    return ADFWI_OK;
}

struct StepMaps { CUtensorMap st[2][3]; CUtensorMap lam[2][3]; };

int acf_make_maps(const FPlan& P, StepMaps* M)
{
    const FGeom& g = P.g;
    for (int b = 0; b < 2; ++b)
        for (int f = 0; f < 3; ++f) {
            int rc = make_tmap_f32(&M->st[b][f], P.st[b][f], 3, g.nxp, g.ld, g.nzp, P.ns, RX, RZ);
            if (rc) return rc;
            if (P.save) { rc = make_tmap_f32(&M->lam[b][f], P.lam[b][f], 3, g.nxp, g.ld, g.nzp, P.ns, RX, RZ); if (rc) return rc; }
        }
    return 0;
}

// one fused forward step of shots [sb,se): reads buffer cur, writes buffer cur^1
int acf_forward_step(const FPlan& P, const StepMaps& M, cudaStream_t st, int cur, int sb, int se, int it, bool save, int tl,
                     const float* src_v, const int64_t* sx, const int64_t* sz,
                     float* rcv_p, float* rcv_u, float* rcv_w, bool illum, int acc_u)
{
    const FGeom& g = P.g;
    FwdArgs a;
    a.cp = acf_pack_ptrs(P);
    a.p_out = P.st[cur ^ 1][0]; a.u_out = P.st[cur ^ 1][1]; a.w_out = P.st[cur ^ 1][2];
    a.src_v = src_v; a.sx = sx; a.sz = sz;
    a.hist = P.hist; a.hist_len = P.K; a.tl = tl; a.it = it;
    a.nr = rcv_p ? P.nr : 0; a.rb = acf_bucket_ptrs(P);
    a.rcv_p = rcv_p; a.rcv_u = rcv_u; a.rcv_w = rcv_w;
    a.ill_p = P.ill_p; a.ill_u = P.ill_u; a.acc_u = acc_u;
    a.s_begin = sb; a.s_end = se;
    const int nitems = g.ntx * g.ntz * (se - sb);
    const int cap = FWD_CTAS_PER_SM * acf_num_sms();
    const int grid = nitems < cap ? nitems : cap;
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
    if (g.nzp >= 32768 || g.nxp >= 65536) return ADFWI_E_DIMS;      // receiver cells are packed as (z<<16)|x
    int rc = acf_init_kernels();
    if (rc) return rc;
    StepMaps M;
    rc = acf_make_maps(P, &M);
    if (rc) return rc;
    rc = acf_setup(P, st, coef, rx, rz);
    if (rc) return rc;
    const bool illum = illum_p || illum_u || illum_w;
    const int nt = g.nt;
    const int csz = cdiv(nt, P.n_segments);
    const int last_chunk_start = (cdiv(nt, csz) - 1) * csz;
    if (illum) {
        ADFWI_CUDA(cudaMemsetAsync(P.ill_p, 0, sizeof(float) * (size_t)P.G * g.plane, st));
        ADFWI_CUDA(cudaMemsetAsync(P.ill_u, 0, sizeof(float) * (size_t)P.G * g.plane, st));
        ADFWI_CUDA(cudaMemsetAsync(P.ill_w, 0, sizeof(float) * g.plane, st));
    }
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
            rc = acf_forward_step(P, M, st, cur, sb, se, it, save, tl, src_v, sx, sz,
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
        acf_illum_finalize<<<dim3(cdiv(nx, 128), nz), 128, 0, st>>>(g.nzp, g.nxp, g.ld, g.nabc, P.G, g.plane, P.ill_p, P.ill_u, P.ill_w,
                                                                     illum_p, illum_u, illum_w);
        ADFWI_LAUNCH_CHECK();
    }
    return ADFWI_OK;
}

int acf_backward(const adfwi_acoustic_desc* d, const float* const* coef, const float* src_v, const int64_t* sx, const int64_t* sz,
                 const int64_t* rx, const int64_t* rz, const float* gp, const float* gu, const float* gw,
                 float* g_alpha1, float* g_src, void* ws, cudaStream_t st)
{
    (void)coef; (void)rx; (void)rz;
    FPlan P;
    acf_make_plan(d, ws, &P);
    const FGeom& g = P.g;
    int rc = acf_init_kernels();
    if (rc) return rc;
    StepMaps M;
    rc = acf_make_maps(P, &M);
    if (rc) return rc;
    // the coefficient pack and the receiver buckets were left in the workspace by the forward call
    const int nt = g.nt;
    ADFWI_CUDA(cudaMemsetAsync(P.g1part, 0, sizeof(float) * (size_t)P.G * g.plane, st));
    for (int sb = 0; sb < P.ns; sb += P.G) {
        const int se = sb + P.G < P.ns ? sb + P.G : P.ns;
        const size_t off = (size_t)sb * g.plane, cnt = (size_t)(se - sb) * g.plane * sizeof(float);
        for (int b = 0; b < 2; ++b) for (int f = 0; f < 3; ++f) ADFWI_CUDA(cudaMemsetAsync(P.lam[b][f] + off, 0, cnt, st));
        int lcur = 0;
        const int nitems = g.ntx * g.ntz * (se - sb);
        const int cap = ADJ_CTAS_PER_SM * acf_num_sms();
        const int grid = nitems < cap ? nitems : cap;
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
                    rc = acf_forward_step(P, M, st, cur, sb, se, it, true, it - t0, src_v, sx, sz, nullptr, nullptr, nullptr, false, 0);
                    if (rc) return rc;
                    cur ^= 1;
                }
            }
            for (int it = t1 - 1; it >= t0; --it) {
                AdjArgs a;
                a.cp = acf_pack_ptrs(P);
                a.lp_out = P.lam[lcur ^ 1][0]; a.lu_out = P.lam[lcur ^ 1][1]; a.lw_out = P.lam[lcur ^ 1][2];
                a.sx = sx; a.sz = sz; a.hist = P.hist; a.hist_len = P.K; a.tl = it - t0; a.it = it;
                a.nr = P.nr; a.rb = acf_bucket_ptrs(P); a.gp = gp; a.gu = gu; a.gw = gw;
                a.g1part = P.g1part; a.g_src = g_src; a.s_begin = sb; a.s_end = se;
                {
                    TimedLaunch tl_(KC_AC_ADJ_FUSED, st);
                    if (P.FS) ac_adj_fused<true><<<grid, NTHREADS, ADJ_SMEM, st>>>(M.lam[lcur][0], M.lam[lcur][1], M.lam[lcur][2], g, a);
                    else      ac_adj_fused<false><<<grid, NTHREADS, ADJ_SMEM, st>>>(M.lam[lcur][0], M.lam[lcur][1], M.lam[lcur][2], g, a);
                }
                ADFWI_LAUNCH_CHECK();
                lcur ^= 1;
            }
        }
    }
    acf_reduce_parts<<<dim3(cdiv(g.nxp, 128), g.nzp), 128, 0, st>>>(g.nzp, g.nxp, g.ld, g.fs, P.G, P.g1part, g_alpha1);
    ADFWI_LAUNCH_CHECK();
    return ADFWI_OK;
}

}  // namespace adfwi
#endif  // !ADFWI_HOST_EMUL
