// acoustic_fused.cu -- fused, TMA-staged iso-acoustic time step for sm_100a (the fast path).
//
// One launch per time step per shot group:
//   ac_fwd_fused : P update + source + free-surface mirror + U/W update + free surface + receiver
//                  sampling + illumination + stencil-history store   (acoustic_kernels.py:115-174)
//   ac_adj_fused : receiver-cotangent injection + the whole reverse step 7T..1T of SURVEY.md
//                  Appendix A.1 + g_alpha1 accumulation + g_src
//
// Design (v3, "tile-persistent"):
//   * A CTA of 128 threads owns one 64(x) x 32(z) tile and walks through a chunk of the group's
//     shots.  Everything that does not depend on the shot stays on chip for the whole walk: the
//     coefficients of the thread's cells live in REGISTERS (loaded once per tile), the
//     illumination / gradient sums accumulate in registers and are flushed once per tile with
//     128-bit reductions (red.global.add.v4.f32).
//   * Per shot, the old p,u,w (or lambda) tiles with their halos (4 cells in x, 3 in z) are
//     brought into shared memory by TMA (cp.async.bulk.tensor.3d, hardware zero fill outside the
//     grid) into a double buffer: the loads of shot s+1 are in flight while shot s is computed,
//     so the inner loop never waits on a global load.
//   * Phase 1 computes the intermediate field (new pressure / pressure cotangent) on the tile plus
//     the ring phase 2 needs (halo recompute); every thread owns a float4 of 4 cells in x and a
//     block of 5 rows in z, so all shared-memory traffic is 64/128-bit and the z neighbours are
//     reused from registers.  Phase 2 (4 cells x 4 rows per thread) produces the new fields and
//     writes them with 128-bit stores to the other buffer of a ping-pong pair.
//   * No per-cell region predicates: the update regions of Appendix A.1 are folded into a private,
//     zero-padded copy of the coefficient planes ("coefficient pack": alpha = 0 and kappa = 0
//     outside a field's update region make the update the identity, bit for bit).  Tiles whose
//     neighbourhood has no damping at all run a variant without the kappa terms (1.0f*x == x).
//   * Receivers are bucketed per tile once per call; a tile only touches its own receivers.
// Same arithmetic and association as the generic kernels in acoustic.cu (-fmad=false): forward
// records stay bit-identical to the CPU reference.
//
// The density gradient (g_alpha2) rides on the same pair: ac_fwd_fused<.., SAVE2> stores D+x p and D+z p of the new pressure,
// ac_adj_fused<.., G2> accumulates -(mu_u * D+x p + mu_w * D+z p); the division by alpha2 happens once, in acf_reduce_g2.
#include "common.cuh"
#ifndef ADFWI_HOST_EMUL
#include "tma.cuh"
#include "acoustic_fused.h"
#include <stdlib.h>
#include <stdio.h>

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
constexpr int CMAX = 32;                    // max shots per chunk (per-shot scalars staged in shared memory)
static_assert((TZ + 3) % RB == 0, "phase-1 row blocks must tile rows [-1,TZ+2)");
static_assert(NGP * NB1 <= NTHREADS, "phase 1 must fit one pass");
static_assert(RX <= NTHREADS && CMAX <= NTHREADS, "row fix-ups / scalar staging use one thread per element");
constexpr int RECT_BYTES = ((RZ * RX * 4 + 127) / 128) * 128;   // 11008
constexpr int STAGE_BYTES = 3 * RECT_BYTES;                      // p,u,w of one shot
constexpr int CPX = 8, CPZ = 4;             // apron of the coefficient pack (cells)
constexpr int FWD_STAGES = 2, ADJ_STAGES = 2;                     // TMA pipeline depth (shots in flight + 1)
constexpr int TSM_BYTES = 12 * NTHREADS * 16;                    // adjoint: (1-kappa) factors parked in shared memory

struct FGeom {
    int nzp, nxp, ld, fs, zlo, nabc, nt;
    int ntx, ntz;
    int cpld;                // pitch of the coefficient pack
    size_t plane;            // nzp*ld
    float c1, c2, dt;
};

// coefficient pack: masked planes, pointers pre-offset to logical cell (0,0)
struct CoefPack { const float *a1, *k1, *a2u, *k2, *a2w, *k3, *ew; };   // ew: adjoint scale of the w cotangent (see adj_tile)

struct RcvBuckets { const int* start; const int* id; const int* zx; const unsigned char* nbr; };

struct FwdArgs {
    CoefPack cp; const unsigned char* tflags;
    float *p_out, *u_out, *w_out;
    const float* src_v; const int64_t *sx, *sz;
    float* hist; int hist_len, tl, it;
    float *hist_dx, *hist_dz;         // density gradient: D+x p and D+z p of the post-step pressure (same indexing as hist)
    int nr; RcvBuckets rb;
    float *rcv_p, *rcv_u, *rcv_w;
    float *ill_p, *ill_u; int acc_u;
    int s_begin, s_end, chunk, nchunks;
};

struct AdjArgs {
    CoefPack cp; const unsigned char* tflags;
    float *lp_out, *lu_out, *lw_out;
    const int64_t *sx, *sz;
    const float* hist; int hist_len, tl, it;
    const float *hist_dx, *hist_dz;
    int nr; RcvBuckets rb;
    const float *gp, *gu, *gw;
    float* g1part; float* g2part; float* g_src;
    int s_begin, s_end, chunk, nchunks;
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
// coefficient-pack loads: read-only path, kept in L2 (evict_last) while the wavefields stream through
__device__ __forceinline__ uint64_t l2_keep_policy()
{
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
__device__ __forceinline__ float4 ldk4(const float* p, uint64_t pol)
{
    float4 v;
    asm volatile("ld.global.nc.L2::cache_hint.v4.f32 {%0,%1,%2,%3}, [%4], %5;" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p), "l"(pol));
    return v;
}

// (tile, shot) sequence of one persistent CTA: items blockIdx.x, +gridDim.x, ...; each item is one
// tile and one chunk of the group's shots.  The TMA producer runs two shots ahead of the compute
// along this flat sequence, across tile boundaries.
// Item order: the adjoint deals the items tile-major (item = tile * nchunks + chunk), the forward kernel chunk-major (item = chunk *
// ntiles + tile: CTA k and CTA k + 1 then work on neighbouring tiles of the SAME shots).  Measured on the C2 grid: forward 0.0521 ->
// 0.0505 ms chunk-major, adjoint 0.0587 -> 0.0595 ms (so it keeps tile-major).
template <bool CM> __device__ __forceinline__ void item_decode(int it, int ntiles, int nchunks, int& tile, int& ch)
{
    if (CM) { ch = it / ntiles; tile = it - ch * ntiles; }
    else    { tile = it / nchunks; ch = it - tile * nchunks; }
}
template <bool CM>
struct Cursor {
    int item, s, s_hi, X0, Z0; bool valid;
    __device__ __forceinline__ void set(int it, const FGeom& g, int s_begin, int s_end, int chunk, int nchunks)
    {
        item = it; valid = it < g.ntx * g.ntz * nchunks;
        if (valid) {
            int tile, ch;
            item_decode<CM>(it, g.ntx * g.ntz, nchunks, tile, ch);
            const int tzi = tile / g.ntx, txi = tile - tzi * g.ntx;
            X0 = txi * TX; Z0 = g.zlo + tzi * TZ;
            s = s_begin + ch * chunk; s_hi = min(s + chunk, s_end);
        }
    }
    __device__ __forceinline__ void next(const FGeom& g, int s_begin, int s_end, int chunk, int nchunks)
    {
        if (valid && ++s >= s_hi) set(item + gridDim.x, g, s_begin, s_end, chunk, nchunks);
    }
};
template <bool CM> __device__ __forceinline__ void issue_stage(const Cursor<CM>& c, unsigned char* smem_raw, uint64_t* bar, int k,
                                            const CUtensorMap* t0, const CUtensorMap* t1, const CUtensorMap* t2)
{
    float* st = (float*)(smem_raw + k * STAGE_BYTES);
    fence_proxy_async();
    mbar_expect_tx(bar + k, 3 * RZ * RX * 4);
    tma_load_3d(st, t0, c.X0 - HX, c.Z0 - HZ, c.s, bar + k);
    tma_load_3d(st + RECT_BYTES / 4, t1, c.X0 - HX, c.Z0 - HZ, c.s, bar + k);
    tma_load_3d(st + 2 * RECT_BYTES / 4, t2, c.X0 - HX, c.Z0 - HZ, c.s, bar + k);
}

// TMA prefetch of a box into L2 (no shared-memory destination): the adjoint reads its history planes with plain 128-bit loads,
// which then hit L2 instead of paying a DRAM round trip on the critical path
__device__ __forceinline__ void acf_prefetch_3d(const CUtensorMap* tm, int c0, int c1, int c2)
{
    asm volatile("cp.async.bulk.prefetch.tensor.3d.L2.global.tile [%0, {%1, %2, %3}];" ::"l"(tm), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
struct HistMaps { const CUtensorMap *s, *dx, *dz; };
template <bool G2> __device__ __forceinline__ void adj_prefetch_hist(const Cursor<false>& c, const HistMaps& hm, int hist_len, int tl)
{
    const int pl = c.s * hist_len + tl;
    acf_prefetch_3d(hm.s, c.X0, c.Z0, pl);
    if (G2) { acf_prefetch_3d(hm.dx, c.X0, c.Z0, pl); acf_prefetch_3d(hm.dz, c.X0, c.Z0, pl); }
}

// programmatic dependent launch: the next time step's grid may become resident while this one drains
// (its prologue -- barrier init, coefficient loads -- overlaps our tail); it must not touch anything we
// produce before griddep_wait() returns, i.e. before this grid has completed and flushed.
__device__ __forceinline__ void griddep_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void griddep_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

__device__ __forceinline__ void red4(float* p, const float4& v)
{
    asm volatile("red.global.add.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ float4 one_minus(const float4& k) { return make_float4(1.0f - k.x, 1.0f - k.y, 1.0f - k.z, 1.0f - k.w); }
__device__ __forceinline__ float4 neg4(const float4& k) { return make_float4(-k.x, -k.y, -k.z, -k.w); }

// thread roles shared by both kernels
struct Roles {
    int b1, gi1, c01, r01, so1; bool p1_active;    // phase 1: float4 group gi1 in [-1,NG], row block b1
    int q2, l2, c02, r02, so2;                     // phase 2: float4 group l2, row quad q2
    __device__ __forceinline__ explicit Roles(int tid)
    {
        b1 = tid / NGP; gi1 = tid - b1 * NGP - 1; p1_active = tid < NGP * NB1;
        c01 = 4 * gi1; r01 = RB * b1 - 1; so1 = (r01 + HZ) * RX + c01 + HX;
        q2 = tid / NG; l2 = tid - q2 * NG; c02 = 4 * l2; r02 = 4 * q2; so2 = (r02 + HZ) * RX + c02 + HX;
    }
};

// ------------------------------------------------------------------------------------------
// forward: one tile, shots [s_lo, s_hi)
// ------------------------------------------------------------------------------------------
template <bool FS, bool SAVE, bool SAVE2, bool ILLUM, bool PML>
__device__ __forceinline__ void fwd_tile(const CUtensorMap* tm_p, const CUtensorMap* tm_u, const CUtensorMap* tm_w,
                                         const FGeom& g, const FwdArgs& a, unsigned char* smem_raw, uint64_t* bar,
                                         uint32_t& par, int& stage, Cursor<true>& pc, int* s_sz, int* s_sx, float* s_sv,
                                         const Roles& R, int tid, int tile, int s_lo, int s_hi, int chunk, bool first)
{
    float* pn = (float*)(smem_raw + FWD_STAGES * STAGE_BYTES);
    const uint64_t pol = l2_keep_policy();
    const int ld = g.ld, cpld = g.cpld;
    const float c1 = g.c1, c2 = g.c2;
    const int tzi = tile / g.ntx, txi = tile - tzi * g.ntx;
    const int X0 = txi * TX, Z0 = g.zlo + tzi * TZ;
    // ---- prologue: stage per-shot scalars, load this tile's coefficients into registers ----------
    if (tid < s_hi - s_lo) {
        const int s = s_lo + tid;
        s_sz[tid] = (int)a.sz[s]; s_sx[tid] = (int)a.sx[s];
        s_sv[tid] = g.dt * a.src_v[(size_t)s * g.nt + a.it];
    }
    const int gz1 = Z0 + R.r01, gx1 = X0 + R.c01;
    const int gz2 = Z0 + R.r02, gx2 = X0 + R.c02;
    float4 A1[RB], T1[RB];
    if (R.p1_active) {
        const ptrdiff_t cpo = (ptrdiff_t)gz1 * cpld + gx1;       // may be negative: the pack has an apron
#pragma unroll
        for (int j = 0; j < RB; ++j) {
            A1[j] = ldk4(a.cp.a1 + cpo + (ptrdiff_t)j * cpld, pol);
            if (PML) T1[j] = one_minus(ldk4(a.cp.k1 + cpo + (ptrdiff_t)j * cpld, pol));
        }
    }
    float4 AU[4], AW[4], T2[4], T3[4];
    {
        const ptrdiff_t cpo = (ptrdiff_t)gz2 * cpld + gx2;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            AU[j] = ldk4(a.cp.a2u + cpo + (ptrdiff_t)j * cpld, pol);
            if (PML) {
                AW[j] = ldk4(a.cp.a2w + cpo + (ptrdiff_t)j * cpld, pol);
                T2[j] = one_minus(ldk4(a.cp.k2 + cpo + (ptrdiff_t)j * cpld, pol));
                T3[j] = one_minus(ldk4(a.cp.k3 + cpo + (ptrdiff_t)j * cpld, pol));
            }
        }
    }
    float4 accp[4], accu[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) { accp[j] = make_float4(0.f, 0.f, 0.f, 0.f); accu[j] = make_float4(0.f, 0.f, 0.f, 0.f); }
    const int rcv_lo = a.nr > 0 ? a.rb.start[tile] : 0, rcv_hi = a.nr > 0 ? a.rb.start[tile + 1] : 0;
    const bool has_rcv = rcv_hi > rcv_lo;
    const bool col1ok = (R.gi1 >= 0) && (R.gi1 < NG) && (gx1 < ld);
    const bool col2ok = gx2 < ld;
    if (first) {      // first tile of this CTA: the coefficient loads above are in flight; now wait for the previous
                      // step's grid and start the TMA producer (two stages ahead)
        griddep_wait();
#pragma unroll
        for (int k = 0; k < FWD_STAGES; ++k)
            if (pc.valid) { if (tid == 0) issue_stage(pc, smem_raw, bar, k, tm_p, tm_u, tm_w); pc.next(g, a.s_begin, a.s_end, a.chunk, a.nchunks); }
    }
    __syncthreads();          // s_sz/s_sx/s_sv visible

    for (int s = s_lo; s < s_hi; ++s) {
        const int k = stage;
        const float* ps = (const float*)(smem_raw + k * STAGE_BYTES);
        float* us = (float*)(smem_raw + k * STAGE_BYTES + RECT_BYTES);
        float* ws = (float*)(smem_raw + k * STAGE_BYTES + 2 * RECT_BYTES);
        const int szs = s_sz[s - s_lo], sxs = s_sx[s - s_lo];
        const bool has_src = (szs >= Z0 - 1) && (szs < Z0 + TZ + 2) && (sxs >= X0 - 1) && (sxs < X0 + TX + 2);
        while (!mbar_try(bar + k, (par >> k) & 1u)) {}
        par ^= 1u << k;
        // ---- phase 1: new pressure on rows [-1,TZ+2) x cols [-4,TX+4) (cols [-1,TX+2) are used) ----
        if (R.p1_active) {
            float4 wq[RB + 3];
#pragma unroll
            for (int q = 0; q < RB + 3; ++q) wq[q] = ld4(ws + R.so1 + (q - 2) * RX);
            float4 pv[RB];
#pragma unroll
            for (int j = 0; j < RB; ++j) {
                const float* ur = us + R.so1 + j * RX;
                const float2 ul = ld2(ur - 2);
                const float4 um = ld4(ur);
                const float uR = ur[4];
                const float4 po = ld4(ps + R.so1 + j * RX);
                const float4 w0 = wq[j + 2], wm1 = wq[j + 1], wp1 = wq[j + 3], wm2 = wq[j];
                float4 S;
                S.x = c1 * (((um.x - ul.y) + w0.x) - wm1.x) + c2 * (((um.y - ul.x) + wp1.x) - wm2.x);
                S.y = c1 * (((um.y - um.x) + w0.y) - wm1.y) + c2 * (((um.z - ul.y) + wp1.y) - wm2.y);
                S.z = c1 * (((um.z - um.y) + w0.z) - wm1.z) + c2 * (((um.w - um.x) + wp1.z) - wm2.z);
                S.w = c1 * (((um.w - um.z) + w0.w) - wm1.w) + c2 * (((uR - um.y) + wp1.w) - wm2.w);
                if (SAVE) {
                    const int r = R.r01 + j, gz = gz1 + j;
                    if (col1ok && r >= 0 && r < TZ && gz < g.nzp)
                        __stcs(reinterpret_cast<float4*>(a.hist + ((size_t)s * a.hist_len + a.tl) * g.plane + (size_t)gz * ld + gx1), S);
                }
                if (PML) {
                    pv[j].x = T1[j].x * po.x - A1[j].x * S.x; pv[j].y = T1[j].y * po.y - A1[j].y * S.y;
                    pv[j].z = T1[j].z * po.z - A1[j].z * S.z; pv[j].w = T1[j].w * po.w - A1[j].w * S.w;
                } else {
                    pv[j].x = po.x - A1[j].x * S.x; pv[j].y = po.y - A1[j].y * S.y;
                    pv[j].z = po.z - A1[j].z * S.z; pv[j].w = po.w - A1[j].w * S.w;
                }
            }
            if (has_src) {     // p[sz,sx] += dt*src   (acoustic_kernels.py:131-132)
                const int dr = szs - gz1, dc = sxs - gx1;
                if (dr >= 0 && dr < RB && dc >= 0 && dc < 4) {
                    const float srcval = s_sv[s - s_lo];
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
            if (FS && tzi == 0 && R.b1 == 0) {      // p[fs-1] = -p[fs+1]: tile rows 0 and 2 = block rows 1 and 3
                pv[1].x = -pv[3].x; pv[1].y = -pv[3].y; pv[1].z = -pv[3].z; pv[1].w = -pv[3].w;
            }
#pragma unroll
            for (int j = 0; j < RB; ++j) st4(pn + R.so1 + j * RX, pv[j]);
        }
        __syncthreads();
        // ---- phase 2: velocities on the interior, stores, illumination -----------------------------
        {
            float4 P[7];
#pragma unroll
            for (int q = 0; q < 7; ++q) P[q] = ld4(pn + R.so2 + (q - 1) * RX);
            float4 uv[4], wv[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float* pr = pn + R.so2 + j * RX;
                const float pL = pr[-1];
                const float2 pR = ld2(pr + 4);
                const float4 p0 = P[j + 1], pm1 = P[j], pp1 = P[j + 2], pp2 = P[j + 3];
                const float4 uo = ld4(us + R.so2 + j * RX), wo = ld4(ws + R.so2 + j * RX);
                const float4 au = AU[j], aw = PML ? AW[j] : AU[j];
                float4 dxp, dzp, du, dw;      // D+x p, D+z p of the new pressure (acoustic_kernels.py:139-160)
                dxp.x = c1 * (p0.y - p0.x) + c2 * (p0.z - pL);
                dxp.y = c1 * (p0.z - p0.y) + c2 * (p0.w - p0.x);
                dxp.z = c1 * (p0.w - p0.z) + c2 * (pR.x - p0.y);
                dxp.w = c1 * (pR.x - p0.w) + c2 * (pR.y - p0.z);
                dzp.x = c1 * (pp1.x - p0.x) + c2 * (pp2.x - pm1.x);
                dzp.y = c1 * (pp1.y - p0.y) + c2 * (pp2.y - pm1.y);
                dzp.z = c1 * (pp1.z - p0.z) + c2 * (pp2.z - pm1.z);
                dzp.w = c1 * (pp1.w - p0.w) + c2 * (pp2.w - pm1.w);
                du.x = au.x * dxp.x; du.y = au.y * dxp.y; du.z = au.z * dxp.z; du.w = au.w * dxp.w;
                dw.x = aw.x * dzp.x; dw.y = aw.y * dzp.y; dw.z = aw.z * dzp.z; dw.w = aw.w * dzp.w;
                if (SAVE && SAVE2) {          // what the density gradient needs: g_alpha2 -= lambda_u * D+x p + lambda_w * D+z p
                    const int gz = gz2 + j;
                    if (col2ok && gz < g.nzp) {
                        const size_t ho = ((size_t)s * a.hist_len + a.tl) * g.plane + (size_t)gz * ld + gx2;
                        __stcs(reinterpret_cast<float4*>(a.hist_dx + ho), dxp);
                        __stcs(reinterpret_cast<float4*>(a.hist_dz + ho), dzp);
                    }
                }
                if (PML) {
                    uv[j].x = T2[j].x * uo.x - du.x; uv[j].y = T2[j].y * uo.y - du.y; uv[j].z = T2[j].z * uo.z - du.z; uv[j].w = T2[j].w * uo.w - du.w;
                    wv[j].x = T3[j].x * wo.x - dw.x; wv[j].y = T3[j].y * wo.y - dw.y; wv[j].z = T3[j].z * wo.z - dw.z; wv[j].w = T3[j].w * wo.w - dw.w;
                } else {
                    uv[j].x = uo.x - du.x; uv[j].y = uo.y - du.y; uv[j].z = uo.z - du.z; uv[j].w = uo.w - du.w;
                    wv[j].x = wo.x - dw.x; wv[j].y = wo.y - dw.y; wv[j].z = wo.z - dw.z; wv[j].w = wo.w - dw.w;
                }
            }
            if (FS && tzi == 0 && R.q2 == 0) wv[0] = wv[1];        // w[fs-1] = w[fs]   (acoustic_kernels.py:163-164)
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int gz = gz2 + j;
                const float4 p0 = P[j + 1];
                if (col2ok && gz < g.nzp) {
                    const size_t o = (size_t)s * g.plane + (size_t)gz * ld + gx2;
                    st4(a.p_out + o, p0);
                    st4(a.u_out + o, uv[j]);
                    st4(a.w_out + o, wv[j]);
                }
                if (ILLUM) {      // summed over the chunk's shots in registers; cropped when finalised
                    accp[j].x += p0.x * p0.x; accp[j].y += p0.y * p0.y; accp[j].z += p0.z * p0.z; accp[j].w += p0.w * p0.w;
                    accu[j].x += uv[j].x * uv[j].x; accu[j].y += uv[j].y * uv[j].y; accu[j].z += uv[j].z * uv[j].z; accu[j].w += uv[j].w * uv[j].w;
                }
                if (has_rcv) { st4(us + R.so2 + j * RX, uv[j]); st4(ws + R.so2 + j * RX, wv[j]); }
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
            fence_proxy_async();      // generic-proxy writes to the stage precede its TMA refill
        }
        __syncthreads();
        // refill this stage with the (tile, shot) two steps ahead while the next one is computed
        if (pc.valid) { if (tid == 0) issue_stage(pc, smem_raw, bar, k, tm_p, tm_u, tm_w); pc.next(g, a.s_begin, a.s_end, a.chunk, a.nchunks); }
        stage = (stage + 1 == FWD_STAGES) ? 0 : stage + 1;
    }
    if (ILLUM) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int gz = gz2 + j;
            if (col2ok && gz < g.nzp) {
                const size_t cg = (size_t)chunk * g.plane + (size_t)gz * ld + gx2;
                red4(a.ill_p + cg, accp[j]);
                if (a.acc_u) red4(a.ill_u + cg, accu[j]);
            }
        }
    }
}

template <bool FS, bool SAVE, bool SAVE2, bool ILLUM>
__global__ void __launch_bounds__(NTHREADS, 2)
ac_fwd_fused(const __grid_constant__ CUtensorMap tm_p, const __grid_constant__ CUtensorMap tm_u,
             const __grid_constant__ CUtensorMap tm_w, const FGeom g, const FwdArgs a)
{
    extern __shared__ __align__(128) unsigned char smem_raw[];
    uint64_t* bar = (uint64_t*)(smem_raw + FWD_STAGES * STAGE_BYTES + RECT_BYTES);
    int* s_sz = (int*)(bar + 4);
    int* s_sx = s_sz + CMAX;
    float* s_sv = (float*)(s_sx + CMAX);
    const int tid = threadIdx.x;
    if (tid == 0) { for (int k = 0; k < FWD_STAGES; ++k) mbar_init(bar + k, 1); }
    __syncthreads();
    const Roles R(tid);
    uint32_t par = 0;
    int stage = 0;
    const int nitems = g.ntx * g.ntz * a.nchunks;
    Cursor<true> pc;
    pc.set(blockIdx.x, g, a.s_begin, a.s_end, a.chunk, a.nchunks);
    griddep_launch_dependents();
    for (int item = blockIdx.x; item < nitems; item += gridDim.x) {
        int tile, chunk;
        item_decode<true>(item, g.ntx * g.ntz, a.nchunks, tile, chunk);
        const int s_lo = a.s_begin + chunk * a.chunk;
        const int s_hi = min(s_lo + a.chunk, a.s_end);
        const bool first = item == (int)blockIdx.x;
        if (a.tflags[tile]) fwd_tile<FS, SAVE, SAVE2, ILLUM, true>(&tm_p, &tm_u, &tm_w, g, a, smem_raw, bar, par, stage, pc, s_sz, s_sx, s_sv, R, tid, tile, s_lo, s_hi, chunk, first);
        else                fwd_tile<FS, SAVE, SAVE2, ILLUM, false>(&tm_p, &tm_u, &tm_w, g, a, smem_raw, bar, par, stage, pc, s_sz, s_sx, s_sv, R, tid, tile, s_lo, s_hi, chunk, first);
        __syncthreads();      // per-shot scalars and stages are reused by the next item
    }
}

// ------------------------------------------------------------------------------------------
// adjoint: one tile, shots [s_lo, s_hi)
//
// State carried between reverse steps: lambda_p and the SCALED velocity cotangents
//     mu_u = alpha2u * lambda_u,   mu_w = ew * lambda_w      (ew = alpha2w on the W region; 1 on row fs-1
//                                                            with a free surface, where lambda_w is parked raw)
// alpha2 is time-invariant, so mu obeys  mu_new = (1-kappa)*mu + alpha2 * D^T(m): the coefficient is needed
// only at the thread's own cells (phase 2) and phase 1 -- lambda_p1 = lambda_p + D+z^T(-mu_w) + D+x^T(-mu_u)
// -- needs no coefficient at all.  Cells outside a region have alpha2 = 0 there and never feed back.
// ------------------------------------------------------------------------------------------
template <bool FS, bool PML, bool G2>
__device__ __forceinline__ void adj_tile(const CUtensorMap* tm_lp, const CUtensorMap* tm_lu, const CUtensorMap* tm_lw, const HistMaps& hm,
                                         const FGeom& g, const AdjArgs& a, unsigned char* smem_raw, uint64_t* bar,
                                         uint32_t& par, int& stage, Cursor<false>& pc, int* s_sz, int* s_sx,
                                         const Roles& R, int tid, int tile, int s_lo, int s_hi, int chunk, bool first)
{
    float* lp1 = (float*)(smem_raw + ADJ_STAGES * STAGE_BYTES);                  // lambda_p after undoing W,U (and 3T)
    float* mps = (float*)(smem_raw + ADJ_STAGES * STAGE_BYTES + RECT_BYTES);     // m = -alpha1 * lambda_p1
    float4* tsm = (float4*)(smem_raw + ADJ_STAGES * STAGE_BYTES + 2 * RECT_BYTES);   // per-thread (1-kappa) factors, [12][NTHREADS]
    const uint64_t pol = l2_keep_policy();
    const int ld = g.ld, cpld = g.cpld;
    const float c1 = g.c1, c2 = g.c2;
    const int tzi = tile / g.ntx, txi = tile - tzi * g.ntx;
    const int X0 = txi * TX, Z0 = g.zlo + tzi * TZ;
    const bool have_g = a.nr > 0 && (a.gp || a.gu || a.gw);
    if (a.g_src && tid < s_hi - s_lo) { s_sz[tid] = (int)a.sz[s_lo + tid]; s_sx[tid] = (int)a.sx[s_lo + tid]; }
    const int gz1 = Z0 + R.r01, gx1 = X0 + R.c01;
    const int gz2 = Z0 + R.r02, gx2 = X0 + R.c02;
    float4 NA1[RB];
    if (R.p1_active) {
        const ptrdiff_t cpo = (ptrdiff_t)gz1 * cpld + gx1;
#pragma unroll
        for (int j = 0; j < RB; ++j) NA1[j] = neg4(ldk4(a.cp.a1 + cpo + (ptrdiff_t)j * cpld, pol));
    }
    float4 EU[4], EW[4];
    {
        const ptrdiff_t cpo = (ptrdiff_t)gz2 * cpld + gx2;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            EU[j] = ldk4(a.cp.a2u + cpo + (ptrdiff_t)j * cpld, pol);
            if (PML) {       // (1-kappa) of the thread's phase-2 cells: parked in shared memory (register budget)
                EW[j] = ldk4(a.cp.ew + cpo + (ptrdiff_t)j * cpld, pol);
                tsm[(3 * j + 0) * NTHREADS + tid] = one_minus(ldk4(a.cp.k1 + cpo + (ptrdiff_t)j * cpld, pol));
                tsm[(3 * j + 1) * NTHREADS + tid] = one_minus(ldk4(a.cp.k2 + cpo + (ptrdiff_t)j * cpld, pol));
                tsm[(3 * j + 2) * NTHREADS + tid] = one_minus(ldk4(a.cp.k3 + cpo + (ptrdiff_t)j * cpld, pol));
            }
        }
    }
    float4 gacc[4], g2acc[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) { gacc[j] = make_float4(0.f, 0.f, 0.f, 0.f); g2acc[j] = make_float4(0.f, 0.f, 0.f, 0.f); }
    const bool inject = have_g && a.rb.nbr[tile];
    const bool col2ok = gx2 < ld;
    if (first) {
        griddep_wait();
#pragma unroll
        for (int k = 0; k < ADJ_STAGES; ++k)
            if (pc.valid) { if (tid == 0) { issue_stage(pc, smem_raw, bar, k, tm_lp, tm_lu, tm_lw); adj_prefetch_hist<G2>(pc, hm, a.hist_len, a.tl); } pc.next(g, a.s_begin, a.s_end, a.chunk, a.nchunks); }
    }
    __syncthreads();

    for (int s = s_lo; s < s_hi; ++s) {
        const int k = stage;
        float* lps = (float*)(smem_raw + k * STAGE_BYTES);
        float* lus = (float*)(smem_raw + k * STAGE_BYTES + RECT_BYTES);
        float* lws = (float*)(smem_raw + k * STAGE_BYTES + 2 * RECT_BYTES);
        // the stencil history of this step does not depend on anything computed here: fetch it early
        float4 Sh[4];
        {
            const float* Hs = a.hist + ((size_t)s * a.hist_len + a.tl) * g.plane;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int gz = gz2 + j;
                Sh[j] = (col2ok && gz < g.nzp) ? __ldcs(reinterpret_cast<const float4*>(Hs + (size_t)gz * ld + gx2)) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
        }
        while (!mbar_try(bar + k, (par >> k) & 1u)) {}
        par ^= 1u << k;
        // ---- 7T: receiver cotangents into the staged rectangle (duplicates legal -> shared atomics);
        //      the u / w cotangents enter scaled by the cell's alpha2 (mu = alpha2 * lambda)
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
                        const int rz = zx >> 16, rx = zx & 0xffff;
                        const int z = rz - (Z0 - HZ), x = rx - (X0 - HX);
                        if (z >= 0 && z < RZ && x >= 0 && x < RX) {
                            const size_t o = ((size_t)s * g.nt + a.it) * a.nr + a.rb.id[i];
                            const ptrdiff_t co = (ptrdiff_t)rz * cpld + rx;
                            if (a.gp) atomicAdd(lps + z * RX + x, a.gp[o]);
                            if (a.gu) atomicAdd(lus + z * RX + x, a.cp.a2u[co] * a.gu[o]);
                            if (a.gw) atomicAdd(lws + z * RX + x, a.cp.ew[co] * a.gw[o]);
                        }
                    }
                }
            }
            __syncthreads();
        }
        if (FS && tzi == 0) {        // 6T: lambda_w[fs] += lambda_w[fs-1]; lambda_w[fs-1] = 0  (staged rows HZ+1 and HZ;
                                     //     row fs-1 holds the raw cotangent, row fs the scaled one)
            if (tid < RX) {
                const float e = a.cp.a2w[(ptrdiff_t)(Z0 + 1) * cpld + (X0 - HX + tid)];
                lws[(HZ + 1) * RX + tid] += e * lws[HZ * RX + tid]; lws[HZ * RX + tid] = 0.f;
            }
            __syncthreads();
        }
        // ---- phase 1: lambda_p after undoing W and U (5T, 4T), 3T, and m = -alpha1*lambda_p ------------
        if (R.p1_active) {
            float4 qw[RB + 3];     // mu_w rows r01-2 .. r01+5
#pragma unroll
            for (int q = 0; q < RB + 3; ++q) qw[q] = ld4(lws + R.so1 + (q - 2) * RX);
            float4 acc[RB];
#pragma unroll
            for (int j = 0; j < RB; ++j) {
                const float* ur = lus + R.so1 + j * RX;
                const float2 ul = ld2(ur - 2); const float4 um = ld4(ur); const float uR = ur[4];
                // mu_u at columns c0-2 .. c0+4
                const float q0 = ul.x, q1 = ul.y, q2_ = um.x, q3 = um.y, q4 = um.z, q5 = um.w, q6 = uR;
                const float4 w1 = qw[j + 1], w2 = qw[j + 2], w0 = qw[j], w3 = qw[j + 3];
                float4 v = ld4(lps + R.so1 + j * RX);
                // transposes of D+z and D+x applied to -mu:  -c1 mu[z-1] + c1 mu[z] - c2 mu[z-2] + c2 mu[z+1]
                v.x += c1 * w2.x - c1 * w1.x + c2 * w3.x - c2 * w0.x;
                v.y += c1 * w2.y - c1 * w1.y + c2 * w3.y - c2 * w0.y;
                v.z += c1 * w2.z - c1 * w1.z + c2 * w3.z - c2 * w0.z;
                v.w += c1 * w2.w - c1 * w1.w + c2 * w3.w - c2 * w0.w;
                v.x += c1 * q2_ - c1 * q1 + c2 * q3 - c2 * q0;
                v.y += c1 * q3 - c1 * q2_ + c2 * q4 - c2 * q1;
                v.z += c1 * q4 - c1 * q3 + c2 * q5 - c2 * q2_;
                v.w += c1 * q5 - c1 * q4 + c2 * q6 - c2 * q3;
                acc[j] = v;
            }
            if (FS && tzi == 0 && R.b1 == 0) {   // 3T: lambda_p[fs+1] -= lambda_p[fs-1]; lambda_p[fs-1] = 0 (block rows 3 and 1)
                acc[3].x -= acc[1].x; acc[3].y -= acc[1].y; acc[3].z -= acc[1].z; acc[3].w -= acc[1].w;
                acc[1] = make_float4(0.f, 0.f, 0.f, 0.f);
            }
#pragma unroll
            for (int j = 0; j < RB; ++j) {
                float4 m;
                m.x = NA1[j].x * acc[j].x; m.y = NA1[j].y * acc[j].y; m.z = NA1[j].z * acc[j].z; m.w = NA1[j].w * acc[j].w;
                st4(lp1 + R.so1 + j * RX, acc[j]);
                st4(mps + R.so1 + j * RX, m);
            }
        }
        __syncthreads();
        // ---- phase 2: new mu_u, mu_w, lambda_p on the interior (5T,4T,1T), g_alpha1, g_src ------------
        {
            float4 M[7];
#pragma unroll
            for (int q = 0; q < 7; ++q) M[q] = ld4(mps + R.so2 + (q - 1) * RX);
            // density gradient: the two extra history planes are loaded one row ahead of their use (register double buffer)
            const size_t hrow0 = ((size_t)s * a.hist_len + a.tl) * g.plane + (size_t)gz2 * ld + gx2;
            float4 hu_n = make_float4(0.f, 0.f, 0.f, 0.f), hw_n = hu_n;
            if (G2 && col2ok && gz2 < g.nzp) {
                hu_n = __ldcs(reinterpret_cast<const float4*>(a.hist_dx + hrow0)); hw_n = __ldcs(reinterpret_cast<const float4*>(a.hist_dz + hrow0));
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int gz = gz2 + j;
                const float4 hu = hu_n, hw = hw_n;
                if (G2 && j < 3) {
                    hu_n = make_float4(0.f, 0.f, 0.f, 0.f); hw_n = hu_n;
                    if (col2ok && gz + 1 < g.nzp) {
                        hu_n = __ldcs(reinterpret_cast<const float4*>(a.hist_dx + hrow0 + (size_t)(j + 1) * ld));
                        hw_n = __ldcs(reinterpret_cast<const float4*>(a.hist_dz + hrow0 + (size_t)(j + 1) * ld));
                    }
                }
                const float* mr = mps + R.so2 + j * RX;
                const float mL = mr[-1];
                const float2 mR = ld2(mr + 4);
                const float4 m0 = M[j + 1], mm1 = M[j], mp1 = M[j + 2], mp2 = M[j + 3];
                const float4 qp = ld4(lp1 + R.so2 + j * RX);
                const float4 luo = ld4(lus + R.so2 + j * RX), lwo = ld4(lws + R.so2 + j * RX);
                const float4 eu = EU[j], ew = PML ? EW[j] : EU[j];
                float4 du, dw, nu, nw, np;
                // transpose of D-x:  +c1 m[x] - c1 m[x+1] + c2 m[x-1] - c2 m[x+2]
                du.x = eu.x * (c1 * m0.x - c1 * m0.y + c2 * mL - c2 * m0.z);
                du.y = eu.y * (c1 * m0.y - c1 * m0.z + c2 * m0.x - c2 * m0.w);
                du.z = eu.z * (c1 * m0.z - c1 * m0.w + c2 * m0.y - c2 * mR.x);
                du.w = eu.w * (c1 * m0.w - c1 * mR.x + c2 * m0.z - c2 * mR.y);
                // transpose of D-z:  +c1 m[z] - c1 m[z+1] + c2 m[z-1] - c2 m[z+2]
                dw.x = ew.x * (c1 * m0.x - c1 * mp1.x + c2 * mm1.x - c2 * mp2.x);
                dw.y = ew.y * (c1 * m0.y - c1 * mp1.y + c2 * mm1.y - c2 * mp2.y);
                dw.z = ew.z * (c1 * m0.z - c1 * mp1.z + c2 * mm1.z - c2 * mp2.z);
                dw.w = ew.w * (c1 * m0.w - c1 * mp1.w + c2 * mm1.w - c2 * mp2.w);
                if (PML) {
                    const float4 t1 = tsm[(3 * j + 0) * NTHREADS + tid], t2 = tsm[(3 * j + 1) * NTHREADS + tid], t3 = tsm[(3 * j + 2) * NTHREADS + tid];
                    nu.x = t2.x * luo.x + du.x; nu.y = t2.y * luo.y + du.y; nu.z = t2.z * luo.z + du.z; nu.w = t2.w * luo.w + du.w;
                    nw.x = t3.x * lwo.x + dw.x; nw.y = t3.y * lwo.y + dw.y; nw.z = t3.z * lwo.z + dw.z; nw.w = t3.w * lwo.w + dw.w;
                    np.x = t1.x * qp.x; np.y = t1.y * qp.y; np.z = t1.z * qp.z; np.w = t1.w * qp.w;
                } else {
                    nu.x = luo.x + du.x; nu.y = luo.y + du.y; nu.z = luo.z + du.z; nu.w = luo.w + du.w;
                    nw.x = lwo.x + dw.x; nw.y = lwo.y + dw.y; nw.z = lwo.z + dw.z; nw.w = lwo.w + dw.w;
                    np = qp;
                }
                if (col2ok && gz < g.nzp) {
                    const size_t o = (size_t)s * g.plane + (size_t)gz * ld + gx2;
                    st4(a.lu_out + o, nu);
                    st4(a.lw_out + o, nw);
                    st4(a.lp_out + o, np);
                    if (a.g_src) {
                        const int dz = s_sz[s - s_lo] - gz, dx = s_sx[s - s_lo] - gx2;
                        if (dz == 0 && dx >= 0 && dx < 4)
                            a.g_src[(size_t)s * g.nt + a.it] = g.dt * (dx == 0 ? qp.x : dx == 1 ? qp.y : dx == 2 ? qp.z : qp.w);
                    }
                }
                // g_alpha1 -= lambda_p1 * S  (masked to the P region when reduced)
                gacc[j].x = gacc[j].x - qp.x * Sh[j].x; gacc[j].y = gacc[j].y - qp.y * Sh[j].y;
                gacc[j].z = gacc[j].z - qp.z * Sh[j].z; gacc[j].w = gacc[j].w - qp.w * Sh[j].w;
                if (G2) {
                    // density gradient (4T, 5T): g_alpha2 -= lambda_u * D+x p + lambda_w * D+z p with the post-injection, post-6T
                    // cotangents.  The staged state is mu = alpha2 * lambda (zero outside a field's region), so the sum of
                    // mu * D p is accumulated here and divided by alpha2 once, when the partial planes are reduced.
                    g2acc[j].x = g2acc[j].x - (luo.x * hu.x + lwo.x * hw.x); g2acc[j].y = g2acc[j].y - (luo.y * hu.y + lwo.y * hw.y);
                    g2acc[j].z = g2acc[j].z - (luo.z * hu.z + lwo.z * hw.z); g2acc[j].w = g2acc[j].w - (luo.w * hu.w + lwo.w * hw.w);
                }
            }
        }
        if (inject || (FS && tzi == 0)) fence_proxy_async();   // generic-proxy writes to the stage precede its TMA refill
        __syncthreads();
        if (pc.valid) { if (tid == 0) { issue_stage(pc, smem_raw, bar, k, tm_lp, tm_lu, tm_lw); adj_prefetch_hist<G2>(pc, hm, a.hist_len, a.tl); } pc.next(g, a.s_begin, a.s_end, a.chunk, a.nchunks); }
        stage = (stage + 1 == ADJ_STAGES) ? 0 : stage + 1;
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const int gz = gz2 + j;
        if (col2ok && gz < g.nzp) {
            red4(a.g1part + (size_t)chunk * g.plane + (size_t)gz * ld + gx2, gacc[j]);
            if (G2) red4(a.g2part + (size_t)chunk * g.plane + (size_t)gz * ld + gx2, g2acc[j]);
        }
    }
}

template <bool FS, bool G2>
__global__ void __launch_bounds__(NTHREADS, 2)
ac_adj_fused(const __grid_constant__ CUtensorMap tm_lp, const __grid_constant__ CUtensorMap tm_lu,
             const __grid_constant__ CUtensorMap tm_lw, const __grid_constant__ CUtensorMap tm_hs,
             const __grid_constant__ CUtensorMap tm_hx, const __grid_constant__ CUtensorMap tm_hz, const FGeom g, const AdjArgs a)
{
    HistMaps hm; hm.s = &tm_hs; hm.dx = &tm_hx; hm.dz = &tm_hz;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    uint64_t* bar = (uint64_t*)(smem_raw + ADJ_STAGES * STAGE_BYTES + 2 * RECT_BYTES + TSM_BYTES);
    int* s_sz = (int*)(bar + 4);
    int* s_sx = s_sz + CMAX;
    const int tid = threadIdx.x;
    if (tid == 0) { for (int k = 0; k < ADJ_STAGES; ++k) mbar_init(bar + k, 1); }
    __syncthreads();
    const Roles R(tid);
    uint32_t par = 0;
    int stage = 0;
    const int nitems = g.ntx * g.ntz * a.nchunks;
    Cursor<false> pc;
    pc.set(blockIdx.x, g, a.s_begin, a.s_end, a.chunk, a.nchunks);
    griddep_launch_dependents();
    for (int item = blockIdx.x; item < nitems; item += gridDim.x) {
        const int tile = item / a.nchunks, chunk = item - tile * a.nchunks;
        const int s_lo = a.s_begin + chunk * a.chunk;
        const int s_hi = min(s_lo + a.chunk, a.s_end);
        const bool first = item == (int)blockIdx.x;
        if (a.tflags[tile]) adj_tile<FS, true, G2>(&tm_lp, &tm_lu, &tm_lw, hm, g, a, smem_raw, bar, par, stage, pc, s_sz, s_sx, R, tid, tile, s_lo, s_hi, chunk, first);
        else                adj_tile<FS, false, G2>(&tm_lp, &tm_lu, &tm_lw, hm, g, a, smem_raw, bar, par, stage, pc, s_sz, s_sx, R, tid, tile, s_lo, s_hi, chunk, first);
        __syncthreads();
    }
}

// ---- set-up kernels ------------------------------------------------------------------------------
// coefficient pack: seven planes [cprows][cpld], logical cell (z,x) at [(z+CPZ)*cpld + x+CPX]; each
// plane is the caller's coefficient inside the update region of the field it drives and 0 elsewhere
// (update regions: SURVEY.md Appendix A.1 / acoustic_kernels.py:115,139,151).
__global__ void acf_pack_coefs(int nzp, int nxp, int fs, int free_surface, int cprows, int cpld, size_t cpplane,
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
    // adjoint scale of the w cotangent: alpha2 on the W region, 1 on the free-surface row fs-1
    pack[6 * cpplane + o] = inW ? a2[c] : ((free_surface && in && z == fs - 1) ? 1.0f : 0.f);
}

// tile flag = 1 when the tile's neighbourhood (everything either kernel reads from the pack) has a
// non-zero damping term or the U / W masks differ; 0 selects the variant without the kappa terms.
__global__ void acf_tile_flags(const FGeom g, size_t cpplane, const float* __restrict__ pack, unsigned char* __restrict__ flags)
{
    const int tile = blockIdx.x;
    const int tzi = tile / g.ntx, txi = tile - tzi * g.ntx;
    const int X0 = txi * TX, Z0 = g.zlo + tzi * TZ;
    const int W = TX + 2 * CPX, H = TZ + 2 * CPZ;
    int bad = 0;
    for (int i = threadIdx.x; i < W * H; i += blockDim.x) {
        const int zz = Z0 + i / W, xx = X0 + i % W;            // pack coordinates (apron included)
        const size_t o = (size_t)zz * g.cpld + xx;
        if (pack[1 * cpplane + o] != 0.f || pack[3 * cpplane + o] != 0.f || pack[5 * cpplane + o] != 0.f) bad = 1;
        if (pack[2 * cpplane + o] != pack[4 * cpplane + o] || pack[2 * cpplane + o] != pack[6 * cpplane + o]) bad = 1;
    }
    bad = __syncthreads_or(bad);
    if (threadIdx.x == 0) flags[tile] = (unsigned char)(bad ? 1 : 0);
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

// density gradient: the partial planes hold -sum_t (mu_u * D+x p + mu_w * D+z p) with mu = alpha2 * lambda; divide by the
// caller's alpha2 on the union of the U and W update regions (mu is identically zero outside a field's region)
__global__ void acf_reduce_g2(int nzp, int nxp, int ld, int fs, int nparts, const float* __restrict__ part, const float* __restrict__ a2,
                              float* __restrict__ out)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x, z = blockIdx.y;
    if (x >= nxp || z >= nzp) return;
    const bool inU = (z >= fs) && (z < nzp - 1) && (x >= 1) && (x < nxp - 2);
    const bool inW = (z >= fs) && (z < nzp - 2) && (x >= 1) && (x < nxp - 1);
    float acc = 0.f;
    if (inU || inW) {
        for (int k = 0; k < nparts; ++k) acc += part[((size_t)k * nzp + z) * ld + x];
        acc = acc / a2[(size_t)z * nxp + x];
    }
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

constexpr int FWD_SMEM = FWD_STAGES * STAGE_BYTES + RECT_BYTES + 32 + 3 * CMAX * 4 + 64;
constexpr int ADJ_SMEM = ADJ_STAGES * STAGE_BYTES + 2 * RECT_BYTES + TSM_BYTES + 32 + 2 * CMAX * 4 + 64;
static_assert(FWD_STAGES <= 4 && ADJ_STAGES <= 4, "four mbarrier slots are reserved");
static_assert(2 * (FWD_SMEM + 1024) <= 233472 && 2 * (ADJ_SMEM + 1024) <= 233472, "two CTAs per SM must fit in shared memory");
constexpr int CTAS_PER_SM = 2;

// SM count of the CURRENT device (cached per device ordinal: one process may drive several GPUs)
int acf_num_sms()
{
    static int cache[kMaxDevices] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= kMaxDevices) dev = 0;
    int n = cache[dev];
    if (!n) { cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev); if (n <= 0) n = 148; cache[dev] = n; }
    return n;
}

struct FPlan {
    FGeom g;
    int ns, nr, FS, save, need_g2, n_segments;
    int K, nseg, nckpt, G;
    int chunk, nchunks;             // shots one CTA walks through per tile; chunks per full group (forward kernel)
    int chunk_b, nchunks_b;         // the same for the adjoint kernel
    int cprows; size_t cpplane;
    float* pack;                    // 7 masked coefficient planes
    unsigned char* tflags;
    float *st[2][3];                // p,u,w ping-pong
    float *lam[2][3];               // lambda ping-pong
    float *hist, *hist_dx, *hist_dz, *ckpt, *g1part, *g2part, *ill_p, *ill_u, *ill_w;
    int *rcv_cnt, *rcv_start, *rcv_cursor, *rcv_id, *rcv_zx; unsigned char* rcv_nbr;
    size_t bytes;
};

// Shots per (tile, chunk) item of the ADJOINT kernel.  Its items are dealt round-robin, tile-major, to the resident CTAs; with few
// rounds the deal is simulated for every candidate chunk length c (chunks c, c, ..., remainder) and the cheapest taken: cost = the
// busiest CTA's shot-steps plus c0 = 0.78 shot-steps per item (fit of tools/c2_chunk_sweep.sh).  C2, 10 shots per launch: 4 + 4 + 2
// (14 steps for the busiest CTA) instead of 5 + 5 (15): 0.0599 -> 0.0578 ms per launch.  The forward kernel (chunk-major order) is
// faster with the plain rule below (0.0506 against 0.0516 ms), so it keeps it.
inline int acf_pick_chunk(int G, int ntiles, int ncta)
{
    const float c0 = 0.78f;
    const int cmax = G < CMAX ? G : CMAX;
    int best = cmax; float tbest = 3.0e38f;
    for (int c = cmax; c >= 1; --c) {
        const int nch = cdiv(G, c);
        if (c < cmax && cdiv(G, c + 1) == nch) continue;             // same chunk count as the longer candidate, worse remainder
        const long long nitems = (long long)ntiles * nch;
        float t;
        if (nitems > 8LL * ncta) {
            t = (float)ntiles * ((float)G + c0 * (float)nch) / (float)ncta + (float)c + c0;
        } else {
            const int last = G - (nch - 1) * c;
            t = 0.f;
            for (int k = 0; k < ncta && k < nitems; ++k) {
                float load = 0.f;
                for (long long it = k; it < nitems; it += ncta) load += (float)((int)(it % nch) == nch - 1 ? last : c) + c0;
                if (load > t) t = load;
            }
        }
        if (t < tbest) { tbest = t; best = c; }
    }
    return best;
}

int acf_make_plan(const adfwi_acoustic_desc* d, void* ws, FPlan* P, int nsm)
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
    P->need_g2 = (d->save_history && d->need_g_alpha2) ? 1 : 0;
    P->n_segments = d->n_segments > 0 ? d->n_segments : 1;
    int K = d->ckpt_interval;
    if (K <= 0 || K >= d->nt) K = d->nt;
    P->K = K; P->nseg = cdiv(d->nt, K); P->nckpt = P->nseg > 2 ? P->nseg - 2 : 0;
    // shots per launch: all of them by default (the wavefields stream through HBM; the tile-persistent
    // CTAs amortise their coefficient loads over the shots they walk through)
    int G = d->shots_per_group;
    if (G <= 0 || G > d->ns) G = d->ns;
    P->G = G;
    // shots per CTA walk: enough (tile, chunk) items for ~3 rounds over the resident CTAs
    const int ntiles = g.ntx * g.ntz;
    int nchunks = d->reserved[1] > 0 ? cdiv(G, d->reserved[1]) : (3 * CTAS_PER_SM * nsm + ntiles / 2) / ntiles;
    if (nchunks < 1) nchunks = 1;
    if (nchunks > G) nchunks = G;
    int chunk = cdiv(G, nchunks);
    if (chunk > CMAX) chunk = CMAX;
    P->chunk = chunk; P->nchunks = cdiv(G, chunk);
    int chunk_b = d->reserved[2] > 0 ? d->reserved[2] : d->reserved[1] > 0 ? d->reserved[1] : acf_pick_chunk(G, ntiles, CTAS_PER_SM * nsm);
    if (chunk_b > G) chunk_b = G;
    if (chunk_b > CMAX) chunk_b = CMAX;
    P->chunk_b = chunk_b; P->nchunks_b = cdiv(G, chunk_b);
    Carver cv(ws);
    const size_t sp = (size_t)d->ns * g.plane;
    P->pack = cv.take<float>(7 * P->cpplane);
    P->tflags = cv.take<unsigned char>(ntiles);
    for (int b = 0; b < 2; ++b) for (int f = 0; f < 3; ++f) P->st[b][f] = cv.take<float>(sp);
    P->ill_p = cv.take<float>((size_t)P->nchunks * g.plane); P->ill_u = cv.take<float>((size_t)P->nchunks * g.plane); P->ill_w = cv.take<float>(g.plane);
    P->rcv_cnt = cv.take<int>(ntiles + 1); P->rcv_start = cv.take<int>(ntiles + 1); P->rcv_cursor = cv.take<int>(ntiles + 1);
    P->rcv_id = cv.take<int>(d->nr > 0 ? d->nr : 1); P->rcv_zx = cv.take<int>(d->nr > 0 ? d->nr : 1);
    P->rcv_nbr = cv.take<unsigned char>(ntiles);
    P->hist = P->hist_dx = P->hist_dz = P->ckpt = P->g1part = P->g2part = nullptr;
    for (int b = 0; b < 2; ++b) for (int f = 0; f < 3; ++f) P->lam[b][f] = nullptr;
    if (P->save) {
        for (int b = 0; b < 2; ++b) for (int f = 0; f < 3; ++f) P->lam[b][f] = cv.take<float>(sp);
        P->g1part = cv.take<float>((size_t)P->nchunks_b * g.plane);
        if (P->need_g2) P->g2part = cv.take<float>((size_t)P->nchunks_b * g.plane);
        if (P->nckpt) P->ckpt = cv.take<float>((size_t)P->nckpt * 3 * sp);
        P->hist = cv.take<float>((size_t)K * sp);
        if (P->need_g2) { P->hist_dx = cv.take<float>((size_t)K * sp); P->hist_dz = cv.take<float>((size_t)K * sp); }
    }
    P->bytes = cv.off;
    return ADFWI_OK;
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
    c.ew = P.pack + 6 * P.cpplane + o;
    return c;
}

RcvBuckets acf_bucket_ptrs(const FPlan& P)
{
    RcvBuckets b; b.start = P.rcv_start; b.id = P.rcv_id; b.zx = P.rcv_zx; b.nbr = P.rcv_nbr;
    return b;
}

// per-call set-up left in the workspace for the matching backward call: coefficient pack, tile
// flags, receiver buckets
int acf_setup(const FPlan& P, cudaStream_t st, const float* const* coef, const int64_t* rx, const int64_t* rz)
{
    const FGeom& g = P.g;
    acf_pack_coefs<<<dim3(cdiv(g.cpld, 128), P.cprows), 128, 0, st>>>(g.nzp, g.nxp, g.fs, P.FS, P.cprows, g.cpld, P.cpplane,
                                                                      coef[0], coef[1], coef[2], coef[3], coef[4], P.pack);
    ADFWI_LAUNCH_CHECK();
    const int ntiles = g.ntx * g.ntz;
    acf_tile_flags<<<ntiles, 128, 0, st>>>(g, P.cpplane, P.pack, P.tflags);
    ADFWI_LAUNCH_CHECK();
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

struct StepMaps { CUtensorMap st[2][3]; CUtensorMap lam[2][3]; CUtensorMap hist[3]; };

// launch with the programmatic-stream-serialization attribute (PDL): see griddep_wait() in the kernels
template <typename Kern, typename... Args>
cudaError_t acf_launch(Kern kern, int grid, int smem, cudaStream_t st, bool pdl, Args... args)
{
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid); cfg.blockDim = dim3(NTHREADS); cfg.dynamicSmemBytes = (size_t)smem; cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = pdl ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kern, args...);
}
inline bool acf_use_pdl()
{
    static int v = -1;
    if (v < 0) { const char* e = getenv("ADFWI_B200_PDL"); v = (e && e[0] == '0') ? 0 : 1; }
    return v != 0;
}

int acf_make_maps(const FPlan& P, StepMaps* M)
{
    const FGeom& g = P.g;
    for (int b = 0; b < 2; ++b)
        for (int f = 0; f < 3; ++f) {
            int rc = make_tmap_f32(&M->st[b][f], P.st[b][f], 3, g.nxp, g.ld, g.nzp, P.ns, RX, RZ);
            if (rc) return rc;
            if (P.save) { rc = make_tmap_f32(&M->lam[b][f], P.lam[b][f], 3, g.nxp, g.ld, g.nzp, P.ns, RX, RZ); if (rc) return rc; }
        }
    if (P.save) {        // history planes [ns*K][nzp][ld] for the L2 prefetch of the adjoint (boxes of one tile, no halo)
        const float* h[3] = {P.hist, P.need_g2 ? P.hist_dx : P.hist, P.need_g2 ? P.hist_dz : P.hist};
        for (int f = 0; f < 3; ++f) {
            const int rc = make_tmap_f32(&M->hist[f], h[f], 3, g.ld, g.ld, g.nzp, (uint64_t)P.ns * P.K, TX, TZ);
            if (rc) return rc;
        }
    }
    return 0;
}

inline int acf_grid(const FPlan& P, int nshots, int* nchunks, bool adjoint = false)
{
    *nchunks = cdiv(nshots, adjoint ? P.chunk_b : P.chunk);
    const int nitems = P.g.ntx * P.g.ntz * *nchunks;
    const int cap = CTAS_PER_SM * acf_num_sms();
    return nitems < cap ? nitems : cap;
}

// one fused forward step of shots [sb,se): reads buffer cur, writes buffer cur^1
int acf_forward_step(const FPlan& P, const StepMaps& M, cudaStream_t st, int cur, int sb, int se, int it, bool save, int tl,
                     const float* src_v, const int64_t* sx, const int64_t* sz,
                     float* rcv_p, float* rcv_u, float* rcv_w, bool illum, int acc_u)
{
    const FGeom& g = P.g;
    FwdArgs a;
    a.cp = acf_pack_ptrs(P); a.tflags = P.tflags;
    a.p_out = P.st[cur ^ 1][0]; a.u_out = P.st[cur ^ 1][1]; a.w_out = P.st[cur ^ 1][2];
    a.src_v = src_v; a.sx = sx; a.sz = sz;
    a.hist = P.hist; a.hist_len = P.K; a.tl = tl; a.it = it;
    a.hist_dx = P.hist_dx; a.hist_dz = P.hist_dz;
    a.nr = rcv_p ? P.nr : 0; a.rb = acf_bucket_ptrs(P);
    a.rcv_p = rcv_p; a.rcv_u = rcv_u; a.rcv_w = rcv_w;
    a.ill_p = P.ill_p; a.ill_u = P.ill_u; a.acc_u = acc_u;
    a.s_begin = sb; a.s_end = se; a.chunk = P.chunk;
    const int grid = acf_grid(P, se - sb, &a.nchunks);
    TimedLaunch tl_(KC_AC_FWD_FUSED, st);
#define LF(FSv, SVv, S2v, ILv) ADFWI_CUDA(acf_launch(ac_fwd_fused<FSv, SVv, S2v, ILv>, grid, FWD_SMEM, st, acf_use_pdl(), M.st[cur][0], M.st[cur][1], M.st[cur][2], g, a))
#define LF2(FSv, ILv) do { if (!save) LF(FSv, false, false, ILv); else if (P.need_g2) LF(FSv, true, true, ILv); else LF(FSv, true, false, ILv); } while (0)
    if (P.FS) { if (illum) LF2(true, true); else LF2(true, false); }
    else      { if (illum) LF2(false, true); else LF2(false, false); }
#undef LF2
#undef LF
    ADFWI_LAUNCH_CHECK();
    return ADFWI_OK;
}


#include "acoustic_persist.inl"

// ---- cluster-persistent small-grid path: host side ------------------------------------------------------------------
constexpr int PP_SMEM_MAX = 227 * 1024 - 1024;      // dynamic shared memory one CTA may take (1 KB left to the system)

inline bool acf_use_persist()
{
    static int v = -1;
    if (v < 0) { const char* e = getenv("ADFWI_B200_PERSIST"); v = (e && e[0] == '0') ? 0 : 1; }
    return v != 0;
}

struct PPlan { PGeom g; int kmax; size_t smem_fwd, smem_adj; };

template <bool FS, int KMAX> int pp_set_attrs()
{
    int rc = (int)cudaFuncSetAttribute(acp_fwd<FS, KMAX>, cudaFuncAttributeMaxDynamicSharedMemorySize, PP_SMEM_MAX);
    rc |= (int)cudaFuncSetAttribute(acp_adj<FS, KMAX>, cudaFuncAttributeMaxDynamicSharedMemorySize, PP_SMEM_MAX);
    return rc;
}
int pp_init_kernels()
{
    static bool done_dev[kMaxDevices] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= kMaxDevices) dev = 0;
    if (done_dev[dev]) return 0;
    int rc = 0;
#define PPK(K) rc |= pp_set_attrs<true, K>(); rc |= pp_set_attrs<false, K>();
    PPK(1) PPK(2) PPK(3) PPK(4) PPK(6) PPK(8)
#undef PPK
    if (!rc) done_dev[dev] = true;
    return rc;
}
inline int pp_round_kmax(int k) { return k <= 4 ? k : (k <= 6 ? 6 : 8); }

// co-resident clusters of `nc` CTAs with `smem` bytes each on the current device (0 = cannot launch); the driver query costs
// milliseconds, so its answers are cached per (device, cluster size, shared-memory size)
template <typename Kern> int pp_max_clusters_query(Kern kern, int nc, size_t smem);
template <typename Kern> int pp_max_clusters(Kern kern, int nc, size_t smem)
{
    struct Ent { int dev, nc; size_t smem; int n; };
    static Ent cache[64]; static int ncache = 0;
    int dev = 0;
    cudaGetDevice(&dev);
    for (int i = 0; i < ncache; ++i) if (cache[i].dev == dev && cache[i].nc == nc && cache[i].smem == smem) return cache[i].n;
    const int n = pp_max_clusters_query(kern, nc, smem);
    if (ncache < 64) { cache[ncache].dev = dev; cache[ncache].nc = nc; cache[ncache].smem = smem; cache[ncache].n = n; ++ncache; }
    return n;
}
template <typename Kern> int pp_max_clusters_query(Kern kern, int nc, size_t smem)
{
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(nc * 64); cfg.blockDim = dim3(PNT); cfg.dynamicSmemBytes = smem;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = nc; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    int n = 0;
    if (cudaOccupancyMaxActiveClusters(&n, kern, &cfg) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

// Is the problem small enough for the persistent path, and with which decomposition?  One cluster per shot, z strips of R rows.
bool pp_make_plan(const adfwi_acoustic_desc* d, const FPlan& P, PPlan* Q)
{
    if (!acf_use_persist() || (d->reserved[0] & 2)) return false;
    if (P.need_g2) return false;                              // the density gradient stays on the per-step kernels
    if (P.save && P.K < P.g.nt) return false;                 // store-all only (small grids: the history fits)
    const FGeom& f = P.g;
    const int nrows = f.nzp - f.zlo;
    const int nxg = cdiv(f.nxp, 4);
    const int pitch = 4 * nxg + 2 * PXPAD;
    if (pp_init_kernels()) return false;
    double best = 1e30; int best_nc = 0;
    static const int force_nc = getenv("ADFWI_B200_PERSIST_NC") ? atoi(getenv("ADFWI_B200_PERSIST_NC")) : 0;     // tuning / A-B switch
    static const bool debug = getenv("ADFWI_B200_DEBUG") != nullptr;
    for (int nc = 1; nc <= 8; ++nc) {
        const int R = cdiv(nrows, nc);
        if (R < 4) break;
        if (force_nc && nc != force_nc) continue;
        const size_t sm_f = pp_smem_floats(R, pitch, 3, 0) * 4, sm_a = pp_smem_floats(R, pitch, 2, 2) * 4;
        if (sm_f > (size_t)PP_SMEM_MAX || sm_a > (size_t)PP_SMEM_MAX) continue;
        const int kraw = cdiv(R * nxg, PNT);
        if (kraw > 8) continue;
        const int ncl = P.FS ? pp_max_clusters(acp_adj<true, 1>, nc, sm_a) : pp_max_clusters(acp_adj<false, 1>, nc, sm_a);
        if (ncl < 1) continue;
        // cost model: waves x (two cluster barriers + halo pulls ~ 0.6 us, ~0.2 us per float4 group a thread owns)
        const double cost = (double)cdiv(P.ns, ncl) * (0.6 + 0.2 * kraw);
        if (debug) fprintf(stderr, "adfwi_b200: persistent plan candidate NC=%d R=%d groups/thread=%d co-resident clusters=%d smem=%zu cost=%.2f\n", nc, R, kraw, ncl, sm_a, cost);
        if (cost < best) { best = cost; best_nc = nc; }
    }
    if (!best_nc) return false;
    PGeom& g = Q->g;
    g.nzp = f.nzp; g.nxp = f.nxp; g.ld = f.ld; g.fs = f.fs; g.zlo = f.zlo; g.nt = f.nt; g.nabc = f.nabc;
    g.NC = best_nc; g.R = cdiv(nrows, best_nc); g.nxg = nxg; g.pitch = pitch; g.cpld = f.cpld; g.plane = f.plane;
    g.c1 = f.c1; g.c2 = f.c2; g.dt = f.dt;
    Q->kmax = pp_round_kmax(cdiv(g.R * nxg, PNT));
    Q->smem_fwd = pp_smem_floats(g.R, pitch, 3, 0) * 4; Q->smem_adj = pp_smem_floats(g.R, pitch, 2, 2) * 4;
    return true;
}

template <typename Kern, typename Args>
cudaError_t pp_launch_k(Kern kern, int nshots, const PGeom& g, size_t smem, cudaStream_t st, const Args& a)
{
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(nshots * g.NC); cfg.blockDim = dim3(PNT); cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = g.NC; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kern, g, a);
}
#define PP_DISPATCH(KERN, FSv, kmax, ...)                                                            \
    ((kmax) == 1 ? pp_launch_k(KERN<FSv, 1>, __VA_ARGS__) : (kmax) == 2 ? pp_launch_k(KERN<FSv, 2>, __VA_ARGS__) : \
     (kmax) == 3 ? pp_launch_k(KERN<FSv, 3>, __VA_ARGS__) : (kmax) == 4 ? pp_launch_k(KERN<FSv, 4>, __VA_ARGS__) : \
     (kmax) == 6 ? pp_launch_k(KERN<FSv, 6>, __VA_ARGS__) : pp_launch_k(KERN<FSv, 8>, __VA_ARGS__))

// function attributes are per device: the >48 KB dynamic shared-memory opt-in is made once per device ordinal
int acf_init_kernels()
{
    static bool done_dev[kMaxDevices] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= kMaxDevices) dev = 0;
    bool& done = done_dev[dev];
    if (done) return 0;
    int rc = 0;
#define SF(FSv, SVv, S2v) do { rc |= acf_set_smem(ac_fwd_fused<FSv, SVv, S2v, true>, FWD_SMEM); rc |= acf_set_smem(ac_fwd_fused<FSv, SVv, S2v, false>, FWD_SMEM); } while (0)
    SF(true, true, true); SF(true, true, false); SF(true, false, false);
    SF(false, true, true); SF(false, true, false); SF(false, false, false);
#undef SF
    rc |= acf_set_smem(ac_adj_fused<true, true>, ADJ_SMEM);  rc |= acf_set_smem(ac_adj_fused<true, false>, ADJ_SMEM);
    rc |= acf_set_smem(ac_adj_fused<false, true>, ADJ_SMEM); rc |= acf_set_smem(ac_adj_fused<false, false>, ADJ_SMEM);
    if (!rc) done = true;
    return rc;
}

}  // namespace

size_t acf_workspace_bytes(const adfwi_acoustic_desc* d)
{
    FPlan P;
    acf_make_plan(d, nullptr, &P, 148);     // workspace size must not depend on the device queried
    return P.bytes;
}

int acf_group_size(const adfwi_acoustic_desc* d)
{
    FPlan P;
    acf_make_plan(d, nullptr, &P, 148);
    return P.G;
}

int acf_forward(const adfwi_acoustic_desc* d, const float* const* coef, const float* src_v, const int64_t* sx, const int64_t* sz,
                const int64_t* rx, const int64_t* rz, float* rcv_p, float* rcv_u, float* rcv_w,
                float* illum_p, float* illum_u, float* illum_w, void* ws, cudaStream_t st)
{
    FPlan P;
    acf_make_plan(d, ws, &P, 148);
    const FGeom& g = P.g;
    if (g.nzp >= 32768 || g.nxp >= 65536) return ADFWI_E_DIMS;      // unreachable: ac_use_fused() sends such grids to the generic kernels
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
        ADFWI_CUDA(cudaMemsetAsync(P.ill_p, 0, sizeof(float) * (size_t)P.nchunks * g.plane, st));
        ADFWI_CUDA(cudaMemsetAsync(P.ill_u, 0, sizeof(float) * (size_t)P.nchunks * g.plane, st));
        ADFWI_CUDA(cudaMemsetAsync(P.ill_w, 0, sizeof(float) * g.plane, st));
    }
    PPlan Q;
    if (pp_make_plan(d, P, &Q)) {
        // small grid: the whole time loop of every shot in ONE launch, state resident in the shared memory of a cluster per shot
        PFwdArgs a;
        a.cp = acf_pack_ptrs(P); a.src_v = src_v; a.sx = sx; a.sz = sz; a.hist = P.hist; a.hist_len = P.K;
        a.nr = P.nr > 0 && rcv_p ? P.nr : 0; a.rx = rx; a.rz = rz; a.rcv_p = rcv_p; a.rcv_u = rcv_u; a.rcv_w = rcv_w;
        a.ill_p = P.ill_p; a.ill_u = P.ill_u; a.ill_w = P.ill_w; a.illum = illum ? 1 : 0; a.last_chunk_start = last_chunk_start;
        a.s_begin = 0; a.save = P.save;
        {
            TimedLaunch tl_(KC_AC_FWD_PERSIST, st);
            if (P.FS) ADFWI_CUDA(PP_DISPATCH(acp_fwd, true, Q.kmax, P.ns, Q.g, Q.smem_fwd, st, a));
            else      ADFWI_CUDA(PP_DISPATCH(acp_fwd, false, Q.kmax, P.ns, Q.g, Q.smem_fwd, st, a));
        }
        ADFWI_LAUNCH_CHECK();
        if (illum) {
            const int nx = g.nxp - 2 * g.nabc, nz = g.nzp - 2 * g.nabc;
            acf_illum_finalize<<<dim3(cdiv(nx, 128), nz), 128, 0, st>>>(g.nzp, g.nxp, g.ld, g.nabc, P.nchunks, g.plane, P.ill_p, P.ill_u, P.ill_w,
                                                                         illum_p, illum_u, illum_w);
            ADFWI_LAUNCH_CHECK();
        }
        return ADFWI_OK;
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
        acf_illum_finalize<<<dim3(cdiv(nx, 128), nz), 128, 0, st>>>(g.nzp, g.nxp, g.ld, g.nabc, P.nchunks, g.plane, P.ill_p, P.ill_u, P.ill_w,
                                                                     illum_p, illum_u, illum_w);
        ADFWI_LAUNCH_CHECK();
    }
    return ADFWI_OK;
}

int acf_backward(const adfwi_acoustic_desc* d, const float* const* coef, const float* src_v, const int64_t* sx, const int64_t* sz,
                 const int64_t* rx, const int64_t* rz, const float* gp, const float* gu, const float* gw,
                 float* g_alpha1, float* g_alpha2, float* g_src, void* ws, cudaStream_t st)
{
    FPlan P;
    acf_make_plan(d, ws, &P, 148);
    const FGeom& g = P.g;
    int rc = acf_init_kernels();
    if (rc) return rc;
    StepMaps M;
    rc = acf_make_maps(P, &M);
    if (rc) return rc;
    // the coefficient pack, tile flags and receiver buckets were left in the workspace by the forward call
    const int nt = g.nt;
    ADFWI_CUDA(cudaMemsetAsync(P.g1part, 0, sizeof(float) * (size_t)P.nchunks_b * g.plane, st));
    if (P.need_g2) ADFWI_CUDA(cudaMemsetAsync(P.g2part, 0, sizeof(float) * (size_t)P.nchunks_b * g.plane, st));
    PPlan Q;
    if (pp_make_plan(d, P, &Q)) {
        PAdjArgs a;
        a.cp = acf_pack_ptrs(P); a.sx = sx; a.sz = sz; a.hist = P.hist; a.hist_len = P.K;
        a.nr = P.nr; a.rx = rx; a.rz = rz; a.gp = gp; a.gu = gu; a.gw = gw;
        a.g1part = P.g1part; a.g_src = g_src; a.s_begin = 0;
        {
            TimedLaunch tl_(KC_AC_ADJ_PERSIST, st);
            if (P.FS) ADFWI_CUDA(PP_DISPATCH(acp_adj, true, Q.kmax, P.ns, Q.g, Q.smem_adj, st, a));
            else      ADFWI_CUDA(PP_DISPATCH(acp_adj, false, Q.kmax, P.ns, Q.g, Q.smem_adj, st, a));
        }
        ADFWI_LAUNCH_CHECK();
        acf_reduce_parts<<<dim3(cdiv(g.nxp, 128), g.nzp), 128, 0, st>>>(g.nzp, g.nxp, g.ld, g.fs, P.nchunks_b, P.g1part, g_alpha1);
        ADFWI_LAUNCH_CHECK();
        return ADFWI_OK;
    }
    for (int sb = 0; sb < P.ns; sb += P.G) {
        const int se = sb + P.G < P.ns ? sb + P.G : P.ns;
        const size_t off = (size_t)sb * g.plane, cnt = (size_t)(se - sb) * g.plane * sizeof(float);
        for (int b = 0; b < 2; ++b) for (int f = 0; f < 3; ++f) ADFWI_CUDA(cudaMemsetAsync(P.lam[b][f] + off, 0, cnt, st));
        int lcur = 0;
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
                a.cp = acf_pack_ptrs(P); a.tflags = P.tflags;
                a.lp_out = P.lam[lcur ^ 1][0]; a.lu_out = P.lam[lcur ^ 1][1]; a.lw_out = P.lam[lcur ^ 1][2];
                a.sx = sx; a.sz = sz; a.hist = P.hist; a.hist_len = P.K; a.tl = it - t0; a.it = it;
                a.hist_dx = P.hist_dx; a.hist_dz = P.hist_dz;
                a.nr = P.nr; a.rb = acf_bucket_ptrs(P); a.gp = gp; a.gu = gu; a.gw = gw;
                a.g1part = P.g1part; a.g2part = P.g2part; a.g_src = g_src; a.s_begin = sb; a.s_end = se; a.chunk = P.chunk_b;
                const int grid = acf_grid(P, se - sb, &a.nchunks, true);
                {
                    TimedLaunch tl_(KC_AC_ADJ_FUSED, st);
#define LA(FSv, G2v) ADFWI_CUDA(acf_launch(ac_adj_fused<FSv, G2v>, grid, ADJ_SMEM, st, acf_use_pdl(), M.lam[lcur][0], M.lam[lcur][1], M.lam[lcur][2], M.hist[0], M.hist[1], M.hist[2], g, a))
                    if (P.FS) { if (P.need_g2) LA(true, true); else LA(true, false); }
                    else      { if (P.need_g2) LA(false, true); else LA(false, false); }
#undef LA
                }
                ADFWI_LAUNCH_CHECK();
                lcur ^= 1;
            }
        }
    }
    acf_reduce_parts<<<dim3(cdiv(g.nxp, 128), g.nzp), 128, 0, st>>>(g.nzp, g.nxp, g.ld, g.fs, P.nchunks_b, P.g1part, g_alpha1);
    ADFWI_LAUNCH_CHECK();
    if (P.need_g2) {
        acf_reduce_g2<<<dim3(cdiv(g.nxp, 128), g.nzp), 128, 0, st>>>(g.nzp, g.nxp, g.ld, g.fs, P.nchunks_b, P.g2part, coef[2], g_alpha2);
        ADFWI_LAUNCH_CHECK();
    }
    return ADFWI_OK;
}

}  // namespace adfwi
#endif  // !ADFWI_HOST_EMUL
