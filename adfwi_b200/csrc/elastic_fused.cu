// elastic_fused.cu -- TMA-staged, tile-persistent split-PML elastic time step for sm_100a (fast path).
//
// Semantics: ADFWI/propagator/elastic_kernels.py:339-418 (O(2,4)) / :495-575 (O(2,6)) and the
// reverse-mode derivative of that loop; same arithmetic and association as
// the generic kernels of elastic.cu (-fmad=false): forward records stay bit-identical to the CPU
// reference.  Used for abc_type "PML"; the sponge (ABL) variants run the generic kernels.
//
// Default: ONE launch per forward step (elf_f) and ONE per reverse step (elf_b, O(2,4)); each fuses two dependent stencil
// layers by recomputing the first layer on a ring around the tile, so the intermediate sums / cotangents never travel to
// HBM (see the banners of elf_f and elf_b).  The single-layer kernels below remain as switchable fallbacks and as the
// O(2,6) reverse step:
//   elf_s  : stress update.  Reads the 4 velocity split fields with a halo (sums and the free-surface
//            velocity rows are formed on chip), updates the 6 stress split fields IN PLACE (own cell
//            only), writes the 3 stress sums and, in recording mode, the 4 velocity derivatives the
//            adjoint needs (history; own cell, so the adjoint needs no halo and no recomputation).
//   elf_v  : velocity update.  Reads the 3 stress sums with a halo (free-surface mirrors applied on
//            chip), updates the 4 velocity split fields in place, samples the receivers, writes the 4
//            stress derivatives of the history.                      (ADFWI_B200_EL_SPLIT=1: elf_s + elf_v)
//   elf_k1 : adjoint of the velocity update: record cotangents + free-surface transpose on chip,
//            own-cell transpose on tile + ring (halo recompute), gather of the operator transposes
//            into the stress-sum cotangents; g_bx, g_bz accumulate in registers.
//   elf_k2 : adjoint of the stress update: same structure; g_C11, g_C13, g_C33, g_C55 in registers,
//            gather into the velocity-sum cotangents; g_src.        (ADFWI_B200_EL_ADJ_SPLIT=1 or O(2,6): elf_k1 + elf_k2)
// Common skeleton (the one of acoustic_fused.cu): a CTA of 256 threads owns one 64(x) x 16(z) tile (one float4 of 4
// x-cells per thread) and walks through a chunk of the launch's shots.  Coefficients, PML factors and gradient sums of
// the thread's cells stay in REGISTERS for the whole walk; per shot every field a kernel touches is brought into shared
// memory by TMA (hardware zero fill outside the grid), the loads of the next shot in flight while this one is computed;
// all shared-memory traffic is 128-bit.  Tiles whose neighbourhood has no damping run a variant without the PML factors
// (x*1 == x exactly) and, in the adjoint, stage only one half of each split pair (elf_tile_class).  (tile, chunk) items
// beyond a CTA's first are drawn from a per-launch atomic counter.  Launches are chained with programmatic dependent launch.
#include "common.cuh"
#ifndef ADFWI_HOST_EMUL
#include "tma.cuh"
#include "elastic_fused.h"
#include <stdlib.h>

namespace adfwi {

namespace {

constexpr int TX = 64, TZ = 16;             // tile interior
constexpr int HX = 4;                       // x halo of the staged rectangles (one float4 group)
constexpr int RXH = TX + 2 * HX;            // 72 floats per halo row
constexpr int NG = TX / 4;                  // float4 groups per tile row
constexpr int RPT = 1;                      // tile rows per thread
constexpr int NTH = NG * (TZ / RPT);        // threads per CTA: one float4 group x RPT rows each
constexpr int CMAX = 32;                    // max shots per chunk
constexpr int CPX = 8, CPZ = 8;             // apron of the coefficient pack (>= the 2NN ring of the fused reverse kernels, NN <= 3)
constexpr int FLX = 4, FLZ = 4;             // neighbourhood of a tile that decides its class (PML factors present or not)
constexpr int NSTAGE = 2;
constexpr int CF = TZ * TX;                 // floats of a core (no halo) rectangle
constexpr int CB = CF * 4;

template <int NN> struct Geo {
    static constexpr int RZH = TZ + 2 * NN;
    static constexpr int HF = RZH * RXH;                       // floats of a halo rectangle
    static constexpr int HB = (HF * 4 + 127) / 128 * 128;      // bytes, 128-B aligned
    static constexpr int NRING = 2 * NN * (NG + 2) + 2 * TZ;   // float4 groups of the ring around the tile
    static constexpr int S_STAGE = 4 * HB + 6 * CB;
    static constexpr int V_STAGE = 3 * HB + 4 * CB;
    static constexpr int K1_STAGE = 6 * HB;
    static constexpr int K2_STAGE = 9 * HB;
    // fused forward kernel: velocity splits staged with a halo of 2NN (the stress is recomputed on a ring of NN)
    static constexpr int HX2 = NN == 2 ? 4 : 8;
    static constexpr int RX2 = TX + 2 * HX2;
    static constexpr int RZ2 = TZ + 4 * NN;
    static constexpr int HF2 = RZ2 * RX2;
    static constexpr int HB2 = (HF2 * 4 + 127) / 128 * 128;
    static constexpr int F_VSTAGE = 4 * HB2;                   // double-buffered
    static constexpr int F_SMEM = NSTAGE * F_VSTAGE + 6 * HB + 3 * HB;   // + stress splits (single buffer) + stress sums
};
constexpr int TAIL_BYTES = 64 + 2 * CMAX * 4 + 3 * CMAX * 4 + 64;   // elf_s: mbarriers + per-shot scalars
constexpr int TAIL_SMALL = 64 + 2 * CMAX * 4;                       // other kernels: mbarriers + source cells

// plane index of field f, shot s in the workspace's plane array: f*ns + s
enum { P_VXX = 0, P_VXZ, P_VZX, P_VZZ, P_S0, P_S1, P_S2, P_S3, P_S4, P_S5, P_TXX, P_TZZ, P_TXZ, P_FWD_COUNT,
       P_B0 = P_FWD_COUNT,          // second set of the 10 split fields (ping-pong of the fused forward kernel elf_f)
       P_FWD2_COUNT = P_B0 + 10,
       P_LV = P_FWD2_COUNT,         // 2 x 4 velocity-split cotangents (ping-pong)
       P_LS = P_LV + 8,             // 2 x 6 stress-split cotangents (ping-pong)
       P_LVX = P_LS + 12, P_LVZ,    // cotangents of the velocity sums
       P_MXX, P_MZZ, P_MXZ,         // cotangents of the stress sums (between elf_k1 and elf_k2)
       P_LVX2, P_LVZ2,              // second set of the velocity-sum cotangents (ping-pong of the fused reverse kernel elf_b)
       P_COUNT };
constexpr int NHIST = 8;            // history planes per step: dxb_vx, dzb_vz, dxf_vz, dzf_vx, dxf_txx, dzb_txz, dxb_txz, dzf_tzz

struct EGeom {
    int nzp, nxp, ld, fs, nt, ns, ntx, ntz, cpld;
    int merge;               // 1: lean adjoint for damping-free tiles (tile classes 0/2, see elf_tile_class)
    size_t plane;
    float dt, dx, dz, dt_dx, dt_dz, half_dt;
    float rdx, rdz;          // RN(1/dx), RN(1/dz) for fdivs()
    float c[3];
};
struct ECoef { const float *c11, *c13, *c33, *c55, *bx, *bz, *bcx, *bcz; };   // pack planes, pre-offset to cell (0,0)
struct RcvB { const int* start; const int* id; const int* zx; const unsigned char* nbr; };
struct Walk { int s_begin, s_end, chunk, nchunks; int* counter; };

struct SArgs { ECoef cp; const unsigned char* tflags; float* planes; const float* mt; const float* src_v;
               const int64_t *sx, *sz; float* hist; int hist_len, tl, it; Walk w; };
struct VArgs { ECoef cp; const unsigned char* tflags; float* planes; float* hist; int hist_len, tl, it;
               int nr; RcvB rb; float* rcv[5]; Walk w; };
struct K1Args { ECoef cp; const unsigned char* tflags; float* planes; const float* hist; int hist_len, tl, it, lcur;
                int nr; RcvB rb; const float* g[5]; float* gpart; Walk w; };
struct K2Args { ECoef cp; const unsigned char* tflags; float* planes; const float* hist; int hist_len, tl, it, lcur;
                const float* mt; const int64_t *sx, *sz; float* g_src; float* gpart; Walk w; };

// ---- small device helpers ------------------------------------------------------------------------
__device__ __forceinline__ bool mbar_try(uint64_t* bar, uint32_t parity)
{
    uint32_t ok;
    asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
                 : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ float4 ld4(const float* p) { return *reinterpret_cast<const float4*>(p); }
__device__ __forceinline__ void st4(float* p, const float4& v) { *reinterpret_cast<float4*>(p) = v; }
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
__device__ __forceinline__ void red4(float* p, const float4& v)
{
    asm volatile("red.global.add.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
// TMA prefetch of a box into L2 (no shared-memory destination): used for the history planes, which the adjoint
// kernels then read with plain 128-bit loads -- an L2 hit instead of a DRAM round trip on the critical path
__device__ __forceinline__ void tma_prefetch_3d(const CUtensorMap* tm, int c0, int c1, int c2)
{
    asm volatile("cp.async.bulk.prefetch.tensor.3d.L2.global.tile [%0, {%1, %2, %3}];" ::"l"(tm), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
__device__ __forceinline__ void griddep_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void griddep_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

#define F4OP(name, expr)                                                                        \
    __device__ __forceinline__ float4 name(const float4& a, const float4& b)                    \
    { float4 r; { const float x = a.x, y = b.x; r.x = (expr); } { const float x = a.y, y = b.y; r.y = (expr); } \
      { const float x = a.z, y = b.z; r.z = (expr); } { const float x = a.w, y = b.w; r.w = (expr); } return r; }
F4OP(add4, x + y)
F4OP(sub4, x - y)
F4OP(mul4, x * y)
F4OP(div4, x / y)
#undef F4OP
__device__ __forceinline__ float4 muls(const float4& a, float s) { return make_float4(a.x * s, a.y * s, a.z * s, a.w * s); }
__device__ __forceinline__ float4 smul(float s, const float4& a) { return make_float4(s * a.x, s * a.y, s * a.z, s * a.w); }
// Correctly rounded a / b from a precomputed correctly rounded reciprocal rb = RN(1/b): one multiply and two
// Newton corrections with exact FMA remainders.  After the first correction q is a faithful rounding of a/b,
// so the second one returns RN(a/b) (Markstein's theorem) -- the same bits as the IEEE division of the eager
// reference -- at 5 instructions and with no slow path for zero numerators (the compiler's division sequence
// takes its out-of-line path for every zero operand, i.e. for every cell the wave has not reached yet).
// The remainders are exact only away from the underflow range: non-zero numerators below 2^-100 in magnitude
// take the scaled sequence fdiv1_tiny (DivGuard).  Divisors are grid spacings and 1 + dt/2*profile (moderate magnitudes).
__device__ __forceinline__ float fdiv1(float a, float b, float rb)
{
    float q = a * rb;
    q = __fmaf_rn(__fmaf_rn(-b, q, a), rb, q);
    q = __fmaf_rn(__fmaf_rn(-b, q, a), rb, q);
    return q;
}
// guard accumulator: ok() is true when every numerator added is zero or at least 2^-100 in magnitude
// (t = 2*bits - 1 wraps zero to the top of the unsigned range; one IADD3 + one unsigned min per value)
struct DivGuard {
    unsigned mn;
    __device__ __forceinline__ DivGuard() : mn(0xffffffffu) {}
    __device__ __forceinline__ void add(const float4& a)
    {
        const unsigned t0 = 2u * __float_as_uint(a.x) - 1u, t1 = 2u * __float_as_uint(a.y) - 1u;
        const unsigned t2 = 2u * __float_as_uint(a.z) - 1u, t3 = 2u * __float_as_uint(a.w) - 1u;
        mn = min(min(mn, min(t0, t1)), min(t2, t3));
    }
    __device__ __forceinline__ bool ok() const { return mn >= ((27u << 24) - 1u); }
};
// RN(a/b) for |a| < 2^-100 (zero and denormals included), b > 0 of moderate magnitude, rb = RN(1/b) -- the numerators of the band
// ahead of every wavefront, where the field decays through the underflow range.  The compiler's IEEE division takes its slowest
// out-of-line path for exactly these operands (ncu: 11 % of all executed instructions of ela_f on an expanding wavefield).  Here:
// scale the numerator by 2^80 (exact), take the Markstein quotient q = RN(2^80 a / b) with its exact remainder r, scale back.
// t = RN(q 2^-80) is exact when the result is normal; when it is denormal it is the correct rounding of a/b unless q sits exactly
// on a rounding boundary of the denormal grid (|q - t 2^80| = 2^-70) while a/b does not (r != 0): then the true quotient lies on
// the side of r, and the hardware's ties-to-even choice is moved by one denormal step when it went the other way.  Checked
// bit for bit against the IEEE division on 3.2e8 operand pairs (ties, denormal numerators, denormal quotients).
__device__ __forceinline__ float fdiv1_tiny(float a, float b, float rb)
{
    const float a2 = a * 0x1p80f;
    float q = a2 * rb;
    q = __fmaf_rn(__fmaf_rn(-b, q, a2), rb, q);
    q = __fmaf_rn(__fmaf_rn(-b, q, a2), rb, q);
    const float r = __fmaf_rn(-b, q, a2);
    float t = q * 0x1p-80f;
    const float e = __fmaf_rn(-t, 0x1p80f, q);
    if (fabsf(e) == 0x1p-70f && r != 0.f && ((r > 0.f) == (e > 0.f))) t += copysignf(0x1p-149f, e);
    return t;
}
__device__ __forceinline__ float fdiv1_any(float a, float b, float rb) { return fabsf(a) < 0x1p-100f ? fdiv1_tiny(a, b, rb) : fdiv1(a, b, rb); }
// the guarded path of the forward kernels, kept out of line (code size, registers): correctly rounded for every numerator
__device__ __noinline__ float4 safe_div4(float4 a, float4 b, float4 rb)
{
    return make_float4(fdiv1_any(a.x, b.x, rb.x), fdiv1_any(a.y, b.y, rb.y), fdiv1_any(a.z, b.z, rb.z), fdiv1_any(a.w, b.w, rb.w));
}
__device__ __forceinline__ float4 safe_divs(const float4& a, float b, float rb) { return safe_div4(a, make_float4(b, b, b, b), make_float4(rb, rb, rb, rb)); }
// unguarded fast divisions: the caller adds the numerator to a DivGuard and redoes the work with safe_div4 when !ok()
__device__ __forceinline__ float4 fdiv4(const float4& a, const float4& b, const float4& rb)
{
    return make_float4(fdiv1(a.x, b.x, rb.x), fdiv1(a.y, b.y, rb.y), fdiv1(a.z, b.z, rb.z), fdiv1(a.w, b.w, rb.w));
}
__device__ __forceinline__ float4 fdivs(const float4& a, float b, float rb)
{
    return make_float4(fdiv1(a.x, b, rb), fdiv1(a.y, b, rb), fdiv1(a.z, b, rb), fdiv1(a.w, b, rb));
}
__device__ __forceinline__ float4 zero4() { return make_float4(0.f, 0.f, 0.f, 0.f); }
__device__ __forceinline__ void opaque4(float4& v) { asm volatile("" : "+f"(v.x), "+f"(v.y), "+f"(v.z), "+f"(v.w)); }
// correctly rounded reciprocals: the same bits as 1.0f / a under -prec-div=true
__device__ __forceinline__ float4 rcp4(const float4& a) { return make_float4(__frcp_rn(a.x), __frcp_rn(a.y), __frcp_rn(a.z), __frcp_rn(a.w)); }
__device__ __forceinline__ float4 one4() { return make_float4(1.f, 1.f, 1.f, 1.f); }
// per-component select: bit x of m set -> a, else b
__device__ __forceinline__ float4 sel4(unsigned m, const float4& a, const float4& b)
{ return make_float4((m & 1u) ? a.x : b.x, (m & 2u) ? a.y : b.y, (m & 4u) ? a.z : b.z, (m & 8u) ? a.w : b.w); }
__device__ __forceinline__ float comp4(const float4& a, int i) { return i == 0 ? a.x : i == 1 ? a.y : i == 2 ? a.z : a.w; }
__device__ __forceinline__ void addc4(float4& a, int i, float v)
{ if (i == 0) a.x = a.x + v; else if (i == 1) a.y = a.y + v; else if (i == 2) a.z = a.z + v; else a.w = a.w + v; }

// 12-float row segment [c0-4, c0+8) around the thread's float4 (p points at the float4)
__device__ __forceinline__ void ldseg(const float* p, float* s)
{
    const float4 a = ld4(p - 4), b = ld4(p), c = ld4(p + 4);
    s[0] = a.x; s[1] = a.y; s[2] = a.z; s[3] = a.w; s[4] = b.x; s[5] = b.y; s[6] = b.z; s[7] = b.w; s[8] = c.x; s[9] = c.y; s[10] = c.z; s[11] = c.w;
}
__device__ __forceinline__ void ldseg2(const float* p, const float* q, float* s)     // sum of two split fields (p + q, in that order)
{
    float a[12], b[12];
    ldseg(p, a); ldseg(q, b);
#pragma unroll
    for (int i = 0; i < 12; ++i) s[i] = a[i] + b[i];
}
// x operators on a segment, cell x of the float4 at seg[4+x].
//   OFF = 0: backward operator D-x (elastic_kernels.py:86-96): c[k]*(a[k] - a[-k-1])
//   OFF = 1: forward operator  D+x (:66-76):                   c[k]*(a[k+1] - a[-k])
template <int NN, int OFF> __device__ __forceinline__ float4 xdiff(const float* s, const float* c)
{
    float r[4];
#pragma unroll
    for (int x = 0; x < 4; ++x) {
        float v = c[0] * (s[4 + x + OFF] - s[4 + x - 1 + OFF]);
#pragma unroll
        for (int k = 1; k < NN; ++k) v = v + c[k] * (s[4 + x + k + OFF] - s[4 + x - k - 1 + OFF]);
        r[x] = v;
    }
    return make_float4(r[0], r[1], r[2], r[3]);
}
// gathers of the operator transposes (m is zero outside the update region):
//   OFF = 0: (D+x)^T : c[k]*(m[j-k-1] - m[j+k]);   OFF = 1: (D-x)^T : c[k]*(m[j-k] - m[j+k+1])
template <int NN, int OFF> __device__ __forceinline__ float4 xgath(const float* s, const float* c)
{
    float r[4];
#pragma unroll
    for (int x = 0; x < 4; ++x) {
        float v = 0.f;
#pragma unroll
        for (int k = 0; k < NN; ++k) v += c[k] * (s[4 + x - k - 1 + OFF] - s[4 + x + k + OFF]);
        r[x] = v;
    }
    return make_float4(r[0], r[1], r[2], r[3]);
}
// z operators on a window of 2NN rows w[0..2NN): D-z with w[q] = row i-NN+q, D+z with w[q] = row i-NN+1+q
template <int NN> __device__ __forceinline__ float4 zdiff(const float4* w, const float* c)
{
    float4 v = smul(c[0], sub4(w[NN], w[NN - 1]));
#pragma unroll
    for (int k = 1; k < NN; ++k) v = add4(v, smul(c[k], sub4(w[NN + k], w[NN - 1 - k])));
    return v;
}
// (D+z)^T with w[q] = row i-NN+q;  (D-z)^T with w[q] = row i-NN+1+q
template <int NN> __device__ __forceinline__ float4 zgath(const float4* w, const float* c)
{
    float4 v = zero4();
#pragma unroll
    for (int k = 0; k < NN; ++k) v = add4(v, smul(c[k], sub4(w[NN - 1 - k], w[NN + k])));
    return v;
}

// (tile, shot) sequence of one persistent CTA.  Items = (tile, chunk of the launch's shots); the first item of a CTA
// is blockIdx.x, the following ones are drawn from the launch's atomic counter (dynamic balancing: PML tiles and
// ragged chunks cost differently).  Only thread 0 -- the TMA producer, running NSTAGE shots ahead of the compute --
// keeps a cursor; the ids it draws are handed to the consumers through a 4-entry ring in shared memory.
struct Cursor {
    int item, tile, s, s_hi, X0, Z0; bool valid; unsigned wr;
    __device__ __forceinline__ void set(int it, const EGeom& g, const Walk& w)
    {
        item = it; valid = it < g.ntx * g.ntz * w.nchunks;
        if (valid) {
            tile = it / w.nchunks;
            const int ch = it - tile * w.nchunks;
            const int tzi = tile / g.ntx, txi = tile - tzi * g.ntx;
            X0 = txi * TX; Z0 = tzi * TZ;
            s = w.s_begin + ch * w.chunk; s_hi = min(s + w.chunk, w.s_end);
        }
    }
    __device__ __forceinline__ void next(const EGeom& g, const Walk& w, int* ring)
    {
        if (valid && ++s >= s_hi) {
            const int it = (int)gridDim.x + atomicAdd(w.counter, 1);
            ring[wr++ & 3u] = it;
            set(it, g, w);
        }
    }
    // a second producer cursor that trails the drawing one reads the ids from the ring (wr is its read index)
    __device__ __forceinline__ void next_follow(const EGeom& g, const Walk& w, const int* ring)
    {
        if (valid && ++s >= s_hi) set(ring[wr++ & 3u], g, w);
    }
};

static_assert(sizeof(Cursor) == 32, "two cursors fit the last 64 bytes of TAIL_BYTES (elf_f keeps them in shared memory)");

// thread roles: float4 group l, rows 2q and 2q+1 of the tile
struct Roles {
    int q, l, r0, c0;
    __device__ __forceinline__ explicit Roles(int tid) { q = tid / NG; l = tid - q * NG; r0 = RPT * q; c0 = 4 * l; }
};
// ring float4 group i in [0, NRING): rows [-NN,0) and [TZ,TZ+NN) x groups [-1,NG], rows [0,TZ) x groups {-1,NG}
template <int NN> __device__ __forceinline__ void ring_cell(int i, int& r, int& gi)
{
    constexpr int W = NG + 2;
    if (i < NN * W) { r = -NN + i / W; gi = i % W - 1; }
    else if (i < 2 * NN * W) { const int ii = i - NN * W; r = TZ + ii / W; gi = ii % W - 1; }
    else { const int ii = i - 2 * NN * W; r = ii >> 1; gi = (ii & 1) ? NG : -1; }
}
// update-region masks (elastic_kernels.py:303-310): rows/cols [NN, n-NN)
template <int NN> __device__ __forceinline__ unsigned col_mask(int gx, int nxp)
{
    unsigned m = 0;
#pragma unroll
    for (int x = 0; x < 4; ++x) if (gx + x >= NN && gx + x < nxp - NN) m |= 1u << x;
    return m;
}
template <int NN> __device__ __forceinline__ bool row_in(int gz, int nzp) { return gz >= NN && gz < nzp - NN; }

#define ELF_SEQ() asm volatile("" ::: "memory")     // compiler-level ordering point (no instruction)
#define ELF_WAIT_STAGE(k) do { while (!mbar_try(bar + (k), (par >> (k)) & 1u)) {} par ^= 1u << (k); } while (0)

// ==========================================================================================
// elf_s : stress update (:341-384 / :497-540)
// ==========================================================================================
template <int NN> __device__ __forceinline__ void s_issue(const Cursor& c, unsigned char* smem, uint64_t* bar, int k,
                                                          const CUtensorMap* th, const CUtensorMap* tc, int ns)
{
    using G = Geo<NN>;
    unsigned char* st = smem + k * G::S_STAGE;
    fence_proxy_async();
    mbar_expect_tx(bar + k, 4 * G::HF * 4 + 6 * CB);
#pragma unroll
    for (int f = 0; f < 4; ++f) tma_load_3d(st + f * G::HB, th, c.X0 - HX, c.Z0 - NN, (P_VXX + f) * ns + c.s, bar + k);
#pragma unroll
    for (int f = 0; f < 6; ++f) tma_load_3d(st + 4 * G::HB + f * CB, tc, c.X0, c.Z0, (P_S0 + f) * ns + c.s, bar + k);
}

template <int NN, bool PML, bool FS, bool SAVE>
__device__ __forceinline__ void s_tile(const CUtensorMap* th, const CUtensorMap* tc, const EGeom& g, const SArgs& a,
                                       unsigned char* smem, uint64_t* bar, uint32_t& par, int& stage, Cursor& pc, int* ring,
                                       int* s_sz, int* s_sx, float* s_sxx, float* s_szz, float* s_sxz,
                                       const Roles& R, int tid, int tile, int s_lo, int s_hi, bool first)
{
    using G = Geo<NN>;
    const uint64_t pol = l2_keep_policy();
    const int tzi = tile / g.ntx, txi = tile - tzi * g.ntx;
    const int X0 = txi * TX, Z0 = tzi * TZ;
    const int gx = X0 + R.c0, gz0 = Z0 + R.r0;
    const unsigned cm = col_mask<NN>(gx, g.nxp);
    const bool col_ok = gx < g.ld;
    const bool lean = !PML && g.merge;     // merged history planes for damping-free tiles (see k1_tile / k2_tile)
    if (tid < s_hi - s_lo) {         // per-shot scalars (:341-346: -(M/2)*src)
        const int s = s_lo + tid;
        const float* M = a.mt + (size_t)s * 9;
        const float v = a.src_v[(size_t)s * g.nt + a.it];
        s_sz[tid] = (int)a.sz[s]; s_sx[tid] = (int)a.sx[s];
        s_sxx[tid] = (-(M[0] / 2.0f)) * v; s_szz[tid] = (-(M[8] / 2.0f)) * v; s_sxz[tid] = (-(M[2] / 2.0f)) * v;
    }
    float4 C11[RPT], C13[RPT], C33[RPT], C55[RPT], PXN[RPT], PXI[RPT], PZN[RPT], PZI[RPT];
#pragma unroll
    for (int j = 0; j < RPT; ++j) {
        const ptrdiff_t o = (ptrdiff_t)(gz0 + j) * g.cpld + gx;
        C11[j] = ldk4(a.cp.c11 + o, pol); C13[j] = ldk4(a.cp.c13 + o, pol);
        C33[j] = ldk4(a.cp.c33 + o, pol); C55[j] = ldk4(a.cp.c55 + o, pol);
        if (PML) {
            const float4 bx_ = ldk4(a.cp.bcx + o, pol), bz_ = ldk4(a.cp.bcz + o, pol);
            const float4 pxd = add4(one4(), smul(g.half_dt, bx_)), pzd = add4(one4(), smul(g.half_dt, bz_));
            PXN[j] = sub4(one4(), smul(g.half_dt, bx_)); PZN[j] = sub4(one4(), smul(g.half_dt, bz_));
            PXI[j] = div4(one4(), pxd); PZI[j] = div4(one4(), pzd);
        }
    }
    if (first) {
        griddep_wait();
#pragma unroll
        for (int k = 0; k < NSTAGE; ++k)
            if (tid == 0 && pc.valid) { s_issue<NN>(pc, smem, bar, k, th, tc, g.ns); pc.next(g, a.w, ring); }
    }
    __syncthreads();

    for (int s = s_lo; s < s_hi; ++s) {
        const int k = stage;
        float* hv = (float*)(smem + k * G::S_STAGE);                    // 4 halo rects: vxx, vxz, vzx, vzz
        const float* cs = (const float*)(smem + k * G::S_STAGE + 4 * G::HB);   // 6 core rects
        float* vxx = hv; float* vxz = hv + G::HB / 4; float* vzx = hv + 2 * (G::HB / 4); float* vzz = hv + 3 * (G::HB / 4);
        const int szs = s_sz[s - s_lo], sxs = s_sx[s - s_lo];
        ELF_WAIT_STAGE(k);
        if (FS && tzi == 0) {
            // free-surface velocity rows (:399-402), formed from the sums of rows h-1, h of the previous step:
            // vz[h-2] = vz[h-3] = vz[h-1]; vx[h-2] = vz[h-2,j+1] - vz[h-2,j] + vz[h-1,j+1] - vz[h-1,j] + vx[h,j]
            // (rows h-2, h-3 are outside the update region: their split fields are zero, so the edit
            //  is parked in the first split of each pair)
            const int t = tid;
            if (t >= 1 && t < RXH - 1) {
                const int j = X0 + t - HX;
                if (j >= NN && j < g.nxp - NN) {
                    const int rh1 = (2 * NN) * RXH + t, rh = (2 * NN + 1) * RXH + t, rh2 = (2 * NN - 1) * RXH + t, rh3 = (2 * NN - 2) * RXH + t;
                    const float vz1 = vzx[rh1] + vzz[rh1];
                    const float vz1n = (j + 1 < g.nxp - NN) ? vzx[rh1 + 1] + vzz[rh1 + 1] : 0.f;
                    const float vxh = vxx[rh] + vxz[rh];
                    const float nvx = (((vz1n - vz1) + vz1n) - vz1) + vxh;
                    vzx[rh2] = vz1; vzz[rh2] = 0.f; vzx[rh3] = vz1; vzz[rh3] = 0.f;
                    vxx[rh2] = nvx; vxz[rh2] = 0.f;
                }
            }
            __syncthreads();
        }
        const bool has_src = (szs >= Z0) && (szs < Z0 + TZ) && (sxs >= X0) && (sxs < X0 + TX);
#pragma unroll
        for (int j = 0; j < RPT; ++j) {
            const int r = R.r0 + j, gz = gz0 + j;
            const int hb = (r + NN) * RXH + R.c0 + HX;
            float sx_[12], sz_[12];
            ldseg2(vxx + hb, vxz + hb, sx_);
            ldseg2(vzx + hb, vzz + hb, sz_);
            float4 wzb[2 * NN], wzf[2 * NN];
#pragma unroll
            for (int q = 0; q < 2 * NN; ++q) {
                wzb[q] = add4(ld4(vzx + hb + (q - NN) * RXH), ld4(vzz + hb + (q - NN) * RXH));           // vz rows r-NN .. r+NN-1
                wzf[q] = add4(ld4(vxx + hb + (q - NN + 1) * RXH), ld4(vxz + hb + (q - NN + 1) * RXH));   // vx rows r-NN+1 .. r+NN
            }
            const float4 dxb_vx = xdiff<NN, 0>(sx_, g.c), dxf_vz = xdiff<NN, 1>(sz_, g.c);
            const float4 dzb_vz = zdiff<NN>(wzb, g.c), dzf_vx = zdiff<NN>(wzf, g.c);
            const int co = r * TX + R.c0;
            float4 a0 = ld4(cs + co), a1 = ld4(cs + CF + co), a2 = ld4(cs + 2 * CF + co);
            float4 a3 = ld4(cs + 3 * CF + co), a4 = ld4(cs + 4 * CF + co), a5 = ld4(cs + 5 * CF + co);
            const unsigned m = row_in<NN>(gz, g.nzp) ? cm : 0u;
            float4 n0, n1, n2, n3, n4, n5;
            if (PML) {
                n0 = mul4(add4(mul4(PXN[j], a0), smul(g.dt_dx, mul4(C11[j], dxb_vx))), PXI[j]);
                n1 = mul4(add4(mul4(PZN[j], a1), smul(g.dt_dz, mul4(C13[j], dzb_vz))), PZI[j]);
                n2 = mul4(add4(mul4(PXN[j], a2), smul(g.dt_dx, mul4(C13[j], dxb_vx))), PXI[j]);
                n3 = mul4(add4(mul4(PZN[j], a3), smul(g.dt_dz, mul4(C33[j], dzb_vz))), PZI[j]);
                n4 = mul4(add4(mul4(PXN[j], a4), smul(g.dt_dx, mul4(C55[j], dxf_vz))), PXI[j]);
                n5 = mul4(add4(mul4(PZN[j], a5), smul(g.dt_dz, mul4(C55[j], dzf_vx))), PZI[j]);
            } else {
                n0 = add4(a0, smul(g.dt_dx, mul4(C11[j], dxb_vx)));
                n1 = add4(a1, smul(g.dt_dz, mul4(C13[j], dzb_vz)));
                n2 = add4(a2, smul(g.dt_dx, mul4(C13[j], dxb_vx)));
                n3 = add4(a3, smul(g.dt_dz, mul4(C33[j], dzb_vz)));
                n4 = add4(a4, smul(g.dt_dx, mul4(C55[j], dxf_vz)));
                n5 = add4(a5, smul(g.dt_dz, mul4(C55[j], dzf_vx)));
            }
            a0 = sel4(m, n0, a0); a1 = sel4(m, n1, a1); a2 = sel4(m, n2, a2);
            a3 = sel4(m, n3, a3); a4 = sel4(m, n4, a4); a5 = sel4(m, n5, a5);
            if (has_src && szs == gz && !(FS && gz < NN)) {
                const int dc = sxs - gx;
                if (dc >= 0 && dc < 4) {
                    const float sxx = s_sxx[s - s_lo], szz = s_szz[s - s_lo], sxz = s_sxz[s - s_lo];
                    addc4(a0, dc, sxx); addc4(a1, dc, sxx); addc4(a2, dc, szz); addc4(a3, dc, szz); addc4(a4, dc, sxz); addc4(a5, dc, sxz);
                }
            }
            if (col_ok && gz < g.nzp) {
                const size_t o = (size_t)gz * g.ld + gx;
                float* P = a.planes + (size_t)s * g.plane + o;
                const size_t fp = (size_t)g.ns * g.plane;
                st4(P + (P_S0 + 0) * fp, a0); st4(P + (P_S0 + 1) * fp, a1); st4(P + (P_S0 + 2) * fp, a2);
                st4(P + (P_S0 + 3) * fp, a3); st4(P + (P_S0 + 4) * fp, a4); st4(P + (P_S0 + 5) * fp, a5);
                st4(P + P_TXX * fp, add4(a0, a1)); st4(P + P_TZZ * fp, add4(a2, a3)); st4(P + P_TXZ * fp, add4(a4, a5));
                if (SAVE) {
                    float* H = a.hist + ((size_t)s * a.hist_len + a.tl) * NHIST * g.plane + o;
                    __stcs(reinterpret_cast<float4*>(H), sel4(m, dxb_vx, zero4()));
                    __stcs(reinterpret_cast<float4*>(H + g.plane), sel4(m, dzb_vz, zero4()));
                    if (lean) __stcs(reinterpret_cast<float4*>(H + 2 * g.plane), sel4(m, add4(dxf_vz, dzf_vx), zero4()));
                    else {
                        __stcs(reinterpret_cast<float4*>(H + 2 * g.plane), sel4(m, dxf_vz, zero4()));
                        __stcs(reinterpret_cast<float4*>(H + 3 * g.plane), sel4(m, dzf_vx, zero4()));
                    }
                }
            }
        }
        if (FS && tzi == 0) fence_proxy_async();
        __syncthreads();
        if (tid == 0 && pc.valid) { s_issue<NN>(pc, smem, bar, k, th, tc, g.ns); pc.next(g, a.w, ring); }
        stage = (stage + 1 == NSTAGE) ? 0 : stage + 1;
    }
}

template <int NN, bool FS, bool SAVE>
__global__ void __launch_bounds__(NTH, 2)
elf_s(const __grid_constant__ CUtensorMap th, const __grid_constant__ CUtensorMap tc, const EGeom g, const SArgs a)
{
    using G = Geo<NN>;
    extern __shared__ __align__(128) unsigned char smem[];
    uint64_t* bar = (uint64_t*)(smem + NSTAGE * G::S_STAGE);
    int* s_sz = (int*)(bar + 8);
    int* s_sx = s_sz + CMAX;
    float* s_sxx = (float*)(s_sx + CMAX);
    float* s_szz = s_sxx + CMAX;
    float* s_sxz = s_szz + CMAX;
    const int tid = threadIdx.x;
    if (tid == 0) { for (int k = 0; k < NSTAGE; ++k) mbar_init(bar + k, 1); }
    __syncthreads();
    const Roles R(tid);
    uint32_t par = 0;
    int stage = 0;
    const int nitems = g.ntx * g.ntz * a.w.nchunks;
    int* ring = (int*)(bar + 4);
    Cursor pc;
    pc.wr = 0;
    pc.set(blockIdx.x, g, a.w);
    griddep_launch_dependents();
    unsigned rd = 0;
    bool first = true;
    for (int item = blockIdx.x; item < nitems; item = ring[rd++ & 3u], first = false) {
        const int tile = item / a.w.nchunks, chunk = item - tile * a.w.nchunks;
        const int s_lo = a.w.s_begin + chunk * a.w.chunk;
        const int s_hi = min(s_lo + a.w.chunk, a.w.s_end);
        if (a.tflags[tile] == 1) s_tile<NN, true, FS, SAVE>(&th, &tc, g, a, smem, bar, par, stage, pc, ring, s_sz, s_sx, s_sxx, s_szz, s_sxz, R, tid, tile, s_lo, s_hi, first);
        else                s_tile<NN, false, FS, SAVE>(&th, &tc, g, a, smem, bar, par, stage, pc, ring, s_sz, s_sx, s_sxx, s_szz, s_sxz, R, tid, tile, s_lo, s_hi, first);
        __syncthreads();
    }
}

// ==========================================================================================
// elf_v : velocity update + receivers (:387-409 / :543-565)
// ==========================================================================================
template <int NN> __device__ __forceinline__ void v_issue(const Cursor& c, unsigned char* smem, uint64_t* bar, int k,
                                                          const CUtensorMap* th, const CUtensorMap* tc, int ns)
{
    using G = Geo<NN>;
    unsigned char* st = smem + k * G::V_STAGE;
    fence_proxy_async();
    mbar_expect_tx(bar + k, 3 * G::HF * 4 + 4 * CB);
#pragma unroll
    for (int f = 0; f < 3; ++f) tma_load_3d(st + f * G::HB, th, c.X0 - HX, c.Z0 - NN, (P_TXX + f) * ns + c.s, bar + k);
#pragma unroll
    for (int f = 0; f < 4; ++f) tma_load_3d(st + 3 * G::HB + f * CB, tc, c.X0, c.Z0, (P_VXX + f) * ns + c.s, bar + k);
}

template <int NN, bool PML, bool FS, bool SAVE>
__device__ __forceinline__ void v_tile(const CUtensorMap* th, const CUtensorMap* tc, const EGeom& g, const VArgs& a,
                                       unsigned char* smem, uint64_t* bar, uint32_t& par, int& stage, Cursor& pc, int* ring,
                                       const Roles& R, int tid, int tile, int s_lo, int s_hi, bool first)
{
    using G = Geo<NN>;
    const uint64_t pol = l2_keep_policy();
    const int tzi = tile / g.ntx, txi = tile - tzi * g.ntx;
    const int X0 = txi * TX, Z0 = tzi * TZ;
    const int gx = X0 + R.c0, gz0 = Z0 + R.r0;
    const unsigned cm = col_mask<NN>(gx, g.nxp);
    const bool col_ok = gx < g.ld;
    const bool lean = !PML && g.merge;     // merged history planes for damping-free tiles (see k1_tile / k2_tile)
    float4 DBX[RPT], DBZ[RPT], PXN[RPT], PXD[RPT], PZN[RPT], PZD[RPT], RPXD[RPT], RPZD[RPT];
#pragma unroll
    for (int j = 0; j < RPT; ++j) {
        const ptrdiff_t o = (ptrdiff_t)(gz0 + j) * g.cpld + gx;
        DBX[j] = smul(g.dt, ldk4(a.cp.bx + o, pol)); DBZ[j] = smul(g.dt, ldk4(a.cp.bz + o, pol));     // dt*bx, dt*bz (:387-394)
        if (PML) {
            const float4 bx_ = ldk4(a.cp.bcx + o, pol), bz_ = ldk4(a.cp.bcz + o, pol);
            PXD[j] = add4(one4(), smul(g.half_dt, bx_)); PZD[j] = add4(one4(), smul(g.half_dt, bz_));
            PXN[j] = sub4(one4(), smul(g.half_dt, bx_)); PZN[j] = sub4(one4(), smul(g.half_dt, bz_));
            RPXD[j] = div4(one4(), PXD[j]); RPZD[j] = div4(one4(), PZD[j]);
        }
    }
    const int rcv_lo = a.nr > 0 ? a.rb.start[tile] : 0, rcv_hi = a.nr > 0 ? a.rb.start[tile + 1] : 0;
    const bool has_rcv = rcv_hi > rcv_lo;
    if (first) {
        griddep_wait();
#pragma unroll
        for (int k = 0; k < NSTAGE; ++k)
            if (tid == 0 && pc.valid) { v_issue<NN>(pc, smem, bar, k, th, tc, g.ns); pc.next(g, a.w, ring); }
    }

    for (int s = s_lo; s < s_hi; ++s) {
        const int k = stage;
        float* ht = (float*)(smem + k * G::V_STAGE);                     // 3 halo rects: txx, tzz, txz
        float* cv = (float*)(smem + k * G::V_STAGE + 3 * G::HB);         // 4 core rects: vxx, vxz, vzx, vzz
        float* txx = ht; float* tzz = ht + G::HB / 4; float* txz = ht + 2 * (G::HB / 4);
        ELF_WAIT_STAGE(k);
        if (FS && tzi == 0) {          // free-surface stress rows (:380-384) on the staged sums
            const int t = tid;
            if (t < RXH) {
                const int j = X0 + t - HX;
                if (j >= NN && j < g.nxp - NN) {
                    const int rh1 = (2 * NN) * RXH + t, rh = (2 * NN + 1) * RXH + t, rh2 = (2 * NN - 1) * RXH + t, rh3 = (2 * NN - 2) * RXH + t;
                    tzz[rh1] = 0.f;
                    txz[rh2] = -txz[rh1];
                    tzz[rh2] = -tzz[rh];
                    txz[rh3] = -txz[rh];
                }
            }
            __syncthreads();
        }
        float4 nvx[RPT], nvz[RPT];
#pragma unroll
        for (int j = 0; j < RPT; ++j) {
            const int r = R.r0 + j, gz = gz0 + j;
            const int hb = (r + NN) * RXH + R.c0 + HX;
            float sxx_[12], sxz_[12];
            ldseg(txx + hb, sxx_);
            ldseg(txz + hb, sxz_);
            float4 wzb[2 * NN], wzf[2 * NN];
#pragma unroll
            for (int q = 0; q < 2 * NN; ++q) {
                wzb[q] = ld4(txz + hb + (q - NN) * RXH);          // txz rows r-NN .. r+NN-1
                wzf[q] = ld4(tzz + hb + (q - NN + 1) * RXH);      // tzz rows r-NN+1 .. r+NN
            }
            const float4 dxf_txx = xdiff<NN, 1>(sxx_, g.c), dxb_txz = xdiff<NN, 0>(sxz_, g.c);
            const float4 dzb_txz = zdiff<NN>(wzb, g.c), dzf_tzz = zdiff<NN>(wzf, g.c);
            const int co = r * TX + R.c0;
            float4 q0 = ld4(cv + co), q1 = ld4(cv + CF + co), q2 = ld4(cv + 2 * CF + co), q3 = ld4(cv + 3 * CF + co);
            const unsigned m = row_in<NN>(gz, g.nzp) ? cm : 0u;
            float4 n0, n1, n2, n3;
            {
                const float4 A0 = mul4(DBX[j], dxf_txx), A1 = mul4(DBX[j], dzb_txz), A2 = mul4(DBZ[j], dxb_txz), A3 = mul4(DBZ[j], dzf_tzz);
                DivGuard dg;
                dg.add(A0); dg.add(A1); dg.add(A2); dg.add(A3);
                float4 t0 = fdivs(A0, g.dx, g.rdx), t1 = fdivs(A1, g.dz, g.rdz), t2 = fdivs(A2, g.dx, g.rdx), t3 = fdivs(A3, g.dz, g.rdz);
                if (PML) {
                    const float4 B0 = add4(mul4(PXN[j], q0), t0), B1 = add4(mul4(PZN[j], q1), t1);
                    const float4 B2 = add4(mul4(PXN[j], q2), t2), B3 = add4(mul4(PZN[j], q3), t3);
                    dg.add(B0); dg.add(B1); dg.add(B2); dg.add(B3);
                    n0 = fdiv4(B0, PXD[j], RPXD[j]); n1 = fdiv4(B1, PZD[j], RPZD[j]); n2 = fdiv4(B2, PXD[j], RPXD[j]); n3 = fdiv4(B3, PZD[j], RPZD[j]);
                } else {
                    n0 = add4(q0, t0); n1 = add4(q1, t1); n2 = add4(q2, t2); n3 = add4(q3, t3);
                }
                if (!dg.ok()) {          // a numerator in the underflow range (the band ahead of a wavefront): scaled sequence
                    t0 = safe_divs(A0, g.dx, g.rdx); t1 = safe_divs(A1, g.dz, g.rdz); t2 = safe_divs(A2, g.dx, g.rdx); t3 = safe_divs(A3, g.dz, g.rdz);
                    if (PML) {
                        n0 = safe_div4(add4(mul4(PXN[j], q0), t0), PXD[j], RPXD[j]); n1 = safe_div4(add4(mul4(PZN[j], q1), t1), PZD[j], RPZD[j]);
                        n2 = safe_div4(add4(mul4(PXN[j], q2), t2), PXD[j], RPXD[j]); n3 = safe_div4(add4(mul4(PZN[j], q3), t3), PZD[j], RPZD[j]);
                    } else {
                        n0 = add4(q0, t0); n1 = add4(q1, t1); n2 = add4(q2, t2); n3 = add4(q3, t3);
                    }
                }
            }
            q0 = sel4(m, n0, q0); q1 = sel4(m, n1, q1); q2 = sel4(m, n2, q2); q3 = sel4(m, n3, q3);
            nvx[j] = add4(q0, q1); nvz[j] = add4(q2, q3);
            if (col_ok && gz < g.nzp) {
                const size_t o = (size_t)gz * g.ld + gx;
                float* P = a.planes + (size_t)s * g.plane + o;
                const size_t fp = (size_t)g.ns * g.plane;
                st4(P + P_VXX * fp, q0); st4(P + P_VXZ * fp, q1); st4(P + P_VZX * fp, q2); st4(P + P_VZZ * fp, q3);
                if (SAVE) {
                    float* H = a.hist + (((size_t)s * a.hist_len + a.tl) * NHIST + 4) * g.plane + o;
                    if (lean) {          // damping-free tile: the adjoint only needs e1 + e2 and e3 + e4 (see k1_tile)
                        __stcs(reinterpret_cast<float4*>(H), sel4(m, add4(dxf_txx, dzb_txz), zero4()));
                        __stcs(reinterpret_cast<float4*>(H + 2 * g.plane), sel4(m, add4(dxb_txz, dzf_tzz), zero4()));
                    } else {
                        __stcs(reinterpret_cast<float4*>(H), sel4(m, dxf_txx, zero4()));
                        __stcs(reinterpret_cast<float4*>(H + g.plane), sel4(m, dzb_txz, zero4()));
                        __stcs(reinterpret_cast<float4*>(H + 2 * g.plane), sel4(m, dxb_txz, zero4()));
                        __stcs(reinterpret_cast<float4*>(H + 3 * g.plane), sel4(m, dzf_tzz, zero4()));
                    }
                }
            }
        }
        if (has_rcv) {                 // receivers of this tile (:405-409): new sums parked in the vxx / vzx rects
#pragma unroll
            for (int j = 0; j < RPT; ++j) { const int co = (R.r0 + j) * TX + R.c0; st4(cv + co, nvx[j]); st4(cv + 2 * CF + co, nvz[j]); }
            __syncthreads();
            for (int i = rcv_lo + tid; i < rcv_hi; i += NTH) {
                const int r = a.rb.id[i], zx = a.rb.zx[i];
                const int z = (zx >> 16) - Z0, x = (zx & 0xffff) - X0;
                const int oh = (z + NN) * RXH + x + HX, oc = z * TX + x;
                const size_t o = ((size_t)s * g.nt + a.it) * a.nr + r;
                a.rcv[0][o] = txx[oh]; a.rcv[1][o] = tzz[oh]; a.rcv[2][o] = txz[oh];
                a.rcv[3][o] = cv[oc]; a.rcv[4][o] = cv[2 * CF + oc];
            }
        }
        if (has_rcv || (FS && tzi == 0)) fence_proxy_async();
        __syncthreads();
        if (tid == 0 && pc.valid) { v_issue<NN>(pc, smem, bar, k, th, tc, g.ns); pc.next(g, a.w, ring); }
        stage = (stage + 1 == NSTAGE) ? 0 : stage + 1;
    }
}

template <int NN, bool FS, bool SAVE>
__global__ void __launch_bounds__(NTH, 2)
elf_v(const __grid_constant__ CUtensorMap th, const __grid_constant__ CUtensorMap tc, const EGeom g, const VArgs a)
{
    using G = Geo<NN>;
    extern __shared__ __align__(128) unsigned char smem[];
    uint64_t* bar = (uint64_t*)(smem + NSTAGE * G::V_STAGE);
    const int tid = threadIdx.x;
    if (tid == 0) { for (int k = 0; k < NSTAGE; ++k) mbar_init(bar + k, 1); }
    __syncthreads();
    const Roles R(tid);
    uint32_t par = 0;
    int stage = 0;
    const int nitems = g.ntx * g.ntz * a.w.nchunks;
    int* ring = (int*)(bar + 4);
    Cursor pc;
    pc.wr = 0;
    pc.set(blockIdx.x, g, a.w);
    griddep_launch_dependents();
    unsigned rd = 0;
    bool first = true;
    for (int item = blockIdx.x; item < nitems; item = ring[rd++ & 3u], first = false) {
        const int tile = item / a.w.nchunks, chunk = item - tile * a.w.nchunks;
        const int s_lo = a.w.s_begin + chunk * a.w.chunk;
        const int s_hi = min(s_lo + a.w.chunk, a.w.s_end);
        if (a.tflags[tile] == 1) v_tile<NN, true, FS, SAVE>(&th, &tc, g, a, smem, bar, par, stage, pc, ring, R, tid, tile, s_lo, s_hi, first);
        else                v_tile<NN, false, FS, SAVE>(&th, &tc, g, a, smem, bar, par, stage, pc, ring, R, tid, tile, s_lo, s_hi, first);
        __syncthreads();
    }
}


// ==========================================================================================
// elf_f : fused forward step = elf_s + elf_v in one launch (:341-409 / :497-565)
//
// The stress update is recomputed on a ring of NN cells around the tile, so the velocity update of the tile
// finds every stress sum it needs in shared memory: the sums never travel to HBM and the velocity splits are
// read once.  Per shot the 4 velocity splits (halo 2NN, double-buffered stage) and the 6 stress splits
// (tile + ring, single buffer refilled as soon as phase A has consumed it) arrive by TMA; all ten split fields are
// written to the other set of a ping-pong pair (neighbouring tiles still read the old values of the ring).
// HBM traffic per cell-update: 10 fields read + 10 written (+ 8 history planes in recording mode).
// ==========================================================================================
struct FArgs { ECoef cp; const unsigned char* tflags; float* planes; int cur; const float* mt; const float* src_v;
               const int64_t *sx, *sz; float* hist; int hist_len, tl, it; int nr; RcvB rb; float* rcv[5]; Walk w; };

template <int NN> __device__ __forceinline__ void f_issue_v(const Cursor& c, unsigned char* smem, uint64_t* bar, int k,
                                                            const CUtensorMap* th2, int ns, int base)
{
    using G = Geo<NN>;
    unsigned char* st = smem + k * G::F_VSTAGE;
    fence_proxy_async();
    mbar_expect_tx(bar + k, 4 * G::HF2 * 4);
#pragma unroll
    for (int f = 0; f < 4; ++f) tma_load_3d(st + f * G::HB2, th2, c.X0 - G::HX2, c.Z0 - 2 * NN, (base + P_VXX + f) * ns + c.s, bar + k);
}
template <int NN> __device__ __forceinline__ void f_issue_s(const Cursor& c, unsigned char* sst, uint64_t* bar,
                                                            const CUtensorMap* th, int ns, int base)
{
    using G = Geo<NN>;
    fence_proxy_async();
    mbar_expect_tx(bar, 6 * G::HF * 4);
#pragma unroll
    for (int f = 0; f < 6; ++f) tma_load_3d(sst + f * G::HB, th, c.X0 - HX, c.Z0 - NN, (base + P_S0 + f) * ns + c.s, bar);
}

// stress update of one float4 group.  hv = index of the group in the velocity rects (pitch RX2), ssp = the group in the
// first stress-split rect, a[6] = the split stresses of the group (out: new where the mask is set, else old), d[4] = D-x vx, D-z vz, D+x vz, D+z vx
template <int NN, bool PML, bool WAIT_S = false>
__device__ __forceinline__ void f_stress_cell(const EGeom& g, unsigned m, const float* vxx, const float* vxz, const float* vzx, const float* vzz,
                                              int hv, const float* ssp, uint64_t* sbar, uint32_t sparity, float4* a, const float4& c11, const float4& c13, const float4& c33, const float4& c55,
                                              const float4& pxn, const float4& pxi, const float4& pzn, const float4& pzi, float4* d)
{
    constexpr int RX2 = Geo<NN>::RX2, HQs = Geo<NN>::HB / 4;
    // one derivative at a time (ELF_SEQ keeps the compiler from hoisting all 24 shared-memory loads to the top):
    // the kernel runs at the 128-register cap and every spilled value comes back at L2 latency
    {
        float sx_[12];
        ldseg2(vxx + hv, vxz + hv, sx_);
        d[0] = xdiff<NN, 0>(sx_, g.c);
    }
    ELF_SEQ();
    {
        float sz_[12];
        ldseg2(vzx + hv, vzz + hv, sz_);
        d[2] = xdiff<NN, 1>(sz_, g.c);
    }
    ELF_SEQ();
    {
        float4 wzb[2 * NN];
#pragma unroll
        for (int q = 0; q < 2 * NN; ++q) wzb[q] = add4(ld4(vzx + hv + (q - NN) * RX2), ld4(vzz + hv + (q - NN) * RX2));
        d[1] = zdiff<NN>(wzb, g.c);
    }
    ELF_SEQ();
    {
        float4 wzf[2 * NN];
#pragma unroll
        for (int q = 0; q < 2 * NN; ++q) wzf[q] = add4(ld4(vxx + hv + (q - NN + 1) * RX2), ld4(vxz + hv + (q - NN + 1) * RX2));
        d[3] = zdiff<NN>(wzf, g.c);
    }
    ELF_SEQ();
    // the stress splits (single TMA buffer, refilled during phase B of the previous shot) are needed only now: the
    // four derivatives above hide part of that load's latency
    if (WAIT_S) { while (!mbar_try(sbar, sparity)) {} }
#pragma unroll
    for (int f = 0; f < 6; ++f) a[f] = ld4(ssp + f * HQs);
    float4 n0, n1, n2, n3, n4, n5;
    if (PML) {
        n0 = mul4(add4(mul4(pxn, a[0]), smul(g.dt_dx, mul4(c11, d[0]))), pxi);
        n1 = mul4(add4(mul4(pzn, a[1]), smul(g.dt_dz, mul4(c13, d[1]))), pzi);
        n2 = mul4(add4(mul4(pxn, a[2]), smul(g.dt_dx, mul4(c13, d[0]))), pxi);
        n3 = mul4(add4(mul4(pzn, a[3]), smul(g.dt_dz, mul4(c33, d[1]))), pzi);
        n4 = mul4(add4(mul4(pxn, a[4]), smul(g.dt_dx, mul4(c55, d[2]))), pxi);
        n5 = mul4(add4(mul4(pzn, a[5]), smul(g.dt_dz, mul4(c55, d[3]))), pzi);
    } else {
        n0 = add4(a[0], smul(g.dt_dx, mul4(c11, d[0])));
        n1 = add4(a[1], smul(g.dt_dz, mul4(c13, d[1])));
        n2 = add4(a[2], smul(g.dt_dx, mul4(c13, d[0])));
        n3 = add4(a[3], smul(g.dt_dz, mul4(c33, d[1])));
        n4 = add4(a[4], smul(g.dt_dx, mul4(c55, d[2])));
        n5 = add4(a[5], smul(g.dt_dz, mul4(c55, d[3])));
    }
    a[0] = sel4(m, n0, a[0]); a[1] = sel4(m, n1, a[1]); a[2] = sel4(m, n2, a[2]);
    a[3] = sel4(m, n3, a[3]); a[4] = sel4(m, n4, a[4]); a[5] = sel4(m, n5, a[5]);
}

template <int NN, bool PML, bool FS, bool SAVE>
__device__ __forceinline__ void f_tile(const CUtensorMap* th, const CUtensorMap* th2, const EGeom& g, const FArgs& a,
                                       unsigned char* smem, uint64_t* bar, volatile int* ctl, Cursor& pcv, Cursor& pcs, int* ring,
                                       int* s_sz, int* s_sx, float* s_sxx, float* s_szz, float* s_sxz,
                                       const Roles& R, int tid, int tile, int s_lo, int s_hi, bool first)
{
    using G = Geo<NN>;
    constexpr int RX2 = G::RX2, HX2 = G::HX2, HQ = G::HB / 4, HQ2 = G::HB2 / 4;
    // loop control of the shot walk lives in the warp's slot of shared memory (ctl[0] = first shot, ctl[1] = end,
    // ctl[2] = shots this CTA has processed): at the 128-register cap the compiler spills exactly these, and a spilled
    // value comes back at L2 latency on the critical path of every shot
    if ((tid & 31) == 0) { ctl[0] = s_lo; ctl[1] = s_hi; }      // one writer per warp slot (all lanes hold the same values)
    __syncwarp();
    const uint64_t pol = l2_keep_policy();
    const int tzi = tile / g.ntx, txi = tile - tzi * g.ntx;
    const int X0 = txi * TX, Z0 = tzi * TZ;
    const int gx = X0 + R.c0, gz0 = Z0 + R.r0;
    const unsigned cm = col_mask<NN>(gx, g.nxp);
    const bool col_ok = gx < g.ld;
    const bool lean = !PML && g.merge;     // merged history planes for damping-free tiles (see k1_tile / k2_tile)
    const size_t fp = (size_t)g.ns * g.plane;
    const int rbase = a.cur ? P_B0 : 0, wbase = a.cur ? 0 : P_B0;
    unsigned char* sst = smem + NSTAGE * G::F_VSTAGE;
    float* ssp = (float*)sst;                               // 6 stress-split rects (tile + ring)
    float* sum = (float*)(sst + 6 * G::HB);                 // 3 stress-sum rects (tile + ring)
    float* txx = sum; float* tzz = sum + HQ; float* txz = sum + 2 * HQ;
    if (tid < s_hi - s_lo) {
        const int s = s_lo + tid;
        const float* M = a.mt + (size_t)s * 9;
        const float v = a.src_v[(size_t)s * g.nt + a.it];
        s_sz[tid] = (int)a.sz[s]; s_sx[tid] = (int)a.sx[s];
        s_sxx[tid] = (-(M[0] / 2.0f)) * v; s_szz[tid] = (-(M[8] / 2.0f)) * v; s_sxz[tid] = (-(M[2] / 2.0f)) * v;
    }
    // PML tiles keep only the scaled profiles dt/2*bcx, dt/2*bcz of the thread's cells; 1 -+ h and the correctly rounded
    // 1/(1 + h) (same bits as the division) are rebuilt where they are used: six float4 of factors held across the
    // walk were what pushed the kernel over the 128-register cap
    float4 C11[RPT], C13[RPT], C33[RPT], C55[RPT], DBX[RPT], DBZ[RPT], HBX[RPT], HBZ[RPT];
#pragma unroll
    for (int j = 0; j < RPT; ++j) {
        const ptrdiff_t o = (ptrdiff_t)(gz0 + j) * g.cpld + gx;
        C11[j] = ldk4(a.cp.c11 + o, pol); C13[j] = ldk4(a.cp.c13 + o, pol);
        C33[j] = ldk4(a.cp.c33 + o, pol); C55[j] = ldk4(a.cp.c55 + o, pol);
        DBX[j] = smul(g.dt, ldk4(a.cp.bx + o, pol)); DBZ[j] = smul(g.dt, ldk4(a.cp.bz + o, pol));
        if (PML) { HBX[j] = smul(g.half_dt, ldk4(a.cp.bcx + o, pol)); HBZ[j] = smul(g.half_dt, ldk4(a.cp.bcz + o, pol)); }
        else { HBX[j] = HBZ[j] = zero4(); }
    }
    const int rcv_lo = a.nr > 0 ? a.rb.start[tile] : 0, rcv_hi = a.nr > 0 ? a.rb.start[tile + 1] : 0;
    const bool has_rcv = rcv_hi > rcv_lo;
    const bool producer = tid == NTH - 32;              // lane 0 of the last warp, which has no ring cells
    if (first) {
        griddep_wait();
        if (producer) {
#pragma unroll
            for (int k = 0; k < NSTAGE; ++k)
                if (pcv.valid) { f_issue_v<NN>(pcv, smem, bar, k, th2, g.ns, rbase); pcv.next(g, a.w, ring); }
            if (pcs.valid) { f_issue_s<NN>(pcs, sst, bar + NSTAGE, th, g.ns, rbase); pcs.next_follow(g, a.w, ring); }
        }
    }
    __syncthreads();

    for (int s = s_lo; s < ctl[1]; ++s) {
        const int n = ctl[2];
        const int k = n & 1;           // V stage; its barrier completes every second shot, the S barrier every shot
        float* vxx = (float*)(smem + k * G::F_VSTAGE); float* vxz = vxx + HQ2; float* vzx = vxx + 2 * HQ2; float* vzz = vxx + 3 * HQ2;
        const int si = s - ctl[0];
        const int szs = s_sz[si], sxs = s_sx[si];
        const float sxx = s_sxx[si], szz = s_szz[si], sxz = s_sxz[si];
        while (!mbar_try(bar + k, (unsigned)(n >> 1) & 1u)) {}
        if (FS && tzi == 0) {          // free-surface velocity rows (:399-402) formed in the staged rects (see elf_s)
            const int t = tid;
            if (t >= HX2 - NN && t < HX2 + TX + NN) {
                const int j = X0 + t - HX2;
                if (j >= NN && j < g.nxp - NN) {
                    const int rh1 = (3 * NN) * RX2 + t, rh = (3 * NN + 1) * RX2 + t, rh2 = (3 * NN - 1) * RX2 + t, rh3 = (3 * NN - 2) * RX2 + t;
                    const float vz1 = vzx[rh1] + vzz[rh1];
                    const float vz1n = (j + 1 < g.nxp - NN) ? vzx[rh1 + 1] + vzz[rh1 + 1] : 0.f;
                    const float vxh = vxx[rh] + vxz[rh];
                    const float nvx = (((vz1n - vz1) + vz1n) - vz1) + vxh;
                    vzx[rh2] = vz1; vzz[rh2] = 0.f; vzx[rh3] = vz1; vzz[rh3] = 0.f;
                    vxx[rh2] = nvx; vxz[rh2] = 0.f;
                }
            }
            __syncthreads();
        }
        // ---- phase A: stress on the tile (coefficients in registers) and on the ring (coefficients from L2) ----
#pragma unroll
        for (int j = 0; j < RPT; ++j) {
            const int r = R.r0 + j, gz = gz0 + j;
            const int hv = (r + 2 * NN) * RX2 + R.c0 + HX2, hs = (r + NN) * RXH + R.c0 + HX;
            const unsigned m = row_in<NN>(gz, g.nzp) ? cm : 0u;
            float4 sp[6], d[4];
            float4 pxn = one4(), pxi = one4(), pzn = one4(), pzi = one4();
            if (PML) {
                float4 hx = HBX[j], hz = HBZ[j];
                opaque4(hx); opaque4(hz);          // keeps the factors from being hoisted out of the shot loop (back into registers)
                pxn = sub4(one4(), hx); pzn = sub4(one4(), hz); pxi = rcp4(add4(one4(), hx)); pzi = rcp4(add4(one4(), hz));
            }
            f_stress_cell<NN, PML, true>(g, m, vxx, vxz, vzx, vzz, hv, ssp + hs, bar + NSTAGE, (unsigned)n & 1u, sp, C11[j], C13[j], C33[j], C55[j], pxn, pxi, pzn, pzi, d);
            if (szs == gz && !(FS && gz < NN)) {
                const int dc = sxs - gx;
                if (dc >= 0 && dc < 4) { addc4(sp[0], dc, sxx); addc4(sp[1], dc, sxx); addc4(sp[2], dc, szz); addc4(sp[3], dc, szz); addc4(sp[4], dc, sxz); addc4(sp[5], dc, sxz); }
            }
            st4(txx + hs, add4(sp[0], sp[1])); st4(tzz + hs, add4(sp[2], sp[3])); st4(txz + hs, add4(sp[4], sp[5]));
            if (col_ok && gz < g.nzp) {
                const size_t o = (size_t)gz * g.ld + gx;
                float* P = a.planes + (size_t)s * g.plane + o + (size_t)(wbase + P_S0) * fp;
#pragma unroll
                for (int f = 0; f < 6; ++f) st4(P + f * fp, sp[f]);
                if (SAVE) {
                    float* H = a.hist + ((size_t)s * a.hist_len + a.tl) * NHIST * g.plane + o;
                    __stcs(reinterpret_cast<float4*>(H), sel4(m, d[0], zero4()));
                    __stcs(reinterpret_cast<float4*>(H + g.plane), sel4(m, d[1], zero4()));
                    if (lean) __stcs(reinterpret_cast<float4*>(H + 2 * g.plane), sel4(m, add4(d[2], d[3]), zero4()));
                    else {
                        __stcs(reinterpret_cast<float4*>(H + 2 * g.plane), sel4(m, d[2], zero4()));
                        __stcs(reinterpret_cast<float4*>(H + 3 * g.plane), sel4(m, d[3], zero4()));
                    }
                }
            }
        }
        for (int i = tid; i < G::NRING; i += NTH) {
            int r, gi;
            ring_cell<NN>(i, r, gi);
            const int gzr = Z0 + r, gxr = X0 + 4 * gi;
            const int hv = (r + 2 * NN) * RX2 + 4 * gi + HX2, hs = (r + NN) * RXH + 4 * gi + HX;
            const unsigned m = row_in<NN>(gzr, g.nzp) ? col_mask<NN>(gxr, g.nxp) : 0u;
            const ptrdiff_t o = (ptrdiff_t)gzr * g.cpld + gxr;
            float4 pxn = one4(), pxi = one4(), pzn = one4(), pzi = one4();
            if (PML) {
                const float4 bx_ = ldk4(a.cp.bcx + o, pol), bz_ = ldk4(a.cp.bcz + o, pol);
                pxn = sub4(one4(), smul(g.half_dt, bx_)); pzn = sub4(one4(), smul(g.half_dt, bz_));
                pxi = rcp4(add4(one4(), smul(g.half_dt, bx_))); pzi = rcp4(add4(one4(), smul(g.half_dt, bz_)));
            }
            float4 sp[6], d[4];
            f_stress_cell<NN, PML>(g, m, vxx, vxz, vzx, vzz, hv, ssp + hs, nullptr, 0u, sp, ldk4(a.cp.c11 + o, pol), ldk4(a.cp.c13 + o, pol), ldk4(a.cp.c33 + o, pol),
                                   ldk4(a.cp.c55 + o, pol), pxn, pxi, pzn, pzi, d);
            if (szs == gzr && !(FS && gzr < NN)) {
                const int dc = sxs - gxr;
                if (dc >= 0 && dc < 4) { addc4(sp[0], dc, sxx); addc4(sp[1], dc, sxx); addc4(sp[2], dc, szz); addc4(sp[3], dc, szz); addc4(sp[4], dc, sxz); addc4(sp[5], dc, sxz); }
            }
            st4(txx + hs, add4(sp[0], sp[1])); st4(tzz + hs, add4(sp[2], sp[3])); st4(txz + hs, add4(sp[4], sp[5]));
        }
        __syncthreads();
        // the stress-split buffer has been consumed: refill it with the next shot while phase B runs
        if (producer && pcs.valid) { f_issue_s<NN>(pcs, sst, bar + NSTAGE, th, g.ns, rbase); pcs.next_follow(g, a.w, ring); }
        if (FS && tzi == 0) {          // free-surface stress rows (:380-384) on the sums
            const int t = tid;
            if (t < RXH) {
                const int j = X0 + t - HX;
                if (j >= NN && j < g.nxp - NN) {
                    const int rh1 = (2 * NN) * RXH + t, rh = (2 * NN + 1) * RXH + t, rh2 = (2 * NN - 1) * RXH + t, rh3 = (2 * NN - 2) * RXH + t;
                    tzz[rh1] = 0.f;
                    txz[rh2] = -txz[rh1];
                    tzz[rh2] = -tzz[rh];
                    txz[rh3] = -txz[rh];
                }
            }
            __syncthreads();
        }
        // ---- phase B: velocity on the tile, history, receivers ------------------------------------------
        float4 nvx[RPT], nvz[RPT];
#pragma unroll
        for (int j = 0; j < RPT; ++j) {
            const int r = R.r0 + j, gz = gz0 + j;
            const int hv = (r + 2 * NN) * RX2 + R.c0 + HX2, hs = (r + NN) * RXH + R.c0 + HX;
            float4 dxf_txx, dxb_txz, dzb_txz, dzf_tzz;
            {
                float sxx_[12];
                ldseg(txx + hs, sxx_);
                dxf_txx = xdiff<NN, 1>(sxx_, g.c);
            }
            ELF_SEQ();
            {
                float sxz_[12];
                ldseg(txz + hs, sxz_);
                dxb_txz = xdiff<NN, 0>(sxz_, g.c);
            }
            ELF_SEQ();
            {
                float4 wzb[2 * NN], wzf[2 * NN];
#pragma unroll
                for (int q = 0; q < 2 * NN; ++q) {
                    wzb[q] = ld4(txz + hs + (q - NN) * RXH);
                    wzf[q] = ld4(tzz + hs + (q - NN + 1) * RXH);
                }
                dzb_txz = zdiff<NN>(wzb, g.c); dzf_tzz = zdiff<NN>(wzf, g.c);
            }
            ELF_SEQ();
            float4 q0 = ld4(vxx + hv), q1 = ld4(vxz + hv), q2 = ld4(vzx + hv), q3 = ld4(vzz + hv);
            const unsigned m = row_in<NN>(gz, g.nzp) ? cm : 0u;
            float4 n0, n1, n2, n3;
            {
                const float4 A0 = mul4(DBX[j], dxf_txx), A1 = mul4(DBX[j], dzb_txz), A2 = mul4(DBZ[j], dxb_txz), A3 = mul4(DBZ[j], dzf_tzz);
                DivGuard dg;
                dg.add(A0); dg.add(A1); dg.add(A2); dg.add(A3);
                float4 t0 = fdivs(A0, g.dx, g.rdx), t1 = fdivs(A1, g.dz, g.rdz), t2 = fdivs(A2, g.dx, g.rdx), t3 = fdivs(A3, g.dz, g.rdz);
                if (PML) {
                    float4 hx = HBX[j], hz = HBZ[j];
                    opaque4(hx); opaque4(hz);
                    const float4 pxn = sub4(one4(), hx), pzn = sub4(one4(), hz), pxd = add4(one4(), hx), pzd = add4(one4(), hz);
                    const float4 B0 = add4(mul4(pxn, q0), t0), B1 = add4(mul4(pzn, q1), t1);
                    const float4 B2 = add4(mul4(pxn, q2), t2), B3 = add4(mul4(pzn, q3), t3);
                    dg.add(B0); dg.add(B1); dg.add(B2); dg.add(B3);
                    const float4 pxi = rcp4(pxd), pzi = rcp4(pzd);
                    n0 = fdiv4(B0, pxd, pxi); n1 = fdiv4(B1, pzd, pzi); n2 = fdiv4(B2, pxd, pxi); n3 = fdiv4(B3, pzd, pzi);
                } else {
                    n0 = add4(q0, t0); n1 = add4(q1, t1); n2 = add4(q2, t2); n3 = add4(q3, t3);
                }
                if (!dg.ok()) {
                    t0 = safe_divs(A0, g.dx, g.rdx); t1 = safe_divs(A1, g.dz, g.rdz); t2 = safe_divs(A2, g.dx, g.rdx); t3 = safe_divs(A3, g.dz, g.rdz);
                    if (PML) {
                        const float4 pxn = sub4(one4(), HBX[j]), pzn = sub4(one4(), HBZ[j]), pxd = add4(one4(), HBX[j]), pzd = add4(one4(), HBZ[j]);
                        const float4 pxi = rcp4(pxd), pzi = rcp4(pzd);
                        n0 = safe_div4(add4(mul4(pxn, q0), t0), pxd, pxi); n1 = safe_div4(add4(mul4(pzn, q1), t1), pzd, pzi);
                        n2 = safe_div4(add4(mul4(pxn, q2), t2), pxd, pxi); n3 = safe_div4(add4(mul4(pzn, q3), t3), pzd, pzi);
                    } else {
                        n0 = add4(q0, t0); n1 = add4(q1, t1); n2 = add4(q2, t2); n3 = add4(q3, t3);
                    }
                }
            }
            q0 = sel4(m, n0, q0); q1 = sel4(m, n1, q1); q2 = sel4(m, n2, q2); q3 = sel4(m, n3, q3);
            nvx[j] = add4(q0, q1); nvz[j] = add4(q2, q3);
            if (col_ok && gz < g.nzp) {
                const size_t o = (size_t)gz * g.ld + gx;
                float* P = a.planes + (size_t)s * g.plane + o + (size_t)(wbase + P_VXX) * fp;
                st4(P, q0); st4(P + fp, q1); st4(P + 2 * fp, q2); st4(P + 3 * fp, q3);
                if (SAVE) {
                    float* H = a.hist + (((size_t)s * a.hist_len + a.tl) * NHIST + 4) * g.plane + o;
                    if (lean) {          // damping-free tile: the adjoint only needs e1 + e2 and e3 + e4 (see k1_tile)
                        __stcs(reinterpret_cast<float4*>(H), sel4(m, add4(dxf_txx, dzb_txz), zero4()));
                        __stcs(reinterpret_cast<float4*>(H + 2 * g.plane), sel4(m, add4(dxb_txz, dzf_tzz), zero4()));
                    } else {
                        __stcs(reinterpret_cast<float4*>(H), sel4(m, dxf_txx, zero4()));
                        __stcs(reinterpret_cast<float4*>(H + g.plane), sel4(m, dzb_txz, zero4()));
                        __stcs(reinterpret_cast<float4*>(H + 2 * g.plane), sel4(m, dxb_txz, zero4()));
                        __stcs(reinterpret_cast<float4*>(H + 3 * g.plane), sel4(m, dzf_tzz, zero4()));
                    }
                }
            }
        }
        if (has_rcv) {                 // receivers of this tile (:405-409): new sums parked in the vxx / vzx rects
#pragma unroll
            for (int j = 0; j < RPT; ++j) { const int hv = (R.r0 + j + 2 * NN) * RX2 + R.c0 + HX2; st4(vxx + hv, nvx[j]); st4(vzx + hv, nvz[j]); }
            __syncthreads();
            for (int i = rcv_lo + tid; i < rcv_hi; i += NTH) {
                const int r = a.rb.id[i], zx = a.rb.zx[i];
                const int z = (zx >> 16) - Z0, x = (zx & 0xffff) - X0;
                const int oh = (z + NN) * RXH + x + HX, ov = (z + 2 * NN) * RX2 + x + HX2;
                const size_t o = ((size_t)s * g.nt + a.it) * a.nr + r;
                a.rcv[0][o] = txx[oh]; a.rcv[1][o] = tzz[oh]; a.rcv[2][o] = txz[oh];
                a.rcv[3][o] = vxx[ov]; a.rcv[4][o] = vzx[ov];
            }
        }
        if (has_rcv || (FS && tzi == 0)) fence_proxy_async();
        __syncthreads();               // a full barrier: the stress sums are single-buffered, phase A of the next shot overwrites what phase B reads
        if (producer && pcv.valid) { f_issue_v<NN>(pcv, smem, bar, ctl[2] & 1, th2, g.ns, rbase); pcv.next(g, a.w, ring); }
        __syncwarp();
        if ((tid & 31) == 0) ctl[2] = ctl[2] + 1;
        __syncwarp();
    }
}

template <int NN, bool FS, bool SAVE>
__global__ void __launch_bounds__(NTH, NN == 2 ? 2 : 1)      // O(2,6): one CTA per SM (shared memory) -> the full register file per thread
elf_f(const __grid_constant__ CUtensorMap th, const __grid_constant__ CUtensorMap th2, const EGeom g, const FArgs a)
{
    using G = Geo<NN>;
    extern __shared__ __align__(128) unsigned char smem[];
    uint64_t* bar = (uint64_t*)(smem + G::F_SMEM);
    int* s_sz = (int*)(bar + 8);
    int* s_sx = s_sz + CMAX;
    float* s_sxx = (float*)(s_sx + CMAX);
    float* s_szz = s_sxx + CMAX;
    float* s_sxz = s_szz + CMAX;
    const int tid = threadIdx.x;
    if (tid == 0) { for (int k = 0; k < NSTAGE + 1; ++k) mbar_init(bar + k, 1); }
    __syncthreads();
    const Roles R(tid);
    static_assert(NSTAGE == 2, "f_tile derives stage and barrier parities from the shot counter");
    volatile int* ctl = (volatile int*)(smem + G::F_SMEM + TAIL_BYTES) + (tid >> 5) * 4;
    if ((tid & 31) == 0) ctl[2] = 0;
    __syncwarp();
    const int nitems = g.ntx * g.ntz * a.w.nchunks;
    int* ring = (int*)(bar + 4);
    // the two producer cursors live in shared memory: only one thread uses them, registers are scarce
    Cursor& pcv = *(Cursor*)(s_sxz + CMAX);
    Cursor& pcs = *((Cursor*)(s_sxz + CMAX) + 1);
    if (tid == NTH - 32) {
        pcv.wr = 0; pcs.wr = 0;
        pcv.set(blockIdx.x, g, a.w);
        pcs.set(blockIdx.x, g, a.w);
    }
    griddep_launch_dependents();
    unsigned rd = 0;
    bool first = true;
    for (int item = blockIdx.x; item < nitems; item = ring[rd++ & 3u], first = false) {
        const int tile = item / a.w.nchunks, chunk = item - tile * a.w.nchunks;
        const int s_lo = a.w.s_begin + chunk * a.w.chunk;
        const int s_hi = min(s_lo + a.w.chunk, a.w.s_end);
        if (a.tflags[tile] == 1) f_tile<NN, true, FS, SAVE>(&th, &th2, g, a, smem, bar, ctl, pcv, pcs, ring, s_sz, s_sx, s_sxx, s_szz, s_sxz, R, tid, tile, s_lo, s_hi, first);
        else                f_tile<NN, false, FS, SAVE>(&th, &th2, g, a, smem, bar, ctl, pcv, pcs, ring, s_sz, s_sx, s_sxx, s_szz, s_sxz, R, tid, tile, s_lo, s_hi, first);
        __syncthreads();
    }
}

// ==========================================================================================
// elf_k1 : adjoint of the velocity update
// ==========================================================================================
template <int NN> __device__ __forceinline__ void k1_issue(const Cursor& c, unsigned char* smem, uint64_t* bar, int k,
                                                           const CUtensorMap* th, int ns, int lcur, const CUtensorMap* thh, int hist_len, int tl,
                                                           bool lean)
{
    // lean (damping-free tile, see elf_tile_class): the two splits of a pair hold the same bits, only the first one is
    // staged; the history carries the merged derivatives in planes 4 and 6
#pragma unroll
    for (int e = 4; e < 8; ++e) if (!lean || !(e & 1)) tma_prefetch_3d(thh, c.X0, c.Z0, (c.s * hist_len + tl) * NHIST + e);
    using G = Geo<NN>;
    unsigned char* st = smem + k * G::K1_STAGE;
    fence_proxy_async();
    mbar_expect_tx(bar + k, (lean ? 4 : 6) * G::HF * 4);
#pragma unroll
    for (int f = 0; f < 4; ++f) if (!lean || !(f & 1)) tma_load_3d(st + f * G::HB, th, c.X0 - HX, c.Z0 - NN, (P_LV + 4 * lcur + f) * ns + c.s, bar + k);
    tma_load_3d(st + 4 * G::HB, th, c.X0 - HX, c.Z0 - NN, P_LVX * ns + c.s, bar + k);
    tma_load_3d(st + 5 * G::HB, th, c.X0 - HX, c.Z0 - NN, P_LVZ * ns + c.s, bar + k);
}

// own-cell transpose of the velocity update for one float4 group (all in registers)
template <bool PML>
__device__ __forceinline__ void k1_cell(const EGeom& g, unsigned m, const float4& L0, const float4& L1, const float4& L2, const float4& L3,
                                        const float4& lvx, const float4& lvz, const float4& bx, const float4& bz,
                                        const float4& pxn, const float4& pxd, const float4& pzn, const float4& pzd,
                                        const float4& rpxd, const float4& rpzd,
                                        float4& w1, float4& w2, float4& w3, float4& w4,
                                        float4& m1, float4& m2, float4& m3, float4& m4,
                                        float4& N0, float4& N1, float4& N2, float4& N3)
{
    const float4 q1 = add4(L0, lvx), q2 = add4(L1, lvx), q3 = add4(L2, lvz), q4 = add4(L3, lvz);
    // the adjoint does not have to reproduce the eager rounding bit for bit (only the forward records do), so the
    // divisions by dx, dz and (1 + dt/2*profile) are multiplications by their correctly rounded reciprocals here:
    // 1-2 ulp per factor against a 1e-4 gradient tolerance
    w1 = muls(muls(q1, g.dt), g.rdx); w2 = muls(muls(q2, g.dt), g.rdz); w3 = muls(muls(q3, g.dt), g.rdx); w4 = muls(muls(q4, g.dt), g.rdz);
    if (PML) {
        w1 = mul4(w1, rpxd); w2 = mul4(w2, rpzd); w3 = mul4(w3, rpxd); w4 = mul4(w4, rpzd);
        N0 = mul4(mul4(pxn, q1), rpxd); N1 = mul4(mul4(pzn, q2), rpzd); N2 = mul4(mul4(pxn, q3), rpxd); N3 = mul4(mul4(pzn, q4), rpzd);
    } else {
        N0 = q1; N1 = q2; N2 = q3; N3 = q4;
    }
    (void)pxd; (void)pzd;
    w1 = sel4(m, w1, zero4()); w2 = sel4(m, w2, zero4()); w3 = sel4(m, w3, zero4()); w4 = sel4(m, w4, zero4());
    m1 = mul4(w1, bx); m2 = mul4(w2, bx); m3 = mul4(w3, bz); m4 = mul4(w4, bz);
    N0 = sel4(m, N0, L0); N1 = sel4(m, N1, L1); N2 = sel4(m, N2, L2); N3 = sel4(m, N3, L3);
}

template <int NN, bool PML, bool FS>
__device__ __forceinline__ void k1_tile(const CUtensorMap* th, const CUtensorMap* thh, const EGeom& g, const K1Args& a,
                                        unsigned char* smem, uint64_t* bar, uint32_t& par, int& stage, Cursor& pc, int* ring,
                                        const Roles& R, int tid, int tile, int chunk, int s_lo, int s_hi, bool first)
{
    using G = Geo<NN>;
    const uint64_t pol = l2_keep_policy();
    const int tzi = tile / g.ntx, txi = tile - tzi * g.ntx;
    const int X0 = txi * TX, Z0 = tzi * TZ;
    const int gx = X0 + R.c0, gz0 = Z0 + R.r0;
    const unsigned cm = col_mask<NN>(gx, g.nxp);
    const bool col_ok = gx < g.ld;
    const size_t fp = (size_t)g.ns * g.plane;
    float4 BX[RPT], BZ[RPT], PXN[RPT], PXD[RPT], PZN[RPT], PZD[RPT], RPXD[RPT], RPZD[RPT], GBX[RPT], GBZ[RPT];
#pragma unroll
    for (int j = 0; j < RPT; ++j) {
        const ptrdiff_t o = (ptrdiff_t)(gz0 + j) * g.cpld + gx;
        BX[j] = ldk4(a.cp.bx + o, pol); BZ[j] = ldk4(a.cp.bz + o, pol);
        if (PML) {
            const float4 bx_ = ldk4(a.cp.bcx + o, pol), bz_ = ldk4(a.cp.bcz + o, pol);
            PXD[j] = add4(one4(), smul(g.half_dt, bx_)); PZD[j] = add4(one4(), smul(g.half_dt, bz_));
            PXN[j] = sub4(one4(), smul(g.half_dt, bx_)); PZN[j] = sub4(one4(), smul(g.half_dt, bz_));
            RPXD[j] = div4(one4(), PXD[j]); RPZD[j] = div4(one4(), PZD[j]);
        } else { PXD[j] = PZD[j] = PXN[j] = PZN[j] = RPXD[j] = RPZD[j] = one4(); }
        GBX[j] = zero4(); GBZ[j] = zero4();
    }
    const bool lean = !PML && g.merge;                 // splits of a pair are bitwise equal on the whole staged rectangle
    const bool deep = lean && a.tflags[tile] == 2;     // ... and no neighbouring tile reads the second split of this tile's cells
    const bool have_gv = a.nr > 0 && (a.g[3] || a.g[4]);
    const bool have_gs = a.nr > 0 && (a.g[0] || a.g[1] || a.g[2]);
    const bool inject_v = have_gv && a.rb.nbr[tile];
    const int rcv_lo = a.nr > 0 ? a.rb.start[tile] : 0, rcv_hi = a.nr > 0 ? a.rb.start[tile + 1] : 0;
    const bool inject_s = have_gs && rcv_hi > rcv_lo;
    if (first) {
        griddep_wait();
#pragma unroll
        for (int k = 0; k < NSTAGE; ++k)
            if (tid == 0 && pc.valid) { k1_issue<NN>(pc, smem, bar, k, th, g.ns, a.lcur, thh, a.hist_len, a.tl, g.merge && a.tflags[pc.tile] != 1); pc.next(g, a.w, ring); }
    }

    for (int s = s_lo; s < s_hi; ++s) {
        const int k = stage;
        float* L = (float*)(smem + k * G::K1_STAGE);       // rects 0..3: velocity-split cotangents (become m1..m4), 4: lvx, 5: lvz
        float* lvx = L + 4 * (G::HB / 4); float* lvz = L + 5 * (G::HB / 4);
        // history of this step (own cells): e1 = D+x txx, e2 = D-z txz, e3 = D-x txz, e4 = D+z tzz
        float4 E[4][RPT];
#pragma unroll
        for (int j = 0; j < RPT; ++j) {
            const int gz = gz0 + j;
            const float* H = a.hist + (((size_t)s * a.hist_len + a.tl) * NHIST + 4) * g.plane + (size_t)gz * g.ld + gx;
#pragma unroll
            for (int e = 0; e < 4; ++e)
                E[e][j] = (col_ok && gz < g.nzp && !(lean && (e & 1))) ? __ldcs(reinterpret_cast<const float4*>(H + (size_t)e * g.plane)) : zero4();
        }
        ELF_WAIT_STAGE(k);
        if (inject_v) {          // 10T: cotangents of the vx / vz records into the staged sums (duplicates legal)
            for (int dz = -1; dz <= 1; ++dz) {
                const int tz2 = tzi + dz;
                if (tz2 < 0 || tz2 >= g.ntz) continue;
                for (int dx = -1; dx <= 1; ++dx) {
                    const int tx2 = txi + dx;
                    if (tx2 < 0 || tx2 >= g.ntx) continue;
                    const int t2 = tz2 * g.ntx + tx2;
                    const int lo = a.rb.start[t2], hi = a.rb.start[t2 + 1];
                    for (int i = lo + tid; i < hi; i += NTH) {
                        const int zx = a.rb.zx[i];
                        const int z = (zx >> 16) - (Z0 - NN), x = (zx & 0xffff) - (X0 - HX);
                        if (z >= 0 && z < G::RZH && x >= 0 && x < RXH) {
                            const size_t o = ((size_t)s * g.nt + a.it) * a.nr + a.rb.id[i];
                            if (a.g[3]) atomicAdd(lvx + z * RXH + x, a.g[3][o]);
                            if (a.g[4]) atomicAdd(lvz + z * RXH + x, a.g[4][o]);
                        }
                    }
                }
            }
            __syncthreads();
        }
        if (FS && tzi == 0) {    // 9T: transpose of the free-surface velocity edits (adds into rows h-1, h)
            const int t = tid;
            if (t >= 1 && t < RXH) {
                const int j = X0 + t - HX;
                if (j >= NN && j < g.nxp - NN) {
                    const int rh1 = (2 * NN) * RXH + t, rh = (2 * NN + 1) * RXH + t, rh2 = (2 * NN - 1) * RXH + t, rh3 = (2 * NN - 2) * RXH + t;
                    const float qj = lvx[rh2];
                    const float qm = (j - 1 >= NN) ? lvx[rh2 - 1] : 0.f;
                    const float add_vz = lvz[rh2] + lvz[rh3] + 2.0f * (qm - qj);
                    lvz[rh1] += add_vz;
                    lvx[rh] += qj;
                }
            }
            __syncthreads();
        }
        // ---- pass A: own-cell transpose; m1..m4 replace the split cotangents in the staged rects ------
#pragma unroll
        for (int j = 0; j < RPT; ++j) {
            const int r = R.r0 + j, gz = gz0 + j;
            const int hb = (r + NN) * RXH + R.c0 + HX;
            const unsigned m = row_in<NN>(gz, g.nzp) ? cm : 0u;
            const float4 L0 = ld4(L + hb), L2 = ld4(L + 2 * (G::HB / 4) + hb);
            const float4 L1 = lean ? L0 : ld4(L + G::HB / 4 + hb), L3 = lean ? L2 : ld4(L + 3 * (G::HB / 4) + hb);
            float4 w1, w2, w3, w4, m1, m2, m3, m4, N0, N1, N2, N3;
            k1_cell<PML>(g, m, L0, L1, L2, L3, ld4(lvx + hb), ld4(lvz + hb), BX[j], BZ[j], PXN[j], PXD[j], PZN[j], PZD[j], RPXD[j], RPZD[j],
                         w1, w2, w3, w4, m1, m2, m3, m4, N0, N1, N2, N3);
            st4(L + hb, m1); st4(L + G::HB / 4 + hb, m2); st4(L + 2 * (G::HB / 4) + hb, m3); st4(L + 3 * (G::HB / 4) + hb, m4);
            if (lean) {          // w1 == w2, w3 == w4 (dx == dz): the history holds e1 + e2 and e3 + e4
                GBX[j] = add4(GBX[j], mul4(w1, E[0][j]));
                GBZ[j] = add4(GBZ[j], mul4(w3, E[2][j]));
            } else {
                GBX[j] = add4(GBX[j], add4(mul4(w1, E[0][j]), mul4(w2, E[1][j])));
                GBZ[j] = add4(GBZ[j], add4(mul4(w3, E[2][j]), mul4(w4, E[3][j])));
            }
            if (col_ok && gz < g.nzp) {
                float* P = a.planes + (size_t)s * g.plane + (size_t)gz * g.ld + gx + (size_t)(P_LV + 4 * (a.lcur ^ 1)) * fp;
                st4(P, N0); st4(P + 2 * fp, N2);
                if (!deep) { st4(P + fp, N1); st4(P + 3 * fp, N3); }
            }
        }
        for (int i = tid; i < G::NRING; i += NTH) {     // ring: m only (coefficients from the L2-resident pack)
            int r, gi;
            ring_cell<NN>(i, r, gi);
            const int gzr = Z0 + r, gxr = X0 + 4 * gi;
            const int hb = (r + NN) * RXH + 4 * gi + HX;
            const unsigned m = row_in<NN>(gzr, g.nzp) ? col_mask<NN>(gxr, g.nxp) : 0u;
            const ptrdiff_t o = (ptrdiff_t)gzr * g.cpld + gxr;
            float4 pxn = one4(), pxd = one4(), pzn = one4(), pzd = one4(), rpxd = one4(), rpzd = one4();
            if (PML) {
                const float4 bx_ = ldk4(a.cp.bcx + o, pol), bz_ = ldk4(a.cp.bcz + o, pol);
                pxd = add4(one4(), smul(g.half_dt, bx_)); pzd = add4(one4(), smul(g.half_dt, bz_));
                pxn = sub4(one4(), smul(g.half_dt, bx_)); pzn = sub4(one4(), smul(g.half_dt, bz_));
                rpxd = div4(one4(), pxd); rpzd = div4(one4(), pzd);
            }
            const float4 L0 = ld4(L + hb), L2 = ld4(L + 2 * (G::HB / 4) + hb);
            const float4 L1 = lean ? L0 : ld4(L + G::HB / 4 + hb), L3 = lean ? L2 : ld4(L + 3 * (G::HB / 4) + hb);
            float4 w1, w2, w3, w4, m1, m2, m3, m4, N0, N1, N2, N3;
            k1_cell<PML>(g, m, L0, L1, L2, L3, ld4(lvx + hb), ld4(lvz + hb), ldk4(a.cp.bx + o, pol), ldk4(a.cp.bz + o, pol), pxn, pxd, pzn, pzd, rpxd, rpzd,
                         w1, w2, w3, w4, m1, m2, m3, m4, N0, N1, N2, N3);
            st4(L + hb, m1); st4(L + G::HB / 4 + hb, m2); st4(L + 2 * (G::HB / 4) + hb, m3); st4(L + 3 * (G::HB / 4) + hb, m4);
        }
        __syncthreads();
        // ---- pass B: 6T gathers -> cotangents of the stress sums -----------------------------------
        {
            const float* M1 = L; const float* M2 = L + G::HB / 4; const float* M3 = L + 2 * (G::HB / 4); const float* M4 = L + 3 * (G::HB / 4);
#pragma unroll
            for (int j = 0; j < RPT; ++j) {
                const int r = R.r0 + j, gz = gz0 + j;
                const int hb = (r + NN) * RXH + R.c0 + HX;
                float s1[12], s3[12];
                ldseg(M1 + hb, s1); ldseg(M3 + hb, s3);
                float4 w2[2 * NN], w4[2 * NN];
#pragma unroll
                for (int q = 0; q < 2 * NN; ++q) {
                    w2[q] = ld4(M2 + hb + (q - NN + 1) * RXH);      // (D-z)^T: rows i-NN+1 .. i+NN
                    w4[q] = ld4(M4 + hb + (q - NN) * RXH);          // (D+z)^T: rows i-NN .. i+NN-1
                }
                const float4 mxx = xgath<NN, 0>(s1, g.c);
                const float4 mxz = add4(zgath<NN>(w2, g.c), xgath<NN, 1>(s3, g.c));
                const float4 mzz = zgath<NN>(w4, g.c);
                if (col_ok && gz < g.nzp) {
                    float* P = a.planes + (size_t)s * g.plane + (size_t)gz * g.ld + gx;
                    st4(P + P_MXX * fp, mxx); st4(P + P_MZZ * fp, mzz); st4(P + P_MXZ * fp, mxz);
                }
            }
        }
        if (inject_s) {          // 10T: cotangents of the stress records add to the sums' cotangents of this tile's cells
            __syncthreads();
            for (int i = rcv_lo + tid; i < rcv_hi; i += NTH) {
                const int zx = a.rb.zx[i];
                const size_t o = ((size_t)s * g.nt + a.it) * a.nr + a.rb.id[i];
                float* P = a.planes + (size_t)s * g.plane + (size_t)(zx >> 16) * g.ld + (zx & 0xffff);
                if (a.g[0]) atomicAdd(P + P_MXX * fp, a.g[0][o]);
                if (a.g[1]) atomicAdd(P + P_MZZ * fp, a.g[1][o]);
                if (a.g[2]) atomicAdd(P + P_MXZ * fp, a.g[2][o]);
            }
        }
        fence_proxy_async();
        __syncthreads();
        if (tid == 0 && pc.valid) { k1_issue<NN>(pc, smem, bar, k, th, g.ns, a.lcur, thh, a.hist_len, a.tl, g.merge && a.tflags[pc.tile] != 1); pc.next(g, a.w, ring); }
        stage = (stage + 1 == NSTAGE) ? 0 : stage + 1;
    }
#pragma unroll
    for (int j = 0; j < RPT; ++j) {
        const int gz = gz0 + j;
        if (col_ok && gz < g.nzp) {
            float* gp = a.gpart + (size_t)chunk * 6 * g.plane + (size_t)gz * g.ld + gx;
            red4(gp + 4 * g.plane, GBX[j]); red4(gp + 5 * g.plane, GBZ[j]);
        }
    }
}

template <int NN, bool FS>
__global__ void __launch_bounds__(NTH, 2)
elf_k1(const __grid_constant__ CUtensorMap th, const __grid_constant__ CUtensorMap thh, const EGeom g, const K1Args a)
{
    using G = Geo<NN>;
    extern __shared__ __align__(128) unsigned char smem[];
    uint64_t* bar = (uint64_t*)(smem + NSTAGE * G::K1_STAGE);
    const int tid = threadIdx.x;
    if (tid == 0) { for (int k = 0; k < NSTAGE; ++k) mbar_init(bar + k, 1); }
    __syncthreads();
    const Roles R(tid);
    uint32_t par = 0;
    int stage = 0;
    const int nitems = g.ntx * g.ntz * a.w.nchunks;
    int* ring = (int*)(bar + 4);
    Cursor pc;
    pc.wr = 0;
    pc.set(blockIdx.x, g, a.w);
    griddep_launch_dependents();
    unsigned rd = 0;
    bool first = true;
    for (int item = blockIdx.x; item < nitems; item = ring[rd++ & 3u], first = false) {
        const int tile = item / a.w.nchunks, chunk = item - tile * a.w.nchunks;
        const int s_lo = a.w.s_begin + chunk * a.w.chunk;
        const int s_hi = min(s_lo + a.w.chunk, a.w.s_end);
        if (a.tflags[tile] == 1) k1_tile<NN, true, FS>(&th, &thh, g, a, smem, bar, par, stage, pc, ring, R, tid, tile, chunk, s_lo, s_hi, first);
        else                k1_tile<NN, false, FS>(&th, &thh, g, a, smem, bar, par, stage, pc, ring, R, tid, tile, chunk, s_lo, s_hi, first);
        __syncthreads();
    }
}

// ==========================================================================================
// elf_k2 : adjoint of the stress update (Appendix A.2, steps 5T..1T)
// ==========================================================================================
template <int NN> __device__ __forceinline__ void k2_issue(const Cursor& c, unsigned char* smem, uint64_t* bar, int k,
                                                           const CUtensorMap* th, int ns, int lcur, const CUtensorMap* thh, int hist_len, int tl,
                                                           bool lean)
{
    // lean: see k1_issue; the history carries D-x vx, D-z vz and the merged D+x vz + D+z vx in planes 0..2
#pragma unroll
    for (int e = 0; e < 4; ++e) if (!lean || e < 3) tma_prefetch_3d(thh, c.X0, c.Z0, (c.s * hist_len + tl) * NHIST + e);
    using G = Geo<NN>;
    unsigned char* st = smem + k * G::K2_STAGE;
    fence_proxy_async();
    mbar_expect_tx(bar + k, (lean ? 6 : 9) * G::HF * 4);
#pragma unroll
    for (int f = 0; f < 3; ++f) tma_load_3d(st + f * G::HB, th, c.X0 - HX, c.Z0 - NN, (P_MXX + f) * ns + c.s, bar + k);
#pragma unroll
    for (int f = 0; f < 6; ++f) if (!lean || !(f & 1)) tma_load_3d(st + (3 + f) * G::HB, th, c.X0 - HX, c.Z0 - NN, (P_LS + 6 * lcur + f) * ns + c.s, bar + k);
}

template <bool PML>
__device__ __forceinline__ void k2_cell(const EGeom& g, unsigned m, const float4* LS, const float4& mxx, const float4& mzz, const float4& mxz,
                                        const float4& c11, const float4& c13, const float4& c33, const float4& c55,
                                        const float4& pxn, const float4& pxi, const float4& pzn, const float4& pzi,
                                        float4* l, float4* q, float4& nA, float4& nB, float4& nC, float4& nD, float4* N)
{
    l[0] = add4(LS[0], mxx); l[1] = add4(LS[1], mxx); l[2] = add4(LS[2], mzz);
    l[3] = add4(LS[3], mzz); l[4] = add4(LS[4], mxz); l[5] = add4(LS[5], mxz);
    float4 t[6];
    if (PML) {
        t[0] = mul4(l[0], pxi); t[1] = mul4(l[1], pzi); t[2] = mul4(l[2], pxi); t[3] = mul4(l[3], pzi); t[4] = mul4(l[4], pxi); t[5] = mul4(l[5], pzi);
        N[0] = mul4(pxn, t[0]); N[1] = mul4(pzn, t[1]); N[2] = mul4(pxn, t[2]); N[3] = mul4(pzn, t[3]); N[4] = mul4(pxn, t[4]); N[5] = mul4(pzn, t[5]);
    } else {
#pragma unroll
        for (int f = 0; f < 6; ++f) { t[f] = l[f]; N[f] = l[f]; }
    }
    q[0] = sel4(m, muls(t[0], g.dt_dx), zero4()); q[1] = sel4(m, muls(t[1], g.dt_dz), zero4());
    q[2] = sel4(m, muls(t[2], g.dt_dx), zero4()); q[3] = sel4(m, muls(t[3], g.dt_dz), zero4());
    q[4] = sel4(m, muls(t[4], g.dt_dx), zero4()); q[5] = sel4(m, muls(t[5], g.dt_dz), zero4());
    nA = add4(mul4(q[0], c11), mul4(q[2], c13));
    nB = add4(mul4(q[1], c13), mul4(q[3], c33));
    nC = mul4(q[4], c55);
    nD = mul4(q[5], c55);
#pragma unroll
    for (int f = 0; f < 6; ++f) N[f] = sel4(m, N[f], LS[f]);
}

template <int NN, bool PML, bool FS>
__device__ __forceinline__ void k2_tile(const CUtensorMap* th, const CUtensorMap* thh, const EGeom& g, const K2Args& a,
                                        unsigned char* smem, uint64_t* bar, uint32_t& par, int& stage, Cursor& pc, int* ring,
                                        int* s_sz, int* s_sx, const Roles& R, int tid, int tile, int chunk, int s_lo, int s_hi, bool first)
{
    using G = Geo<NN>;
    constexpr int HQ = G::HB / 4;
    const uint64_t pol = l2_keep_policy();
    const int tzi = tile / g.ntx, txi = tile - tzi * g.ntx;
    const int X0 = txi * TX, Z0 = tzi * TZ;
    const int gx = X0 + R.c0, gz0 = Z0 + R.r0;
    const unsigned cm = col_mask<NN>(gx, g.nxp);
    const bool col_ok = gx < g.ld;
    const size_t fp = (size_t)g.ns * g.plane;
    const bool lean = !PML && g.merge;                 // see k1_tile
    const bool deep = lean && a.tflags[tile] == 2;
    if (a.g_src && tid < s_hi - s_lo) { s_sz[tid] = (int)a.sz[s_lo + tid]; s_sx[tid] = (int)a.sx[s_lo + tid]; }
    float4 C11[RPT], C13[RPT], C33[RPT], C55[RPT], PXN[RPT], PXI[RPT], PZN[RPT], PZI[RPT], G11[RPT], G13[RPT], G33[RPT], G55[RPT];
#pragma unroll
    for (int j = 0; j < RPT; ++j) {
        const ptrdiff_t o = (ptrdiff_t)(gz0 + j) * g.cpld + gx;
        C11[j] = ldk4(a.cp.c11 + o, pol); C13[j] = ldk4(a.cp.c13 + o, pol);
        C33[j] = ldk4(a.cp.c33 + o, pol); C55[j] = ldk4(a.cp.c55 + o, pol);
        if (PML) {
            const float4 bx_ = ldk4(a.cp.bcx + o, pol), bz_ = ldk4(a.cp.bcz + o, pol);
            const float4 pxd = add4(one4(), smul(g.half_dt, bx_)), pzd = add4(one4(), smul(g.half_dt, bz_));
            PXN[j] = sub4(one4(), smul(g.half_dt, bx_)); PZN[j] = sub4(one4(), smul(g.half_dt, bz_));
            PXI[j] = div4(one4(), pxd); PZI[j] = div4(one4(), pzd);
        } else { PXN[j] = PZN[j] = PXI[j] = PZI[j] = one4(); }
        G11[j] = zero4(); G13[j] = zero4(); G33[j] = zero4(); G55[j] = zero4();
    }
    if (first) {
        griddep_wait();
#pragma unroll
        for (int k = 0; k < NSTAGE; ++k)
            if (tid == 0 && pc.valid) { k2_issue<NN>(pc, smem, bar, k, th, g.ns, a.lcur, thh, a.hist_len, a.tl, g.merge && a.tflags[pc.tile] != 1); pc.next(g, a.w, ring); }
    }
    __syncthreads();

    for (int s = s_lo; s < s_hi; ++s) {
        const int k = stage;
        float* MS = (float*)(smem + k * G::K2_STAGE);      // rects 0..2: mxx, mzz, mxz; 3..8: stress-split cotangents
        float* mzz_r = MS + HQ; float* mxz_r = MS + 2 * HQ;
        float* LSr = MS + 3 * HQ;                           // rects LS0, LS1, LS4, LS5 become nA, nB, nC, nD
        // history of this step (own cells): D-x vx, D-z vz, D+x vz, D+z vx
        float4 D[4][RPT];
#pragma unroll
        for (int j = 0; j < RPT; ++j) {
            const int gz = gz0 + j;
            const float* H = a.hist + ((size_t)s * a.hist_len + a.tl) * NHIST * g.plane + (size_t)gz * g.ld + gx;
#pragma unroll
            for (int e = 0; e < 4; ++e)
                D[e][j] = (col_ok && gz < g.nzp && !(lean && e == 3)) ? __ldcs(reinterpret_cast<const float4*>(H + (size_t)e * g.plane)) : zero4();
        }
        ELF_WAIT_STAGE(k);
        if (FS && tzi == 0) {    // 5T: transpose of the free-surface stress mirrors
            const int t = tid;
            if (t < RXH) {
                const int j = X0 + t - HX;
                if (j >= NN && j < g.nxp - NN) {
                    const int rh1 = (2 * NN) * RXH + t, rh = (2 * NN + 1) * RXH + t, rh2 = (2 * NN - 1) * RXH + t, rh3 = (2 * NN - 2) * RXH + t;
                    mxz_r[rh] -= mxz_r[rh3];
                    mzz_r[rh] -= mzz_r[rh2];
                    mxz_r[rh1] -= mxz_r[rh2];
                    mzz_r[rh1] = 0.f;
                }
            }
            __syncthreads();
        }
        // ---- pass A: own-cell transpose; nA..nD replace four of the split cotangents in the staged rects ----
#pragma unroll
        for (int j = 0; j < RPT; ++j) {
            const int r = R.r0 + j, gz = gz0 + j;
            const int hb = (r + NN) * RXH + R.c0 + HX;
            const unsigned m = row_in<NN>(gz, g.nzp) ? cm : 0u;
            float4 LS[6], l[6], q[6], N[6], nA, nB, nC, nD;
#pragma unroll
            for (int f = 0; f < 6; f += 2) { LS[f] = ld4(LSr + f * HQ + hb); LS[f + 1] = lean ? LS[f] : ld4(LSr + (f + 1) * HQ + hb); }
            k2_cell<PML>(g, m, LS, ld4(MS + hb), ld4(mzz_r + hb), ld4(mxz_r + hb), C11[j], C13[j], C33[j], C55[j], PXN[j], PXI[j], PZN[j], PZI[j],
                         l, q, nA, nB, nC, nD, N);
            st4(LSr + hb, nA); st4(LSr + HQ + hb, nB); st4(LSr + 4 * HQ + hb, nC); st4(LSr + 5 * HQ + hb, nD);
            G11[j] = add4(G11[j], mul4(q[0], D[0][j]));
            G13[j] = add4(G13[j], add4(mul4(q[1], D[1][j]), mul4(q[2], D[0][j])));
            G33[j] = add4(G33[j], mul4(q[3], D[1][j]));
            if (lean) G55[j] = add4(G55[j], mul4(q[4], D[2][j]));       // q4 == q5 (dt/dx == dt/dz): the history holds d3 + d4
            else      G55[j] = add4(G55[j], add4(mul4(q[4], D[2][j]), mul4(q[5], D[3][j])));
            if (col_ok && gz < g.nzp) {
                float* P = a.planes + (size_t)s * g.plane + (size_t)gz * g.ld + gx + (size_t)(P_LS + 6 * (a.lcur ^ 1)) * fp;
#pragma unroll
                for (int f = 0; f < 6; ++f) if (!deep || !(f & 1)) st4(P + f * fp, N[f]);
                if (a.g_src && m != 0u && s_sz[s - s_lo] == gz) {       // 3T
                    const int dc = s_sx[s - s_lo] - gx;
                    if (dc >= 0 && dc < 4 && ((m >> dc) & 1u)) {
                        const float* M = a.mt + (size_t)s * 9;
                        a.g_src[(size_t)s * g.nt + a.it] = -(M[0] / 2.0f) * (comp4(l[0], dc) + comp4(l[1], dc))
                                                          - (M[8] / 2.0f) * (comp4(l[2], dc) + comp4(l[3], dc))
                                                          - (M[2] / 2.0f) * (comp4(l[4], dc) + comp4(l[5], dc));
                    }
                }
            }
        }
        for (int i = tid; i < G::NRING; i += NTH) {
            int r, gi;
            ring_cell<NN>(i, r, gi);
            const int gzr = Z0 + r, gxr = X0 + 4 * gi;
            const int hb = (r + NN) * RXH + 4 * gi + HX;
            const unsigned m = row_in<NN>(gzr, g.nzp) ? col_mask<NN>(gxr, g.nxp) : 0u;
            const ptrdiff_t o = (ptrdiff_t)gzr * g.cpld + gxr;
            float4 pxn = one4(), pxi = one4(), pzn = one4(), pzi = one4();
            if (PML) {
                const float4 bx_ = ldk4(a.cp.bcx + o, pol), bz_ = ldk4(a.cp.bcz + o, pol);
                const float4 pxd = add4(one4(), smul(g.half_dt, bx_)), pzd = add4(one4(), smul(g.half_dt, bz_));
                pxn = sub4(one4(), smul(g.half_dt, bx_)); pzn = sub4(one4(), smul(g.half_dt, bz_));
                pxi = div4(one4(), pxd); pzi = div4(one4(), pzd);
            }
            float4 LS[6], l[6], q[6], N[6], nA, nB, nC, nD;
#pragma unroll
            for (int f = 0; f < 6; f += 2) { LS[f] = ld4(LSr + f * HQ + hb); LS[f + 1] = lean ? LS[f] : ld4(LSr + (f + 1) * HQ + hb); }
            k2_cell<PML>(g, m, LS, ld4(MS + hb), ld4(mzz_r + hb), ld4(mxz_r + hb), ldk4(a.cp.c11 + o, pol), ldk4(a.cp.c13 + o, pol),
                         ldk4(a.cp.c33 + o, pol), ldk4(a.cp.c55 + o, pol), pxn, pxi, pzn, pzi, l, q, nA, nB, nC, nD, N);
            st4(LSr + hb, nA); st4(LSr + HQ + hb, nB); st4(LSr + 4 * HQ + hb, nC); st4(LSr + 5 * HQ + hb, nD);
        }
        __syncthreads();
        // ---- pass B: 2T/1T gathers -> cotangents of the velocity sums (pre-step) --------------------
        {
            const float* NA = LSr; const float* NB = LSr + HQ; const float* NC = LSr + 4 * HQ; const float* ND = LSr + 5 * HQ;
#pragma unroll
            for (int j = 0; j < RPT; ++j) {
                const int r = R.r0 + j, gz = gz0 + j;
                const int hb = (r + NN) * RXH + R.c0 + HX;
                float sa[12], sc[12];
                ldseg(NA + hb, sa); ldseg(NC + hb, sc);
                float4 wb[2 * NN], wd[2 * NN];
#pragma unroll
                for (int q = 0; q < 2 * NN; ++q) {
                    wb[q] = ld4(NB + hb + (q - NN + 1) * RXH);      // (D-z)^T
                    wd[q] = ld4(ND + hb + (q - NN) * RXH);          // (D+z)^T
                }
                const float4 nvx = add4(xgath<NN, 1>(sa, g.c), zgath<NN>(wd, g.c));
                const float4 nvz = add4(zgath<NN>(wb, g.c), xgath<NN, 0>(sc, g.c));
                if (col_ok && gz < g.nzp) {
                    float* P = a.planes + (size_t)s * g.plane + (size_t)gz * g.ld + gx;
                    st4(P + P_LVX * fp, nvx); st4(P + P_LVZ * fp, nvz);
                }
            }
        }
        fence_proxy_async();
        __syncthreads();
        if (tid == 0 && pc.valid) { k2_issue<NN>(pc, smem, bar, k, th, g.ns, a.lcur, thh, a.hist_len, a.tl, g.merge && a.tflags[pc.tile] != 1); pc.next(g, a.w, ring); }
        stage = (stage + 1 == NSTAGE) ? 0 : stage + 1;
    }
#pragma unroll
    for (int j = 0; j < RPT; ++j) {
        const int gz = gz0 + j;
        if (col_ok && gz < g.nzp) {
            float* gp = a.gpart + (size_t)chunk * 6 * g.plane + (size_t)gz * g.ld + gx;
            red4(gp, G11[j]); red4(gp + g.plane, G13[j]); red4(gp + 2 * g.plane, G33[j]); red4(gp + 3 * g.plane, G55[j]);
        }
    }
}

template <int NN, bool FS>
__global__ void __launch_bounds__(NTH, 2)
elf_k2(const __grid_constant__ CUtensorMap th, const __grid_constant__ CUtensorMap thh, const EGeom g, const K2Args a)
{
    using G = Geo<NN>;
    extern __shared__ __align__(128) unsigned char smem[];
    uint64_t* bar = (uint64_t*)(smem + NSTAGE * G::K2_STAGE);
    int* s_sz = (int*)(bar + 8);
    int* s_sx = s_sz + CMAX;
    const int tid = threadIdx.x;
    if (tid == 0) { for (int k = 0; k < NSTAGE; ++k) mbar_init(bar + k, 1); }
    __syncthreads();
    const Roles R(tid);
    uint32_t par = 0;
    int stage = 0;
    const int nitems = g.ntx * g.ntz * a.w.nchunks;
    int* ring = (int*)(bar + 4);
    Cursor pc;
    pc.wr = 0;
    pc.set(blockIdx.x, g, a.w);
    griddep_launch_dependents();
    unsigned rd = 0;
    bool first = true;
    for (int item = blockIdx.x; item < nitems; item = ring[rd++ & 3u], first = false) {
        const int tile = item / a.w.nchunks, chunk = item - tile * a.w.nchunks;
        const int s_lo = a.w.s_begin + chunk * a.w.chunk;
        const int s_hi = min(s_lo + a.w.chunk, a.w.s_end);
        if (a.tflags[tile] == 1) k2_tile<NN, true, FS>(&th, &thh, g, a, smem, bar, par, stage, pc, ring, s_sz, s_sx, R, tid, tile, chunk, s_lo, s_hi, first);
        else                k2_tile<NN, false, FS>(&th, &thh, g, a, smem, bar, par, stage, pc, ring, s_sz, s_sx, R, tid, tile, chunk, s_lo, s_hi, first);
        __syncthreads();
    }
}

// ==========================================================================================
// elf_b : fused reverse step = elf_k1 + elf_k2 in one launch (O(2,4); Appendix A.2, steps 10T..1T)
//
// The adjoint of the velocity update is evaluated on the tile plus a ring of 2NN cells (own-cell transposes) and
// NN cells (operator-transpose gathers), so the adjoint of the stress update finds the cotangents of the stress
// sums in shared memory: they never travel to HBM, and the cotangents of the velocity sums make one round trip
// per step instead of two.  Per shot two TMA groups arrive in SINGLE buffers that are refilled as soon as their
// consumer phase is over, so each load has half a shot of compute to hide behind:
//   V : 4 velocity-split cotangents + cotangents of the two velocity sums, halo 2NN  (free after the first gather)
//   S : 6 stress-split cotangents, halo NN                                          (free after the second gather)
// plus a scratch group M (3 rects: cotangents of the stress sums on tile + ring).  All split cotangents and the two
// sum cotangents are written to the other set of a ping-pong pair (neighbouring tiles still read the old ring).
// HBM traffic per cell-update, damping-free tiles: 4 + 3 staged planes read, 5 history planes, 2 + 3 + 2 written.
// ==========================================================================================
struct BArgs { ECoef cp; const unsigned char* tflags; float* planes; const float* hist; int hist_len, tl, it, lcur;
               int nr; RcvB rb; const float* g[5]; const float* mt; const int64_t *sx, *sz; float* g_src; float* gpart; Walk w; };

template <int NN> struct GeoB {
    using G = Geo<NN>;
    static constexpr int GX2 = G::HX2 / 4;                    // halo float4 groups on each side of a V rect
    static constexpr int W2 = NG + 2 * GX2;
    static constexpr int NRING2 = 4 * NN * W2 + TZ * 2 * GX2; // float4 groups of the 2NN ring
    static constexpr int V_BYTES = 6 * G::HB2, S_BYTES = 6 * G::HB, M_BYTES = 3 * G::HB;
    // per-tile coefficient tables of the ring cells (filled once per (tile, chunk), read every shot): bx, bz and the
    // scaled PML profiles dt/2*bcx, dt/2*bcz on the 2NN ring, C11, C13, C33, C55 on the NN ring (float4 each, SoA)
    static constexpr int T_BYTES = (4 * NRING2 + 4 * G::NRING) * 16;
    static constexpr int SMEM = V_BYTES + S_BYTES + M_BYTES + T_BYTES;
};
// index of the NN-ring group (r, gi) in the 2NN-ring enumeration of ring2_cell
template <int NN> __device__ __forceinline__ int ring2_index(int r, int gi)
{
    constexpr int GX2 = GeoB<NN>::GX2, W = GeoB<NN>::W2;
    if (r < 0) return (r + 2 * NN) * W + gi + GX2;
    if (r >= TZ) return 2 * NN * W + (r - TZ) * W + gi + GX2;
    return 4 * NN * W + r * (2 * GX2) + (gi < 0 ? gi + GX2 : GX2 + gi - NG);
}

// ring float4 group i in [0, NRING2): rows [-2NN,0) and [TZ,TZ+2NN) x groups [-GX2, NG+GX2), rows [0,TZ) x the side groups
template <int NN> __device__ __forceinline__ void ring2_cell(int i, int& r, int& gi)
{
    constexpr int GX2 = GeoB<NN>::GX2, W = GeoB<NN>::W2;
    if (i < 2 * NN * W) { r = -2 * NN + i / W; gi = i % W - GX2; }
    else if (i < 4 * NN * W) { const int ii = i - 2 * NN * W; r = TZ + ii / W; gi = ii % W - GX2; }
    else { const int ii = i - 4 * NN * W; r = ii / (2 * GX2); const int k = ii - r * (2 * GX2); gi = k < GX2 ? k - GX2 : NG + (k - GX2); }
}
// segment load at the edge of a staged rectangle: a missing neighbour group reads as zero (its values only reach
// ring cells that no consumer uses)
__device__ __forceinline__ void ldseg_e(const float* p, float* s, bool hl, bool hr)
{
    const float4 a = hl ? ld4(p - 4) : zero4(), b = ld4(p), c = hr ? ld4(p + 4) : zero4();
    s[0] = a.x; s[1] = a.y; s[2] = a.z; s[3] = a.w; s[4] = b.x; s[5] = b.y; s[6] = b.z; s[7] = b.w; s[8] = c.x; s[9] = c.y; s[10] = c.z; s[11] = c.w;
}

template <int NN> __device__ __forceinline__ void b_issue_v(const Cursor& c, unsigned char* smem, uint64_t* bar, const CUtensorMap* th2,
                                                            const CUtensorMap* thh, int ns, int lcur, int hist_len, int tl, bool lean)
{
    using G = Geo<NN>;
    // history planes of the shot into L2 (read with plain loads by the consumers): 0..3 for the stress part (lean: the
    // merged plane 2 replaces 2 and 3), 4..7 for the velocity part (lean: merged planes 4 and 6)
#pragma unroll
    for (int e = 0; e < 8; ++e)
        if (!lean || (e < 3) || e == 4 || e == 6) tma_prefetch_3d(thh, c.X0, c.Z0, (c.s * hist_len + tl) * NHIST + e);
    fence_proxy_async();
    mbar_expect_tx(bar, (lean ? 4 : 6) * G::HF2 * 4);
#pragma unroll
    for (int f = 0; f < 4; ++f)
        if (!lean || !(f & 1)) tma_load_3d(smem + f * G::HB2, th2, c.X0 - G::HX2, c.Z0 - 2 * NN, (P_LV + 4 * lcur + f) * ns + c.s, bar);
    tma_load_3d(smem + 4 * G::HB2, th2, c.X0 - G::HX2, c.Z0 - 2 * NN, (lcur ? P_LVX2 : P_LVX) * ns + c.s, bar);
    tma_load_3d(smem + 5 * G::HB2, th2, c.X0 - G::HX2, c.Z0 - 2 * NN, (lcur ? P_LVZ2 : P_LVZ) * ns + c.s, bar);
}
template <int NN> __device__ __forceinline__ void b_issue_s(const Cursor& c, unsigned char* sst, uint64_t* bar, const CUtensorMap* th,
                                                            int ns, int lcur, bool lean)
{
    using G = Geo<NN>;
    fence_proxy_async();
    mbar_expect_tx(bar, (lean ? 3 : 6) * G::HF * 4);
#pragma unroll
    for (int f = 0; f < 6; ++f)
        if (!lean || !(f & 1)) tma_load_3d(sst + f * G::HB, th, c.X0 - HX, c.Z0 - NN, (P_LS + 6 * lcur + f) * ns + c.s, bar);
}

template <int NN, bool PML, bool FS>
__device__ __forceinline__ void b_tile(const CUtensorMap* th, const CUtensorMap* th2, const CUtensorMap* thh, const EGeom& g, const BArgs& a,
                                       unsigned char* smem, uint64_t* bar, uint32_t& par, Cursor& pcv, Cursor& pcs, int* ring,
                                       int* s_sz, int* s_sx, const Roles& R, int tid, int tile, int chunk, int s_lo, int s_hi, bool first)
{
    using G = Geo<NN>;
    using B = GeoB<NN>;
    constexpr int RX2 = G::RX2, HX2 = G::HX2, HQ = G::HB / 4, HQ2 = G::HB2 / 4, GX2 = B::GX2;
    const uint64_t pol = l2_keep_policy();
    const int tzi = tile / g.ntx, txi = tile - tzi * g.ntx;
    const int X0 = txi * TX, Z0 = tzi * TZ;
    const int gx = X0 + R.c0, gz = Z0 + R.r0;
    const unsigned cm = col_mask<NN>(gx, g.nxp);
    const unsigned mt_ = row_in<NN>(gz, g.nzp) ? cm : 0u;          // update-region mask of the thread's own group
    const bool cell_ok = gx < g.ld && gz < g.nzp;
    const size_t fp = (size_t)g.ns * g.plane;
    const bool lean = !PML && g.merge;                 // see k1_tile
    const bool deep = lean && a.tflags[tile] == 2;
    float* V = (float*)smem;                            // rects 0..3: velocity-split cotangents (become m1..m4), 4: lvx, 5: lvz
    float* lvx = V + 4 * HQ2; float* lvz = V + 5 * HQ2;
    unsigned char* sst = smem + B::V_BYTES;
    float* LSr = (float*)sst;                           // 6 stress-split cotangents; rects 0, 1, 4, 5 become nA, nB, nC, nD
    float* mxx_r = (float*)(sst + B::S_BYTES); float* mzz_r = mxx_r + HQ; float* mxz_r = mxx_r + 2 * HQ;
    float4* T_bx = (float4*)(sst + B::S_BYTES + B::M_BYTES); float4* T_bz = T_bx + B::NRING2;
    float4* T_hx = T_bz + B::NRING2; float4* T_hz = T_hx + B::NRING2;
    float4* T_c11 = T_hz + B::NRING2; float4* T_c13 = T_c11 + G::NRING; float4* T_c33 = T_c13 + G::NRING; float4* T_c55 = T_c33 + G::NRING;
    const int hv = (R.r0 + 2 * NN) * RX2 + R.c0 + HX2, hs = (R.r0 + NN) * RXH + R.c0 + HX;
    if (a.g_src && tid < s_hi - s_lo) { s_sz[tid] = (int)a.sz[s_lo + tid]; s_sx[tid] = (int)a.sx[s_lo + tid]; }
    for (int i = tid; i < B::NRING2; i += NTH) {
        int r, gi;
        ring2_cell<NN>(i, r, gi);
        const ptrdiff_t o = (ptrdiff_t)(Z0 + r) * g.cpld + X0 + 4 * gi;
        T_bx[i] = ldk4(a.cp.bx + o, pol); T_bz[i] = ldk4(a.cp.bz + o, pol);
        if (PML) { T_hx[i] = smul(g.half_dt, ldk4(a.cp.bcx + o, pol)); T_hz[i] = smul(g.half_dt, ldk4(a.cp.bcz + o, pol)); }
    }
    for (int i = tid; i < G::NRING; i += NTH) {
        int r, gi;
        ring_cell<NN>(i, r, gi);
        const ptrdiff_t o = (ptrdiff_t)(Z0 + r) * g.cpld + X0 + 4 * gi;
        T_c11[i] = ldk4(a.cp.c11 + o, pol); T_c13[i] = ldk4(a.cp.c13 + o, pol); T_c33[i] = ldk4(a.cp.c33 + o, pol); T_c55[i] = ldk4(a.cp.c55 + o, pol);
    }
    const ptrdiff_t oc = (ptrdiff_t)gz * g.cpld + gx;
    const float4 C11 = ldk4(a.cp.c11 + oc, pol), C13 = ldk4(a.cp.c13 + oc, pol), C33 = ldk4(a.cp.c33 + oc, pol), C55 = ldk4(a.cp.c55 + oc, pol);
    const float4 BX = ldk4(a.cp.bx + oc, pol), BZ = ldk4(a.cp.bz + oc, pol);
    float4 PXN = one4(), PZN = one4(), PXI = one4(), PZI = one4();
    if (PML) {
        const float4 bx_ = ldk4(a.cp.bcx + oc, pol), bz_ = ldk4(a.cp.bcz + oc, pol);
        PXN = sub4(one4(), smul(g.half_dt, bx_)); PZN = sub4(one4(), smul(g.half_dt, bz_));
        PXI = div4(one4(), add4(one4(), smul(g.half_dt, bx_))); PZI = div4(one4(), add4(one4(), smul(g.half_dt, bz_)));
    }
    float4 G11 = zero4(), G13 = zero4(), G33 = zero4(), G55 = zero4(), GBX = zero4(), GBZ = zero4();
    const bool have_gv = a.nr > 0 && (a.g[3] || a.g[4]);
    const bool have_gs = a.nr > 0 && (a.g[0] || a.g[1] || a.g[2]);
    const bool inject_v = have_gv && a.rb.nbr[tile];
    const bool inject_s = have_gs && a.rb.nbr[tile];
    const bool producer = tid == NTH - 32;              // lane 0 of the last warp, which has no ring cells
    if (first) {
        griddep_wait();
        if (producer) {
            if (pcv.valid) { b_issue_v<NN>(pcv, smem, bar, th2, thh, g.ns, a.lcur, a.hist_len, a.tl, g.merge && a.tflags[pcv.tile] != 1); pcv.next(g, a.w, ring); }
            if (pcs.valid) { b_issue_s<NN>(pcs, sst, bar + 1, th, g.ns, a.lcur, g.merge && a.tflags[pcs.tile] != 1); pcs.next_follow(g, a.w, ring); }
        }
    }
    __syncthreads();

    for (int s = s_lo; s < s_hi; ++s) {
        const float* H = a.hist + ((size_t)s * a.hist_len + a.tl) * NHIST * g.plane + (size_t)gz * g.ld + gx;
        // history of the velocity update (own cell): e1 = D+x txx, e2 = D-z txz, e3 = D-x txz, e4 = D+z tzz (lean: e1+e2, e3+e4)
        float4 E0 = zero4(), E1 = zero4(), E2 = zero4(), E3 = zero4();
        if (cell_ok) {
            E0 = __ldcs(reinterpret_cast<const float4*>(H + 4 * g.plane)); E2 = __ldcs(reinterpret_cast<const float4*>(H + 6 * g.plane));
            if (!lean) { E1 = __ldcs(reinterpret_cast<const float4*>(H + 5 * g.plane)); E3 = __ldcs(reinterpret_cast<const float4*>(H + 7 * g.plane)); }
        }
        ELF_WAIT_STAGE(0);
        if (inject_v) {          // 10T: cotangents of the vx / vz records into the staged sums (duplicates legal)
            for (int dz = -1; dz <= 1; ++dz) {
                const int tz2 = tzi + dz;
                if (tz2 < 0 || tz2 >= g.ntz) continue;
                for (int dx = -1; dx <= 1; ++dx) {
                    const int tx2 = txi + dx;
                    if (tx2 < 0 || tx2 >= g.ntx) continue;
                    const int t2 = tz2 * g.ntx + tx2;
                    const int lo = a.rb.start[t2], hi = a.rb.start[t2 + 1];
                    for (int i = lo + tid; i < hi; i += NTH) {
                        const int zx = a.rb.zx[i];
                        const int z = (zx >> 16) - (Z0 - 2 * NN), x = (zx & 0xffff) - (X0 - HX2);
                        if (z >= 0 && z < G::RZ2 && x >= 0 && x < RX2) {
                            const size_t o = ((size_t)s * g.nt + a.it) * a.nr + a.rb.id[i];
                            if (a.g[3]) atomicAdd(lvx + z * RX2 + x, a.g[3][o]);
                            if (a.g[4]) atomicAdd(lvz + z * RX2 + x, a.g[4][o]);
                        }
                    }
                }
            }
            __syncthreads();
        }
        if (FS && tzi == 0) {    // 9T: transpose of the free-surface velocity edits (adds into rows h-1, h)
            const int t = tid;
            if (t >= 1 && t < RX2) {
                const int j = X0 + t - HX2;
                if (j >= NN && j < g.nxp - NN) {
                    const int rh1 = (3 * NN) * RX2 + t, rh = (3 * NN + 1) * RX2 + t, rh2 = (3 * NN - 1) * RX2 + t, rh3 = (3 * NN - 2) * RX2 + t;
                    const float qj = lvx[rh2];
                    const float qm = (j - 1 >= NN) ? lvx[rh2 - 1] : 0.f;
                    const float add_vz = lvz[rh2] + lvz[rh3] + 2.0f * (qm - qj);
                    lvz[rh1] += add_vz;
                    lvx[rh] += qj;
                }
            }
            __syncthreads();
        }
        // ---- phase A: own-cell transpose of the velocity update on tile + 2NN ring; m1..m4 replace the split cotangents ----
        float4 W1, W2 = zero4(), W3, W4 = zero4();
        {
            const float4 L0 = ld4(V + hv), L2 = ld4(V + 2 * HQ2 + hv);
            const float4 L1 = lean ? L0 : ld4(V + HQ2 + hv), L3 = lean ? L2 : ld4(V + 3 * HQ2 + hv);
            float4 w1, w2, w3, w4, m1, m2, m3, m4, N0, N1, N2, N3;
            k1_cell<PML>(g, mt_, L0, L1, L2, L3, ld4(lvx + hv), ld4(lvz + hv), BX, BZ, PXN, PXN, PZN, PZN, PXI, PZI,
                         w1, w2, w3, w4, m1, m2, m3, m4, N0, N1, N2, N3);
            st4(V + hv, m1); st4(V + HQ2 + hv, m2); st4(V + 2 * HQ2 + hv, m3); st4(V + 3 * HQ2 + hv, m4);
            if (cell_ok) {
                float* P = a.planes + (size_t)s * g.plane + (size_t)gz * g.ld + gx + (size_t)(P_LV + 4 * (a.lcur ^ 1)) * fp;
                st4(P, N0); st4(P + 2 * fp, N2);
                if (!deep) { st4(P + fp, N1); st4(P + 3 * fp, N3); }
            }
            W1 = w1; W3 = w3;
            if (!lean) { W2 = w2; W4 = w4; }
        }
        for (int i = tid; i < B::NRING2; i += NTH) {
            int r, gi;
            ring2_cell<NN>(i, r, gi);
            const int gzr = Z0 + r, gxr = X0 + 4 * gi;
            const int h2 = (r + 2 * NN) * RX2 + 4 * gi + HX2;
            const unsigned m = row_in<NN>(gzr, g.nzp) ? col_mask<NN>(gxr, g.nxp) : 0u;
            float4 pxn = one4(), pzn = one4(), rpxd = one4(), rpzd = one4();
            if (PML) {
                const float4 hx_ = T_hx[i], hz_ = T_hz[i];
                pxn = sub4(one4(), hx_); pzn = sub4(one4(), hz_);
                rpxd = rcp4(add4(one4(), hx_)); rpzd = rcp4(add4(one4(), hz_));
            }
            const float4 L0 = ld4(V + h2), L2 = ld4(V + 2 * HQ2 + h2);
            const float4 L1 = lean ? L0 : ld4(V + HQ2 + h2), L3 = lean ? L2 : ld4(V + 3 * HQ2 + h2);
            float4 w1, w2, w3, w4, m1, m2, m3, m4, N0, N1, N2, N3;
            k1_cell<PML>(g, m, L0, L1, L2, L3, ld4(lvx + h2), ld4(lvz + h2), T_bx[i], T_bz[i], pxn, pxn, pzn, pzn, rpxd, rpzd,
                         w1, w2, w3, w4, m1, m2, m3, m4, N0, N1, N2, N3);
            st4(V + h2, m1); st4(V + HQ2 + h2, m2); st4(V + 2 * HQ2 + h2, m3); st4(V + 3 * HQ2 + h2, m4);
        }
        // g_bx, g_bz (after the ring, so that the wait for the history loads does not hold up the ring cells)
        if (lean) {              // w1 == w2, w3 == w4 (dx == dz): the history holds e1 + e2 and e3 + e4
            GBX = add4(GBX, mul4(W1, E0));
            GBZ = add4(GBZ, mul4(W3, E2));
        } else {
            GBX = add4(GBX, add4(mul4(W1, E0), mul4(W2, E1)));
            GBZ = add4(GBZ, add4(mul4(W3, E2), mul4(W4, E3)));
        }
        __syncthreads();
        // history of the stress update (own cell): D-x vx, D-z vz, D+x vz, D+z vx (lean: the last two merged); in flight during phase B
        float4 D0 = zero4(), D1 = zero4(), D2 = zero4(), D3 = zero4();
        if (cell_ok) {
            D0 = __ldcs(reinterpret_cast<const float4*>(H)); D1 = __ldcs(reinterpret_cast<const float4*>(H + g.plane));
            D2 = __ldcs(reinterpret_cast<const float4*>(H + 2 * g.plane));
            if (!lean) D3 = __ldcs(reinterpret_cast<const float4*>(H + 3 * g.plane));
        }
        // ---- phase B: 6T gathers -> cotangents of the stress sums on tile + NN ring (shared memory only) ----
        {
            const float* M1 = V; const float* M2 = V + HQ2; const float* M3 = V + 2 * HQ2; const float* M4 = V + 3 * HQ2;
            for (int i = tid; i < NTH + G::NRING; i += NTH) {
                int r = R.r0, gi = R.l;
                if (i >= NTH) ring_cell<NN>(i - NTH, r, gi);
                const int h2 = (r + 2 * NN) * RX2 + 4 * gi + HX2, h1 = (r + NN) * RXH + 4 * gi + HX;
                float s1[12], s3[12];
                ldseg_e(M1 + h2, s1, gi - 1 >= -GX2, gi + 1 < NG + GX2); ldseg_e(M3 + h2, s3, gi - 1 >= -GX2, gi + 1 < NG + GX2);
                float4 w2[2 * NN], w4[2 * NN];
#pragma unroll
                for (int q = 0; q < 2 * NN; ++q) {
                    w2[q] = ld4(M2 + h2 + (q - NN + 1) * RX2);      // (D-z)^T: rows i-NN+1 .. i+NN
                    w4[q] = ld4(M4 + h2 + (q - NN) * RX2);          // (D+z)^T: rows i-NN .. i+NN-1
                }
                st4(mxx_r + h1, xgath<NN, 0>(s1, g.c));
                st4(mxz_r + h1, add4(zgath<NN>(w2, g.c), xgath<NN, 1>(s3, g.c)));
                st4(mzz_r + h1, zgath<NN>(w4, g.c));
            }
        }
        fence_proxy_async();
        __syncthreads();
        // group V has been consumed: refill it with the next shot while the stress part runs
        if (producer && pcv.valid) { b_issue_v<NN>(pcv, smem, bar, th2, thh, g.ns, a.lcur, a.hist_len, a.tl, g.merge && a.tflags[pcv.tile] != 1); pcv.next(g, a.w, ring); }
        if (inject_s) {          // 10T: cotangents of the stress records add to the sums' cotangents (tile + ring)
            for (int dz = -1; dz <= 1; ++dz) {
                const int tz2 = tzi + dz;
                if (tz2 < 0 || tz2 >= g.ntz) continue;
                for (int dx = -1; dx <= 1; ++dx) {
                    const int tx2 = txi + dx;
                    if (tx2 < 0 || tx2 >= g.ntx) continue;
                    const int t2 = tz2 * g.ntx + tx2;
                    const int lo = a.rb.start[t2], hi = a.rb.start[t2 + 1];
                    for (int i = lo + tid; i < hi; i += NTH) {
                        const int zx = a.rb.zx[i];
                        const int z = (zx >> 16) - (Z0 - NN), x = (zx & 0xffff) - (X0 - HX);
                        if (z >= 0 && z < G::RZH && x >= 0 && x < RXH) {
                            const size_t o = ((size_t)s * g.nt + a.it) * a.nr + a.rb.id[i];
                            if (a.g[0]) atomicAdd(mxx_r + z * RXH + x, a.g[0][o]);
                            if (a.g[1]) atomicAdd(mzz_r + z * RXH + x, a.g[1][o]);
                            if (a.g[2]) atomicAdd(mxz_r + z * RXH + x, a.g[2][o]);
                        }
                    }
                }
            }
            __syncthreads();
        }
        if (FS && tzi == 0) {    // 5T: transpose of the free-surface stress mirrors
            const int t = tid;
            if (t < RXH) {
                const int j = X0 + t - HX;
                if (j >= NN && j < g.nxp - NN) {
                    const int rh1 = (2 * NN) * RXH + t, rh = (2 * NN + 1) * RXH + t, rh2 = (2 * NN - 1) * RXH + t, rh3 = (2 * NN - 2) * RXH + t;
                    mxz_r[rh] -= mxz_r[rh3];
                    mzz_r[rh] -= mzz_r[rh2];
                    mxz_r[rh1] -= mxz_r[rh2];
                    mzz_r[rh1] = 0.f;
                }
            }
            __syncthreads();
        }
        ELF_WAIT_STAGE(1);
        // ---- phase C: own-cell transpose of the stress update on tile + NN ring; nA..nD replace four split cotangents ----
        {
            float4 LS[6], l[6], q[6], N[6], nA, nB, nC, nD;
#pragma unroll
            for (int f = 0; f < 6; f += 2) { LS[f] = ld4(LSr + f * HQ + hs); LS[f + 1] = lean ? LS[f] : ld4(LSr + (f + 1) * HQ + hs); }
            k2_cell<PML>(g, mt_, LS, ld4(mxx_r + hs), ld4(mzz_r + hs), ld4(mxz_r + hs), C11, C13, C33, C55, PXN, PXI, PZN, PZI, l, q, nA, nB, nC, nD, N);
            st4(LSr + hs, nA); st4(LSr + HQ + hs, nB); st4(LSr + 4 * HQ + hs, nC); st4(LSr + 5 * HQ + hs, nD);
            G11 = add4(G11, mul4(q[0], D0));
            G13 = add4(G13, add4(mul4(q[1], D1), mul4(q[2], D0)));
            G33 = add4(G33, mul4(q[3], D1));
            if (lean) G55 = add4(G55, mul4(q[4], D2));       // q4 == q5 (dt/dx == dt/dz): the history holds d3 + d4
            else      G55 = add4(G55, add4(mul4(q[4], D2), mul4(q[5], D3)));
            if (cell_ok) {
                float* P = a.planes + (size_t)s * g.plane + (size_t)gz * g.ld + gx + (size_t)(P_LS + 6 * (a.lcur ^ 1)) * fp;
#pragma unroll
                for (int f = 0; f < 6; ++f) if (!deep || !(f & 1)) st4(P + f * fp, N[f]);
                if (a.g_src && mt_ != 0u && s_sz[s - s_lo] == gz) {       // 3T
                    const int dc = s_sx[s - s_lo] - gx;
                    if (dc >= 0 && dc < 4 && ((mt_ >> dc) & 1u)) {
                        const float* Mt = a.mt + (size_t)s * 9;
                        a.g_src[(size_t)s * g.nt + a.it] = -(Mt[0] / 2.0f) * (comp4(l[0], dc) + comp4(l[1], dc))
                                                          - (Mt[8] / 2.0f) * (comp4(l[2], dc) + comp4(l[3], dc))
                                                          - (Mt[2] / 2.0f) * (comp4(l[4], dc) + comp4(l[5], dc));
                    }
                }
            }
        }
        for (int i = tid; i < G::NRING; i += NTH) {
            int r, gi;
            ring_cell<NN>(i, r, gi);
            const int gzr = Z0 + r, gxr = X0 + 4 * gi;
            const int h1 = (r + NN) * RXH + 4 * gi + HX;
            const unsigned m = row_in<NN>(gzr, g.nzp) ? col_mask<NN>(gxr, g.nxp) : 0u;
            float4 pxn = one4(), pxi = one4(), pzn = one4(), pzi = one4();
            if (PML) {
                const int i2 = ring2_index<NN>(r, gi);
                const float4 hx_ = T_hx[i2], hz_ = T_hz[i2];
                pxn = sub4(one4(), hx_); pzn = sub4(one4(), hz_);
                pxi = rcp4(add4(one4(), hx_)); pzi = rcp4(add4(one4(), hz_));
            }
            float4 LS[6], l[6], q[6], N[6], nA, nB, nC, nD;
#pragma unroll
            for (int f = 0; f < 6; f += 2) { LS[f] = ld4(LSr + f * HQ + h1); LS[f + 1] = lean ? LS[f] : ld4(LSr + (f + 1) * HQ + h1); }
            k2_cell<PML>(g, m, LS, ld4(mxx_r + h1), ld4(mzz_r + h1), ld4(mxz_r + h1), T_c11[i], T_c13[i], T_c33[i], T_c55[i],
                         pxn, pxi, pzn, pzi, l, q, nA, nB, nC, nD, N);
            st4(LSr + h1, nA); st4(LSr + HQ + h1, nB); st4(LSr + 4 * HQ + h1, nC); st4(LSr + 5 * HQ + h1, nD);
        }
        __syncthreads();
        // ---- phase D: 2T/1T gathers -> cotangents of the velocity sums (pre-step), to the other set ----
        {
            const float* NA = LSr; const float* NB = LSr + HQ; const float* NC = LSr + 4 * HQ; const float* ND = LSr + 5 * HQ;
            float sa[12], sc[12];
            ldseg(NA + hs, sa); ldseg(NC + hs, sc);
            float4 wb[2 * NN], wd[2 * NN];
#pragma unroll
            for (int q = 0; q < 2 * NN; ++q) {
                wb[q] = ld4(NB + hs + (q - NN + 1) * RXH);      // (D-z)^T
                wd[q] = ld4(ND + hs + (q - NN) * RXH);          // (D+z)^T
            }
            const float4 nvx = add4(xgath<NN, 1>(sa, g.c), zgath<NN>(wd, g.c));
            const float4 nvz = add4(zgath<NN>(wb, g.c), xgath<NN, 0>(sc, g.c));
            if (cell_ok) {
                float* P = a.planes + (size_t)s * g.plane + (size_t)gz * g.ld + gx;
                st4(P + (size_t)(a.lcur ? P_LVX : P_LVX2) * fp, nvx); st4(P + (size_t)(a.lcur ? P_LVZ : P_LVZ2) * fp, nvz);
            }
        }
        fence_proxy_async();
        // group S has been consumed: refill it with the next shot while the velocity part of that shot runs.  Only the
        // producer's warp waits for the other warps' phase D (named barrier 1); they go straight on to the next shot,
        // whose phases A and B touch neither S nor anything phase D reads
        if (tid >= NTH - 32) {
            asm volatile("bar.sync 1, %0;" ::"n"(NTH) : "memory");
            if (producer && pcs.valid) { b_issue_s<NN>(pcs, sst, bar + 1, th, g.ns, a.lcur, g.merge && a.tflags[pcs.tile] != 1); pcs.next_follow(g, a.w, ring); }
        } else {
            asm volatile("bar.arrive 1, %0;" ::"n"(NTH) : "memory");
        }
    }
    if (cell_ok) {
        float* gp = a.gpart + (size_t)chunk * 6 * g.plane + (size_t)gz * g.ld + gx;
        red4(gp, G11); red4(gp + g.plane, G13); red4(gp + 2 * g.plane, G33); red4(gp + 3 * g.plane, G55);
        red4(gp + 4 * g.plane, GBX); red4(gp + 5 * g.plane, GBZ);
    }
}

template <int NN, bool FS>
__global__ void __launch_bounds__(NTH, NN == 2 ? 2 : 1)      // O(2,6): one CTA per SM (shared memory) -> the full register file per thread
elf_b(const __grid_constant__ CUtensorMap th, const __grid_constant__ CUtensorMap th2, const __grid_constant__ CUtensorMap thh, const EGeom g, const BArgs a)
{
    static_assert(RPT == 1, "elf_b: one float4 group per thread");
    static_assert(Geo<NN>::HX2 <= CPX && 2 * NN <= CPZ, "elf_b: the 2NN ring must stay inside the apron of the coefficient pack");
    using B = GeoB<NN>;
    extern __shared__ __align__(128) unsigned char smem[];
    uint64_t* bar = (uint64_t*)(smem + B::SMEM);
    int* s_sz = (int*)(bar + 8);
    int* s_sx = s_sz + CMAX;
    const int tid = threadIdx.x;
    if (tid == 0) { mbar_init(bar, 1); mbar_init(bar + 1, 1); }
    __syncthreads();
    const Roles R(tid);
    uint32_t par = 0;
    const int nitems = g.ntx * g.ntz * a.w.nchunks;
    int* ring = (int*)(bar + 4);
    // the two producer cursors live in shared memory: only one thread uses them, registers are scarce
    Cursor& pcv = *(Cursor*)(s_sx + CMAX);
    Cursor& pcs = *((Cursor*)(s_sx + CMAX) + 1);
    if (tid == NTH - 32) {
        pcv.wr = 0; pcs.wr = 0;
        pcv.set(blockIdx.x, g, a.w);
        pcs.set(blockIdx.x, g, a.w);
    }
    griddep_launch_dependents();
    unsigned rd = 0;
    bool first = true;
    for (int item = blockIdx.x; item < nitems; item = ring[rd++ & 3u], first = false) {
        const int tile = item / a.w.nchunks, chunk = item - tile * a.w.nchunks;
        const int s_lo = a.w.s_begin + chunk * a.w.chunk;
        const int s_hi = min(s_lo + a.w.chunk, a.w.s_end);
        if (a.tflags[tile] == 1) b_tile<NN, true, FS>(&th, &th2, &thh, g, a, smem, bar, par, pcv, pcs, ring, s_sz, s_sx, R, tid, tile, chunk, s_lo, s_hi, first);
        else                b_tile<NN, false, FS>(&th, &th2, &thh, g, a, smem, bar, par, pcv, pcs, ring, s_sz, s_sx, R, tid, tile, chunk, s_lo, s_hi, first);
        __syncthreads();
    }
}

#include "elastic_abl_fused.inl"

// ---- set-up kernels ------------------------------------------------------------------------------
// coefficient pack: eight planes [cprows][cpld] (C11,C13,C33,C55,bx,bz,bcx,bcz), logical cell (z,x) at
// [(z+CPZ)*cpld + x+CPX], zero outside the grid
struct PackSrc { const float* p[8]; };
__global__ void elf_pack_coefs(int nzp, int nxp, int cprows, int cpld, size_t cpplane, PackSrc src, float* __restrict__ pack)
{
    const int xx = blockIdx.x * blockDim.x + threadIdx.x, zz = blockIdx.y;
    if (xx >= cpld || zz >= cprows) return;
    const int x = xx - CPX, z = zz - CPZ;
    const bool in = (z >= 0) && (z < nzp) && (x >= 0) && (x < nxp);
    const size_t c = in ? (size_t)z * nxp + x : 0;
    const size_t o = (size_t)zz * cpld + xx;
#pragma unroll
    for (int k = 0; k < 8; ++k) pack[k * cpplane + o] = (in && src.p[k]) ? src.p[k][c] : 0.f;
}
// tile flag = 1 when any cell of the tile's neighbourhood (tile + FLX / FLZ cells) has a non-zero PML profile
__global__ void elf_tile_flags(int ntx, int cpld, size_t cpplane, const float* __restrict__ pack, unsigned char* __restrict__ flags)
{
    const int tile = blockIdx.x;
    const int tzi = tile / ntx, txi = tile - tzi * ntx;
    const int X0 = txi * TX, Z0 = tzi * TZ;
    const int W = TX + 2 * FLX, H = TZ + 2 * FLZ;
    int bad = 0;
    for (int i = threadIdx.x; i < W * H; i += blockDim.x) {
        const size_t o = (size_t)(Z0 + CPZ - FLZ + i / W) * cpld + (X0 + CPX - FLX + i % W);
        if (pack[6 * cpplane + o] != 0.f || pack[7 * cpplane + o] != 0.f) bad = 1;
    }
    bad = __syncthreads_or(bad);
    if (threadIdx.x == 0) flags[tile] = (unsigned char)(bad ? 1 : 0);
}
// tile classes for the lean adjoint (EGeom::merge): on a damping-free cell both halves of every split pair go through
// identical arithmetic from identical (zero) initial values -- pmln = pmld = 1 exactly -- so their cotangents are bitwise
// equal for all time.  A tile of class 0 or 2 (no damping on tile + apron) stages only the first half of each pair.
//   1 = PML tile: full split treatment;
//   0 = damping-free tile next to a PML tile: reads first halves, writes both (the PML tile's halo reads them);
//   2 = damping-free tile whose 8 neighbours are damping-free too: reads and writes first halves only.
// In place on the 0/1 flags: a concurrent reader only tests "== 1".
__global__ void elf_tile_class(int ntx, int ntz, unsigned char* __restrict__ flags)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= ntx * ntz) return;
    const volatile unsigned char* f = flags;
    if (f[t] == 1) return;
    const int tz = t / ntx, tx = t - tz * ntx;
    int pml = 0;
    for (int dz = -1; dz <= 1; ++dz)
        for (int dx = -1; dx <= 1; ++dx) {
            const int z2 = tz + dz, x2 = tx + dx;
            if (z2 < 0 || z2 >= ntz || x2 < 0 || x2 >= ntx) continue;
            pml |= (f[z2 * ntx + x2] == 1);
        }
    flags[t] = (unsigned char)(pml ? 0 : 2);
}
// receiver buckets: counting sort of the receivers by tile
__device__ __forceinline__ int rcv_tile(int nzp, int nxp, int ntx, int64_t z, int64_t x)
{
    if (z < 0 || z >= nzp || x < 0 || x >= nxp) return -1;
    return (int)z / TZ * ntx + (int)x / TX;
}
__global__ void elf_rcv_count(int nzp, int nxp, int ntx, int nr, const int64_t* __restrict__ rx, const int64_t* __restrict__ rz, int* __restrict__ cnt)
{
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= nr) return;
    const int t = rcv_tile(nzp, nxp, ntx, rz[r], rx[r]);
    if (t >= 0) atomicAdd(cnt + t, 1);
}
__global__ void elf_rcv_scan(int ntiles, const int* __restrict__ cnt, int* __restrict__ start, int* __restrict__ cursor)
{
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        int acc = 0;
        for (int t = 0; t < ntiles; ++t) { start[t] = acc; cursor[t] = acc; acc += cnt[t]; }
        start[ntiles] = acc;
    }
}
__global__ void elf_rcv_fill(int nzp, int nxp, int ntx, int nr, const int64_t* __restrict__ rx, const int64_t* __restrict__ rz,
                             int* __restrict__ cursor, int* __restrict__ id, int* __restrict__ zx)
{
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= nr) return;
    const int t = rcv_tile(nzp, nxp, ntx, rz[r], rx[r]);
    if (t < 0) return;
    const int i = atomicAdd(cursor + t, 1);
    id[i] = r;
    zx[i] = ((int)rz[r] << 16) | (int)rx[r];
}
__global__ void elf_rcv_nbr(int ntx, int ntz, const int* __restrict__ start, unsigned char* __restrict__ nbr)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= ntx * ntz) return;
    const int tz = t / ntx, tx = t - tz * ntx;
    int any = 0;
    for (int dz = -1; dz <= 1; ++dz)
        for (int dx = -1; dx <= 1; ++dx) {
            const int z2 = tz + dz, x2 = tx + dx;
            if (z2 < 0 || z2 >= ntz || x2 < 0 || x2 >= ntx) continue;
            const int t2 = z2 * ntx + x2;
            any |= (start[t2 + 1] > start[t2]);
        }
    nbr[t] = (unsigned char)any;
}
// pitched partial planes -> dense caller planes
__global__ void elf_reduce_parts(int nzp, int nxp, int ld, size_t plane, int nparts, int k, const float* __restrict__ part, float* __restrict__ out)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x, z = blockIdx.y;
    if (x >= nxp || z >= nzp) return;
    float acc = 0.f;
    for (int p = 0; p < nparts; ++p) acc += part[((size_t)p * 6 + k) * plane + (size_t)z * ld + x];
    out[(size_t)z * nxp + x] = acc;
}

// squares of the five sum fields of the current state, summed over shots (:414-418), physical cells only
__global__ void elf_illum_acc(EGeom g, int NN, int nz, int nx, int nabc, int zoff, int sb, int se, int base, const float* __restrict__ planes, float* __restrict__ ill)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x, z = blockIdx.y;
    if (x >= nx || z >= nz) return;
    const int gz = z + zoff, gxx = x + nabc;
    const size_t c = (size_t)gz * g.ld + gxx, o = (size_t)z * nx + x, n = (size_t)nz * nx;
    const size_t fp = (size_t)g.ns * g.plane;
    float acc[5] = {0.f, 0.f, 0.f, 0.f, 0.f};
    for (int s = sb; s < se; ++s) {
        const float* P = planes + (size_t)s * g.plane + c + (size_t)base * fp;
        float txx = P[P_S0 * fp] + P[P_S1 * fp], tzz = P[P_S2 * fp] + P[P_S3 * fp], txz = P[P_S4 * fp] + P[P_S5 * fp];
        if (g.fs && gz == NN) tzz = 0.f;                        // tzz[h-1] = 0 (:380)
        const float vx = P[P_VXX * fp] + P[P_VXZ * fp], vz = P[P_VZX * fp] + P[P_VZZ * fp];
        acc[0] += txx * txx; acc[1] += tzz * tzz; acc[2] += txz * txz; acc[3] += vx * vx; acc[4] += vz * vz;
    }
    for (int k = 0; k < 5; ++k) ill[k * n + o] += acc[k];
}
__global__ void elf_illum_out(int n, const float* __restrict__ ill, float* o0, float* o1, float* o2, float* o3, float* o4)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float* out[5] = {o0, o1, o2, o3, o4};
    for (int k = 0; k < 5; ++k) if (out[k]) out[k][i] = ill[(size_t)k * n + i];
}

// ---- host side -------------------------------------------------------------------------------------
constexpr int CTAS_PER_SM = 2;

// SM count of the CURRENT device (cached per device ordinal: one process may drive several GPUs)
int elf_num_sms()
{
    static int cache[kMaxDevices] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= kMaxDevices) dev = 0;
    int n = cache[dev];
    if (!n) { cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev); if (n <= 0) n = 148; cache[dev] = n; }
    return n;
}

// lean adjoint on damping-free tiles (needs dx == dz, which the reference asserts: ADFWI/model/base.py:79);
// ADFWI_B200_EL_LEAN=0 keeps the full split treatment everywhere (A/B switch for tests and profiles)
inline bool elf_lean_adjoint()
{
    static int v = -1;
    if (v < 0) { const char* e = getenv("ADFWI_B200_EL_LEAN"); v = (e && e[0] == '0') ? 0 : 1; }
    return v != 0;
}

struct EFPlan {
    EGeom g;
    int ns, nr, NN, FS, save, n_segments, nz, nx, nabc, zoff;
    int K, nseg, nckpt, G;
    int chunk, nchunks;          // shots per (tile, chunk) item of the forward kernels
    int chunk_b, nchunks_b;      // ... of the reverse kernels (longer walks: the per-item gradient read-modify-write is amortised over them)
    int cprows; size_t cpplane;
    float* pack; unsigned char* tflags;
    float* planes; int nfields;
    float *hist, *ckpt, *gpart, *ill;
    int* counters; int ncounters;       // one work counter per launch of a call
    int *rcv_cnt, *rcv_start, *rcv_cursor, *rcv_id, *rcv_zx; unsigned char* rcv_nbr;
    size_t bytes;
};

// Shots per (tile, chunk) item.  A launch of G shots is cut into ntiles * ceil(G / chunk) items which the resident CTAs draw
// dynamically.  Cost model (shot-steps of one CTA; fitted to sweeps on the C3 and C4 grids, profiles/r02w_el_items.md):
//     T(chunk) = ntiles * (G + c0 * nchunks) / ncta  +  (chunk + c0)
// -- the work with a fixed cost c0 per item (coefficient tables, pipeline fill; the reverse kernels add the read-modify-write of
// six gradient partial planes: c0 = 1.5 against 0.5 for the forward kernels), plus a tail of about one item.  Ragged splits
// (15 = 4+4+4+3) pay c0 for the short item too, which is why exact divisors tend to win.
inline int elf_pick_chunk(int G, int ntiles, int ncta, float c0)
{
    int best = 1; float tbest = 3.0e38f;
    const int cmax = G < CMAX ? G : CMAX;
    for (int c = 1; c <= cmax; ++c) {
        const float t = (float)ntiles * ((float)G + c0 * (float)cdiv(G, c)) / (float)ncta + ((float)c + c0);
        if (t < tbest) { tbest = t; best = c; }
    }
    return best;
}
// desc->reserved[1] (both directions) and reserved[2] (reverse kernels only) override the model (tuning sweeps, tests)
inline void elf_plan_chunks(const adfwi_elastic_desc* d, int G, int ntiles, int nsm, int* chunk, int* nchunks, int* chunk_b, int* nchunks_b)
{
    const int ncta = CTAS_PER_SM * nsm;
    int cf = d->reserved[1] > 0 ? d->reserved[1] : elf_pick_chunk(G, ntiles, ncta, 0.5f);
    int cb = d->reserved[2] > 0 ? d->reserved[2] : d->reserved[1] > 0 ? d->reserved[1] : elf_pick_chunk(G, ntiles, ncta, 1.5f);
    cf = cf < 1 ? 1 : (cf > CMAX ? CMAX : cf); if (cf > G) cf = G;
    cb = cb < 1 ? 1 : (cb > CMAX ? CMAX : cb); if (cb > G) cb = G;
    *chunk = cf; *nchunks = cdiv(G, cf);
    *chunk_b = cb; *nchunks_b = cdiv(G, cb);
}

int elf_make_plan(const adfwi_elastic_desc* d, void* ws, EFPlan* P, int nsm)
{
    EGeom& g = P->g;
    const int NN = d->fd_order / 2;
    g.nzp = d->nzp; g.nxp = d->nxp; g.ld = (d->nxp + 31) / 32 * 32; g.fs = d->free_surface ? 1 : 0; g.nt = d->nt; g.ns = d->ns;
    g.ntx = cdiv(g.nxp, TX); g.ntz = cdiv(g.nzp, TZ);
    g.cpld = g.ntx * TX + 2 * CPX;
    g.plane = (size_t)g.nzp * g.ld;
    g.dt = d->dt; g.dx = d->dx; g.dz = d->dz; g.dt_dx = d->dt_dx; g.dt_dz = d->dt_dz; g.half_dt = d->half_dt;
    g.rdx = 1.0f / d->dx; g.rdz = 1.0f / d->dz;
    g.merge = (g.rdx == g.rdz && g.dt_dx == g.dt_dz && elf_lean_adjoint()) ? 1 : 0;
    for (int k = 0; k < 3; ++k) g.c[k] = d->fdc[k];
    P->cprows = g.ntz * TZ + 2 * CPZ;
    P->cpplane = align_up((size_t)P->cprows * g.cpld, 64);
    P->ns = d->ns; P->nr = d->nr; P->NN = NN; P->FS = g.fs; P->save = d->save_history ? 1 : 0;
    P->n_segments = d->n_segments > 0 ? d->n_segments : 1;
    P->nz = d->nz; P->nx = d->nx; P->nabc = d->nabc; P->zoff = d->free_surface ? NN : NN + d->nabc;
    int K = d->ckpt_interval;
    if (K <= 0 || K >= d->nt) K = d->nt;
    P->K = K; P->nseg = cdiv(d->nt, K); P->nckpt = P->nseg > 2 ? P->nseg - 2 : 0;
    int G = d->shots_per_group;
    if (G <= 0 || G > d->ns) G = d->ns;
    P->G = G;
    const int ntiles = g.ntx * g.ntz;
    elf_plan_chunks(d, G, ntiles, nsm, &P->chunk, &P->nchunks, &P->chunk_b, &P->nchunks_b);
    Carver cv(ws);
    const size_t sp = (size_t)d->ns * g.plane;
    P->pack = cv.take<float>(8 * P->cpplane);
    P->tflags = cv.take<unsigned char>(ntiles);
    P->nfields = P->save ? P_COUNT : P_FWD2_COUNT;
    P->planes = cv.take<float>((size_t)P->nfields * sp);
    P->ill = cv.take<float>((size_t)5 * d->nz * d->nx);
    P->rcv_cnt = cv.take<int>(ntiles + 1); P->rcv_start = cv.take<int>(ntiles + 1); P->rcv_cursor = cv.take<int>(ntiles + 1);
    P->rcv_id = cv.take<int>(d->nr > 0 ? d->nr : 1); P->rcv_zx = cv.take<int>(d->nr > 0 ? d->nr : 1);
    P->rcv_nbr = cv.take<unsigned char>(ntiles);
    P->ncounters = 4 * d->nt * cdiv(d->ns, G) + 64;
    P->counters = cv.take<int>(P->ncounters);
    P->hist = P->ckpt = P->gpart = nullptr;
    if (P->save) {
        P->gpart = cv.take<float>((size_t)P->nchunks_b * 6 * g.plane);
        if (P->nckpt) P->ckpt = cv.take<float>((size_t)P->nckpt * 10 * sp);
        P->hist = cv.take<float>((size_t)K * NHIST * sp);
    }
    P->bytes = cv.off;
    return ADFWI_OK;
}

ECoef elf_pack_ptrs(const EFPlan& P)
{
    const size_t o = (size_t)CPZ * P.g.cpld + CPX;
    ECoef c;
    c.c11 = P.pack + 0 * P.cpplane + o; c.c13 = P.pack + 1 * P.cpplane + o; c.c33 = P.pack + 2 * P.cpplane + o; c.c55 = P.pack + 3 * P.cpplane + o;
    c.bx = P.pack + 4 * P.cpplane + o; c.bz = P.pack + 5 * P.cpplane + o; c.bcx = P.pack + 6 * P.cpplane + o; c.bcz = P.pack + 7 * P.cpplane + o;
    return c;
}
RcvB elf_bucket_ptrs(const EFPlan& P) { RcvB b; b.start = P.rcv_start; b.id = P.rcv_id; b.zx = P.rcv_zx; b.nbr = P.rcv_nbr; return b; }

int elf_setup(const EFPlan& P, cudaStream_t st, const float* const* coef, const float* bcx, const float* bcz, const int64_t* rx, const int64_t* rz)
{
    const EGeom& g = P.g;
    PackSrc src;
    for (int k = 0; k < 6; ++k) src.p[k] = coef[k];
    src.p[6] = bcx; src.p[7] = bcz;
    elf_pack_coefs<<<dim3(cdiv(g.cpld, 128), P.cprows), 128, 0, st>>>(g.nzp, g.nxp, P.cprows, g.cpld, P.cpplane, src, P.pack);
    ADFWI_LAUNCH_CHECK();
    const int ntiles = g.ntx * g.ntz;
    elf_tile_flags<<<ntiles, 128, 0, st>>>(g.ntx, g.cpld, P.cpplane, P.pack, P.tflags);
    ADFWI_LAUNCH_CHECK();
    if (g.merge) {
        elf_tile_class<<<cdiv(ntiles, 128), 128, 0, st>>>(g.ntx, g.ntz, P.tflags);
        ADFWI_LAUNCH_CHECK();
    }
    ADFWI_CUDA(cudaMemsetAsync(P.rcv_cnt, 0, sizeof(int) * (ntiles + 1), st));
    if (P.nr > 0) {
        elf_rcv_count<<<cdiv(P.nr, 128), 128, 0, st>>>(g.nzp, g.nxp, g.ntx, P.nr, rx, rz, P.rcv_cnt);
        ADFWI_LAUNCH_CHECK();
    }
    elf_rcv_scan<<<1, 32, 0, st>>>(ntiles, P.rcv_cnt, P.rcv_start, P.rcv_cursor);
    ADFWI_LAUNCH_CHECK();
    if (P.nr > 0) {
        elf_rcv_fill<<<cdiv(P.nr, 128), 128, 0, st>>>(g.nzp, g.nxp, g.ntx, P.nr, rx, rz, P.rcv_cursor, P.rcv_id, P.rcv_zx);
        ADFWI_LAUNCH_CHECK();
    }
    elf_rcv_nbr<<<cdiv(ntiles, 128), 128, 0, st>>>(g.ntx, g.ntz, P.rcv_start, P.rcv_nbr);
    ADFWI_LAUNCH_CHECK();
    return ADFWI_OK;
}

struct EMaps { CUtensorMap halo, halo2, core, hist; };

int elf_make_maps(const EFPlan& P, EMaps* M)
{
    const EGeom& g = P.g;
    int rc = make_tmap_f32(&M->halo, P.planes, 3, g.nxp, g.ld, g.nzp, (uint64_t)P.nfields * P.ns, RXH, TZ + 2 * P.NN);
    if (rc) return rc;
    rc = make_tmap_f32(&M->core, P.planes, 3, g.nxp, g.ld, g.nzp, (uint64_t)P.nfields * P.ns, TX, TZ);
    if (rc) return rc;
    rc = make_tmap_f32(&M->halo2, P.planes, 3, g.nxp, g.ld, g.nzp, (uint64_t)P.nfields * P.ns, TX + 2 * (P.NN == 2 ? 4 : 8), TZ + 4 * P.NN);
    if (rc || !P.save) return rc;
    return make_tmap_f32(&M->hist, P.hist, 3, g.nxp, g.ld, g.nzp, (uint64_t)P.ns * P.K * NHIST, TX, TZ);
}

template <typename Kern, typename... Args>
cudaError_t elf_launch(Kern kern, int grid, int smem, cudaStream_t st, bool pdl, Args... args)
{
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid); cfg.blockDim = dim3(NTH); cfg.dynamicSmemBytes = (size_t)smem; cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = pdl ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kern, args...);
}
inline bool elf_use_pdl()
{
    static int v = -1;
    if (v < 0) { const char* e = getenv("ADFWI_B200_PDL"); v = (e && e[0] == '0') ? 0 : 1; }
    return v != 0;
}
template <typename K> int elf_set_smem(K kern, int bytes) { return (int)cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes); }

template <int NN> constexpr int s_smem() { return NSTAGE * Geo<NN>::S_STAGE + TAIL_BYTES; }
template <int NN> constexpr int v_smem() { return NSTAGE * Geo<NN>::V_STAGE + TAIL_SMALL; }
template <int NN> constexpr int f_smem() { return Geo<NN>::F_SMEM + TAIL_BYTES + (NTH / 32) * 16; }
template <int NN> constexpr int k1_smem() { return NSTAGE * Geo<NN>::K1_STAGE + TAIL_SMALL; }
template <int NN> constexpr int k2_smem() { return NSTAGE * Geo<NN>::K2_STAGE + TAIL_SMALL; }
template <int NN> constexpr int b_smem() { return GeoB<NN>::SMEM + TAIL_SMALL + 2 * (int)sizeof(Cursor); }
static_assert(2 * (s_smem<3>() + 1024) <= 233472 && 2 * (k2_smem<3>() + 1024) <= 233472 && 2 * (f_smem<2>() + 1024) <= 233472 && f_smem<3>() <= 232448 &&
              2 * (b_smem<2>() + 1024) <= 233472 && b_smem<3>() <= 232448,
              "two CTAs per SM must fit in shared memory (the O(2,6) fused forward and reverse kernels run one CTA per SM)");

// function attributes are per device: the >48 KB dynamic shared-memory opt-in is made once per device ordinal
template <int NN> int elf_init_kernels()
{
    static bool done_dev[kMaxDevices] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= kMaxDevices) dev = 0;
    bool& done = done_dev[dev];
    if (done) return 0;
    int rc = 0;
    rc |= elf_set_smem(elf_s<NN, true, true>, s_smem<NN>());   rc |= elf_set_smem(elf_s<NN, true, false>, s_smem<NN>());
    rc |= elf_set_smem(elf_s<NN, false, true>, s_smem<NN>());  rc |= elf_set_smem(elf_s<NN, false, false>, s_smem<NN>());
    rc |= elf_set_smem(elf_v<NN, true, true>, v_smem<NN>());   rc |= elf_set_smem(elf_v<NN, true, false>, v_smem<NN>());
    rc |= elf_set_smem(elf_v<NN, false, true>, v_smem<NN>());  rc |= elf_set_smem(elf_v<NN, false, false>, v_smem<NN>());
    rc |= elf_set_smem(elf_f<NN, true, true>, f_smem<NN>());   rc |= elf_set_smem(elf_f<NN, true, false>, f_smem<NN>());
    rc |= elf_set_smem(elf_f<NN, false, true>, f_smem<NN>());  rc |= elf_set_smem(elf_f<NN, false, false>, f_smem<NN>());
    rc |= elf_set_smem(elf_k1<NN, true>, k1_smem<NN>());       rc |= elf_set_smem(elf_k1<NN, false>, k1_smem<NN>());
    rc |= elf_set_smem(elf_k2<NN, true>, k2_smem<NN>());       rc |= elf_set_smem(elf_k2<NN, false>, k2_smem<NN>());
    rc |= elf_set_smem(elf_b<NN, true>, b_smem<NN>()); rc |= elf_set_smem(elf_b<NN, false>, b_smem<NN>());
    if (!rc) done = true;
    return rc;
}

inline Walk elf_walk(const EFPlan& P, int sb, int se, int* grid, bool reverse = false)
{
    const int chunk = reverse ? P.chunk_b : P.chunk;
    Walk w; w.s_begin = sb; w.s_end = se; w.chunk = chunk; w.nchunks = cdiv(se - sb, chunk); w.counter = nullptr;
    const int nitems = P.g.ntx * P.g.ntz * w.nchunks;
    const int cap = CTAS_PER_SM * elf_num_sms();
    *grid = nitems < cap ? nitems : cap;
    return w;
}

struct EArgs { const float *mt, *src_v; const int64_t *sx, *sz; };

// reverse step: the fused kernel elf_b for O(2,4), the two-launch form elf_k1 + elf_k2 for O(2,6); ADFWI_B200_EL_ADJ_SPLIT=1 / =0 force either
// O(2,6): the fused kernel fits one CTA per SM only (139 KB of shared memory) and measured 0.376 ms per launch against 0.334 ms for the
// pair on the C3 grid (profiles/r02e_o26_reverse.md), so the pair stays the default there; ADFWI_B200_EL_ADJ_SPLIT=0 selects elf_b<3>.
inline bool elf_split_adjoint(int NN)
{
    static int v = -1;
    if (v < 0) { const char* e = getenv("ADFWI_B200_EL_ADJ_SPLIT"); v = !e ? 2 : (e[0] == '1' ? 1 : 0); }
    return v == 2 ? NN == 3 : v == 1;
}

inline bool elf_split_forward()
{
    static int v = -1;
    if (v < 0) { const char* e = getenv("ADFWI_B200_EL_SPLIT"); v = (e && e[0] == '1') ? 1 : 0; }
    return v != 0;
}

// one forward step of shots [sb,se).  Default: the fused kernel elf_f (reads split-field set *cur, writes the other
// one and flips *cur).  ADFWI_B200_EL_SPLIT=1 selects the two-launch form elf_s + elf_v (in place on set 0).
template <int NN>
int elf_forward_step(const EFPlan& P, const EMaps& M, cudaStream_t st, int sb, int se, int it, bool save, int tl,
                     const EArgs& ea, float* const* rcv, int* seq, int* cur)
{
    if (*seq + 2 > P.ncounters) return ADFWI_E_DIMS;
    const EGeom& g = P.g;
    int grid;
    const Walk w = elf_walk(P, sb, se, &grid);
    const bool pdl = elf_use_pdl();
    if (!elf_split_forward()) {
        FArgs a;
        a.cp = elf_pack_ptrs(P); a.tflags = P.tflags; a.planes = P.planes; a.cur = *cur; a.mt = ea.mt; a.src_v = ea.src_v; a.sx = ea.sx; a.sz = ea.sz;
        a.hist = P.hist; a.hist_len = P.K; a.tl = tl; a.it = it;
        a.nr = rcv ? P.nr : 0; a.rb = elf_bucket_ptrs(P);
        for (int k = 0; k < 5; ++k) a.rcv[k] = rcv ? rcv[k] : nullptr;
        a.w = w; a.w.counter = P.counters + (*seq)++;
        {
            TimedLaunch tl_(KC_EL_FWD_FUSED, st);
            if (P.FS) { if (save) ADFWI_CUDA(elf_launch(elf_f<NN, true, true>, grid, f_smem<NN>(), st, pdl, M.halo, M.halo2, g, a));
                        else      ADFWI_CUDA(elf_launch(elf_f<NN, true, false>, grid, f_smem<NN>(), st, pdl, M.halo, M.halo2, g, a)); }
            else      { if (save) ADFWI_CUDA(elf_launch(elf_f<NN, false, true>, grid, f_smem<NN>(), st, pdl, M.halo, M.halo2, g, a));
                        else      ADFWI_CUDA(elf_launch(elf_f<NN, false, false>, grid, f_smem<NN>(), st, pdl, M.halo, M.halo2, g, a)); }
        }
        ADFWI_LAUNCH_CHECK();
        *cur ^= 1;
        return ADFWI_OK;
    }
    {
        SArgs a;
        a.cp = elf_pack_ptrs(P); a.tflags = P.tflags; a.planes = P.planes; a.mt = ea.mt; a.src_v = ea.src_v; a.sx = ea.sx; a.sz = ea.sz;
        a.hist = P.hist; a.hist_len = P.K; a.tl = tl; a.it = it; a.w = w; a.w.counter = P.counters + (*seq)++;
        TimedLaunch tl_(KC_EL_FWD_STRESS, st);
        if (P.FS) { if (save) ADFWI_CUDA(elf_launch(elf_s<NN, true, true>, grid, s_smem<NN>(), st, pdl, M.halo, M.core, g, a));
                    else      ADFWI_CUDA(elf_launch(elf_s<NN, true, false>, grid, s_smem<NN>(), st, pdl, M.halo, M.core, g, a)); }
        else      { if (save) ADFWI_CUDA(elf_launch(elf_s<NN, false, true>, grid, s_smem<NN>(), st, pdl, M.halo, M.core, g, a));
                    else      ADFWI_CUDA(elf_launch(elf_s<NN, false, false>, grid, s_smem<NN>(), st, pdl, M.halo, M.core, g, a)); }
    }
    ADFWI_LAUNCH_CHECK();
    {
        VArgs a;
        a.cp = elf_pack_ptrs(P); a.tflags = P.tflags; a.planes = P.planes; a.hist = P.hist; a.hist_len = P.K; a.tl = tl; a.it = it;
        a.nr = rcv ? P.nr : 0; a.rb = elf_bucket_ptrs(P);
        for (int k = 0; k < 5; ++k) a.rcv[k] = rcv ? rcv[k] : nullptr;
        a.w = w; a.w.counter = P.counters + (*seq)++;
        TimedLaunch tl_(KC_EL_FWD_VEL, st);
        if (P.FS) { if (save) ADFWI_CUDA(elf_launch(elf_v<NN, true, true>, grid, v_smem<NN>(), st, pdl, M.halo, M.core, g, a));
                    else      ADFWI_CUDA(elf_launch(elf_v<NN, true, false>, grid, v_smem<NN>(), st, pdl, M.halo, M.core, g, a)); }
        else      { if (save) ADFWI_CUDA(elf_launch(elf_v<NN, false, true>, grid, v_smem<NN>(), st, pdl, M.halo, M.core, g, a));
                    else      ADFWI_CUDA(elf_launch(elf_v<NN, false, false>, grid, v_smem<NN>(), st, pdl, M.halo, M.core, g, a)); }
    }
    ADFWI_LAUNCH_CHECK();
    return ADFWI_OK;
}

// the 10 persistent split fields of shots [sb,se) <-> checkpoint slot
int elf_copy_state(const EFPlan& P, cudaStream_t st, int sb, int se, float* ck, bool to_ckpt, int cur)
{
    const size_t sp = (size_t)P.ns * P.g.plane;
    const size_t off = (size_t)sb * P.g.plane, cnt = (size_t)(se - sb) * P.g.plane * sizeof(float);
    for (int f = 0; f < 10; ++f) {
        float* a = P.planes + (size_t)((cur ? P_B0 : 0) + f) * sp + off;
        float* b = ck + (size_t)f * sp + off;
        ADFWI_CUDA(cudaMemcpyAsync(to_ckpt ? b : a, to_ckpt ? a : b, cnt, cudaMemcpyDeviceToDevice, st));
    }
    return ADFWI_OK;
}
int elf_zero_fields(const EFPlan& P, cudaStream_t st, int f0, int f1, int sb, int se)
{
    const size_t sp = (size_t)P.ns * P.g.plane;
    if (sb == 0 && se == P.ns) return (int)cudaMemsetAsync(P.planes + (size_t)f0 * sp, 0, (size_t)(f1 - f0) * sp * sizeof(float), st);
    for (int f = f0; f < f1; ++f)
        ADFWI_CUDA(cudaMemsetAsync(P.planes + (size_t)f * sp + (size_t)sb * P.g.plane, 0, (size_t)(se - sb) * P.g.plane * sizeof(float), st));
    return ADFWI_OK;
}

template <int NN>
int elf_forward_t(const EFPlan& P, const EMaps& M, cudaStream_t st, const EArgs& ea, float* const* rcv, float* const* illum)
{
    const EGeom& g = P.g;
    const int nt = g.nt;
    const int csz = cdiv(nt, P.n_segments);
    const int nphys = P.nz * P.nx;
    if (illum) ADFWI_CUDA(cudaMemsetAsync(P.ill, 0, sizeof(float) * 5 * nphys, st));
    ADFWI_CUDA(cudaMemsetAsync(P.counters, 0, sizeof(int) * P.ncounters, st));
    int seq = 0;
    for (int sb = 0; sb < P.ns; sb += P.G) {
        const int se = sb + P.G < P.ns ? sb + P.G : P.ns;
        int rc = elf_zero_fields(P, st, 0, P_FWD2_COUNT, sb, se);
        if (rc) return rc;
        int cur = 0;
        for (int it = 0; it < nt; ++it) {
            const int seg = it / P.K, tl = it - seg * P.K;
            if (P.save && tl == 0 && seg >= 1 && seg <= P.nseg - 2) {
                rc = elf_copy_state(P, st, sb, se, P.ckpt + (size_t)(seg - 1) * 10 * P.ns * g.plane, true, cur);
                if (rc) return rc;
            }
            const bool save = P.save && seg == P.nseg - 1;
            rc = elf_forward_step<NN>(P, M, st, sb, se, it, save, tl, ea, P.nr > 0 ? rcv : nullptr, &seq, &cur);
            if (rc) return rc;
            if (illum && ((it + 1) % csz == 0 || it == nt - 1)) {
                elf_illum_acc<<<dim3(cdiv(P.nx, 128), P.nz), 128, 0, st>>>(g, NN, P.nz, P.nx, P.nabc, P.zoff, sb, se, cur ? P_B0 : 0, P.planes, P.ill);
                ADFWI_LAUNCH_CHECK();
            }
        }
    }
    if (illum) {
        elf_illum_out<<<cdiv(nphys, 256), 256, 0, st>>>(nphys, P.ill, illum[0], illum[1], illum[2], illum[3], illum[4]);
        ADFWI_LAUNCH_CHECK();
    }
    return ADFWI_OK;
}

template <int NN>
int elf_backward_t(const EFPlan& P, const EMaps& M, cudaStream_t st, const EArgs& ea, const float* const* g_rcv, float* const* g_coef, float* g_src)
{
    const EGeom& g = P.g;
    const int nt = g.nt;
    const bool pdl = elf_use_pdl();
    ADFWI_CUDA(cudaMemsetAsync(P.gpart, 0, sizeof(float) * (size_t)P.nchunks_b * 6 * g.plane, st));
    ADFWI_CUDA(cudaMemsetAsync(P.counters, 0, sizeof(int) * P.ncounters, st));
    int seq = 0;
    for (int sb = 0; sb < P.ns; sb += P.G) {
        const int se = sb + P.G < P.ns ? sb + P.G : P.ns;
        int rc = elf_zero_fields(P, st, P_LV, P_COUNT, sb, se);
        if (rc) return rc;
        int grid;
        const Walk w = elf_walk(P, sb, se, &grid, true);
        int lcur = 0;
        for (int seg = P.nseg - 1; seg >= 0; --seg) {
            const int t0 = seg * P.K, t1 = t0 + P.K < nt ? t0 + P.K : nt;
            if (seg != P.nseg - 1) {
                int cur = 0;
                if (seg == 0) rc = elf_zero_fields(P, st, 0, P_FWD2_COUNT, sb, se);
                else          rc = elf_copy_state(P, st, sb, se, P.ckpt + (size_t)(seg - 1) * 10 * P.ns * g.plane, false, 0);
                if (rc) return rc;
                for (int it = t0; it < t1; ++it) {
                    rc = elf_forward_step<NN>(P, M, st, sb, se, it, true, it - t0, ea, nullptr, &seq, &cur);
                    if (rc) return rc;
                }
            }
            for (int it = t1 - 1; it >= t0; --it) {
                if (!elf_split_adjoint(NN)) {
                    BArgs a;
                    a.cp = elf_pack_ptrs(P); a.tflags = P.tflags; a.planes = P.planes; a.hist = P.hist; a.hist_len = P.K; a.tl = it - t0; a.it = it; a.lcur = lcur;
                    a.nr = P.nr; a.rb = elf_bucket_ptrs(P);
                    for (int k = 0; k < 5; ++k) a.g[k] = g_rcv[k];
                    a.mt = ea.mt; a.sx = ea.sx; a.sz = ea.sz; a.g_src = g_src;
                    if (seq + 1 > P.ncounters) return ADFWI_E_DIMS;
                    a.gpart = P.gpart; a.w = w; a.w.counter = P.counters + seq++;
                    {
                        TimedLaunch tl_(KC_EL_ADJ_FUSED, st);
                        if (P.FS) ADFWI_CUDA(elf_launch(elf_b<NN, true>, grid, b_smem<NN>(), st, pdl, M.halo, M.halo2, M.hist, g, a));
                        else      ADFWI_CUDA(elf_launch(elf_b<NN, false>, grid, b_smem<NN>(), st, pdl, M.halo, M.halo2, M.hist, g, a));
                    }
                    ADFWI_LAUNCH_CHECK();
                    lcur ^= 1;
                    continue;
                }
                {
                    K1Args a;
                    a.cp = elf_pack_ptrs(P); a.tflags = P.tflags; a.planes = P.planes; a.hist = P.hist; a.hist_len = P.K; a.tl = it - t0; a.it = it; a.lcur = lcur;
                    a.nr = P.nr; a.rb = elf_bucket_ptrs(P);
                    for (int k = 0; k < 5; ++k) a.g[k] = g_rcv[k];
                    if (seq + 2 > P.ncounters) return ADFWI_E_DIMS;
                    a.gpart = P.gpart; a.w = w; a.w.counter = P.counters + seq++;
                    TimedLaunch tl_(KC_EL_ADJ_VEL, st);
                    if (P.FS) ADFWI_CUDA(elf_launch(elf_k1<NN, true>, grid, k1_smem<NN>(), st, pdl, M.halo, M.hist, g, a));
                    else      ADFWI_CUDA(elf_launch(elf_k1<NN, false>, grid, k1_smem<NN>(), st, pdl, M.halo, M.hist, g, a));
                }
                ADFWI_LAUNCH_CHECK();
                {
                    K2Args a;
                    a.cp = elf_pack_ptrs(P); a.tflags = P.tflags; a.planes = P.planes; a.hist = P.hist; a.hist_len = P.K; a.tl = it - t0; a.it = it; a.lcur = lcur;
                    a.mt = ea.mt; a.sx = ea.sx; a.sz = ea.sz; a.g_src = g_src; a.gpart = P.gpart; a.w = w; a.w.counter = P.counters + seq++;
                    TimedLaunch tl_(KC_EL_ADJ_STRESS, st);
                    if (P.FS) ADFWI_CUDA(elf_launch(elf_k2<NN, true>, grid, k2_smem<NN>(), st, pdl, M.halo, M.hist, g, a));
                    else      ADFWI_CUDA(elf_launch(elf_k2<NN, false>, grid, k2_smem<NN>(), st, pdl, M.halo, M.hist, g, a));
                }
                ADFWI_LAUNCH_CHECK();
                lcur ^= 1;
            }
        }
    }
    for (int k = 0; k < 6; ++k) {
        elf_reduce_parts<<<dim3(cdiv(g.nxp, 128), g.nzp), 128, 0, st>>>(g.nzp, g.nxp, g.ld, g.plane, P.nchunks_b, k, P.gpart, g_coef[k]);
        ADFWI_LAUNCH_CHECK();
    }
    return ADFWI_OK;
}


// ====================================================================================================
// sponge (ABL) pipeline: host side of ela_f / ela_b (elastic_abl_fused.inl)
// ====================================================================================================
struct EAPlan {
    EGeom g;
    int ns, nr, NN, FS, save, n_segments, nz, nx, nabc, zoff;
    int K, nseg, nckpt, G;
    int chunk, nchunks, chunk_b, nchunks_b;      // as in EFPlan
    int cprows; size_t cpplane;
    float* pack;
    float* planes; int nfields;
    float* side;                        // free-surface side buffer: [2 parities][2 rows][ns][ld]
    float *hist, *ckpt, *gpart, *ill;
    int* counters; int ncounters;
    int *rcv_cnt, *rcv_start, *rcv_cursor, *rcv_id, *rcv_zx; unsigned char* rcv_nbr;
    size_t ckpt_stride;                 // floats per checkpoint slot: 5 state planes + the side rows, all shots
    size_t bytes;
};

int ela_make_plan(const adfwi_elastic_desc* d, void* ws, EAPlan* P, int nsm)
{
    EGeom& g = P->g;
    const int NN = d->fd_order / 2;
    g.nzp = d->nzp; g.nxp = d->nxp; g.ld = (d->nxp + 31) / 32 * 32; g.fs = d->free_surface ? 1 : 0; g.nt = d->nt; g.ns = d->ns;
    g.ntx = cdiv(g.nxp, TX); g.ntz = cdiv(g.nzp, TZ);
    g.cpld = g.ntx * TX + 2 * CPX;
    g.plane = (size_t)g.nzp * g.ld;
    g.dt = d->dt; g.dx = d->dx; g.dz = d->dz; g.dt_dx = d->dt_dx; g.dt_dz = d->dt_dz; g.half_dt = d->half_dt;
    g.rdx = 1.0f / d->dx; g.rdz = 1.0f / d->dz;
    g.merge = 0;
    for (int k = 0; k < 3; ++k) g.c[k] = d->fdc[k];
    P->cprows = g.ntz * TZ + 2 * CPZ;
    P->cpplane = align_up((size_t)P->cprows * g.cpld, 64);
    P->ns = d->ns; P->nr = d->nr; P->NN = NN; P->FS = g.fs; P->save = d->save_history ? 1 : 0;
    P->n_segments = d->n_segments > 0 ? d->n_segments : 1;
    P->nz = d->nz; P->nx = d->nx; P->nabc = d->nabc; P->zoff = d->free_surface ? NN : NN + d->nabc;
    int K = d->ckpt_interval;
    if (K <= 0 || K >= d->nt) K = d->nt;
    P->K = K; P->nseg = cdiv(d->nt, K); P->nckpt = P->nseg > 2 ? P->nseg - 2 : 0;
    int G = d->shots_per_group;
    if (G <= 0 || G > d->ns) G = d->ns;
    P->G = G;
    const int ntiles = g.ntx * g.ntz;
    elf_plan_chunks(d, G, ntiles, nsm, &P->chunk, &P->nchunks, &P->chunk_b, &P->nchunks_b);
    Carver cv(ws);
    const size_t sp = (size_t)d->ns * g.plane;
    P->pack = cv.take<float>(8 * P->cpplane);
    P->nfields = P->save ? A_COUNT : A_FWD_COUNT;
    P->planes = cv.take<float>((size_t)P->nfields * sp);
    P->side = cv.take<float>((size_t)4 * d->ns * g.ld);
    P->ill = cv.take<float>((size_t)5 * d->nz * d->nx);
    P->rcv_cnt = cv.take<int>(ntiles + 1); P->rcv_start = cv.take<int>(ntiles + 1); P->rcv_cursor = cv.take<int>(ntiles + 1);
    P->rcv_id = cv.take<int>(d->nr > 0 ? d->nr : 1); P->rcv_zx = cv.take<int>(d->nr > 0 ? d->nr : 1);
    P->rcv_nbr = cv.take<unsigned char>(ntiles);
    P->ncounters = 3 * d->nt * cdiv(d->ns, G) + 64;
    P->counters = cv.take<int>(P->ncounters);
    P->hist = P->ckpt = P->gpart = nullptr;
    P->ckpt_stride = 5 * sp + (size_t)2 * d->ns * g.ld;
    if (P->save) {
        P->gpart = cv.take<float>((size_t)P->nchunks_b * 6 * g.plane);
        if (P->nckpt) P->ckpt = cv.take<float>((size_t)P->nckpt * P->ckpt_stride);
        P->hist = cv.take<float>((size_t)K * ANHIST * sp);
    }
    P->bytes = cv.off;
    return ADFWI_OK;
}

ECoef ela_pack_ptrs(const EAPlan& P)
{
    const size_t o = (size_t)CPZ * P.g.cpld + CPX;
    ECoef c;
    c.c11 = P.pack + 0 * P.cpplane + o; c.c13 = P.pack + 1 * P.cpplane + o; c.c33 = P.pack + 2 * P.cpplane + o; c.c55 = P.pack + 3 * P.cpplane + o;
    c.bx = P.pack + 4 * P.cpplane + o; c.bz = P.pack + 5 * P.cpplane + o; c.bcx = P.pack + 6 * P.cpplane + o; c.bcz = P.pack + 7 * P.cpplane + o;
    return c;
}
RcvB ela_bucket_ptrs(const EAPlan& P) { RcvB b; b.start = P.rcv_start; b.id = P.rcv_id; b.zx = P.rcv_zx; b.nbr = P.rcv_nbr; return b; }

int ela_setup(const EAPlan& P, cudaStream_t st, const float* const* coef, const float* damp, const int64_t* rx, const int64_t* rz)
{
    const EGeom& g = P.g;
    PackSrc src;
    for (int k = 0; k < 6; ++k) src.p[k] = coef[k];
    src.p[6] = damp; src.p[7] = nullptr;
    elf_pack_coefs<<<dim3(cdiv(g.cpld, 128), P.cprows), 128, 0, st>>>(g.nzp, g.nxp, P.cprows, g.cpld, P.cpplane, src, P.pack);
    ADFWI_LAUNCH_CHECK();
    const int ntiles = g.ntx * g.ntz;
    ADFWI_CUDA(cudaMemsetAsync(P.rcv_cnt, 0, sizeof(int) * (ntiles + 1), st));
    if (P.nr > 0) {
        elf_rcv_count<<<cdiv(P.nr, 128), 128, 0, st>>>(g.nzp, g.nxp, g.ntx, P.nr, rx, rz, P.rcv_cnt);
        ADFWI_LAUNCH_CHECK();
    }
    elf_rcv_scan<<<1, 32, 0, st>>>(ntiles, P.rcv_cnt, P.rcv_start, P.rcv_cursor);
    ADFWI_LAUNCH_CHECK();
    if (P.nr > 0) {
        elf_rcv_fill<<<cdiv(P.nr, 128), 128, 0, st>>>(g.nzp, g.nxp, g.ntx, P.nr, rx, rz, P.rcv_cursor, P.rcv_id, P.rcv_zx);
        ADFWI_LAUNCH_CHECK();
    }
    elf_rcv_nbr<<<cdiv(ntiles, 128), 128, 0, st>>>(g.ntx, g.ntz, P.rcv_start, P.rcv_nbr);
    ADFWI_LAUNCH_CHECK();
    return ADFWI_OK;
}

struct EAMaps { CUtensorMap halo, halo2, hist; };

int ela_make_maps(const EAPlan& P, EAMaps* M)
{
    const EGeom& g = P.g;
    int rc = make_tmap_f32(&M->halo, P.planes, 3, g.nxp, g.ld, g.nzp, (uint64_t)P.nfields * P.ns, RXH, TZ + 2 * P.NN);
    if (rc) return rc;
    rc = make_tmap_f32(&M->halo2, P.planes, 3, g.nxp, g.ld, g.nzp, (uint64_t)P.nfields * P.ns, TX + 2 * (P.NN == 2 ? 4 : 8), TZ + 4 * P.NN);
    if (rc || !P.save) return rc;
    return make_tmap_f32(&M->hist, P.hist, 3, g.nxp, g.ld, g.nzp, (uint64_t)P.ns * P.K * ANHIST, TX, TZ);
}

template <int NN> constexpr int af_smem() { return NSTAGE * GeoA<NN>::STAGE + TAIL_BYTES + GeoA<NN>::TF_BYTES; }
template <int NN> constexpr int ab_smem() { return NSTAGE * GeoA<NN>::STAGE + TAIL_SMALL + GeoA<NN>::FS_BYTES + 64; }
static_assert(2 * (af_smem<3>() + 1024) <= 233472 && 2 * (ab_smem<3>() + 1024) <= 233472, "two CTAs per SM must fit in shared memory");

template <int NN> int ela_init_kernels()
{
    static bool done_dev[kMaxDevices] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= kMaxDevices) dev = 0;
    bool& done = done_dev[dev];
    if (done) return 0;
    int rc = 0;
    rc |= elf_set_smem(ela_f<NN, true, true>, af_smem<NN>());   rc |= elf_set_smem(ela_f<NN, true, false>, af_smem<NN>());
    rc |= elf_set_smem(ela_f<NN, false, true>, af_smem<NN>());  rc |= elf_set_smem(ela_f<NN, false, false>, af_smem<NN>());
    rc |= elf_set_smem(ela_b<NN, true>, ab_smem<NN>());         rc |= elf_set_smem(ela_b<NN, false>, ab_smem<NN>());
    if (!rc) done = true;
    return rc;
}

inline Walk ela_walk(const EAPlan& P, int sb, int se, int* grid, bool reverse = false)
{
    const int chunk = reverse ? P.chunk_b : P.chunk;
    Walk w; w.s_begin = sb; w.s_end = se; w.chunk = chunk; w.nchunks = cdiv(se - sb, chunk); w.counter = nullptr;
    const int nitems = P.g.ntx * P.g.ntz * w.nchunks;
    const int cap = CTAS_PER_SM * elf_num_sms();
    *grid = nitems < cap ? nitems : cap;
    return w;
}

// one forward step of shots [sb,se): reads field set *cur (and side-buffer parity *cur), writes the other one, flips *cur
template <int NN>
int ela_forward_step(const EAPlan& P, const EAMaps& M, cudaStream_t st, int sb, int se, int it, bool save, int tl,
                     const EArgs& ea, float* const* rcv, int* seq, int* cur)
{
    if (*seq + 1 > P.ncounters) return ADFWI_E_DIMS;
    const EGeom& g = P.g;
    int grid;
    AFArgs a;
    a.w = ela_walk(P, sb, se, &grid);
    a.w.counter = P.counters + (*seq)++;
    a.cp = ela_pack_ptrs(P); a.planes = P.planes; a.cur = *cur; a.mt = ea.mt; a.src_v = ea.src_v; a.sx = ea.sx; a.sz = ea.sz;
    a.hist = P.hist; a.hist_len = P.K; a.tl = tl; a.it = it;
    a.nr = rcv ? P.nr : 0; a.rb = ela_bucket_ptrs(P);
    for (int k = 0; k < 5; ++k) a.rcv[k] = rcv ? rcv[k] : nullptr;
    a.side = P.side;
    const bool pdl = elf_use_pdl();
    {
        TimedLaunch tl_(KC_EL_FWD_FUSED, st);
        if (P.FS) { if (save) ADFWI_CUDA(elf_launch(ela_f<NN, true, true>, grid, af_smem<NN>(), st, pdl, M.halo, M.halo2, g, a));
                    else      ADFWI_CUDA(elf_launch(ela_f<NN, true, false>, grid, af_smem<NN>(), st, pdl, M.halo, M.halo2, g, a)); }
        else      { if (save) ADFWI_CUDA(elf_launch(ela_f<NN, false, true>, grid, af_smem<NN>(), st, pdl, M.halo, M.halo2, g, a));
                    else      ADFWI_CUDA(elf_launch(ela_f<NN, false, false>, grid, af_smem<NN>(), st, pdl, M.halo, M.halo2, g, a)); }
    }
    ADFWI_LAUNCH_CHECK();
    *cur ^= 1;
    return ADFWI_OK;
}

// the five state planes (set `cur`) and the two side rows (parity `cur`) of shots [sb,se) <-> checkpoint slot
int ela_copy_state(const EAPlan& P, cudaStream_t st, int sb, int se, float* ck, bool to_ckpt, int cur)
{
    const size_t sp = (size_t)P.ns * P.g.plane;
    const size_t off = (size_t)sb * P.g.plane, cnt = (size_t)(se - sb) * P.g.plane * sizeof(float);
    for (int f = 0; f < 5; ++f) {
        float* a = P.planes + (size_t)((cur ? A_SET : 0) + f) * sp + off;
        float* b = ck + (size_t)f * sp + off;
        ADFWI_CUDA(cudaMemcpyAsync(to_ckpt ? b : a, to_ckpt ? a : b, cnt, cudaMemcpyDeviceToDevice, st));
    }
    const size_t srow = (size_t)P.ns * P.g.ld, soff = (size_t)sb * P.g.ld, scnt = (size_t)(se - sb) * P.g.ld * sizeof(float);
    for (int r = 0; r < 2; ++r) {
        float* a = P.side + (size_t)((cur ? 2 : 0) + r) * srow + soff;
        float* b = ck + 5 * sp + (size_t)r * srow + soff;
        ADFWI_CUDA(cudaMemcpyAsync(to_ckpt ? b : a, to_ckpt ? a : b, scnt, cudaMemcpyDeviceToDevice, st));
    }
    return ADFWI_OK;
}
int ela_zero_fields(const EAPlan& P, cudaStream_t st, int f0, int f1, int sb, int se)
{
    const size_t sp = (size_t)P.ns * P.g.plane;
    if (sb == 0 && se == P.ns) return (int)cudaMemsetAsync(P.planes + (size_t)f0 * sp, 0, (size_t)(f1 - f0) * sp * sizeof(float), st);
    for (int f = f0; f < f1; ++f)
        ADFWI_CUDA(cudaMemsetAsync(P.planes + (size_t)f * sp + (size_t)sb * P.g.plane, 0, (size_t)(se - sb) * P.g.plane * sizeof(float), st));
    return ADFWI_OK;
}
int ela_zero_state(const EAPlan& P, cudaStream_t st, int sb, int se)
{
    int rc = ela_zero_fields(P, st, 0, A_FWD_COUNT, sb, se);
    if (rc) return rc;
    const size_t srow = (size_t)P.ns * P.g.ld;
    for (int r = 0; r < 4; ++r)
        ADFWI_CUDA(cudaMemsetAsync(P.side + (size_t)r * srow + (size_t)sb * P.g.ld, 0, (size_t)(se - sb) * P.g.ld * sizeof(float), st));
    return ADFWI_OK;
}

// squares of the five fields of the current state, summed over shots (elastic_kernels.py:769-774), physical cells only
__global__ void ela_illum_acc(EGeom g, int nz, int nx, int nabc, int zoff, int sb, int se, int base, const float* __restrict__ planes, float* __restrict__ ill)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x, z = blockIdx.y;
    if (x >= nx || z >= nz) return;
    const size_t c = (size_t)(z + zoff) * g.ld + (x + nabc), o = (size_t)z * nx + x, n = (size_t)nz * nx;
    const size_t fp = (size_t)g.ns * g.plane;
    const int order[5] = {A_TXX, A_TZZ, A_TXZ, A_VX, A_VZ};
    for (int k = 0; k < 5; ++k) {
        float acc = 0.f;
        for (int s = sb; s < se; ++s) { const float v = planes[(size_t)(base + order[k]) * fp + (size_t)s * g.plane + c]; acc += v * v; }
        ill[k * n + o] += acc;
    }
}

template <int NN>
int ela_forward_t(const EAPlan& P, const EAMaps& M, cudaStream_t st, const EArgs& ea, float* const* rcv, float* const* illum)
{
    const EGeom& g = P.g;
    const int nt = g.nt;
    const int csz = cdiv(nt, P.n_segments);
    const int nphys = P.nz * P.nx;
    if (illum) ADFWI_CUDA(cudaMemsetAsync(P.ill, 0, sizeof(float) * 5 * nphys, st));
    ADFWI_CUDA(cudaMemsetAsync(P.counters, 0, sizeof(int) * P.ncounters, st));
    int seq = 0;
    for (int sb = 0; sb < P.ns; sb += P.G) {
        const int se = sb + P.G < P.ns ? sb + P.G : P.ns;
        int rc = ela_zero_state(P, st, sb, se);
        if (rc) return rc;
        int cur = 0;
        for (int it = 0; it < nt; ++it) {
            const int seg = it / P.K, tl = it - seg * P.K;
            if (P.save && tl == 0 && seg >= 1 && seg <= P.nseg - 2) {
                rc = ela_copy_state(P, st, sb, se, P.ckpt + (size_t)(seg - 1) * P.ckpt_stride, true, cur);
                if (rc) return rc;
            }
            const bool save = P.save && seg == P.nseg - 1;
            rc = ela_forward_step<NN>(P, M, st, sb, se, it, save, tl, ea, P.nr > 0 ? rcv : nullptr, &seq, &cur);
            if (rc) return rc;
            if (illum && ((it + 1) % csz == 0 || it == nt - 1)) {
                ela_illum_acc<<<dim3(cdiv(P.nx, 128), P.nz), 128, 0, st>>>(g, P.nz, P.nx, P.nabc, P.zoff, sb, se, cur ? A_SET : 0, P.planes, P.ill);
                ADFWI_LAUNCH_CHECK();
            }
        }
    }
    if (illum) {
        elf_illum_out<<<cdiv(nphys, 256), 256, 0, st>>>(nphys, P.ill, illum[0], illum[1], illum[2], illum[3], illum[4]);
        ADFWI_LAUNCH_CHECK();
    }
    return ADFWI_OK;
}

template <int NN>
int ela_backward_t(const EAPlan& P, const EAMaps& M, cudaStream_t st, const EArgs& ea, const float* const* g_rcv, float* const* g_coef, float* g_src)
{
    const EGeom& g = P.g;
    const int nt = g.nt;
    const bool pdl = elf_use_pdl();
    ADFWI_CUDA(cudaMemsetAsync(P.gpart, 0, sizeof(float) * (size_t)P.nchunks_b * 6 * g.plane, st));
    ADFWI_CUDA(cudaMemsetAsync(P.counters, 0, sizeof(int) * P.ncounters, st));
    int seq = 0;
    for (int sb = 0; sb < P.ns; sb += P.G) {
        const int se = sb + P.G < P.ns ? sb + P.G : P.ns;
        int rc = ela_zero_fields(P, st, A_L0, A_COUNT, sb, se);
        if (rc) return rc;
        int lcur = 0;
        for (int seg = P.nseg - 1; seg >= 0; --seg) {
            const int t0 = seg * P.K, t1 = t0 + P.K < nt ? t0 + P.K : nt;
            if (seg != P.nseg - 1) {
                int cur = 0;
                if (seg == 0) rc = ela_zero_state(P, st, sb, se);
                else          rc = ela_copy_state(P, st, sb, se, P.ckpt + (size_t)(seg - 1) * P.ckpt_stride, false, 0);
                if (rc) return rc;
                for (int it = t0; it < t1; ++it) {
                    rc = ela_forward_step<NN>(P, M, st, sb, se, it, true, it - t0, ea, nullptr, &seq, &cur);
                    if (rc) return rc;
                }
            }
            for (int it = t1 - 1; it >= t0; --it) {
                if (seq + 1 > P.ncounters) return ADFWI_E_DIMS;
                int grid;
                ABArgs a;
                a.w = ela_walk(P, sb, se, &grid, true);
                a.w.counter = P.counters + seq++;
                a.cp = ela_pack_ptrs(P); a.planes = P.planes; a.hist = P.hist; a.hist_len = P.K; a.tl = it - t0; a.it = it; a.lcur = lcur;
                a.nr = P.nr; a.rb = ela_bucket_ptrs(P);
                for (int k = 0; k < 5; ++k) a.g[k] = g_rcv[k];
                a.mt = ea.mt; a.sx = ea.sx; a.sz = ea.sz; a.g_src = g_src; a.gpart = P.gpart;
                {
                    TimedLaunch tl_(KC_EL_ADJ_FUSED, st);
                    if (P.FS) ADFWI_CUDA(elf_launch(ela_b<NN, true>, grid, ab_smem<NN>(), st, pdl, M.halo, M.halo2, M.hist, g, a));
                    else      ADFWI_CUDA(elf_launch(ela_b<NN, false>, grid, ab_smem<NN>(), st, pdl, M.halo, M.halo2, M.hist, g, a));
                }
                ADFWI_LAUNCH_CHECK();
                lcur ^= 1;
            }
        }
    }
    for (int k = 0; k < 6; ++k) {
        elf_reduce_parts<<<dim3(cdiv(g.nxp, 128), g.nzp), 128, 0, st>>>(g.nzp, g.nxp, g.ld, g.plane, P.nchunks_b, k, P.gpart, g_coef[k]);
        ADFWI_LAUNCH_CHECK();
    }
    return ADFWI_OK;
}

}  // namespace

bool elf_supported(const adfwi_elastic_desc* d)
{
    if (!d || !d->abc_pml) return false;
    if (d->fd_order != 4 && d->fd_order != 6) return false;
    if (d->reserved[0] & 1) return false;
    if (d->nzp >= 32768 || d->nxp >= 65536) return false;      // receiver cells are packed as (z<<16)|x
    if ((uint64_t)P_COUNT * (uint64_t)d->ns >= (1ull << 31)) return false;
    return true;
}

size_t elf_workspace_bytes(const adfwi_elastic_desc* d)
{
    EFPlan P;
    elf_make_plan(d, nullptr, &P, 148);
    return P.bytes;
}

int elf_forward(const adfwi_elastic_desc* d, const float* const* coef, const float* bcx, const float* bcz, const float* mt,
                const float* src_v, const int64_t* sx, const int64_t* sz, const int64_t* rx, const int64_t* rz,
                float* const* rcv, float* const* illum, void* ws, cudaStream_t st)
{
    EFPlan P;
    elf_make_plan(d, ws, &P, 148);
    int rc = P.NN == 2 ? elf_init_kernels<2>() : elf_init_kernels<3>();
    if (rc) return rc;
    EMaps M;
    rc = elf_make_maps(P, &M);
    if (rc) return rc;
    rc = elf_setup(P, st, coef, bcx, bcz, rx, rz);
    if (rc) return rc;
    EArgs ea; ea.mt = mt; ea.src_v = src_v; ea.sx = sx; ea.sz = sz;
    return P.NN == 2 ? elf_forward_t<2>(P, M, st, ea, rcv, illum) : elf_forward_t<3>(P, M, st, ea, rcv, illum);
}

int elf_backward(const adfwi_elastic_desc* d, const float* const* coef, const float* bcx, const float* bcz, const float* mt,
                 const float* src_v, const int64_t* sx, const int64_t* sz, const int64_t* rx, const int64_t* rz,
                 const float* const* g_rcv, float* const* g_coef, float* g_src, void* ws, cudaStream_t st)
{
    (void)coef; (void)bcx; (void)bcz; (void)rx; (void)rz;      // pack, tile flags and receiver buckets were left in the workspace by forward
    EFPlan P;
    elf_make_plan(d, ws, &P, 148);
    int rc = P.NN == 2 ? elf_init_kernels<2>() : elf_init_kernels<3>();
    if (rc) return rc;
    EMaps M;
    rc = elf_make_maps(P, &M);
    if (rc) return rc;
    EArgs ea; ea.mt = mt; ea.src_v = src_v; ea.sx = sx; ea.sz = sz;
    return P.NN == 2 ? elf_backward_t<2>(P, M, st, ea, g_rcv, g_coef, g_src) : elf_backward_t<3>(P, M, st, ea, g_rcv, g_coef, g_src);
}


// ---- sponge (ABL) pipeline entry points ------------------------------------------------------------
bool ela_supported(const adfwi_elastic_desc* d)
{
    if (!d || d->abc_pml) return false;
    if (d->fd_order != 4 && d->fd_order != 6) return false;
    if (d->reserved[0] & 1) return false;
    if (d->nzp >= 32768 || d->nxp >= 65536) return false;      // receiver cells are packed as (z<<16)|x
    if ((uint64_t)A_COUNT * (uint64_t)d->ns >= (1ull << 31)) return false;
    return true;
}

size_t ela_workspace_bytes(const adfwi_elastic_desc* d)
{
    EAPlan P;
    ela_make_plan(d, nullptr, &P, 148);
    return P.bytes;
}

int ela_forward(const adfwi_elastic_desc* d, const float* const* coef, const float* damp, const float* mt,
                const float* src_v, const int64_t* sx, const int64_t* sz, const int64_t* rx, const int64_t* rz,
                float* const* rcv, float* const* illum, void* ws, cudaStream_t st)
{
    EAPlan P;
    ela_make_plan(d, ws, &P, 148);
    int rc = P.NN == 2 ? ela_init_kernels<2>() : ela_init_kernels<3>();
    if (rc) return rc;
    EAMaps M;
    rc = ela_make_maps(P, &M);
    if (rc) return rc;
    rc = ela_setup(P, st, coef, damp, rx, rz);
    if (rc) return rc;
    EArgs ea; ea.mt = mt; ea.src_v = src_v; ea.sx = sx; ea.sz = sz;
    return P.NN == 2 ? ela_forward_t<2>(P, M, st, ea, rcv, illum) : ela_forward_t<3>(P, M, st, ea, rcv, illum);
}

int ela_backward(const adfwi_elastic_desc* d, const float* mt, const float* src_v, const int64_t* sx, const int64_t* sz,
                 const float* const* g_rcv, float* const* g_coef, float* g_src, void* ws, cudaStream_t st)
{
    EAPlan P;                       // pack and receiver buckets were left in the workspace by forward
    ela_make_plan(d, ws, &P, 148);
    int rc = P.NN == 2 ? ela_init_kernels<2>() : ela_init_kernels<3>();
    if (rc) return rc;
    EAMaps M;
    rc = ela_make_maps(P, &M);
    if (rc) return rc;
    EArgs ea; ea.mt = mt; ea.src_v = src_v; ea.sx = sx; ea.sz = sz;
    return P.NN == 2 ? ela_backward_t<2>(P, M, st, ea, g_rcv, g_coef, g_src) : ela_backward_t<3>(P, M, st, ea, g_rcv, g_coef, g_src);
}

}  // namespace adfwi
#endif  // !ADFWI_HOST_EMUL
