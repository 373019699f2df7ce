// acoustic_persist.inl -- cluster-persistent iso-acoustic time loop for SMALL grids (included inside the anonymous namespace of
// acoustic_fused.cu; shares its coefficient pack, workspace plan and history layout).
//
// On small grids (BASELINE.json configs[0]: 148 x 260 padded cells, 40 shots) one time step is ~1.5 M cell-updates: a launch per step
// is bound by launch latency and by the TMA-load -> compute -> store round trip of each (tile, shot) item, not by HBM (the whole
// state sits in L2).  Here the state never leaves the chip:
//
//   * one thread-block CLUSTER of NC CTAs owns one shot for ALL nt time steps; CTA k owns a strip of R consecutive rows of the
//     active region (rows >= zlo).  p, u, w of the strip (plus 3 halo rows on each side) and the five coefficient planes of the
//     strip live in SHARED MEMORY for the whole time loop; shots beyond the co-resident clusters run as further waves;
//   * per time step two cluster barriers (barrier.cluster arrive.release / wait.acquire): after the pressure update each CTA PULLS
//     the pressure rows it needs from its neighbours' shared memory (distributed shared memory, mapa + ld.shared::cluster), after the
//     velocity update the w rows;
//   * HBM sees only what has to leave: the stencil history (4 B per cell-step, streaming stores) and the receiver samples going out,
//     the source samples coming in; the adjoint kernel reads the history back (prefetched one step ahead), takes the receiver
//     cotangents in, and keeps the gradient of alpha1 in REGISTERS across the whole reverse loop (one reduction per shot).
//
// Same arithmetic and association as fwd_tile / adj_tile (-fmad=false): records bit-identical to the per-step kernels and to the
// CPU reference.  Used when the active region fits (pp_supported); grids beyond that keep the per-step TMA kernels.

constexpr int PNT = 512;                 // threads per CTA
constexpr int PHALO = 3;                 // halo rows kept on each side of a strip (w needs 2 above / 1 below, p 1 above / 2 below)
constexpr int PXPAD = 4;                 // zero columns left and right of a staged row

struct PGeom {
    int nzp, nxp, ld, fs, zlo, nt, nabc;
    int NC, R, nxg, pitch;               // CTAs per cluster, rows per strip, float4 groups per row, floats per staged row
    int cpld;
    size_t plane;
    float c1, c2, dt;
};
struct PFwdArgs {
    CoefPack cp;
    const float* src_v; const int64_t *sx, *sz;
    float* hist; int hist_len;
    int nr; const int64_t *rx, *rz;
    float *rcv_p, *rcv_u, *rcv_w;
    float *ill_p, *ill_u, *ill_w; int illum, last_chunk_start;
    int s_begin, save;
};
struct PAdjArgs {
    CoefPack cp;
    const int64_t *sx, *sz;
    const float* hist; int hist_len;
    int nr; const int64_t *rx, *rz;
    const float *gp, *gu, *gw;
    float* g1part; float* g_src;
    int s_begin;
};

__device__ __forceinline__ unsigned cluster_rank() { unsigned r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ unsigned cluster_id_x() { unsigned r; asm volatile("mov.u32 %0, %%clusterid.x;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync_all()
{
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// 128-bit load from the same shared-memory offset of another CTA of the cluster
__device__ __forceinline__ float4 ld_dsmem4(const float* local_ptr, unsigned rank)
{
    uint32_t ra;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(smem_u32(local_ptr)), "r"(rank));
    float4 v;
    asm volatile("ld.shared::cluster.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(ra) : "memory");
    return v;
}
// rows [lo, hi) (local row numbers, may be negative) of a staged field <- the neighbour strips that own them
__device__ __forceinline__ void pull_halo(float* f, const PGeom& g, unsigned k, int lo_above, int hi_below, int tid)
{
    // above: local rows [-lo_above, 0) = rows [R - lo_above, R) of CTA k-1; below: local rows [R, R + hi_below) = rows [0, hi_below) of CTA k+1
    const int n_above = (k > 0) ? lo_above * g.nxg : 0;
    const int n_below = (k + 1 < (unsigned)g.NC) ? hi_below * g.nxg : 0;
    for (int i = tid; i < n_above + n_below; i += PNT) {
        int l, xg; unsigned src; int lsrc;
        if (i < n_above) { const int q = i / g.nxg; xg = i - q * g.nxg; l = -lo_above + q; src = k - 1; lsrc = g.R + l; }
        else { const int j = i - n_above; const int q = j / g.nxg; xg = j - q * g.nxg; l = g.R + q; src = k + 1; lsrc = q; }
        const int o_dst = (l + PHALO) * g.pitch + PXPAD + 4 * xg, o_src = (lsrc + PHALO) * g.pitch + PXPAD + 4 * xg;
        st4(f + o_dst, ld_dsmem4(f + o_src, src));
    }
}

// shared-memory carve-up of one CTA: 3 state planes with halo rows, 5 coefficient planes (own rows): a1 (masked alpha1), t1 = 1 - kappa1,
// a2 (alpha2 on the union of the U and W regions), t2 = 1 - kappa2, t3 = 1 - kappa3 (kappas masked to their regions)
struct PSmem {
    float *f0, *f1, *f2, *f3, *a1, *t1, *a2, *t2, *t3;
};
__host__ __device__ inline size_t pp_smem_floats(int R, int pitch, int nstate_halo, int nstate_own)
{
    return (size_t)(nstate_halo * (R + 2 * PHALO) + (nstate_own + 5) * R) * pitch;
}

template <bool FS>
__device__ __forceinline__ void pp_load_coefs(const PGeom& g, const CoefPack& cp, const PSmem& S, int z0, int rows_own, int tid)
{
    // own rows, every column group (columns beyond nxp read the pack's zero apron -> alpha = 0, 1 - kappa = 1)
    for (int i = tid; i < g.R * g.nxg * 4; i += PNT) {
        const int l = i / (g.nxg * 4), x = i - l * (g.nxg * 4);
        const int o = l * g.pitch + PXPAD + x;
        float a1 = 0.f, k1 = 0.f, a2 = 0.f, k2 = 0.f, k3 = 0.f;
        if (l < rows_own) {
            const ptrdiff_t c = (ptrdiff_t)(z0 + l) * g.cpld + x;
            a1 = cp.a1[c]; k1 = cp.k1[c]; k2 = cp.k2[c]; k3 = cp.k3[c];
            const float au = cp.a2u[c], aw = cp.a2w[c];
            a2 = au != 0.f ? au : aw;
        }
        S.a1[o] = a1; S.t1[o] = 1.0f - k1; S.a2[o] = a2; S.t2[o] = 1.0f - k2; S.t3[o] = 1.0f - k3;
    }
}
// per-thread work list: float4 group q of this thread = (local row l, column group xg); offsets precomputed once
struct PItem {                              // o: offset in a haloed state plane (< 0 = no work), hz: (z0+l)*ld + 4*xg, lg: (l << 16) | 4*xg
    int o, hz, lg;
    __device__ __forceinline__ int l() const { return lg >> 16; }
    __device__ __forceinline__ int gx() const { return lg & 0xffff; }
};
__device__ __forceinline__ PItem pp_item(const PGeom& g, int gi, int ngroups, int z0)
{
    PItem it; it.o = -1; it.hz = 0; it.lg = 0;
    if (gi < ngroups) {
        const int l = gi / g.nxg, xg = gi - l * g.nxg;
        it.lg = (l << 16) | (4 * xg);
        it.o = (l + PHALO) * g.pitch + PXPAD + 4 * xg; it.hz = (z0 + l) * g.ld + 4 * xg;
    }
    return it;
}
// The staged alpha2 plane holds alpha2 on the UNION of the U and W update regions (acoustic_kernels.py:139,151).  The two regions
// differ only on row nzp-2 (U only) and on column nxp-2 (W only):
__device__ __forceinline__ void pp_split_a2(const PGeom& g, int gz, int gx, const float4& a2, float4& au, float4& aw)
{
    au = a2; aw = a2;
    if (gz == g.nzp - 2) aw = make_float4(0.f, 0.f, 0.f, 0.f);
    const int dc = g.nxp - 2 - gx;
    if (dc == 0) au.x = 0.f; else if (dc == 1) au.y = 0.f; else if (dc == 2) au.z = 0.f; else if (dc == 3) au.w = 0.f;
}
__device__ __forceinline__ void pp_addc(float4& a, int i, float v)
{ if (i == 0) a.x = a.x + v; else if (i == 1) a.y = a.y + v; else if (i == 2) a.z = a.z + v; else a.w = a.w + v; }

// ------------------------------------------------------------------------------------------------------------------
// forward: all nt steps of one shot (acoustic_kernels.py:113-174)
// ------------------------------------------------------------------------------------------------------------------
template <bool FS, int KMAX>
__global__ void __launch_bounds__(PNT, 1)
acp_fwd(const PGeom g, const PFwdArgs a)
{
    extern __shared__ __align__(16) float psm[];
    const int tid = threadIdx.x;
    const unsigned k = cluster_rank();
    const int s = a.s_begin + (int)cluster_id_x();
    const int rows_s = g.R + 2 * PHALO;
    const size_t sp = (size_t)rows_s * g.pitch, cpn = (size_t)g.R * g.pitch;
    float* p = psm; float* u = p + sp; float* w = u + sp;
    PSmem S; S.a1 = w + sp; S.t1 = S.a1 + cpn; S.a2 = S.t1 + cpn; S.t2 = S.a2 + cpn; S.t3 = S.t2 + cpn;
    const int z0 = g.zlo + (int)k * g.R;
    const int rows_own = max(0, min(g.R, g.nzp - z0));
    const int ngroups = rows_own * g.nxg;
    const float c1 = g.c1, c2 = g.c2;
    const int pitch = g.pitch;
    for (int i = tid; i < 3 * (int)sp; i += PNT) psm[i] = 0.f;
    pp_load_coefs<FS>(g, a.cp, S, z0, rows_own, tid);
    const int szs = (int)a.sz[s], sxs = (int)a.sx[s];
    const bool fs_top = FS && k == 0;             // strip 0 starts at row fs-1: local rows 0, 1, 2 = fs-1, fs, fs+1
    PItem W[KMAX];
    float4 accp[KMAX], accu[KMAX];
    int srcq = -1, srcc = 0;                      // which of this thread's groups holds the source cell, and its component
#pragma unroll
    for (int q = 0; q < KMAX; ++q) {
        W[q] = pp_item(g, tid + q * PNT, ngroups, z0);
        accp[q] = make_float4(0.f, 0.f, 0.f, 0.f); accu[q] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (W[q].o >= 0 && z0 + W[q].l() == szs && sxs >= W[q].gx() && sxs < W[q].gx() + 4 && sxs < g.nxp) { srcq = q; srcc = sxs - W[q].gx(); }
    }
    // receivers of this strip handled by this thread: r = tid, tid + PNT, ...  (staged offset, -1 = not in this strip)
    constexpr int RMAX = 4;
    int roff[RMAX];
#pragma unroll
    for (int j = 0; j < RMAX; ++j) {
        const int r = tid + j * PNT;
        roff[j] = -1;
        if (r < a.nr) {
            const int rz = (int)a.rz[r], rx = (int)a.rx[r];
            if (rz >= z0 && rz < z0 + rows_own && rx >= 0 && rx < g.nxp) roff[j] = (rz - z0 + PHALO) * pitch + PXPAD + rx;
        }
    }
    float* hist_s = a.save ? a.hist + (size_t)s * a.hist_len * g.plane : nullptr;
    __syncthreads();
    cluster_sync_all();

    float src_next = srcq >= 0 ? g.dt * a.src_v[(size_t)s * g.nt] : 0.f;
    for (int it = 0; it < g.nt; ++it) {
        const float src_cur = src_next;
        if (srcq >= 0 && it + 1 < g.nt) src_next = g.dt * a.src_v[(size_t)s * g.nt + it + 1];
        // ---- pressure update (:115-128) + source (:131-132) + free-surface mirror (:135-136), stencil history ---------------
#pragma unroll
        for (int q = 0; q < KMAX; ++q) {
            if (W[q].o >= 0) {
                const int o = W[q].o, oc = o - PHALO * pitch;
                const float* ur = u + o;
                const float2 ul = ld2(ur - 2);
                const float4 um = ld4(ur);
                const float uR = ur[4];
                const float4 w0 = ld4(w + o), wm1 = ld4(w + o - pitch), wp1 = ld4(w + o + pitch), wm2 = ld4(w + o - 2 * pitch);
                // row fs-1 of the first strip is written by the owner of row fs+1 (mirror): its own owner must not touch it
                const bool mirrored = fs_top && W[q].l() == 0;
                const float4 po = mirrored ? make_float4(0.f, 0.f, 0.f, 0.f) : ld4(p + o), A1 = ld4(S.a1 + oc), T1 = ld4(S.t1 + oc);
                float4 Sv, pv;
                Sv.x = c1 * (((um.x - ul.y) + w0.x) - wm1.x) + c2 * (((um.y - ul.x) + wp1.x) - wm2.x);
                Sv.y = c1 * (((um.y - um.x) + w0.y) - wm1.y) + c2 * (((um.z - ul.y) + wp1.y) - wm2.y);
                Sv.z = c1 * (((um.z - um.y) + w0.z) - wm1.z) + c2 * (((um.w - um.x) + wp1.z) - wm2.z);
                Sv.w = c1 * (((um.w - um.z) + w0.w) - wm1.w) + c2 * (((uR - um.y) + wp1.w) - wm2.w);
                if (hist_s) __stcs(reinterpret_cast<float4*>(hist_s + (size_t)it * g.plane + W[q].hz), Sv);
                pv.x = T1.x * po.x - A1.x * Sv.x; pv.y = T1.y * po.y - A1.y * Sv.y;
                pv.z = T1.z * po.z - A1.z * Sv.z; pv.w = T1.w * po.w - A1.w * Sv.w;
                if (q == srcq) pp_addc(pv, srcc, src_cur);
                // the owner of row fs+1 also writes the mirrored row fs-1; the owner of row fs-1 leaves it alone
                if (fs_top && W[q].l() == 2) st4(p + o - 2 * pitch, make_float4(-pv.x, -pv.y, -pv.z, -pv.w));
                if (!mirrored) st4(p + o, pv);
            }
        }
        cluster_sync_all();
        pull_halo(p, g, k, 1, 2, tid);
        __syncthreads();
        // ---- velocity updates (:139-160) + free surface (:163-164), illumination ----------------------------------------------
        // (row fs-1 of the first strip lies outside the U and W regions; its w is written by the owner of row fs)
#pragma unroll
        for (int q = 0; q < KMAX; ++q) {
            if (W[q].o >= 0 && !(fs_top && W[q].l() == 0)) {
                const int o = W[q].o, oc = o - PHALO * pitch;
                const float* pr = p + o;
                const float pL = pr[-1];
                const float2 pR = ld2(pr + 4);
                const float4 p0 = ld4(pr), pm1 = ld4(pr - pitch), pp1 = ld4(pr + pitch), pp2 = ld4(pr + 2 * pitch);
                const float4 uo = ld4(u + o), wo = ld4(w + o);
                const float4 T2 = ld4(S.t2 + oc), T3 = ld4(S.t3 + oc);
                float4 au, aw;
                pp_split_a2(g, z0 + W[q].l(), W[q].gx(), ld4(S.a2 + oc), au, aw);
                float4 du, dw, uv, wv;
                du.x = au.x * (c1 * (p0.y - p0.x) + c2 * (p0.z - pL));
                du.y = au.y * (c1 * (p0.z - p0.y) + c2 * (p0.w - p0.x));
                du.z = au.z * (c1 * (p0.w - p0.z) + c2 * (pR.x - p0.y));
                du.w = au.w * (c1 * (pR.x - p0.w) + c2 * (pR.y - p0.z));
                dw.x = aw.x * (c1 * (pp1.x - p0.x) + c2 * (pp2.x - pm1.x));
                dw.y = aw.y * (c1 * (pp1.y - p0.y) + c2 * (pp2.y - pm1.y));
                dw.z = aw.z * (c1 * (pp1.z - p0.z) + c2 * (pp2.z - pm1.z));
                dw.w = aw.w * (c1 * (pp1.w - p0.w) + c2 * (pp2.w - pm1.w));
                uv.x = T2.x * uo.x - du.x; uv.y = T2.y * uo.y - du.y; uv.z = T2.z * uo.z - du.z; uv.w = T2.w * uo.w - du.w;
                wv.x = T3.x * wo.x - dw.x; wv.y = T3.y * wo.y - dw.y; wv.z = T3.z * wo.z - dw.z; wv.w = T3.w * wo.w - dw.w;
                st4(u + o, uv);
                // w[fs-1] = w[fs]: the owner of row fs also writes row fs-1
                if (fs_top && W[q].l() == 1) st4(w + o - pitch, wv);
                st4(w + o, wv);
                if (a.illum) {
                    accp[q].x += p0.x * p0.x; accp[q].y += p0.y * p0.y; accp[q].z += p0.z * p0.z; accp[q].w += p0.w * p0.w;
                    if (it >= a.last_chunk_start) { accu[q].x += uv.x * uv.x; accu[q].y += uv.y * uv.y; accu[q].z += uv.z * uv.z; accu[q].w += uv.w * uv.w; }
                }
            }
        }
        cluster_sync_all();
        // ---- receivers of this strip (:167-169); w rows for the next pressure update ------------------------------------------
        {
            const size_t rbase = ((size_t)s * g.nt + it) * a.nr;
#pragma unroll
            for (int j = 0; j < RMAX; ++j) {
                if (roff[j] >= 0) {
                    const size_t ro = rbase + tid + j * PNT;
                    a.rcv_p[ro] = p[roff[j]];
                    if (a.rcv_u) a.rcv_u[ro] = u[roff[j]];
                    if (a.rcv_w) a.rcv_w[ro] = w[roff[j]];
                }
            }
            for (int r = tid + RMAX * PNT; r < a.nr; r += PNT) {      // more than RMAX*PNT receivers: the general path
                const int rz = (int)a.rz[r], rx = (int)a.rx[r];
                if (rz >= z0 && rz < z0 + rows_own && rx >= 0 && rx < g.nxp) {
                    const int o = (rz - z0 + PHALO) * pitch + PXPAD + rx;
                    a.rcv_p[rbase + r] = p[o];
                    if (a.rcv_u) a.rcv_u[rbase + r] = u[o];
                    if (a.rcv_w) a.rcv_w[rbase + r] = w[o];
                }
            }
        }
        pull_halo(w, g, k, 2, 1, tid);
        __syncthreads();
    }
    cluster_sync_all();          // no CTA may exit (and give up its shared memory) while a neighbour is still pulling rows from it
    if (a.illum) {
#pragma unroll
        for (int q = 0; q < KMAX; ++q) {
            if (W[q].o >= 0) {
                const float4 wf = ld4(w + W[q].o);
                red4(a.ill_p + W[q].hz, accp[q]);
                red4(a.ill_u + W[q].hz, accu[q]);
                red4(a.ill_w + W[q].hz, make_float4(wf.x * wf.x, wf.y * wf.y, wf.z * wf.z, wf.w * wf.w));
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------------------------
// adjoint: all nt reverse steps of one shot (state = lambda_p, mu_u = alpha2u*lambda_u,
// mu_w = ew*lambda_w as in adj_tile)
// ------------------------------------------------------------------------------------------------------------------
// lambda_p after undoing W and U (5T, 4T) for the float4 group at haloed offset oh / own-rows offset oc
__device__ __forceinline__ float4 pp_adj_phase1(const float* lw, const float* lu, const float* lp, int oh, int oc, int pitch, float c1, float c2)
{
    const float4 w0 = ld4(lw + oh - 2 * pitch), w1 = ld4(lw + oh - pitch), w2 = ld4(lw + oh), w3 = ld4(lw + oh + pitch);
    const float* ur = lu + oc;
    const float2 ul = ld2(ur - 2); const float4 um = ld4(ur); const float uR = ur[4];
    const float q0 = ul.x, q1 = ul.y, q2_ = um.x, q3 = um.y, q4 = um.z, q5 = um.w, q6 = uR;
    float4 v = ld4(lp + oc);
    v.x += c1 * w2.x - c1 * w1.x + c2 * w3.x - c2 * w0.x;
    v.y += c1 * w2.y - c1 * w1.y + c2 * w3.y - c2 * w0.y;
    v.z += c1 * w2.z - c1 * w1.z + c2 * w3.z - c2 * w0.z;
    v.w += c1 * w2.w - c1 * w1.w + c2 * w3.w - c2 * w0.w;
    v.x += c1 * q2_ - c1 * q1 + c2 * q3 - c2 * q0;
    v.y += c1 * q3 - c1 * q2_ + c2 * q4 - c2 * q1;
    v.z += c1 * q4 - c1 * q3 + c2 * q5 - c2 * q2_;
    v.w += c1 * q5 - c1 * q4 + c2 * q6 - c2 * q3;
    return v;
}

template <bool FS, int KMAX>
__global__ void __launch_bounds__(PNT, 1)
acp_adj(const PGeom g, const PAdjArgs a)
{
    extern __shared__ __align__(16) float psm[];
    const int tid = threadIdx.x;
    const unsigned k = cluster_rank();
    const int s = a.s_begin + (int)cluster_id_x();
    const int rows_s = g.R + 2 * PHALO;
    const size_t sp = (size_t)rows_s * g.pitch, cpn = (size_t)g.R * g.pitch;
    float* lw = psm; float* ms = lw + sp;             // mu_w and m = -alpha1*lambda_p1: neighbours' rows are read -> halo rows
    float* lp = ms + sp; float* lu = lp + cpn;        // lambda_p, mu_u: own rows only
    PSmem S; S.a1 = lu + cpn; S.t1 = S.a1 + cpn; S.a2 = S.t1 + cpn; S.t2 = S.a2 + cpn; S.t3 = S.t2 + cpn;
    const int z0 = g.zlo + (int)k * g.R;
    const int rows_own = max(0, min(g.R, g.nzp - z0));
    const int ngroups = rows_own * g.nxg;
    const float c1 = g.c1, c2 = g.c2;
    const int pitch = g.pitch;
    for (int i = tid; i < 2 * (int)sp + 2 * (int)cpn; i += PNT) psm[i] = 0.f;
    pp_load_coefs<FS>(g, a.cp, S, z0, rows_own, tid);
    const int szs = a.g_src ? (int)a.sz[s] : -1, sxs = a.g_src ? (int)a.sx[s] : -1;
    const bool have_g = a.nr > 0 && (a.gp || a.gu || a.gw);
    const bool fs_top = FS && k == 0;
    PItem W[KMAX];
    float4 gacc[KMAX], Sh[KMAX];
    const float* hist_s = a.hist + (size_t)s * a.hist_len * g.plane;
#pragma unroll
    for (int q = 0; q < KMAX; ++q) {
        W[q] = pp_item(g, tid + q * PNT, ngroups, z0);
        gacc[q] = make_float4(0.f, 0.f, 0.f, 0.f);
        Sh[q] = W[q].o >= 0 ? __ldcs(reinterpret_cast<const float4*>(hist_s + (size_t)(g.nt - 1) * g.plane + W[q].hz)) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    // receivers of this strip handled by this thread (own-rows offset; scale of the u / w cotangents)
    constexpr int RMAX = 4;
    int roff[RMAX]; float rau[RMAX], rew[RMAX];
#pragma unroll
    for (int j = 0; j < RMAX; ++j) {
        const int r = tid + j * PNT;
        roff[j] = -1; rau[j] = 0.f; rew[j] = 0.f;
        if (have_g && r < a.nr) {
            const int rz = (int)a.rz[r], rx = (int)a.rx[r];
            if (rz >= z0 && rz < z0 + rows_own && rx >= 0 && rx < g.nxp) {
                roff[j] = (rz - z0) * pitch + PXPAD + rx;
                const ptrdiff_t co = (ptrdiff_t)rz * g.cpld + rx;
                rau[j] = a.cp.a2u[co]; rew[j] = a.cp.ew[co];
            }
        }
    }
    __syncthreads();
    cluster_sync_all();

    for (int it = g.nt - 1; it >= 0; --it) {
        // ---- 7T: receiver cotangents into the staged cotangents (u / w enter scaled by the cell's alpha2 / ew) -----------
        if (have_g) {
            const size_t rbase = ((size_t)s * g.nt + it) * a.nr;
#pragma unroll
            for (int j = 0; j < RMAX; ++j) {
                if (roff[j] >= 0) {
                    const size_t ro = rbase + tid + j * PNT;
                    if (a.gp) atomicAdd(lp + roff[j], a.gp[ro]);
                    if (a.gu) atomicAdd(lu + roff[j], rau[j] * a.gu[ro]);
                    if (a.gw) atomicAdd(lw + roff[j] + PHALO * pitch, rew[j] * a.gw[ro]);
                }
            }
            for (int r = tid + RMAX * PNT; r < a.nr; r += PNT) {
                const int rz = (int)a.rz[r], rx = (int)a.rx[r];
                if (rz >= z0 && rz < z0 + rows_own && rx >= 0 && rx < g.nxp) {
                    const int l = rz - z0;
                    const ptrdiff_t co = (ptrdiff_t)rz * g.cpld + rx;
                    if (a.gp) atomicAdd(lp + l * pitch + PXPAD + rx, a.gp[rbase + r]);
                    if (a.gu) atomicAdd(lu + l * pitch + PXPAD + rx, a.cp.a2u[co] * a.gu[rbase + r]);
                    if (a.gw) atomicAdd(lw + (l + PHALO) * pitch + PXPAD + rx, a.cp.ew[co] * a.gw[rbase + r]);
                }
            }
            if (fs_top) __syncthreads();
        }
        if (fs_top) {            // 6T: lambda_w[fs] += lambda_w[fs-1]; lambda_w[fs-1] = 0 (row fs-1 holds the raw cotangent, row fs the scaled one)
            for (int x = tid; x < 4 * g.nxg; x += PNT) {
                const float e = a.cp.a2w[(ptrdiff_t)(z0 + 1) * g.cpld + x];
                float* r1 = lw + (1 + PHALO) * pitch + PXPAD + x; float* r0 = lw + (0 + PHALO) * pitch + PXPAD + x;
                *r1 += e * *r0; *r0 = 0.f;
            }
        }
        cluster_sync_all();
        pull_halo(lw, g, k, 2, 1, tid);
        __syncthreads();
        // ---- phase 1: lambda_p after undoing W and U (5T, 4T), 3T, and m = -alpha1*lambda_p1; in place (own cell) ---------------
#pragma unroll
        for (int q = 0; q < KMAX; ++q) {
            if (W[q].o >= 0) {
                const int oh = W[q].o, oc = oh - PHALO * pitch;
                float4 v = pp_adj_phase1(lw, lu, lp, oh, oc, pitch, c1, c2);
                if (fs_top && W[q].l() == 2) {         // 3T: lambda_p[fs+1] -= lambda_p[fs-1] (recomputed here); the owner of row fs-1 writes its zero
                    const float4 v0 = pp_adj_phase1(lw, lu, lp, oh - 2 * pitch, oc - 2 * pitch, pitch, c1, c2);
                    v.x -= v0.x; v.y -= v0.y; v.z -= v0.z; v.w -= v0.w;
                }
                // row fs-1: lambda_p1 = 0 after 3T.  Its owner must NOT store that yet -- the owner of row fs+1 is reading the old value
                // in this very phase; phase 2 (after the barriers) treats it as zero and writes the zero back.
                const bool zeroed = fs_top && W[q].l() == 0;
                if (zeroed) v = make_float4(0.f, 0.f, 0.f, 0.f);
                const float4 A1 = ld4(S.a1 + oc);
                st4(ms + oh, make_float4((-A1.x) * v.x, (-A1.y) * v.y, (-A1.z) * v.z, (-A1.w) * v.w));
                if (!zeroed) st4(lp + oc, v);        // lambda_p1, read again by phase 2 (own cell)
            }
        }
        cluster_sync_all();
        pull_halo(ms, g, k, 1, 2, tid);
        __syncthreads();
        // ---- phase 2: new mu_u, mu_w, lambda_p (5T, 4T, 1T), g_alpha1, g_src ------------------------------------------------------
#pragma unroll
        for (int q = 0; q < KMAX; ++q) {
            float4 Snext = make_float4(0.f, 0.f, 0.f, 0.f);
            if (W[q].o >= 0) {
                const int oh = W[q].o, oc = oh - PHALO * pitch;
                const int gz = z0 + W[q].l(), gx = W[q].gx();
                if (it > 0) Snext = __ldcs(reinterpret_cast<const float4*>(hist_s + (size_t)(it - 1) * g.plane + W[q].hz));
                const float* mr = ms + oh;
                const float mL = mr[-1];
                const float2 mR = ld2(mr + 4);
                const float4 m0 = ld4(mr), mm1 = ld4(mr - pitch), mp1 = ld4(mr + pitch), mp2 = ld4(mr + 2 * pitch);
                const float4 qp = (fs_top && W[q].l() == 0) ? make_float4(0.f, 0.f, 0.f, 0.f) : ld4(lp + oc), luo = ld4(lu + oc), lwo = ld4(lw + oh);
                const float4 t1 = ld4(S.t1 + oc), t2 = ld4(S.t2 + oc), t3 = ld4(S.t3 + oc);
                float4 eu, ew;
                pp_split_a2(g, gz, gx, ld4(S.a2 + oc), eu, ew);
                if (FS && gz == g.fs - 1) {          // raw cotangent parked on row fs-1: scale 1 inside the grid
                    ew.x = gx < g.nxp ? 1.f : 0.f; ew.y = gx + 1 < g.nxp ? 1.f : 0.f; ew.z = gx + 2 < g.nxp ? 1.f : 0.f; ew.w = gx + 3 < g.nxp ? 1.f : 0.f;
                }
                float4 du, dw, nu, nw, np;
                du.x = eu.x * (c1 * m0.x - c1 * m0.y + c2 * mL - c2 * m0.z);
                du.y = eu.y * (c1 * m0.y - c1 * m0.z + c2 * m0.x - c2 * m0.w);
                du.z = eu.z * (c1 * m0.z - c1 * m0.w + c2 * m0.y - c2 * mR.x);
                du.w = eu.w * (c1 * m0.w - c1 * mR.x + c2 * m0.z - c2 * mR.y);
                dw.x = ew.x * (c1 * m0.x - c1 * mp1.x + c2 * mm1.x - c2 * mp2.x);
                dw.y = ew.y * (c1 * m0.y - c1 * mp1.y + c2 * mm1.y - c2 * mp2.y);
                dw.z = ew.z * (c1 * m0.z - c1 * mp1.z + c2 * mm1.z - c2 * mp2.z);
                dw.w = ew.w * (c1 * m0.w - c1 * mp1.w + c2 * mm1.w - c2 * mp2.w);
                nu.x = t2.x * luo.x + du.x; nu.y = t2.y * luo.y + du.y; nu.z = t2.z * luo.z + du.z; nu.w = t2.w * luo.w + du.w;
                nw.x = t3.x * lwo.x + dw.x; nw.y = t3.y * lwo.y + dw.y; nw.z = t3.z * lwo.z + dw.z; nw.w = t3.w * lwo.w + dw.w;
                np.x = t1.x * qp.x; np.y = t1.y * qp.y; np.z = t1.z * qp.z; np.w = t1.w * qp.w;
                st4(lu + oc, nu); st4(lw + oh, nw); st4(lp + oc, np);
                if (a.g_src && gz == szs) {
                    const int dx = sxs - gx;
                    if (dx >= 0 && dx < 4) a.g_src[(size_t)s * g.nt + it] = g.dt * (dx == 0 ? qp.x : dx == 1 ? qp.y : dx == 2 ? qp.z : qp.w);
                }
                gacc[q].x = gacc[q].x - qp.x * Sh[q].x; gacc[q].y = gacc[q].y - qp.y * Sh[q].y;
                gacc[q].z = gacc[q].z - qp.z * Sh[q].z; gacc[q].w = gacc[q].w - qp.w * Sh[q].w;
            }
            Sh[q] = Snext;
        }
        __syncthreads();
    }
    cluster_sync_all();          // no CTA may exit (and give up its shared memory) while a neighbour is still pulling rows from it
#pragma unroll
    for (int q = 0; q < KMAX; ++q)
        if (W[q].o >= 0) red4(a.g1part + W[q].hz, gacc[q]);
}
