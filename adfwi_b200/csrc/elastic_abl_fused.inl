// elastic_abl_fused.inl -- TMA-staged, tile-persistent elastic time step with the multiplicative sponge (ABL) boundary.
// Included inside the anonymous namespace of elastic_fused.cu (shares its helpers, tile geometry and work distribution).
//
// Semantics: ADFWI/propagator/elastic_kernels.py:709-774 (step_forward_ABL_4order) / :844-905 (_6order) and the reverse-mode
// derivative of that loop ("ABL step" / "Adjoint of the ABL step"); same arithmetic and association as the
// generic kernels of elastic.cu with PML = false (-fmad=false): forward records stay bit-identical to the CPU reference.
//
//   ela_f : ONE launch per forward step.  Five unsplit fields.  Per shot the two velocities (halo 2NN) and the three stresses (halo NN)
//           arrive by TMA in a double-buffered stage.  Phase A updates the stresses on the tile plus a ring of NN cells IN PLACE in the
//           staged rectangles (own-cell operation), adds the source and applies the free-surface stress edits (they persist in the
//           state here, unlike the split-PML case); phase B stores the stresses, updates the velocities on the tile, multiplies by the
//           sponge plane, samples the receivers and stores the history as five own-cell planes
//               D-x vx,  D-z vz,  D+x vz/dx + D+z vx/dz,  D+x txx/dx + D-z txz/dz,  D-x txz/dx + D+z tzz/dz.
//           Free-surface velocity rows h-2, h-3 (elastic_kernels.py:753-756) depend on the UNDAMPED rows h-1, h of the x-neighbour
//           tile: the tiles of the top row park their undamped rows h-1 (vz) and h (vx) in a two-row side buffer (ping-pong over time
//           steps), and form rows h-2, h-3 (times the sponge) in their staged velocity rectangles at the start of the next step.
//   ela_b : ONE launch per reverse step.  Cotangents of the five fields, same staging.  Phase A: q = sponge * lambda_v with the
//           free-surface transposes, g_bx, g_bz, and m = dt*b*q on tile + 2NN ring, in place in the velocity rectangles; phase B:
//           lambda_tau += transposed operators of m on tile + NN ring, in place in the stress rectangles; free-surface stress
//           transposes; phase C: new lambda_tau to HBM, g_C11..g_C55, g_src, n = dt*C*lambda_tau on tile + NN ring in place; phase D:
//           lambda_v = transposed operators of n + q, to the other set of the ping-pong pair.
// Algorithmic HBM bytes per cell-update: forward 64 (5 fields r+w 40, 6 coefficient planes 24) + 20 recording = 84; reverse 84
// (5 cotangents r+w 40, coefficients 24, 5 history planes 20).

struct AFArgs { ECoef cp; float* planes; int cur; const float* mt; const float* src_v; const int64_t *sx, *sz;
                float* hist; int hist_len, tl, it; int nr; RcvB rb; float* rcv[5]; float* side; Walk w; };
struct ABArgs { ECoef cp; float* planes; const float* hist; int hist_len, tl, it, lcur; int nr; RcvB rb; const float* g[5];
                const float* mt; const int64_t *sx, *sz; float* g_src; float* gpart; Walk w; };

// plane index of field f, shot s in the ABL workspace: f*ns + s.  Two sets of the five fields (ping-pong of ela_f), two sets of
// their cotangents (ping-pong of ela_b).  Field order inside a set: vx, vz, txx, tzz, txz.
enum { A_VX = 0, A_VZ, A_TXX, A_TZZ, A_TXZ, A_SET = 5, A_FWD_COUNT = 10, A_L0 = 10, A_COUNT = 20 };
constexpr int ANHIST = 5;

template <int NN> struct GeoA {
    using G = Geo<NN>;
    static constexpr int STAGE = 2 * G::HB2 + 3 * G::HB;       // vx, vz (halo 2NN) + txx, tzz, txz (halo NN)
    static constexpr int GX2 = G::HX2 / 4;
    static constexpr int W2 = NG + 2 * GX2;
    static constexpr int NRING2 = 4 * NN * W2 + TZ * 2 * GX2;  // float4 groups of the 2NN ring (same enumeration as GeoB)
    static constexpr int FS_BYTES = 2 * G::RX2 * 4;            // reverse kernel: free-surface addends of rows h-1 (vz) and h (vx)
    // forward kernel: C11, C13, C33, C55 of the ring cells, filled once per (tile, chunk) by the thread that reads them every shot
    // (float4 SoA).  ela_f 0.80 -> 0.83 of its roofline; the same tables bought nothing in ela_b and cost elf_f 1 % (its two CTAs
    // already take 217 KB of the SM's shared memory), so those two keep reading the L2-resident pack (profiles/r02x_el_forward.md)
    static constexpr int TF_BYTES = 4 * G::NRING * 16;
};

template <int NN> __device__ __forceinline__ void a_issue(const Cursor& c, unsigned char* smem, uint64_t* bar, int k,
                                                          const CUtensorMap* th, const CUtensorMap* th2, int ns, int base)
{
    using G = Geo<NN>;
    unsigned char* st = smem + k * GeoA<NN>::STAGE;
    fence_proxy_async();
    mbar_expect_tx(bar + k, 2 * G::HF2 * 4 + 3 * G::HF * 4);
    tma_load_3d(st, th2, c.X0 - G::HX2, c.Z0 - 2 * NN, (base + A_VX) * ns + c.s, bar + k);
    tma_load_3d(st + G::HB2, th2, c.X0 - G::HX2, c.Z0 - 2 * NN, (base + A_VZ) * ns + c.s, bar + k);
#pragma unroll
    for (int f = 0; f < 3; ++f) tma_load_3d(st + 2 * G::HB2 + f * G::HB, th, c.X0 - HX, c.Z0 - NN, (base + A_TXX + f) * ns + c.s, bar + k);
}
// the reverse kernel also pulls the five history planes of the shot into L2 (plain loads by the consumers then hit there)
template <int NN> __device__ __forceinline__ void ab_issue(const Cursor& c, unsigned char* smem, uint64_t* bar, int k,
                                                           const CUtensorMap* th, const CUtensorMap* th2, const CUtensorMap* thh,
                                                           int ns, int base, int hist_len, int tl)
{
#pragma unroll
    for (int e = 0; e < ANHIST; ++e) tma_prefetch_3d(thh, c.X0, c.Z0, (c.s * hist_len + tl) * ANHIST + e);
    a_issue<NN>(c, smem, bar, k, th, th2, ns, base);
}

// ------------------------------------------------------------------------------------------------------------------
// forward: stress update of one float4 group (elastic_kernels.py:720-725).  vx, vz: velocity rects (pitch RX2), hv = index of the
// group there; sp[3] = txx, tzz, txz of the group (in: old, out: new where the mask is set); d[3] = D-x vx, D-z vz and the
// history combination D+x vz/dx + D+z vx/dz
// ------------------------------------------------------------------------------------------------------------------
template <int NN>
__device__ __forceinline__ void a_stress_cell(const EGeom& g, unsigned m, const float* vx, const float* vz, int hv, float4* sp,
                                              const float4& c11, const float4& c13, const float4& c33, const float4& c55, float4* d)
{
    constexpr int RX2 = Geo<NN>::RX2;
    float4 dxb_vx, dzb_vz, dxf_vz, dzf_vx;
    {
        float sx_[12];
        ldseg(vx + hv, sx_);
        dxb_vx = xdiff<NN, 0>(sx_, g.c);
    }
    ELF_SEQ();
    {
        float sz_[12];
        ldseg(vz + hv, sz_);
        dxf_vz = xdiff<NN, 1>(sz_, g.c);
    }
    ELF_SEQ();
    {
        float4 wzb[2 * NN];
#pragma unroll
        for (int q = 0; q < 2 * NN; ++q) wzb[q] = ld4(vz + hv + (q - NN) * RX2);
        dzb_vz = zdiff<NN>(wzb, g.c);
    }
    ELF_SEQ();
    {
        float4 wzf[2 * NN];
#pragma unroll
        for (int q = 0; q < 2 * NN; ++q) wzf[q] = ld4(vx + hv + (q - NN + 1) * RX2);
        dzf_vx = zdiff<NN>(wzf, g.c);
    }
    ELF_SEQ();
    // txx + dt*((C11*D-x vx)/dx + (C13*D-z vz)/dz), tzz, txz likewise: true divisions, see fdiv1 / fdiv1_tiny / DivGuard
    const float4 A0 = mul4(c11, dxb_vx), A1 = mul4(c13, dzb_vz), A2 = mul4(c13, dxb_vx), A3 = mul4(c33, dzb_vz);
    const float4 A4 = mul4(c55, dxf_vz), A5 = mul4(c55, dzf_vx);
    DivGuard dg;
    dg.add(A0); dg.add(A1); dg.add(A2); dg.add(A3); dg.add(A4); dg.add(A5);
    float4 t0 = fdivs(A0, g.dx, g.rdx), t1 = fdivs(A1, g.dz, g.rdz), t2 = fdivs(A2, g.dx, g.rdx);
    float4 t3 = fdivs(A3, g.dz, g.rdz), t4 = fdivs(A4, g.dx, g.rdx), t5 = fdivs(A5, g.dz, g.rdz);
    if (!dg.ok()) {          // a numerator in the underflow range (the band ahead of a wavefront): scaled sequence
        t0 = safe_divs(A0, g.dx, g.rdx); t1 = safe_divs(A1, g.dz, g.rdz); t2 = safe_divs(A2, g.dx, g.rdx);
        t3 = safe_divs(A3, g.dz, g.rdz); t4 = safe_divs(A4, g.dx, g.rdx); t5 = safe_divs(A5, g.dz, g.rdz);
    }
    const float4 n0 = add4(sp[0], smul(g.dt, add4(t0, t1)));
    const float4 n1 = add4(sp[1], smul(g.dt, add4(t2, t3)));
    const float4 n2 = add4(sp[2], smul(g.dt, add4(t4, t5)));
    sp[0] = sel4(m, n0, sp[0]); sp[1] = sel4(m, n1, sp[1]); sp[2] = sel4(m, n2, sp[2]);
    d[0] = dxb_vx; d[1] = dzb_vz;
    d[2] = add4(muls(dxf_vz, g.rdx), muls(dzf_vx, g.rdz));     // history only (the adjoint is not bit-faithful): reciprocal multiplies
}

template <int NN, bool FS, bool SAVE>
__device__ __forceinline__ void af_tile(const CUtensorMap* th, const CUtensorMap* th2, const EGeom& g, const AFArgs& a,
                                        unsigned char* smem, uint64_t* bar, uint32_t& par, int& stage, Cursor& pc, int* ring,
                                        int* s_sz, int* s_sx, float* s_sxx, float* s_szz, float* s_sxz, float4* tab,
                                        const Roles& R, int tid, int tile, int s_lo, int s_hi, bool first)
{
    using G = Geo<NN>;
    constexpr int RX2 = G::RX2, HX2 = G::HX2, HQ = G::HB / 4, HQ2 = G::HB2 / 4;
    const uint64_t pol = l2_keep_policy();
    float4* T_c11 = tab; float4* T_c13 = tab + G::NRING; float4* T_c33 = tab + 2 * G::NRING; float4* T_c55 = tab + 3 * G::NRING;
    const int tzi = tile / g.ntx, txi = tile - tzi * g.ntx;
    const int X0 = txi * TX, Z0 = tzi * TZ;
    const int gx = X0 + R.c0, gz = Z0 + R.r0;
    const unsigned mt_ = row_in<NN>(gz, g.nzp) ? col_mask<NN>(gx, g.nxp) : 0u;
    const bool cell_ok = gx < g.ld && gz < g.nzp;
    const size_t fp = (size_t)g.ns * g.plane;
    const int rbase = a.cur ? A_SET : 0, wbase = a.cur ? 0 : A_SET;
    const int h = NN + 1;
    if (tid < s_hi - s_lo) {         // per-shot scalars: (-1/3 * M) * src  (elastic_kernels.py:703, :733-735)
        const int s = s_lo + tid;
        const float* M = a.mt + (size_t)s * 9;
        const float v = a.src_v[(size_t)s * g.nt + a.it];
        const float sc = (float)(-1.0 / 3.0);
        s_sz[tid] = (int)a.sz[s]; s_sx[tid] = (int)a.sx[s];
        s_sxx[tid] = (sc * M[0]) * v; s_szz[tid] = (sc * M[8]) * v; s_sxz[tid] = (sc * M[2]) * v;
    }
    const ptrdiff_t oc = (ptrdiff_t)gz * g.cpld + gx;
    const float4 C11 = ldk4(a.cp.c11 + oc, pol), C13 = ldk4(a.cp.c13 + oc, pol), C33 = ldk4(a.cp.c33 + oc, pol), C55 = ldk4(a.cp.c55 + oc, pol);
    const float4 DBX = smul(g.dt, ldk4(a.cp.bx + oc, pol)), DBZ = smul(g.dt, ldk4(a.cp.bz + oc, pol));     // dt*bx, dt*bz (:749-750)
    const float4 DMP = ldk4(a.cp.bcx + oc, pol);                                                              // sponge plane
    for (int i = tid; i < G::NRING; i += NTH) {          // ring coefficients: written and read by the same thread (no barrier needed)
        int r, gi;
        ring_cell<NN>(i, r, gi);
        const ptrdiff_t o = (ptrdiff_t)(Z0 + r) * g.cpld + X0 + 4 * gi;
        T_c11[i] = ldk4(a.cp.c11 + o, pol); T_c13[i] = ldk4(a.cp.c13 + o, pol); T_c33[i] = ldk4(a.cp.c33 + o, pol); T_c55[i] = ldk4(a.cp.c55 + o, pol);
    }
    const int rcv_lo = a.nr > 0 ? a.rb.start[tile] : 0, rcv_hi = a.nr > 0 ? a.rb.start[tile + 1] : 0;
    const bool has_rcv = rcv_hi > rcv_lo;
    const bool producer = tid == NTH - 32;              // lane 0 of the last warp, which has no ring cells
    const int hv = (R.r0 + 2 * NN) * RX2 + R.c0 + HX2, hs = (R.r0 + NN) * RXH + R.c0 + HX;
    // side buffer of the free-surface rows: [parity][row: 0 = undamped vz[h-1], 1 = undamped vx[h]][shot][ld]
    const size_t side_row = (size_t)g.ns * g.ld;
    const float* side_rd = a.side + (size_t)(a.cur ? 2 : 0) * side_row;
    float* side_wr = a.side + (size_t)(a.cur ? 0 : 2) * side_row;
    if (first) {
        griddep_wait();
#pragma unroll
        for (int k = 0; k < NSTAGE; ++k)
            if (producer && pc.valid) { a_issue<NN>(pc, smem, bar, k, th, th2, g.ns, rbase); pc.next(g, a.w, ring); }
    }
    __syncthreads();

    for (int s = s_lo; s < s_hi; ++s) {
        const int k = stage;
        float* vx = (float*)(smem + k * GeoA<NN>::STAGE); float* vz = vx + HQ2;
        float* txx = vx + 2 * HQ2; float* tzz = txx + HQ; float* txz = txx + 2 * HQ;
        const int si = s - s_lo;
        const int szs = s_sz[si], sxs = s_sx[si];
        const float sxx = s_sxx[si], szz = s_szz[si], sxz = s_sxz[si];
        ELF_WAIT_STAGE(k);
        if (FS && tzi == 0) {
            // free-surface velocity rows of the PREVIOUS step (:753-756), formed from its undamped rows h-1 (vz), h (vx) and damped:
            // vz[h-2] = vz[h-3] = vz[h-1];  vx[h-2] = vz[h-2,j+1] - vz[h-2,j] + vz[h-1,j+1] - vz[h-1,j] + vx[h,j]
            const int t = tid;
            if (t >= HX2 - 2 * NN && t < HX2 + TX + 2 * NN) {
                const int j = X0 + t - HX2;
                if (j >= NN && j < g.nxp - NN) {
                    const float* sv = side_rd + (size_t)s * g.ld;
                    const float vz1 = sv[j];
                    const float vz1n = (j + 1 < g.nxp - NN) ? sv[j + 1] : 0.f;
                    const float vxh = sv[side_row + j];
                    const float nvx = (((vz1n - vz1) + vz1n) - vz1) + vxh;
                    const float d2 = a.cp.bcx[(ptrdiff_t)(h - 2) * g.cpld + j], d3 = a.cp.bcx[(ptrdiff_t)(h - 3) * g.cpld + j];
                    vz[(h - 2 + 2 * NN) * RX2 + t] = vz1 * d2;
                    vx[(h - 2 + 2 * NN) * RX2 + t] = nvx * d2;
                    vz[(h - 3 + 2 * NN) * RX2 + t] = vz1 * d3;
                }
            }
            __syncthreads();
        }
        // ---- phase A: stresses on the tile (coefficients in registers) and on the ring (coefficients from L2), in place ----
        float4 D0, D1, D2;
        {
            float4 sp[3], d[3];
            sp[0] = ld4(txx + hs); sp[1] = ld4(tzz + hs); sp[2] = ld4(txz + hs);
            a_stress_cell<NN>(g, mt_, vx, vz, hv, sp, C11, C13, C33, C55, d);
            if (szs == gz && !(FS && gz < NN)) {
                const int dc = sxs - gx;
                if (dc >= 0 && dc < 4) { addc4(sp[0], dc, sxx); addc4(sp[1], dc, szz); addc4(sp[2], dc, sxz); }
            }
            st4(txx + hs, sp[0]); st4(tzz + hs, sp[1]); st4(txz + hs, sp[2]);
            D0 = sel4(mt_, d[0], zero4()); D1 = sel4(mt_, d[1], zero4()); D2 = sel4(mt_, d[2], zero4());
        }
        if (SAVE && cell_ok) {
            float* H = a.hist + ((size_t)s * a.hist_len + a.tl) * ANHIST * g.plane + (size_t)gz * g.ld + gx;
            __stcs(reinterpret_cast<float4*>(H), D0);
            __stcs(reinterpret_cast<float4*>(H + g.plane), D1);
            __stcs(reinterpret_cast<float4*>(H + 2 * g.plane), D2);
        }
        for (int i = tid; i < G::NRING; i += NTH) {
            int r, gi;
            ring_cell<NN>(i, r, gi);
            const int gzr = Z0 + r, gxr = X0 + 4 * gi;
            const int hv2 = (r + 2 * NN) * RX2 + 4 * gi + HX2, hs2 = (r + NN) * RXH + 4 * gi + HX;
            const unsigned m = row_in<NN>(gzr, g.nzp) ? col_mask<NN>(gxr, g.nxp) : 0u;
            float4 sp[3], d[3];
            sp[0] = ld4(txx + hs2); sp[1] = ld4(tzz + hs2); sp[2] = ld4(txz + hs2);
            a_stress_cell<NN>(g, m, vx, vz, hv2, sp, T_c11[i], T_c13[i], T_c33[i], T_c55[i], d);
            if (szs == gzr && !(FS && gzr < NN)) {
                const int dc = sxs - gxr;
                if (dc >= 0 && dc < 4) { addc4(sp[0], dc, sxx); addc4(sp[1], dc, szz); addc4(sp[2], dc, sxz); }
            }
            st4(txx + hs2, sp[0]); st4(tzz + hs2, sp[1]); st4(txz + hs2, sp[2]);
        }
        __syncthreads();
        if (FS && tzi == 0) {          // free-surface stress rows (:737-741), on the state itself
            const int t = tid;
            if (t < RXH) {
                const int j = X0 + t - HX;
                if (j >= NN && j < g.nxp - NN) {
                    const int rh1 = (2 * NN) * RXH + t, rh = (2 * NN + 1) * RXH + t, rh2 = (2 * NN - 1) * RXH + t, rh3 = (2 * NN - 2) * RXH + t;
                    tzz[rh1] = 0.f;
                    tzz[rh2] = -tzz[rh];
                    txz[rh2] = -txz[rh1];
                    txz[rh3] = -txz[rh];
                }
            }
            __syncthreads();
        }
        // ---- phase B: stresses out, velocity on the tile, sponge, history, receivers ---------------------------------
        float4 nvx, nvz;
        {
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
            const float4 q0 = ld4(vx + hv), q1 = ld4(vz + hv);
            DivGuard dg;
            dg.add(dxf_txx); dg.add(dzb_txz); dg.add(dxb_txz); dg.add(dzf_tzz);
            float4 ex = add4(fdivs(dxf_txx, g.dx, g.rdx), fdivs(dzb_txz, g.dz, g.rdz));      // D+x txx/dx + D-z txz/dz
            float4 ez = add4(fdivs(dxb_txz, g.dx, g.rdx), fdivs(dzf_tzz, g.dz, g.rdz));      // D-x txz/dx + D+z tzz/dz
            if (!dg.ok()) {
                ex = add4(safe_divs(dxf_txx, g.dx, g.rdx), safe_divs(dzb_txz, g.dz, g.rdz));
                ez = add4(safe_divs(dxb_txz, g.dx, g.rdx), safe_divs(dzf_tzz, g.dz, g.rdz));
            }
            const float4 ux = sel4(mt_, add4(q0, mul4(DBX, ex)), q0);        // undamped new velocities (:749-750)
            const float4 uz = sel4(mt_, add4(q1, mul4(DBZ, ez)), q1);
            // sponge (:759-760) on the update region; cells outside it hold zero or the free-surface rows formed (and damped) above
            nvx = sel4(mt_, mul4(ux, DMP), q0);
            nvz = sel4(mt_, mul4(uz, DMP), q1);
            if (cell_ok) {
                const size_t o = (size_t)gz * g.ld + gx;
                float* P = a.planes + (size_t)s * g.plane + o + (size_t)wbase * fp;
                st4(P + A_VX * fp, nvx); st4(P + A_VZ * fp, nvz);
                st4(P + A_TXX * fp, ld4(txx + hs)); st4(P + A_TZZ * fp, ld4(tzz + hs)); st4(P + A_TXZ * fp, ld4(txz + hs));
                if (FS && tzi == 0) {      // undamped rows for the free-surface rows of the next step
                    if (gz == h - 1) st4(side_wr + (size_t)s * g.ld + gx, uz);
                    if (gz == h) st4(side_wr + side_row + (size_t)s * g.ld + gx, ux);
                }
                if (SAVE) {
                    float* H = a.hist + (((size_t)s * a.hist_len + a.tl) * ANHIST + 3) * g.plane + o;
                    __stcs(reinterpret_cast<float4*>(H), sel4(mt_, ex, zero4()));
                    __stcs(reinterpret_cast<float4*>(H + g.plane), sel4(mt_, ez, zero4()));
                }
            }
        }
        if (has_rcv) {                 // receivers of this tile (:763-767): the new velocities parked in the staged rects
            st4(vx + hv, nvx); st4(vz + hv, nvz);
            __syncthreads();
            for (int i = rcv_lo + tid; i < rcv_hi; i += NTH) {
                const int r = a.rb.id[i], zx = a.rb.zx[i];
                const int z = (zx >> 16) - Z0, x = (zx & 0xffff) - X0;
                const int oh = (z + NN) * RXH + x + HX, ov = (z + 2 * NN) * RX2 + x + HX2;
                const size_t o = ((size_t)s * g.nt + a.it) * a.nr + r;
                a.rcv[0][o] = txx[oh]; a.rcv[1][o] = tzz[oh]; a.rcv[2][o] = txz[oh];
                a.rcv[3][o] = vx[ov]; a.rcv[4][o] = vz[ov];
            }
        }
        fence_proxy_async();
        // stage k has been consumed: refill it with the shot after the next.  Only the producer's warp waits for the other warps'
        // phase B (named barrier 1); they go straight on to the next shot, which lives in the other stage
        if (tid >= NTH - 32) {
            asm volatile("bar.sync 1, %0;" ::"n"(NTH) : "memory");
            if (producer && pc.valid) { a_issue<NN>(pc, smem, bar, k, th, th2, g.ns, rbase); pc.next(g, a.w, ring); }
        } else {
            asm volatile("bar.arrive 1, %0;" ::"n"(NTH) : "memory");
        }
        stage = (stage + 1 == NSTAGE) ? 0 : stage + 1;
    }
}

template <int NN, bool FS, bool SAVE>
__global__ void __launch_bounds__(NTH, 2)
ela_f(const __grid_constant__ CUtensorMap th, const __grid_constant__ CUtensorMap th2, const EGeom g, const AFArgs a)
{
    static_assert(RPT == 1, "ela_f: one float4 group per thread");
    extern __shared__ __align__(128) unsigned char smem[];
    uint64_t* bar = (uint64_t*)(smem + NSTAGE * GeoA<NN>::STAGE);
    int* s_sz = (int*)(bar + 8);
    int* s_sx = s_sz + CMAX;
    float* s_sxx = (float*)(s_sx + CMAX);
    float* s_szz = s_sxx + CMAX;
    float* s_sxz = s_szz + CMAX;
    float4* tab = (float4*)(smem + NSTAGE * GeoA<NN>::STAGE + TAIL_BYTES);
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
        af_tile<NN, FS, SAVE>(&th, &th2, g, a, smem, bar, par, stage, pc, ring, s_sz, s_sx, s_sxx, s_szz, s_sxz, tab, R, tid, tile, s_lo, s_hi, first);
        __syncthreads();
    }
}

// ------------------------------------------------------------------------------------------------------------------
// reverse step
// ------------------------------------------------------------------------------------------------------------------
// phase A of one float4 group: q = sponge * lambda_v (+ free-surface addend), masked to the update region; m = dt*b*q
__device__ __forceinline__ void ab_vel_cell(const EGeom& g, unsigned m, const float4& lvx, const float4& lvz, const float4& dmp,
                                            const float4& bx, const float4& bz, const float4& addx, const float4& addz,
                                            float4& qx, float4& qz, float4& wx, float4& wz, float4& mx, float4& mz)
{
    qx = sel4(m, add4(mul4(dmp, lvx), addx), zero4());
    qz = sel4(m, add4(mul4(dmp, lvz), addz), zero4());
    wx = muls(qx, g.dt); wz = muls(qz, g.dt);
    mx = mul4(wx, bx); mz = mul4(wz, bz);
}

template <int NN, bool FS>
__device__ __forceinline__ void ab_tile(const CUtensorMap* th, const CUtensorMap* th2, const CUtensorMap* thh, const EGeom& g, const ABArgs& a,
                                        unsigned char* smem, uint64_t* bar, uint32_t& par, int& stage, Cursor& pc, int* ring,
                                        int* s_sz, int* s_sx, float* fsadd, const Roles& R, int tid, int tile, int chunk, int s_lo, int s_hi, bool first)
{
    using G = Geo<NN>;
    using A = GeoA<NN>;
    constexpr int RX2 = G::RX2, HX2 = G::HX2, HQ = G::HB / 4, HQ2 = G::HB2 / 4, GX2 = A::GX2;
    const uint64_t pol = l2_keep_policy();
    const int tzi = tile / g.ntx, txi = tile - tzi * g.ntx;
    const int X0 = txi * TX, Z0 = tzi * TZ;
    const int gx = X0 + R.c0, gz = Z0 + R.r0;
    const unsigned mt_ = row_in<NN>(gz, g.nzp) ? col_mask<NN>(gx, g.nxp) : 0u;
    const bool cell_ok = gx < g.ld && gz < g.nzp;
    const size_t fp = (size_t)g.ns * g.plane;
    const int rbase = A_L0 + (a.lcur ? A_SET : 0), wbase = A_L0 + (a.lcur ? 0 : A_SET);
    const int h = NN + 1;
    const int hv = (R.r0 + 2 * NN) * RX2 + R.c0 + HX2, hs = (R.r0 + NN) * RXH + R.c0 + HX;
    float* addz_r = fsadd; float* addx_r = fsadd + RX2;      // free-surface addends per column of the velocity rects
    if (a.g_src && tid < s_hi - s_lo) { s_sz[tid] = (int)a.sz[s_lo + tid]; s_sx[tid] = (int)a.sx[s_lo + tid]; }
    const ptrdiff_t oc = (ptrdiff_t)gz * g.cpld + gx;
    const float4 C11 = ldk4(a.cp.c11 + oc, pol), C13 = ldk4(a.cp.c13 + oc, pol), C33 = ldk4(a.cp.c33 + oc, pol), C55 = ldk4(a.cp.c55 + oc, pol);
    const float4 BX = ldk4(a.cp.bx + oc, pol), BZ = ldk4(a.cp.bz + oc, pol), DMP = ldk4(a.cp.bcx + oc, pol);
    float4 G11 = zero4(), G13 = zero4(), G33 = zero4(), G55 = zero4(), GBX = zero4(), GBZ = zero4();
    const bool have_g = a.nr > 0 && (a.g[0] || a.g[1] || a.g[2] || a.g[3] || a.g[4]);
    const bool inject = have_g && a.rb.nbr[tile];
    const bool producer = tid == NTH - 32;              // lane 0 of the last warp
    if (first) {
        griddep_wait();
#pragma unroll
        for (int k = 0; k < NSTAGE; ++k)
            if (producer && pc.valid) { ab_issue<NN>(pc, smem, bar, k, th, th2, thh, g.ns, rbase, a.hist_len, a.tl); pc.next(g, a.w, ring); }
    }
    __syncthreads();

    for (int s = s_lo; s < s_hi; ++s) {
        const int k = stage;
        float* lvx = (float*)(smem + k * A::STAGE); float* lvz = lvx + HQ2;          // become mx, mz in phase A
        float* ltxx = lvx + 2 * HQ2; float* ltzz = ltxx + HQ; float* ltxz = ltxx + 2 * HQ;   // become nA, nB, nS in phase C
        const float* H = a.hist + ((size_t)s * a.hist_len + a.tl) * ANHIST * g.plane + (size_t)gz * g.ld + gx;
        float4 EX = zero4(), EZ = zero4();
        if (cell_ok) { EX = __ldcs(reinterpret_cast<const float4*>(H + 3 * g.plane)); EZ = __ldcs(reinterpret_cast<const float4*>(H + 4 * g.plane)); }
        ELF_WAIT_STAGE(k);
        if (inject) {            // 9T: record cotangents into the staged cotangents (duplicates legal -> shared-memory atomics)
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
                        const size_t o = ((size_t)s * g.nt + a.it) * a.nr + a.rb.id[i];
                        const int z2 = (zx >> 16) - (Z0 - 2 * NN), x2 = (zx & 0xffff) - (X0 - HX2);
                        if (z2 >= 0 && z2 < G::RZ2 && x2 >= 0 && x2 < RX2) {
                            if (a.g[3]) atomicAdd(lvx + z2 * RX2 + x2, a.g[3][o]);
                            if (a.g[4]) atomicAdd(lvz + z2 * RX2 + x2, a.g[4][o]);
                        }
                        const int z1 = (zx >> 16) - (Z0 - NN), x1 = (zx & 0xffff) - (X0 - HX);
                        if (z1 >= 0 && z1 < G::RZH && x1 >= 0 && x1 < RXH) {
                            if (a.g[0]) atomicAdd(ltxx + z1 * RXH + x1, a.g[0][o]);
                            if (a.g[1]) atomicAdd(ltzz + z1 * RXH + x1, a.g[1][o]);
                            if (a.g[2]) atomicAdd(ltxz + z1 * RXH + x1, a.g[2][o]);
                        }
                    }
                }
            }
            __syncthreads();
        }
        if (FS && tzi == 0) {
            // 8T + 7T for the free-surface rows: what rows h-2, h-3 of sponge*lambda_v hand to rows h-1 (vz) and h (vx)
            const int t = tid;
            if (t < RX2) {
                const int j = X0 + t - HX2;
                float az = 0.f, ax = 0.f;
                if (t >= 1 && j >= NN && j < g.nxp - NN && t >= HX2 - 2 * NN && t < HX2 + TX + 2 * NN) {
                    const int r2 = (h - 2 + 2 * NN) * RX2 + t, r3 = (h - 3 + 2 * NN) * RX2 + t;
                    const float d2 = a.cp.bcx[(ptrdiff_t)(h - 2) * g.cpld + j], d3 = a.cp.bcx[(ptrdiff_t)(h - 3) * g.cpld + j];
                    const float qj = d2 * lvx[r2];
                    const float qm = (j - 1 >= NN) ? a.cp.bcx[(ptrdiff_t)(h - 2) * g.cpld + j - 1] * lvx[r2 - 1] : 0.f;
                    az = d2 * lvz[r2] + d3 * lvz[r3] + 2.0f * (qm - qj);
                    ax = qj;
                }
                addz_r[t] = az; addx_r[t] = ax;
            }
            __syncthreads();
        }
        // ---- phase A: own-cell transpose of sponge + velocity update on tile + 2NN ring; m replaces lambda_v in the rects ----
        float4 QX, QZ;
        {
            float4 ax4 = zero4(), az4 = zero4();
            if (FS && tzi == 0) {
                if (gz == h) ax4 = ld4(addx_r + R.c0 + HX2);
                if (gz == h - 1) az4 = ld4(addz_r + R.c0 + HX2);
            }
            float4 wx, wz, mx, mz;
            ab_vel_cell(g, mt_, ld4(lvx + hv), ld4(lvz + hv), DMP, BX, BZ, ax4, az4, QX, QZ, wx, wz, mx, mz);
            st4(lvx + hv, mx); st4(lvz + hv, mz);
            GBX = add4(GBX, mul4(wx, EX));          // g_bx += q*dt*(D+x txx/dx + D-z txz/dz)
            GBZ = add4(GBZ, mul4(wz, EZ));
        }
        for (int i = tid; i < A::NRING2; i += NTH) {
            int r, gi;
            ring2_cell<NN>(i, r, gi);
            const int gzr = Z0 + r, gxr = X0 + 4 * gi;
            const int h2 = (r + 2 * NN) * RX2 + 4 * gi + HX2;
            const unsigned m = row_in<NN>(gzr, g.nzp) ? col_mask<NN>(gxr, g.nxp) : 0u;
            float4 mx = zero4(), mz = zero4();
            if (m) {
                const ptrdiff_t o = (ptrdiff_t)gzr * g.cpld + gxr;
                float4 ax4 = zero4(), az4 = zero4();
                if (FS && tzi == 0) {
                    if (gzr == h) ax4 = ld4(addx_r + 4 * gi + HX2);
                    if (gzr == h - 1) az4 = ld4(addz_r + 4 * gi + HX2);
                }
                float4 qx, qz, wx, wz;
                ab_vel_cell(g, m, ld4(lvx + h2), ld4(lvz + h2), ldk4(a.cp.bcx + o, pol), ldk4(a.cp.bx + o, pol), ldk4(a.cp.bz + o, pol), ax4, az4,
                            qx, qz, wx, wz, mx, mz);
            }
            st4(lvx + h2, mx); st4(lvz + h2, mz);
        }
        __syncthreads();
        // history of the stress update (own cell), in flight during phase B
        float4 D0 = zero4(), D1 = zero4(), D2 = zero4();
        if (cell_ok) {
            D0 = __ldcs(reinterpret_cast<const float4*>(H)); D1 = __ldcs(reinterpret_cast<const float4*>(H + g.plane));
            D2 = __ldcs(reinterpret_cast<const float4*>(H + 2 * g.plane));
        }
        // ---- phase B: lambda_tau += transposed operators of m on tile + NN ring, in place ----------------------------------
        for (int i = tid; i < NTH + G::NRING; i += NTH) {
            int r = R.r0, gi = R.l;
            if (i >= NTH) ring_cell<NN>(i - NTH, r, gi);
            const int h2 = (r + 2 * NN) * RX2 + 4 * gi + HX2, h1 = (r + NN) * RXH + 4 * gi + HX;
            float sxm[12], szm[12];
            ldseg_e(lvx + h2, sxm, gi - 1 >= -GX2, gi + 1 < NG + GX2); ldseg_e(lvz + h2, szm, gi - 1 >= -GX2, gi + 1 < NG + GX2);
            float4 wxb[2 * NN], wzf[2 * NN];
#pragma unroll
            for (int q = 0; q < 2 * NN; ++q) {
                wxb[q] = ld4(lvx + h2 + (q - NN + 1) * RX2);      // (D-z)^T of mx: rows i-NN+1 .. i+NN
                wzf[q] = ld4(lvz + h2 + (q - NN) * RX2);          // (D+z)^T of mz: rows i-NN .. i+NN-1
            }
            st4(ltxx + h1, add4(ld4(ltxx + h1), muls(xgath<NN, 0>(sxm, g.c), g.rdx)));
            st4(ltxz + h1, add4(ld4(ltxz + h1), add4(muls(zgath<NN>(wxb, g.c), g.rdz), muls(xgath<NN, 1>(szm, g.c), g.rdx))));
            st4(ltzz + h1, add4(ld4(ltzz + h1), muls(zgath<NN>(wzf, g.c), g.rdz)));
        }
        // phase C is an own-cell operation on what THIS thread just stored (tile cell and ring cell alike): no barrier between the
        // two, except around the free-surface transposes of the top tile row, which mix rows
        if (FS && tzi == 0) {    // 4T: transpose of the free-surface stress edits, on the state cotangents (rows h-2, h-3 end up zero)
            __syncthreads();
            const int t = tid;
            if (t < RXH) {
                const int j = X0 + t - HX;
                if (j >= NN && j < g.nxp - NN) {
                    const int rh1 = (2 * NN) * RXH + t, rh = (2 * NN + 1) * RXH + t, rh2 = (2 * NN - 1) * RXH + t, rh3 = (2 * NN - 2) * RXH + t;
                    ltxz[rh] -= ltxz[rh3];
                    ltzz[rh] -= ltzz[rh2];
                    ltxz[rh1] -= ltxz[rh2];
                    ltzz[rh1] = 0.f;
                    ltxz[rh2] = 0.f; ltzz[rh2] = 0.f; ltxz[rh3] = 0.f;
                }
            }
            __syncthreads();
        }
        // ---- phase C: new lambda_tau out; own-cell transpose of the stress update on tile + NN ring; n replaces lambda_tau ----
        {
            const float4 mxx = ld4(ltxx + hs), mzz = ld4(ltzz + hs), mxz = ld4(ltxz + hs);
            if (cell_ok) {
                float* P = a.planes + (size_t)s * g.plane + (size_t)gz * g.ld + gx + (size_t)wbase * fp;
                st4(P + A_TXX * fp, mxx); st4(P + A_TZZ * fp, mzz); st4(P + A_TXZ * fp, mxz);
                if (a.g_src && mt_ != 0u && s_sz[s - s_lo] == gz) {       // 3T
                    const int dc = s_sx[s - s_lo] - gx;
                    if (dc >= 0 && dc < 4 && ((mt_ >> dc) & 1u)) {
                        const float* Mt = a.mt + (size_t)s * 9;
                        const float sc = (float)(-1.0 / 3.0);
                        a.g_src[(size_t)s * g.nt + a.it] = sc * (Mt[0] * comp4(mxx, dc) + Mt[8] * comp4(mzz, dc) + Mt[2] * comp4(mxz, dc));
                    }
                }
            }
            const float4 qx = sel4(mt_, muls(mxx, g.dt), zero4()), qz = sel4(mt_, muls(mzz, g.dt), zero4()), qs = sel4(mt_, muls(mxz, g.dt), zero4());
            G11 = add4(G11, muls(mul4(qx, D0), g.rdx));
            G13 = add4(G13, add4(muls(mul4(qx, D1), g.rdz), muls(mul4(qz, D0), g.rdx)));
            G33 = add4(G33, muls(mul4(qz, D1), g.rdz));
            G55 = add4(G55, mul4(qs, D2));
            st4(ltxx + hs, add4(mul4(qx, C11), mul4(qz, C13)));
            st4(ltzz + hs, add4(mul4(qx, C13), mul4(qz, C33)));
            st4(ltxz + hs, mul4(qs, C55));
        }
        for (int i = tid; i < G::NRING; i += NTH) {
            int r, gi;
            ring_cell<NN>(i, r, gi);
            const int gzr = Z0 + r, gxr = X0 + 4 * gi;
            const int h1 = (r + NN) * RXH + 4 * gi + HX;
            const unsigned m = row_in<NN>(gzr, g.nzp) ? col_mask<NN>(gxr, g.nxp) : 0u;
            float4 nA = zero4(), nB = zero4(), nS = zero4();
            if (m) {
                const ptrdiff_t o = (ptrdiff_t)gzr * g.cpld + gxr;
                const float4 c11 = ldk4(a.cp.c11 + o, pol), c13 = ldk4(a.cp.c13 + o, pol), c33 = ldk4(a.cp.c33 + o, pol), c55 = ldk4(a.cp.c55 + o, pol);
                const float4 qx = sel4(m, muls(ld4(ltxx + h1), g.dt), zero4()), qz = sel4(m, muls(ld4(ltzz + h1), g.dt), zero4());
                const float4 qs = sel4(m, muls(ld4(ltxz + h1), g.dt), zero4());
                nA = add4(mul4(qx, c11), mul4(qz, c13)); nB = add4(mul4(qx, c13), mul4(qz, c33)); nS = mul4(qs, c55);
            }
            st4(ltxx + h1, nA); st4(ltzz + h1, nB); st4(ltxz + h1, nS);
        }
        __syncthreads();
        // ---- phase D: lambda_v (pre-step) = transposed operators of n + q, to the other set ---------------------------------
        {
            float sa[12], ss[12];
            ldseg(ltxx + hs, sa); ldseg(ltxz + hs, ss);
            float4 wb[2 * NN], wd[2 * NN];
#pragma unroll
            for (int q = 0; q < 2 * NN; ++q) {
                wb[q] = ld4(ltzz + hs + (q - NN + 1) * RXH);      // (D-z)^T of nB
                wd[q] = ld4(ltxz + hs + (q - NN) * RXH);          // (D+z)^T of nS
            }
            const float4 nvx = add4(add4(muls(xgath<NN, 1>(sa, g.c), g.rdx), muls(zgath<NN>(wd, g.c), g.rdz)), QX);
            const float4 nvz = add4(add4(muls(zgath<NN>(wb, g.c), g.rdz), muls(xgath<NN, 0>(ss, g.c), g.rdx)), QZ);
            if (cell_ok) {
                float* P = a.planes + (size_t)s * g.plane + (size_t)gz * g.ld + gx + (size_t)wbase * fp;
                st4(P + A_VX * fp, nvx); st4(P + A_VZ * fp, nvz);
            }
        }
        fence_proxy_async();
        // as in ela_f: only the producer's warp waits for the other warps' phase D before it refills stage k
        if (tid >= NTH - 32) {
            asm volatile("bar.sync 1, %0;" ::"n"(NTH) : "memory");
            if (producer && pc.valid) { ab_issue<NN>(pc, smem, bar, k, th, th2, thh, g.ns, rbase, a.hist_len, a.tl); pc.next(g, a.w, ring); }
        } else {
            asm volatile("bar.arrive 1, %0;" ::"n"(NTH) : "memory");
        }
        stage = (stage + 1 == NSTAGE) ? 0 : stage + 1;
    }
    if (cell_ok) {
        float* gp = a.gpart + (size_t)chunk * 6 * g.plane + (size_t)gz * g.ld + gx;
        red4(gp, G11); red4(gp + g.plane, G13); red4(gp + 2 * g.plane, G33); red4(gp + 3 * g.plane, G55);
        red4(gp + 4 * g.plane, GBX); red4(gp + 5 * g.plane, GBZ);
    }
}

template <int NN, bool FS>
__global__ void __launch_bounds__(NTH, 2)
ela_b(const __grid_constant__ CUtensorMap th, const __grid_constant__ CUtensorMap th2, const __grid_constant__ CUtensorMap thh, const EGeom g, const ABArgs a)
{
    static_assert(RPT == 1, "ela_b: one float4 group per thread");
    static_assert(Geo<NN>::HX2 <= CPX && 2 * NN <= CPZ, "ela_b: the 2NN ring must stay inside the apron of the coefficient pack");
    extern __shared__ __align__(128) unsigned char smem[];
    uint64_t* bar = (uint64_t*)(smem + NSTAGE * GeoA<NN>::STAGE);
    int* s_sz = (int*)(bar + 8);
    int* s_sx = s_sz + CMAX;
    float* fsadd = (float*)(s_sx + CMAX);
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
        ab_tile<NN, FS>(&th, &th2, &thh, g, a, smem, bar, par, stage, pc, ring, s_sz, s_sx, fsadd, R, tid, tile, chunk, s_lo, s_hi, first);
        __syncthreads();
    }
}
