// Register-resident fast path of the batched step() for small systems
// (NODOFS D known at compile time, D <= LPC lanes per conductor).
//
// Same arithmetic as step_kernels.cuh (and therefore as the reference, file
// and line references there); what changes is where the numbers live:
//   * K_G2 stores, per element and per matrix ROW, the element-matrix values
//     themselves (S*dz/3, K/dz, A/2, dz/3*M, dz/6*M, load halves), so the
//     sweep does no per-entry divisions and every lane reads one contiguous
//     record with compile-time offsets;
//   * K_F2 keeps the lane's block row [L D U | rhs] in registers, prefetches
//     the next element's record while the current node is eliminated, and
//     publishes finished pivot rows (with the pivot's reciprocal) through
//     shared memory;
//   * K_B2 keeps the lane's factor row in registers, prefetches the next
//     node's row, and exchanges the solved unknowns through shared memory.
#pragma once

#include <type_traits>

#include "step_kernels.cuh"
#include "osc2_div.cuh"

// Row record r of one element in `gm` (fast path): D + 9 doubles.
struct GmSlots2 {
    int d, M, S;
    __host__ __device__ int width() const { return d + 9; }
    __host__ __device__ int sd(int r, int c) const { return r * (d + 9) + c; }      // S[r][c]*dz/3
    __host__ __device__ int kb(int r) const { return r * (d + 9) + d; }             // K[r]/dz
    __host__ __device__ int md(int r) const { return r * (d + 9) + d + 1; }         // dz/3*M[r]
    __host__ __device__ int mo(int r) const { return r * (d + 9) + d + 2; }         // dz/6*M[r]
    __host__ __device__ int ahd(int r) const { return r * (d + 9) + d + 3; }        // A[r][r]/2
    __host__ __device__ int aho(int r) const { return r * (d + 9) + d + 4; }        // A[r][offcol(r)]/2
    __host__ __device__ int lod(int r, int which) const { return r * (d + 9) + d + 5 + which; }
    __host__ __device__ int count() const { return d * (d + 9); }
};
enum { LOD_UP = 0, LOD_LO = 1, LOD_UP_PREV = 2, LOD_LO_PREV = 3 };

// ---------------------------------------------------------------------------
// K_G2: same statement order as osc2_gauss_kernel for SMAT, then converts the
// Gauss-point values into the element-matrix values (smc:1015-1216).
// ---------------------------------------------------------------------------
template <bool SMEM_S>
__global__ void osc2_gauss2_kernel(const osc2_topology *__restrict__ tp, StepDev p)
{
    OSC2_DYN_SMEM(sm_dyn);
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (long long)p.E * p.B) return;
    const size_t B = (size_t)p.B;
    const size_t b = (size_t)(idx % p.B);
    const int e = (int)(idx / p.B);
    const int M = p.M, Sn = p.S, d = p.d, E = p.E;
    const GmSlots2 sl{d, M, Sn};
    double *g = p.gm + ((size_t)e * p.NS) * B + b;
    static_assert(SMEM_S, "the fast path always stages SMAT in shared memory");
    double *Sb = sm_dyn + threadIdx.x;
    const size_t Ss = blockDim.x;
    const unsigned char *__restrict__ smap = p.smap;
#define SM_(r, c) Sb[(size_t)smap[(r) * d + (c)] * Ss]
#define FG(q, j) p.fl_gauss[((((size_t)(q) * M + (j)) * E) + e) * B + b]
#define GOUT(slot) g[(size_t)(slot) * B]
    for (int i = 0; i < p.s_slots; ++i) Sb[(size_t)i * Ss] = 0.0;
    const double dz = p.dz[(size_t)e * B + b];
    const double alpha = 0.0;
    const double cdiag = dz * (1. / 3. + alpha), coff = dz * (1. / 6. - alpha);

    for (int j = 0; j < M; ++j) {
        const int iv = j, ip = M + j, it = 2 * M + j;
        const double v = FG(FL_V, j), rho = FG(FL_RHO, j), c0 = FG(FL_C0, j);
        const double grun = FG(FL_GRUN, j);
        const double kd = dz * 1.0 * fabs(v) / 2.0;
        const double kb = kd / dz;
        const double ah_off[3] = {(1. / rho) / 2., (rho * (c0 * c0)) / 2., (grun * FG(FL_T, j)) / 2.};
        for (int q = 0; q < 3; ++q) {
            const int row = q * M + j;              // velocity, pressure, temperature equation
            GOUT(sl.ahd(row)) = v / 2.;
            GOUT(sl.aho(row)) = ah_off[q];
            GOUT(sl.kb(row)) = kb;
            GOUT(sl.md(row)) = cdiag;              // MMAT = 1 on the fluid equations
            GOUT(sl.mo(row)) = coff;
            for (int w = 0; w < 4; ++w) GOUT(sl.lod(row, w)) = 0.0;
        }
        const double svv = 2.0 * FG(FL_F, j) * fabs(v) / tp->ch_dh[j];
        SM_(iv, iv) = svv;
        SM_(ip, iv) = -svv * grun * rho * v;
        SM_(it, iv) = -svv / FG(FL_CV, j) * v;
    }
    for (int k = 0; k < p.nff; ++k) {
        const int ca = tp->ff[k][0], cb = tp->ff[k][1];
        const double pa = FG(FL_P, ca), pb = FG(FL_P, cb);
        double delta_p = fabs(pa - pb);
        if (delta_p < tp->delta_p_min) delta_p = tp->delta_p_min;
        const int cu = (pb < pa) ? ca : cb;
        const double vel = FG(FL_V, cu);
        const double peri_o = p.peri_ff[(((size_t)0 * p.nff + k) * E + e) * B + b];
        const double peri_c = p.peri_ff[(((size_t)1 * p.nff + k) * E + e) * B + b];
        const double K1 = peri_o * sqrt(2. * FG(FL_RHO, cu) / (tp->k_loc * delta_p));
        const double K2 = K1 * tp->lambda_v * vel;
        const double vl = vel * tp->lambda_v;
        const double K3 = K1 * (FG(FL_H, cu) + .5 * (vl * vl));
        p.k123[(((size_t)0 * p.nff + k) * E + e) * B + b] = K1;
        p.k123[(((size_t)1 * p.nff + k) * E + e) * B + b] = K2;
        p.k123[(((size_t)2 * p.nff + k) * E + e) * B + b] = K3;
        const double coef_htc = peri_o * p.htc_ff[(((size_t)0 * p.nff + k) * E + e) * B + b]
                              + peri_c * p.htc_ff[(((size_t)1 * p.nff + k) * E + e) * B + b];
        for (int side = 0; side < 2; ++side) {
            const int c1 = side ? cb : ca, c2 = side ? ca : cb;
            const int v1 = c1, p1 = M + c1, t1 = 2 * M + c1, p2 = M + c2, t2 = 2 * M + c2;
            const double v = FG(FL_V, c1), rho = FG(FL_RHO, c1), A = tp->ch_area[c1];
            const double phi = FG(FL_GRUN, c1), h = FG(FL_H, c1), cv = FG(FL_CV, c1);
            const double c0 = FG(FL_C0, c1), T = FG(FL_T, c1);
            const double s_vj_pj = (K1 * v - K2) / (A * rho);
            SM_(v1, p1) -= s_vj_pj;
            SM_(v1, p2) = s_vj_pj;
            const double cga = phi / A;
            const double s_pj_pj = cga * (K3 - v * K2 - (h - (v * v) / 2. - (c0 * c0) / phi) * K1);
            SM_(p1, p1) += s_pj_pj;
            SM_(p1, p2) = -s_pj_pj;
            const double s_pj_tj = cga * coef_htc;
            SM_(p1, t1) += s_pj_tj;
            SM_(p1, t2) = -s_pj_tj;
            const double crca = 1. / (rho * cv * A);
            const double s_tj_pj = crca * (K3 - v * K2 - (h - (v * v) / 2. - phi * cv * T) * K1);
            SM_(t1, p1) += s_tj_pj;
            SM_(t1, p2) = -s_tj_pj;
            const double s_tj_tj = crca * coef_htc;
            SM_(t1, t1) += s_tj_tj;
            SM_(t1, t2) = -s_tj_tj;
        }
    }
    for (int k = 0; k < p.nfs; ++k) {
        const int j = tp->fs[k][0], l = tp->fs[k][1];
        const int ip = M + j, it = 2 * M + j, is = 3 * M + l;
        const double A = tp->ch_area[j];
        const double cga = FG(FL_GRUN, j) / A;
        const double coef_htc = p.peri_fs[((size_t)k * E + e) * B + b] * p.htc_fs[((size_t)k * E + e) * B + b];
        const double s_pj_tj = cga * coef_htc;
        SM_(ip, it) += s_pj_tj;
        SM_(ip, is) = -s_pj_tj;
        const double crca = 1. / (FG(FL_RHO, j) * FG(FL_CV, j) * A);
        const double s_tj_tj = crca * coef_htc;
        SM_(it, it) += s_tj_tj;
        SM_(it, is) = -s_tj_tj;
        SM_(is, is) += coef_htc;
        SM_(is, it) = -coef_htc;
    }
    const bool first = (p.num_step == 1);
    const double dz_6 = dz / 6.;
    for (int k = 0; k < p.nss; ++k) {
        const int a = 3 * M + tp->ss[k][0], c = 3 * M + tp->ss[k][1];
        const double coef = p.peri_ss[((size_t)k * E + e) * B + b] * p.htc_ss[((size_t)k * E + e) * B + b];
        SM_(a, a) += coef;
        SM_(a, c) = -coef;
        SM_(c, c) += coef;
        SM_(c, a) = -coef;
    }
    // per-solid env heat (added to SVEC after build_svec, smc:950-1013); only one
    // environment interface per solid exists in valid decks, but keep the order
    for (int l = 0; l < Sn; ++l) {
        const int is = 3 * M + l;
        const double A = tp->sol_area[l], cs = tp->sol_costeta[l];
        const double rho = p.sol_gauss[(((size_t)0 * Sn + l) * E + e) * B + b];
        const double cp = p.sol_gauss[(((size_t)1 * Sn + l) * E + e) * B + b];
        const double kk = p.sol_gauss[(((size_t)2 * Sn + l) * E + e) * B + b];
        const double mm = A * rho * cp / cs;
        GOUT(sl.md(is)) = cdiag * mm;
        GOUT(sl.mo(is)) = coff * mm;
        GOUT(sl.kb(is)) = (A * kk / cs) / dz;
        GOUT(sl.ahd(is)) = 0.0;
        GOUT(sl.aho(is)) = 0.0;
        const size_t q1 = ((((size_t)0 * Sn + l) * E + e) * 2) * B + b;
        const size_t q2 = ((((size_t)1 * Sn + l) * E + e) * 2) * B + b;
        const double qs0 = p.qsource[((size_t)e * Sn + l) * B + b];
        const double qs1 = p.qsource[((size_t)(e + 1) * Sn + l) * B + b];
        double sv0 = p.sol_q[q1] - qs0, sv1 = p.sol_q[q2] - qs1;
        double sp0 = first ? p.sol_q[q1 + B] - qs0 : 0.0, sp1 = first ? p.sol_q[q2 + B] - qs1 : 0.0;
        for (int k = 0; k < p.nes; ++k) {
            if (tp->es[k] != l) continue;
            const double h0 = p.htc_es[(((size_t)k * 3 + 0) * E + e) * B + b];
            const double peri = p.peri_es[((size_t)k * E + e) * B + b];
            const bool rect = tp->es_choice2[k] && tp->is_rectangular;
            double coef_htc;
            if (rect) {
                const double h1 = p.htc_es[(((size_t)k * 3 + 1) * E + e) * B + b];
                const double h2 = p.htc_es[(((size_t)k * 3 + 2) * E + e) * B + b];
                coef_htc = 2. * tp->height * h0 + tp->width * (h1 + h2);
            } else {
                coef_htc = h0 * peri;
            }
            SM_(is, is) += coef_htc;
            if (!rect) {
                const double env_heat = (peri * h0) * p.t_env;
                sv0 += env_heat; sv1 += env_heat;
                if (first) { sp0 += env_heat; sp1 += env_heat; }
            }
        }
        GOUT(sl.lod(is, LOD_UP)) = (2. * sv0 + sv1) * dz_6;
        GOUT(sl.lod(is, LOD_LO)) = (sv0 + 2. * sv1) * dz_6;
        GOUT(sl.lod(is, LOD_UP_PREV)) = (2. * sp0 + sp1) * dz_6;
        GOUT(sl.lod(is, LOD_LO_PREV)) = (sp0 + 2. * sp1) * dz_6;
    }
    // S * dz / 3 (smc:1126-1163): 3 is a shared divisor, so its reciprocal is refined once (bit-identical
    // quotients, osc2_div.cuh); entries the topology cannot touch are +0 and need no arithmetic at all
    const Osc2Rcp r3 = osc2_rcp_prepare(3.0);
    for (int r = 0; r < d; ++r)
        for (int c = 0; c < d; ++c) {
            const int slot = smap[r * d + c];
            GOUT(sl.sd(r, c)) = slot ? osc2_div(Sb[(size_t)slot * Ss] * dz, r3) : 0.0;
        }
#undef SM_
#undef FG
#undef GOUT
}

// shared-memory doubles per conductor of the fast forward sweep
template <int D>
struct FastFwdSmem {
    static constexpr int PW = 2 * D + 2;                 // published row: D' | U | rhs | 1/pivot
    static constexpr int PUB = 2 * D * PW;               // two nodes (double buffer)
    static constexpr int XS = 3 * D;                     // old solution ring
    static constexpr int STRIP = 5 * D + 1;              // per-lane product strip with zero pads
    static constexpr int RAW = PUB + XS + D * STRIP;
    static constexpr int PER_COND = RAW + ((18 - (RAW % 16)) % 16);   // == 2 (mod 16): bank spread
};

// ---------------------------------------------------------------------------
// K_F2.  BE: theta == 1 (Backward Euler; the Known coefficients off the block
// diagonals vanish and, for D % 8 == 0, numpy's pairwise sum of the three
// surviving products is (pL + pD) + pU for every row).  CAPTURE: also write
// the pre-solve system in the reference band layout.
// ---------------------------------------------------------------------------
template <int D, int LPC, bool BE, bool CAPTURE>
__global__ void osc2_forward_fast_kernel(const osc2_topology *__restrict__ tp, StepDev p)
{
    OSC2_DYN_SMEM(sm_dyn);
    using SM = FastFwdSmem<D>;
    constexpr int PW = SM::PW, STRIP = SM::STRIP, RW = D + 9;
    constexpr bool CLOSED_KNOWN = BE && (D % 8 == 0);
    const int M = p.M, N = p.N;
    const size_t B = (size_t)p.B;
    const int cpb = blockDim.x / LPC;
    const int grp = threadIdx.x / LPC, r = threadIdx.x % LPC;
    const long long b_ll = (long long)blockIdx.x * cpb + grp;
    const bool act = (b_ll < (long long)p.B) && (r < D);
    const size_t b = act ? (size_t)b_ll : 0;
    const int rr_ = (r < D) ? r : 0;
    const long long Total = (long long)D * N;
    constexpr int Half = 2 * D;
    const double theta = tp->theta, omt = 1.0 - tp->theta;
    const Osc2Rcp rdt = osc2_rcp_prepare(p.dt);
    const bool first = (p.num_step == 1);
    const bool shift = (p.num_step > 1);
    int dt_flag = 0;
    osc2_rcp_check(rdt, dt_flag);

    double *base = sm_dyn + (size_t)grp * SM::PER_COND;
    double *pub = base;                                  // [2][D][PW]
    double *xs = base + SM::PUB;                         // [3][D]
    double *strip = xs + SM::XS + (size_t)rr_ * STRIP;   // [5D+1], products live at [D, 4D)

    // column of the off-diagonal AMAT entry of this row (smc:62-101)
    int kind = 3, j = 0, aoffcol = -1;
    if (r < M) { kind = 0; j = r; aoffcol = M + j; }
    else if (r < 2 * M) { kind = 1; j = r - M; aoffcol = j; }
    else if (r < 3 * M) { kind = 2; j = r - 2 * M; aoffcol = j; }
    else { kind = 3; j = r - 3 * M; }
    bool bc_first = false, bc_last = false;
    double val_first = 0.0, val_last = 0.0;
    if (act && kind < 3) {
        const bool fwd = tp->ch_forward[j] != 0;
        const int intial = tp->ch_intial[j];
#define FBC(q) p.fl_bc[((size_t)(q) * M + j) * B + b]
#define FND(q) p.fl_node[((size_t)(q) * M + j) * B + b]
        if (kind == 1) {
            if (intial == 1 || intial == 2) { if (fwd) { bc_first = true; val_first = FBC(0); } else { bc_last = true; val_last = FBC(0); } }
            if (intial == 1 || intial == 3) { if (fwd) { bc_last = true; val_last = FBC(1); } else { bc_first = true; val_first = FBC(1); } }
        } else if (kind == 0) {
            const double area = tp->ch_area[j];
            if (intial == 2) {
                if (fwd) { bc_last = true; val_last = FBC(5) / FND(3) / area; }
                else { bc_first = true; val_first = FBC(5) / FND(2) / area; }
            } else if (intial == 3) {
                if (fwd) { bc_first = true; val_first = FBC(4) / FND(2) / area; }
                else { bc_last = true; val_last = FBC(4) / FND(3) / area; }
            }
        } else {
            const double v_first = FND(0), v_last = FND(1);
            if (fwd) {
                if (v_first > 0) { bc_first = true; val_first = FBC(2); }
                if (v_last < 0) { bc_last = true; val_last = FBC(3); }
            } else {
                if (v_last < 0) { bc_last = true; val_last = FBC(2); }
                if (v_first > 0) { bc_first = true; val_first = FBC(3); }
            }
        }
#undef FBC
#undef FND
    }
    long long bad_row = -1;
    int range_flag = 0;     // set when a value leaves the exact range of osc2_div_flag

    // this lane's row record of one element
    struct Elem {
        double sd[D];
        double kb, md, mo, ahd, aho, lodu, lodl, lodup, lodlp;
    };
    const size_t gstep = (size_t)p.NS * B;                       // one element further
    const double *gptr = p.gm + (size_t)rr_ * RW * B + b;        // row record of element 0
    auto load_elem = [&](const double *g, Elem &el) {
#pragma unroll
        for (int c = 0; c < D; ++c) el.sd[c] = g[(size_t)c * B];
        el.kb = g[(size_t)(D + 0) * B];
        el.md = g[(size_t)(D + 1) * B];
        el.mo = g[(size_t)(D + 2) * B];
        el.ahd = g[(size_t)(D + 3) * B];
        el.aho = g[(size_t)(D + 4) * B];
        el.lodu = g[(size_t)(D + 5) * B];
        el.lodl = g[(size_t)(D + 6) * B];
        if (first) { el.lodup = g[(size_t)(D + 7) * B]; el.lodlp = g[(size_t)(D + 8) * B]; }
        else { el.lodup = 0.0; el.lodlp = 0.0; }
    };
    auto zero_elem = [&](Elem &el) {
#pragma unroll
        for (int c = 0; c < D; ++c) el.sd[c] = 0.0;
        el.md = el.mo = el.kb = el.ahd = el.aho = el.lodu = el.lodl = el.lodup = el.lodlp = 0.0;
    };

    Elem eL, eR, eN;
    zero_elem(eL); zero_elem(eR); zero_elem(eN);
    double x_next = 0.0, sl0_next = 0.0, sl1_next = 0.0;
    const size_t xstep = (size_t)D * B;
    const double *xptr = p.sysvar + (size_t)rr_ * B + b;                 // x[node][r]
    double *lodptr = p.syslod + (size_t)rr_ * 2 * B + b;                 // SYSLOD[node][r][0]
    double *spptr = p.spill + (size_t)rr_ * (2 * D + 2) * B + b;         // spill[node][r][0]
    if (act) {
        load_elem(gptr, eR);
        xs[0 * D + r] = xptr[0];
        xs[1 * D + r] = xptr[xstep];
        sl0_next = lodptr[0];
        sl1_next = lodptr[B];
    }
    if (r < D) for (int q = 0; q < STRIP; ++q) strip[q] = 0.0;   // zero pads, once
    double *xa = xs, *xb = xs + D, *xc = xs + 2 * D;             // nodes n-1, n, n+1 (rotating)
    {   // at n = 0: node -1 does not exist; ring holds nodes 0, 1 in slots 0, 1
        double *t = xa; xa = xc; xc = xb; xb = t;                 // xb = slot0 (node 0), xc = slot1 (node 1)
    }

    // One node of the sweep.  The first and the last node are compiled separately
    // (no left / right element, boundary rows possible), so the interior loop has
    // neither of those branches.
    auto node = [&](auto HL, auto HR, const int n) {
        constexpr bool hasL = decltype(HL)::value, hasR = decltype(HR)::value;
        double *pub_cur = pub + (size_t)(n & 1) * D * PW;
        const double *pub_prev = pub + (size_t)((n + 1) & 1) * D * PW;
        const long long i_row = (long long)n * D + r;
        double sl0 = sl0_next, sl1 = sl1_next;
        // ---- prefetch for node n+1: element n+1, x[n+2], SYSLOD row ---------
        if (act) {
            if (n + 1 < N - 1) load_elem(gptr + gstep, eN);
            if (n + 2 < N) x_next = xptr[2 * xstep];
            if (n + 1 < N) { sl0_next = lodptr[2 * xstep]; sl1_next = lodptr[2 * xstep + B]; }
        }
        __syncwarp();   // old-solution ring of this node visible; pub reuse ordered

        double rowL[D], rowD[D], rowU[D], rhs = 0.0;
        if (act) {
            // SYSLOD (smc:1165-1318); fluid rows carry zero loads
            if (shift) { sl1 = sl0; sl0 = 0.0; }
            if (hasL) { sl0 += eL.lodl; if (first) sl1 += eL.lodlp; }
            if (hasR) { sl0 += eR.lodu; if (first) sl1 += eR.lodup; }
            lodptr[0] = sl0;
            lodptr[B] = sl1;

            const bool is_bc = (!hasL && bc_first) || (!hasR && bc_last);
            if (is_bc) {
#pragma unroll
                for (int c = 0; c < D; ++c) { rowL[c] = 0.0; rowD[c] = (c == r) ? 1.0 : 0.0; rowU[c] = 0.0; }
                rhs = (!hasL) ? val_first : val_last;
            } else {
                double mx = 0.0;
                // MASMAT/dt is non-zero on the diagonal entry of each block only
                const double masL = hasL ? osc2_div_flag(eL.mo, rdt, range_flag) : 0.0;
                const double masD = osc2_div_flag((hasL ? eL.md : 0.0) + (hasR ? eR.md : 0.0), rdt, range_flag);
                const double masU = hasR ? osc2_div_flag(eR.mo, rdt, range_flag) : 0.0;
#pragma unroll
                for (int c = 0; c < D; ++c) {
                    const bool dg = (c == r);
                    const double avL = dg ? eL.ahd : (c == aoffcol ? eL.aho : 0.0);
                    const double avR = dg ? eR.ahd : (c == aoffcol ? eR.aho : 0.0);
                    // L block: lower-left of element n-1
                    {
                        double sys = 0.0;
                        if (hasL) {
                            const double fds = (-avL + (dg ? -eL.kb : 0.0)) + eL.sd[c] / 2.;
                            const double m = dg ? masL : 0.0;
                            sys = BE ? m + fds : m + theta * fds;
                            if (!CLOSED_KNOWN) strip[D + c] = (m - omt * fds) * xa[c];
                        } else if (!CLOSED_KNOWN) strip[D + c] = 0.0;
                        rowL[c] = sys;
                        { const double a_ = fabs(sys); if (a_ > mx) mx = a_; }
                    }
                    // D block: lower-right of element n-1 + upper-left of element n
                    {
                        double flx = 0.0, dif = 0.0, sor = 0.0;
                        if (hasL) { if (dg) dif += eL.kb; flx += avL; sor += eL.sd[c]; }
                        if (hasR) { if (dg) dif += eR.kb; flx += -avR; sor += eR.sd[c]; }
                        const double fds = (flx + dif) + sor;
                        const double m = dg ? masD : 0.0;
                        const double sys = BE ? m + fds : m + theta * fds;
                        if (!CLOSED_KNOWN) strip[2 * D + c] = (m - omt * fds) * xb[c];
                        rowD[c] = sys;
                        { const double a_ = fabs(sys); if (a_ > mx) mx = a_; }
                    }
                    // U block: upper-right of element n
                    {
                        double sys = 0.0;
                        if (hasR) {
                            const double fds = (avR + (dg ? -eR.kb : 0.0)) + eR.sd[c] / 2.;
                            const double m = dg ? masU : 0.0;
                            sys = BE ? m + fds : m + theta * fds;
                            if (!CLOSED_KNOWN) strip[3 * D + c] = (m - omt * fds) * xc[c];
                        } else if (!CLOSED_KNOWN) strip[3 * D + c] = 0.0;
                        rowU[c] = sys;
                        { const double a_ = fabs(sys); if (a_ > mx) mx = a_; }
                    }
                }
                // Known (smc:1366-1443)
                double known;
                if (CLOSED_KNOWN) {
                    // only MASMAT/dt * x_old survives; the three products sit 8k band
                    // positions apart, i.e. in one numpy accumulator (or its tail)
                    known = 0.0;
                    if (hasL) known = masL * xa[rr_];
                    known = hasL ? known + masD * xb[rr_] : masD * xb[rr_];
                    if (hasR) known = known + masU * xc[rr_];
                } else {
                    long long lo, hi;
                    if (i_row <= Half - 1) { lo = 0; hi = Half + i_row; }
                    else if (i_row >= Total - (Half - 1)) { lo = i_row - (Half - 1); hi = Total; }
                    else { lo = i_row - (Half - 1); hi = i_row + Half; }
                    const int Lfull = (int)(hi - lo);
                    // band position jj <-> strip index D + (lo + jj - (n-1)*D)
                    const double *w = strip + D + (int)(lo - (long long)(n - 1) * D);
                    if (Lfull < 8) {
                        known = 0.0;
                        for (int q = 0; q < Lfull; ++q) known += w[q];
                    } else {
                        double a0 = w[0], a1 = w[1], a2 = w[2], a3 = w[3], a4 = w[4], a5 = w[5], a6 = w[6], a7 = w[7];
                        int q = 8;
                        for (; q < Lfull - (Lfull % 8); q += 8) {
                            a0 += w[q]; a1 += w[q + 1]; a2 += w[q + 2]; a3 += w[q + 3];
                            a4 += w[q + 4]; a5 += w[q + 5]; a6 += w[q + 6]; a7 += w[q + 7];
                        }
                        known = ((a0 + a1) + (a2 + a3)) + ((a4 + a5) + (a6 + a7));
                        for (; q < Lfull; ++q) known += w[q];
                    }
                }
                known += BE ? sl0 : (theta * sl0 + omt * sl1);
                // row scaling (tsf:538-551)
                const Osc2Rcp rmx = osc2_rcp_prepare(mx);
                osc2_rcp_check(rmx, range_flag);
#pragma unroll
                for (int c = 0; c < D; ++c) {
                    rowL[c] = osc2_div_flag(rowL[c], rmx, range_flag);
                    rowD[c] = osc2_div_flag(rowD[c], rmx, range_flag);
                    rowU[c] = osc2_div_flag(rowU[c], rmx, range_flag);
                }
                rhs = osc2_div_flag(known, rmx, range_flag);
            }
            if (CAPTURE) {
#pragma unroll
                for (int c = 0; c < 3 * D; ++c) {
                    const long long col = (long long)(n - 1) * D + c;
                    if (col < 0 || col >= Total) continue;
                    const long long kb_ = (2 * D - 1) + (col - i_row);
                    const double v = (c < D) ? rowL[c % D] : (c < 2 * D ? rowD[c % D] : rowU[c % D]);
                    p.cap_sysmat[((size_t)kb_ * Total + i_row) * B + b] = v;
                }
                p.cap_known[(size_t)i_row * B + b] = rhs;
            }

            // pivots of node n-1 (gredub / gbacsb forward part)
            if (hasL) {
#pragma unroll
                for (int k = 0; k < D; ++k) {
                    const double *pk = pub_prev + k * PW;
                    const double q = osc2_div_flag(rowL[k], osc2_rcp_make(pk[k], pk[2 * D + 1]), range_flag);
#pragma unroll
                    for (int c = k + 1; c < D; ++c) rowL[c] = rowL[c] - pk[c] * q;
#pragma unroll
                    for (int c = 0; c < D; ++c) rowD[c] = rowD[c] - pk[D + c] * q;
                    rhs = rhs - pk[2 * D] * q;
                }
            }
        } else {
#pragma unroll
            for (int c = 0; c < D; ++c) { rowL[c] = 0.0; rowD[c] = 0.0; rowU[c] = 0.0; }
        }
        // pivots inside node n
        Osc2Rcp my_rcp = osc2_rcp_make(1.0, 1.0);
#pragma unroll
        for (int k = 0; k < D; ++k) {
            if (act && r == k) {
                double *pk = pub_cur + k * PW;
#pragma unroll
                for (int c = k; c < D; ++c) pk[c] = rowD[c];
#pragma unroll
                for (int c = 0; c < D; ++c) pk[D + c] = rowU[c];
                pk[2 * D] = rhs;
                my_rcp = osc2_rcp_prepare(rowD[k]);
                pk[2 * D + 1] = my_rcp.y;
                if (fabs(rowD[k]) <= OSC2_TINY_PIVOT) { if (bad_row < 0) bad_row = i_row; }
                else osc2_rcp_check(my_rcp, range_flag);
            }
            __syncwarp();
            {
                // rows above the pivot (and idle lanes) take the multiplier 0: x - p * 0 leaves x unchanged
                // bit for bit, and ONE select on the multiplier replaces a select per updated entry
                const double *pk = pub_cur + k * PW;
                int flag_k = 0;
                double q = osc2_div_flag(rowD[k], osc2_rcp_make(pk[k], pk[2 * D + 1]), flag_k);
                const bool on = act && r > k;
                if (on) range_flag |= flag_k;
                q = on ? q : 0.0;
#pragma unroll
                for (int c = k + 1; c < D; ++c) rowD[c] = rowD[c] - pk[c] * q;
#pragma unroll
                for (int c = 0; c < D; ++c) rowU[c] = rowU[c] - pk[D + c] * q;
                rhs = rhs - pk[2 * D] * q;
            }
        }
        if (act) {
#pragma unroll
            for (int c = 0; c < D; ++c) spptr[(size_t)c * B] = rowD[c];
#pragma unroll
            for (int c = 0; c < D; ++c) spptr[(size_t)(D + c) * B] = rowU[c];
            spptr[(size_t)(2 * D) * B] = rhs;
            spptr[(size_t)(2 * D + 1) * B] = my_rcp.y;
            // rotate the element records and the old-solution ring
            eL = eR;
            eR = eN;
            if (n + 2 < N) xa[r] = x_next;     // slot of node n-1 is free now: holds node n+2
        }
        { double *t = xa; xa = xb; xb = xc; xc = t; }
        gptr += gstep;
        xptr += xstep;
        lodptr += 2 * xstep;
        spptr += (size_t)D * (2 * D + 2) * B;
    };
    node(std::false_type(), std::true_type(), 0);
    for (int n = 1; n < N - 1; ++n) node(std::true_type(), std::true_type(), n);
    node(std::true_type(), std::false_type(), N - 1);
    __syncwarp();
    double *scratch = base;
    if (r < D) { scratch[r] = (double)bad_row; scratch[D + r] = (double)(range_flag | dt_flag); }
    __syncwarp();
    if (act && r == 0) {
        long long worst = -1;
        bool range = false;
        for (int q = 0; q < D; ++q) {
            const long long v = (long long)scratch[q];
            if (v >= 0 && (worst < 0 || v < worst)) worst = v;
            range = range || (scratch[D + q] != 0.0);
        }
        // a singular pivot wins (the reference raises there); otherwise report the range problem
        p.status[b] = (worst >= 0) ? -(int)(worst + 1) : (range ? OSC2_STATUS_RANGE : 0);
    }
}

// ---------------------------------------------------------------------------
// K_B2: back substitution (gbacsb, tsf:772-802), node N-1 down to 0.  Every lane keeps the
// solved unknowns of the current and of the next node in registers; a freshly solved x travels
// to the other lanes of its conductor by warp shuffle, so the serial chain of a node
// (the subtractions of one row must follow the reference's column order) carries no
// shared-memory round trip.  All lanes run the same arithmetic on their own factor row; only
// lane rr's quotient of turn rr is kept.
// ---------------------------------------------------------------------------
#ifdef OSC2_EMU
inline double osc2_bcast(double v, int src_lane) { return emu_shfl(v, src_lane); }
#else
__device__ __forceinline__ double osc2_bcast(double v, int src_lane) { return __shfl_sync(0xffffffffu, v, src_lane); }
#endif

template <int D, int LPC>
__global__ void osc2_backward_fast_kernel(StepDev p)
{
    const int N = p.N;
    const size_t B = (size_t)p.B;
    const int cpb = blockDim.x / LPC;
    const int grp = threadIdx.x / LPC, r = threadIdx.x % LPC;
    const long long b_ll = (long long)blockIdx.x * cpb + grp;
    const bool act = (b_ll < (long long)p.B) && (r < D);
    const size_t b = act ? (size_t)b_ll : 0;
    const int lane0 = (threadIdx.x & 31) - r;            // first lane of this conductor's group

    constexpr int W = 2 * D + 2;      // D' | U | rhs | 1/pivot
    int range_flag = 0;
    double f[W], fn[W], xc[D], xn[D];
#pragma unroll
    for (int c = 0; c < W; ++c) { f[c] = 0.0; fn[c] = 0.0; }
#pragma unroll
    for (int c = 0; c < D; ++c) { xc[c] = 0.0; xn[c] = 0.0; }
    if (act) {
        const double *sp = p.spill + (((size_t)(N - 1) * D + r) * W) * B + b;
#pragma unroll
        for (int c = 0; c < W; ++c) fn[c] = sp[(size_t)c * B];
    } else {
        fn[r < D ? r : 0] = 1.0;      // idle lanes divide by 1
        fn[2 * D + 1] = 1.0;
    }
    for (int n = N - 1; n >= 0; --n) {
#pragma unroll
        for (int c = 0; c < W; ++c) f[c] = fn[c];
        if (act && n > 0) {
            const double *sp = p.spill + (((size_t)(n - 1) * D + r) * W) * B + b;
#pragma unroll
            for (int c = 0; c < W; ++c) fn[c] = sp[(size_t)c * B];
        }
        // products with the already known node n+1 are off the dependency chain
        double pu[D];
        const bool hasU = n < N - 1;
#pragma unroll
        for (int c = 0; c < D; ++c) pu[c] = hasU ? f[D + c] * xn[c] : 0.0;
        double mine = 0.0;
#pragma unroll
        for (int rr = D - 1; rr >= 0; --rr) {
            double q = f[2 * D];
#pragma unroll
            for (int c = rr + 1; c < D; ++c) q = q - f[c] * xc[c];
            if (hasU) {
#pragma unroll
                for (int c = 0; c < D; ++c) q = q - pu[c];
            }
            int flag = 0;
            const double x = osc2_div_flag(q, osc2_rcp_make(f[rr], f[2 * D + 1]), flag);
            if (r == rr) { mine = x; range_flag |= flag; }
            xc[rr] = osc2_bcast(x, lane0 + rr);
        }
        if (act) p.sysvar_new[((size_t)n * D + r) * B + b] = mine;
#pragma unroll
        for (int c = 0; c < D; ++c) xn[c] = xc[c];
    }
    if (act && range_flag && p.status[b] == 0) p.status[b] = OSC2_STATUS_RANGE;
}
