// Batched OPENSC2 step() kernels (sm_100a, FP64, no FMA contraction: build with -fmad=false).
//
// Reference being reproduced (every expression keeps its operand order):
//   step            transient_solution_functions.py:186-599
//   builders        step_matrix_construction.py:62-1013
//   element stencil step_matrix_construction.py:1015-1216
//   band assembly   step_matrix_construction.py:1218-1318
//   SYSMAT / Known  step_matrix_construction.py:1320-1474
//   boundary rows   fluid_component.py:117-565
//   gredub / gbacsb transient_solution_functions.py:602-808
//   norms, EQTEIG   transient_solution_functions.py:810-900
//
// The reference's banded matrix of half-width 2d-1 is block tridiagonal with
// d x d blocks (a node couples only to its neighbours); entries of the band
// outside the three blocks are structurally zero and stay zero in a no-pivot
// elimination, so working on [L D U] block rows performs exactly the
// reference's floating-point operations on every non-zero entry.
#pragma once

#include "osc2_dev.h"

#define OSC2_TINY_PIVOT 1.0e-20

#ifndef OSC2_DYN_SMEM
#define OSC2_DYN_SMEM(name) extern __shared__ double name[]
#endif

enum { FL_V = 0, FL_P, FL_T, FL_RHO, FL_C0, FL_GRUN, FL_H, FL_CV, FL_F };

// ---------------------------------------------------------------------------
// K_G: Gauss-point matrices, one thread per (element, conductor).
// SMAT is accumulated in the reference's statement order (`=` vs `+=`
// matters); it is staged per thread either in shared memory ([slot][thread],
// conflict free) or directly in the output record.
// ---------------------------------------------------------------------------
template <bool SMEM_S>
__global__ void osc2_gauss_kernel(const osc2_topology *__restrict__ tp, StepDev p)
{
    OSC2_DYN_SMEM(sm_dyn);
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (long long)p.E * p.B) return;
    const size_t B = (size_t)p.B;
    const size_t b = (size_t)(idx % p.B);
    const int e = (int)(idx / p.B);
    const int M = p.M, Sn = p.S, d = p.d, E = p.E;
    const GmSlots sl{d, M, Sn};
    double *g = p.gm + ((size_t)e * p.NS) * B + b;
    double *Sb;
    size_t Ss;
    if (SMEM_S) { Sb = sm_dyn + threadIdx.x; Ss = blockDim.x; }
    else { Sb = g; Ss = B; }
#define SM_(r, c) Sb[(size_t)((r) * d + (c)) * Ss]
#define FG(q, j) p.fl_gauss[((((size_t)(q) * M + (j)) * E) + e) * B + b]
#define GOUT(slot) g[(size_t)(slot) * B]
    for (int i = 0; i < d * d; ++i) Sb[(size_t)i * Ss] = 0.0;
    const double dz = p.dz[(size_t)e * B + b];

    for (int j = 0; j < M; ++j) {
        const int iv = j, ip = M + j, it = 2 * M + j;
        const double v = FG(FL_V, j), rho = FG(FL_RHO, j), c0 = FG(FL_C0, j);
        const double grun = FG(FL_GRUN, j);
        // build_amat (smc:62-101)
        GOUT(sl.a(j, 0)) = v;
        GOUT(sl.a(j, 1)) = 1. / rho;
        GOUT(sl.a(j, 2)) = rho * (c0 * c0);
        GOUT(sl.a(j, 3)) = grun * FG(FL_T, j);
        // build_kmat_fluid (smc:220-260), UPWEQT = 1
        const double kd = dz * 1.0 * fabs(v) / 2.0;
        GOUT(sl.k(iv)) = kd;
        GOUT(sl.k(ip)) = kd;
        GOUT(sl.k(it)) = kd;
        // build_smat_fluid (smc:262-311)
        const double svv = 2.0 * FG(FL_F, j) * fabs(v) / tp->ch_dh[j];
        SM_(iv, iv) = svv;
        SM_(ip, iv) = -svv * grun * rho * v;
        SM_(it, iv) = -svv / FG(FL_CV, j) * v;
    }
    // fluid-fluid interfaces (smc:103-218 transport coefficients, smc:313-557)
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
    // fluid-solid interfaces (smc:559-670)
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
    // solids (smc:672-730, 888-948)
    const bool first = (p.num_step == 1);
    for (int l = 0; l < Sn; ++l) {
        const int is = 3 * M + l;
        const double A = tp->sol_area[l], cs = tp->sol_costeta[l];
        const double rho = p.sol_gauss[(((size_t)0 * Sn + l) * E + e) * B + b];
        const double cp = p.sol_gauss[(((size_t)1 * Sn + l) * E + e) * B + b];
        const double kk = p.sol_gauss[(((size_t)2 * Sn + l) * E + e) * B + b];
        GOUT(sl.msol(l)) = A * rho * cp / cs;
        GOUT(sl.k(is)) = A * kk / cs;
        const size_t q1 = ((((size_t)0 * Sn + l) * E + e) * 2) * B + b;
        const size_t q2 = ((((size_t)1 * Sn + l) * E + e) * 2) * B + b;
        const double qs0 = p.qsource[((size_t)e * Sn + l) * B + b];
        const double qs1 = p.qsource[((size_t)(e + 1) * Sn + l) * B + b];
        GOUT(sl.sv(l, 0)) = p.sol_q[q1] - qs0;
        GOUT(sl.sv(l, 1)) = p.sol_q[q2] - qs1;
        GOUT(sl.svp(l, 0)) = first ? p.sol_q[q1 + B] - qs0 : 0.0;
        GOUT(sl.svp(l, 1)) = first ? p.sol_q[q2 + B] - qs1 : 0.0;
    }
    // solid-solid interfaces (smc:732-828)
    for (int k = 0; k < p.nss; ++k) {
        const int a = 3 * M + tp->ss[k][0], c = 3 * M + tp->ss[k][1];
        const double coef = p.peri_ss[((size_t)k * E + e) * B + b] * p.htc_ss[((size_t)k * E + e) * B + b];
        SM_(a, a) += coef;
        SM_(a, c) = -coef;
        SM_(c, c) += coef;
        SM_(c, a) = -coef;
    }
    // environment-solid interfaces (smc:830-886, 950-1013)
    for (int k = 0; k < p.nes; ++k) {
        const int l = tp->es[k], is = 3 * M + l;
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
        if (!rect) {  // the rectangular branch never adds env_heat (smc:987-1011)
            const double env_heat = (peri * h0) * p.t_env;
            GOUT(sl.sv(l, 0)) += env_heat;
            GOUT(sl.sv(l, 1)) += env_heat;
            if (first) {
                GOUT(sl.svp(l, 0)) += env_heat;
                GOUT(sl.svp(l, 1)) += env_heat;
            }
        }
    }
    if (SMEM_S) {
        for (int i = 0; i < d * d; ++i) g[(size_t)i * B] = Sb[(size_t)i * Ss];
    }
#undef SM_
#undef FG
#undef GOUT
}

// ---------------------------------------------------------------------------
// Thread-group helpers: LPC lanes own one conductor, lane r owns row r of the
// current block row.  LPC <= 32: groups live inside a warp; LPC == 64: one
// conductor per CTA.
// ---------------------------------------------------------------------------
template <int LPC>
__device__ __forceinline__ void osc2_group_sync()
{
    if (LPC <= 32) __syncwarp();
    else __syncthreads();
}

__host__ __device__ inline int osc2_odd(int x) { return x | 1; }

// Shared-memory doubles per conductor for the forward sweep.
__host__ __device__ inline int osc2_fwd_smem_per_conductor(int d)
{
    const int RS = osc2_odd(3 * d + 1);
    return osc2_odd(2 * d * RS + 3 * d + 15 * d + 1);
}
__host__ __device__ inline int osc2_bwd_smem_per_conductor(int d)
{
    const int PS = osc2_odd(2 * d + 1);
    return osc2_odd(d * PS + 2 * d);
}

// numpy pairwise sum (loops_utils.h.src) of a virtual array of length n whose
// entry j is strip[j - off] when 0 <= j - off < len and 0 otherwise.
__device__ inline double osc2_window_at(const double *strip, int len, long long off, long long jj)
{
    const long long c = jj - off;
    return (c >= 0 && c < len) ? strip[c] : 0.0;
}
__device__ inline double osc2_pairwise_window(const double *strip, int len, long long off, long long j0, long long n)
{
    if (n < 8) {
        double res = 0.0;
        for (long long i = 0; i < n; ++i) res += osc2_window_at(strip, len, off, j0 + i);
        return res;
    } else if (n <= 128) {
        double rr[8];
        for (int k = 0; k < 8; ++k) rr[k] = osc2_window_at(strip, len, off, j0 + k);
        long long i;
        for (i = 8; i < n - (n % 8); i += 8)
            for (int k = 0; k < 8; ++k) rr[k] += osc2_window_at(strip, len, off, j0 + i + k);
        double res = ((rr[0] + rr[1]) + (rr[2] + rr[3])) + ((rr[4] + rr[5]) + (rr[6] + rr[7]));
        for (; i < n; ++i) res += osc2_window_at(strip, len, off, j0 + i);
        return res;
    } else {
        long long n2 = n / 2;
        n2 -= n2 % 8;
        return osc2_pairwise_window(strip, len, off, j0, n2) + osc2_pairwise_window(strip, len, off, j0 + n2, n - n2);
    }
}

// ---------------------------------------------------------------------------
// K_F: block-row assembly + boundary rows + row scaling + no-pivot block
// elimination with forward substitution folded in.  Marches node by node;
// spills D' (upper part), U and the reduced right-hand side for K_B.
// ---------------------------------------------------------------------------
// y[i] = y[i] - x[i] * q for i < n, eight entries per trip: the row being updated and the pivot row are
// different rows of the shared-memory block, but the compiler cannot know that and would order every
// load behind the previous store -- one exposed shared-memory round trip per entry.  Same arithmetic
// per entry, so the result is unchanged bit for bit.
__device__ __forceinline__ void osc2_row_update(double *__restrict__ y, const double *__restrict__ x, const double q, const int n)
{
    int i = 0;
    for (; i + 8 <= n; i += 8) {
        double a[8], b[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) { a[u] = y[i + u]; b[u] = x[i + u]; }
#pragma unroll
        for (int u = 0; u < 8; ++u) y[i + u] = a[u] - b[u] * q;
    }
    for (; i < n; ++i) y[i] = y[i] - x[i] * q;
}

template <int LPC>
__global__ void osc2_forward_kernel(const osc2_topology *__restrict__ tp, StepDev p)
{
    OSC2_DYN_SMEM(sm_dyn);
    const int d = p.d, M = p.M, N = p.N;
    const size_t B = (size_t)p.B;
    const int cpb = blockDim.x / LPC;
    const int grp = threadIdx.x / LPC, r = threadIdx.x % LPC;
    const long long b_ll = (long long)blockIdx.x * cpb + grp;
    const bool act = (b_ll < (long long)p.B) && (r < d);
    const size_t b = act ? (size_t)b_ll : 0;
    const GmSlots sl{d, M, p.S};
    const int RS = osc2_odd(3 * d + 1);
    const long long Total = (long long)d * N;
    const int Half = 2 * d, Main = 2 * d - 1;
    const double dt = p.dt, theta = tp->theta;
    const bool first = (p.num_step == 1);

    double *base = sm_dyn + (size_t)grp * osc2_fwd_smem_per_conductor(d);
    double *rows = base;                    // [2][d][RS]
    double *xs = base + 2 * d * RS;         // [3][d] old solution at nodes n-1, n, n+1
    double *myacc = xs + 3 * d + 15 * (r < d ? r : 0);   // 8 accumulators + 7 tail slots

    // ---- per-lane row kind, AMAT pattern and boundary rows -----------------
    int kind = 3, j = 0, aoffcol = -1, aoffq = 0;
    if (r < M) { kind = 0; j = r; aoffcol = M + j; aoffq = 1; }
    else if (r < 2 * M) { kind = 1; j = r - M; aoffcol = j; aoffq = 2; }
    else if (r < 3 * M) { kind = 2; j = r - 2 * M; aoffcol = j; aoffq = 3; }
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
            if (intial == 2) {   // outlet velocity from MDTOUT
                if (fwd) { bc_last = true; val_last = FBC(5) / FND(3) / area; }
                else { bc_first = true; val_first = FBC(5) / FND(2) / area; }
            } else if (intial == 3) {   // inlet velocity from MDTIN
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

    if (act) {
        xs[0 * d + r] = p.sysvar[((size_t)0 * d + r) * B + b];
        xs[1 * d + r] = p.sysvar[((size_t)1 * d + r) * B + b];
    }

    for (int n = 0; n < N; ++n) {
        double *cur = rows + (size_t)(n & 1) * d * RS;
        const double *prev = rows + (size_t)((n + 1) & 1) * d * RS;
        double *row = cur + (size_t)(r < d ? r : 0) * RS;
        if (act && n + 1 < N && n >= 1)
            xs[((n + 1) % 3) * d + r] = p.sysvar[((size_t)(n + 1) * d + r) * B + b];
        osc2_group_sync<LPC>();

        double rhs = 0.0;
        if (act) {
            const bool hasL = n > 0, hasR = n < N - 1;
            const long long i_row = (long long)n * d + r;
            const double *gL = p.gm + ((size_t)(hasL ? n - 1 : 0) * p.NS) * B + b;   // only dereferenced if hasL
            const double *gR = p.gm + ((size_t)(hasR ? n : 0) * p.NS) * B + b;       // only dereferenced if hasR
#define GL(slot) gL[(size_t)(slot) * B]
#define GR(slot) gR[(size_t)(slot) * B]
            double dzL = 1.0, dzR = 1.0, mL = 0.0, mR = 0.0, kbL = 0.0, kbR = 0.0;
            double adL = 0.0, aoL = 0.0, adR = 0.0, aoR = 0.0;
            if (hasL) {
                dzL = p.dz[(size_t)(n - 1) * B + b];
                mL = (kind < 3) ? 1.0 : GL(sl.msol(j));
                kbL = GL(sl.k(r)) / dzL;
                if (kind < 3) { adL = GL(sl.a(j, 0)); aoL = GL(sl.a(j, aoffq)); }
            }
            if (hasR) {
                dzR = p.dz[(size_t)n * B + b];
                mR = (kind < 3) ? 1.0 : GR(sl.msol(j));
                kbR = GR(sl.k(r)) / dzR;
                if (kind < 3) { adR = GR(sl.a(j, 0)); aoR = GR(sl.a(j, aoffq)); }
            }
            const double alpha = 0.0;   // ALFA (tsf:265): consistent mass
            const double cdL = dzL * (1. / 3. + alpha), coL = dzL * (1. / 6. - alpha);
            const double cdR = dzR * (1. / 3. + alpha), coR = dzR * (1. / 6. - alpha);

            // SYSLOD (smc:1165-1318) -- done for every row, boundary rows included
            double sl0 = p.syslod[((size_t)i_row * 2 + 0) * B + b];
            double sl1 = p.syslod[((size_t)i_row * 2 + 1) * B + b];
            if (p.num_step > 1) { sl1 = sl0; sl0 = 0.0; }
            if (kind == 3) {
                if (hasL) {
                    const double dz_6 = dzL / 6.;
                    sl0 += (GL(sl.sv(j, 0)) + 2. * GL(sl.sv(j, 1))) * dz_6;
                    if (first) sl1 += (GL(sl.svp(j, 0)) + 2. * GL(sl.svp(j, 1))) * dz_6;
                }
                if (hasR) {
                    const double dz_6 = dzR / 6.;
                    sl0 += (2. * GR(sl.sv(j, 0)) + GR(sl.sv(j, 1))) * dz_6;
                    if (first) sl1 += (2. * GR(sl.svp(j, 0)) + GR(sl.svp(j, 1))) * dz_6;
                }
            }
            p.syslod[((size_t)i_row * 2 + 0) * B + b] = sl0;
            p.syslod[((size_t)i_row * 2 + 1) * B + b] = sl1;

            const bool is_bc = (n == 0 && bc_first) || (n == N - 1 && bc_last);
            if (is_bc) {
                for (int c = 0; c < 3 * d; ++c) row[c] = 0.0;
                row[d + r] = 1.0;
                rhs = (n == 0) ? val_first : val_last;
            } else {
                // Known (smc:1366-1443): numpy pairwise sum over the band positions
                long long lo, hi;
                if (i_row <= Half - 1) { lo = 0; hi = Half + i_row; }
                else if (i_row >= Total - (Half - 1)) { lo = i_row - (Half - 1); hi = Total; }
                else { lo = i_row - (Half - 1); hi = i_row + Half; }
                const int Lfull = (int)(hi - lo);
                // one band entry: SYSMAT value and the Known coefficient (smc:1343-1346, 1432-1441)
                auto entry = [&](int part, int c, double &sys, double &ck) {
                    const bool dg = (c == r);
                    double mas, flx, dif, sor;
                    if (part == 0) {
                        const double aval = dg ? adL : (c == aoffcol ? aoL : 0.0);
                        mas = dg ? coL * mL : 0.0;
                        flx = -(aval / 2.);
                        dif = dg ? -kbL : 0.0;
                        sor = (GL(sl.s(r, c)) * dzL / 3.) / 2.;
                    } else if (part == 2) {
                        const double aval = dg ? adR : (c == aoffcol ? aoR : 0.0);
                        mas = dg ? coR * mR : 0.0;
                        flx = aval / 2.;
                        dif = dg ? -kbR : 0.0;
                        sor = (GR(sl.s(r, c)) * dzR / 3.) / 2.;
                    } else {
                        mas = 0.0; flx = 0.0; dif = 0.0; sor = 0.0;
                        if (hasL) {
                            const double aval = dg ? adL : (c == aoffcol ? aoL : 0.0);
                            if (dg) { mas += cdL * mL; dif += kbL; }
                            flx += aval / 2.;
                            sor += GL(sl.s(r, c)) * dzL / 3.;
                        }
                        if (hasR) {
                            const double aval = dg ? adR : (c == aoffcol ? aoR : 0.0);
                            if (dg) { mas += cdR * mR; dif += kbR; }
                            flx += -(aval / 2.);
                            sor += GR(sl.s(r, c)) * dzR / 3.;
                        }
                    }
                    const double fds = (flx + dif) + sor;
                    sys = mas / dt + theta * fds;
                    ck = mas / dt - (1.0 - theta) * fds;
                };
                double known, mx = 0.0;
                if (Lfull <= 128) {
                    // single numpy block: 8 running accumulators + sequential tail, filled on the fly
                    const int nblk = (Lfull < 8) ? 0 : Lfull - (Lfull % 8);
                    for (int q = 0; q < 15; ++q) myacc[q] = 0.0;
                    double seq = 0.0;
                    for (int part = 0; part < 3; ++part) {
                        const bool present = (part == 0) ? hasL : (part == 2 ? hasR : true);
                        if (!present) {
                            for (int c = 0; c < d; ++c) row[part * d + c] = 0.0;
                            continue;
                        }
                        const double *xo = xs + ((n - 1 + part + 3) % 3) * d;
                        for (int c = 0; c < d; ++c) {
                            double sys, ck;
                            entry(part, c, sys, ck);
                            row[part * d + c] = sys;
                            const double a = fabs(sys);
                            if (a > mx) mx = a;
                            const double prod = ck * xo[c];
                            const int jpos = (int)((long long)(n - 1 + part) * d + c - lo);
                            if (Lfull < 8) seq += prod;
                            else if (jpos < nblk) myacc[jpos & 7] += prod;
                            else myacc[8 + (jpos - nblk)] = prod;
                        }
                    }
                    if (Lfull < 8) known = seq;
                    else {
                        known = ((myacc[0] + myacc[1]) + (myacc[2] + myacc[3])) + ((myacc[4] + myacc[5]) + (myacc[6] + myacc[7]));
                        for (int q = 0; q < Lfull - nblk; ++q) known += myacc[8 + q];
                    }
                } else {
                    // wide bands (d > 32): numpy splits the sum recursively; stage the products
                    // in the row strip, reduce, then overwrite the strip with the SYSMAT entries
                    for (int part = 0; part < 3; ++part) {
                        const bool present = (part == 0) ? hasL : (part == 2 ? hasR : true);
                        const double *xo = xs + ((n - 1 + part + 3) % 3) * d;
                        for (int c = 0; c < d; ++c) {
                            double sys = 0.0, ck = 0.0;
                            if (present) entry(part, c, sys, ck);
                            row[part * d + c] = present ? ck * xo[c] : 0.0;
                        }
                    }
                    const long long off = (long long)(n - 1) * d - lo;
                    known = osc2_pairwise_window(row, 3 * d, off, 0, Lfull);
                    for (int part = 0; part < 3; ++part) {
                        const bool present = (part == 0) ? hasL : (part == 2 ? hasR : true);
                        for (int c = 0; c < d; ++c) {
                            double sys = 0.0, ck = 0.0;
                            if (present) entry(part, c, sys, ck);
                            row[part * d + c] = sys;
                            const double a = fabs(sys);
                            if (a > mx) mx = a;
                        }
                    }
                }
                known += (theta * sl0 + (1.0 - theta) * sl1);
                // row scaling (tsf:538-551)
                for (int c = 0; c < 3 * d; ++c) row[c] = row[c] / mx;
                rhs = known / mx;
            }
            if (p.cap_sysmat) {
                for (int c = 0; c < 3 * d; ++c) {
                    const long long col = (long long)(n - 1) * d + c;
                    if (col < 0 || col >= Total) continue;
                    const long long kb = Main + (col - i_row);
                    p.cap_sysmat[((size_t)kb * Total + i_row) * B + b] = row[c];
                }
                p.cap_known[(size_t)i_row * B + b] = rhs;
            }

            // gredub + gbacsb forward part: pivots of node n-1 (rows final in `prev`)
            if (hasL) {
                for (int k = 0; k < d; ++k) {
                    const double *pk = prev + (size_t)k * RS;
                    const double q = row[k] / pk[d + k];
                    osc2_row_update(row + k + 1, pk + d + k + 1, q, d - k - 1);
                    osc2_row_update(row + d, pk + 2 * d, q, d);
                    rhs = rhs - pk[3 * d] * q;
                }
            }
#undef GL
#undef GR
        }
        // pivots inside node n: row k is final once lanes > k have seen rows < k
        for (int k = 0; k < d; ++k) {
            if (act && r == k) {
                row[3 * d] = rhs;
                if (fabs(row[d + r]) <= OSC2_TINY_PIVOT && bad_row < 0) bad_row = (long long)n * d + r;
            }
            osc2_group_sync<LPC>();
            if (act && r > k) {
                const double *pk = cur + (size_t)k * RS;
                const double q = row[d + k] / pk[d + k];
                osc2_row_update(row + d + k + 1, pk + d + k + 1, q, d - k - 1);
                osc2_row_update(row + 2 * d, pk + 2 * d, q, d);
                rhs = rhs - pk[3 * d] * q;
            }
        }
        if (act) {
            double *sp = p.spill + (((size_t)n * d + r) * (2 * d + 1)) * B + b;
            for (int c = 0; c < 2 * d + 1; ++c) sp[(size_t)c * B] = row[d + c];
        }
    }
    // first singular row of the conductor (gredub/gbacsb raise on abs(pivot) <= 1e-20)
    osc2_group_sync<LPC>();
    double *scratch = base;   // rows are dead now
    if (r < d) scratch[r] = (double)bad_row;
    osc2_group_sync<LPC>();
    if (act && r == 0) {
        long long worst = -1;
        for (int q = 0; q < d; ++q) {
            const long long v = (long long)scratch[q];
            if (v >= 0 && (worst < 0 || v < worst)) worst = v;
        }
        if (worst >= 0 && p.status[b] == 0) p.status[b] = -(int)(worst + 1);   // sticky: the first failing step is kept
    }
}

// ---------------------------------------------------------------------------
// K_B: back substitution (gbacsb, tsf:772-802), node N-1 down to 0.  Inside a
// node the rows depend on each other top-down, so lanes take turns.
// ---------------------------------------------------------------------------
template <int LPC>
__global__ void osc2_backward_kernel(StepDev p)
{
    OSC2_DYN_SMEM(sm_dyn);
    const int d = p.d, N = p.N;
    const size_t B = (size_t)p.B;
    const int cpb = blockDim.x / LPC;
    const int grp = threadIdx.x / LPC, r = threadIdx.x % LPC;
    const long long b_ll = (long long)blockIdx.x * cpb + grp;
    const bool act = (b_ll < (long long)p.B) && (r < d);
    const size_t b = act ? (size_t)b_ll : 0;
    const int PS = osc2_odd(2 * d + 1);
    double *base = sm_dyn + (size_t)grp * osc2_bwd_smem_per_conductor(d);
    double *fac = base;                 // [d][PS]
    double *xcur = base + d * PS;       // [d]
    double *xnext = xcur + d;           // [d]
    double *my = fac + (size_t)(r < d ? r : 0) * PS;

    for (int n = N - 1; n >= 0; --n) {
        if (act) {
            const double *sp = p.spill + (((size_t)n * d + r) * (2 * d + 1)) * B + b;
            for (int c = 0; c < 2 * d + 1; ++c) my[c] = sp[(size_t)c * B];
        }
        for (int rr = d - 1; rr >= 0; --rr) {
            if (act && r == rr) {
                double q = my[2 * d];
                for (int c = r + 1; c < d; ++c) q = q - my[c] * xcur[c];
                if (n < N - 1)
                    for (int c = 0; c < d; ++c) q = q - my[d + c] * xnext[c];
                const double x = q / my[r];
                xcur[r] = x;
                p.sysvar_new[((size_t)n * d + r) * B + b] = x;
            }
            osc2_group_sync<LPC>();
        }
        if (act) xnext[r] = xcur[r];
        osc2_group_sync<LPC>();
    }
}

// ---------------------------------------------------------------------------
// K_N: norms and EQTEIG with the reference's observable quirks
// (tsf:574-591, 810-900): Known is squared in place, so the "change" is
// x^2 - x and EQTEIG[pressure] takes the velocity column.
// ---------------------------------------------------------------------------
struct Osc2NormSolution { __device__ double operator()(double x, double) const { return x * x; } };
struct Osc2NormChange { __device__ double operator()(double x, double) const { const double c = x * x - x; return c * c; } };

template <class F>
__device__ double osc2_pairwise(const double *a, long long n, size_t stride, F f, double dt)
{
    // numpy pairwise summation (loops_utils.h.src, PW_BLOCKSIZE = 128)
    if (n < 8) {
        double res = 0.0;
        for (long long i = 0; i < n; ++i) res += f(a[(size_t)i * stride], dt);
        return res;
    } else if (n <= 128) {
        double r0 = f(a[0 * stride], dt), r1 = f(a[1 * stride], dt), r2 = f(a[2 * stride], dt), r3 = f(a[3 * stride], dt);
        double r4 = f(a[4 * stride], dt), r5 = f(a[5 * stride], dt), r6 = f(a[6 * stride], dt), r7 = f(a[7 * stride], dt);
        long long i;
        for (i = 8; i < n - (n % 8); i += 8) {
            r0 += f(a[(size_t)(i + 0) * stride], dt); r1 += f(a[(size_t)(i + 1) * stride], dt);
            r2 += f(a[(size_t)(i + 2) * stride], dt); r3 += f(a[(size_t)(i + 3) * stride], dt);
            r4 += f(a[(size_t)(i + 4) * stride], dt); r5 += f(a[(size_t)(i + 5) * stride], dt);
            r6 += f(a[(size_t)(i + 6) * stride], dt); r7 += f(a[(size_t)(i + 7) * stride], dt);
        }
        double res = ((r0 + r1) + (r2 + r3)) + ((r4 + r5) + (r6 + r7));
        for (; i < n; ++i) res += f(a[(size_t)i * stride], dt);
        return res;
    } else {
        long long n2 = n / 2;
        n2 -= n2 % 8;
        return osc2_pairwise(a, n2, stride, f, dt) + osc2_pairwise(a + (size_t)n2 * stride, n - n2, stride, f, dt);
    }
}

// numpy's pairwise sum of N values is a fixed binary tree whose leaves are runs
// of at most 128 values (8 running accumulators + sequential tail).  The tree
// only depends on N, so the leaves are summed by independent threads
// (K_N1: one thread per leaf, equation and conductor -- coalesced over the
// batch) and K_N2 replays the tree's additions in numpy's order from a postfix
// program built on the host (osc2_build_norm_plan in step_launch.h).
// part: [leaf][3: solution, change, eig max][d][B].
__global__ void osc2_norm_leaves_kernel(StepDev p)
{
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long per_leaf = (long long)p.d * p.B;
    if (idx >= per_leaf * p.n_leaves) return;
    const size_t B = (size_t)p.B;
    const int leaf = (int)(idx / per_leaf);
    const size_t qb = (size_t)(idx % per_leaf);          // q * B + b
    const int d = p.d;
    const size_t stride = (size_t)d * B;
    const int start = p.norm_leaf[2 * leaf], len = p.norm_leaf[2 * leaf + 1];
    const double *x = p.sysvar_new + (size_t)start * stride + qb;
    double *out = p.norm_part + ((size_t)leaf * 3) * stride + qb;
    out[0] = osc2_pairwise(x, len, stride, Osc2NormSolution(), p.dt);
    out[stride] = osc2_pairwise(x, len, stride, Osc2NormChange(), p.dt);
    double mxe = 0.0;
    for (int n = 0; n < len; ++n) {
        const double v = x[(size_t)n * stride];
        const double sq = v * v;
        const double eig = fabs((sq - v) / p.dt) / (fabs(sq) + 1.0e-5);
        if (n == 0 || eig > mxe) mxe = eig;
    }
    out[2 * stride] = mxe;
}

__global__ void osc2_norm_combine_kernel(StepDev p)
{
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (long long)p.d * p.B) return;
    const size_t B = (size_t)p.B;
    const size_t b = (size_t)(idx % p.B);
    const int q = (int)(idx / p.B);
    const int d = p.d, M = p.M;
    const size_t stride = (size_t)d * B;
    const double *part = p.norm_part + (size_t)q * B + b;
    for (int which = 0; which < 2; ++which) {
        double st[24];
        int sp = 0;
        for (int i = 0; i < p.n_prog; ++i) {
            const int op = p.norm_prog[i];
            if (op >= 0) st[sp++] = part[((size_t)op * 3 + which) * stride];
            else { st[sp - 2] = st[sp - 2] + st[sp - 1]; --sp; }
        }
        (which ? p.norm_change : p.norm_solution)[(size_t)q * B + b] = sqrt(st[0]);
    }
    const int src = (q >= M && q < 2 * M) ? q - M : q;   // pressure <- velocity (tsf:886-888)
    const double *pe = p.norm_part + (size_t)src * B + b + 2 * stride;
    double mxe = 0.0;
    for (int leaf = 0; leaf < p.n_leaves; ++leaf) {
        const double eig = pe[(size_t)leaf * 3 * stride];
        if (leaf == 0 || eig > mxe) mxe = eig;
    }
    p.eqteig[(size_t)q * B + b] = mxe;
}

// ---------------------------------------------------------------------------
// Layout change between the reference's per-conductor arrays ([B][K], host
// order) and the batch-innermost device layout ([K][B]).  32x32 tiles through
// shared memory so both sides are coalesced.
// ---------------------------------------------------------------------------
__global__ void osc2_transpose_kernel(const double *__restrict__ in, double *__restrict__ out,
                                      long long rows, long long cols)
{
    // in: [rows][cols] -> out: [cols][rows]
    __shared__ double tile[32][33];
    const long long tiles_c = (cols + 31) / 32;
    const long long tr = (long long)blockIdx.x / tiles_c, tc = (long long)blockIdx.x % tiles_c;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;   // 32 x 8 threads
    for (int k = ty; k < 32; k += 8) {
        const long long rr = tr * 32 + k, cc = tc * 32 + tx;
        if (rr < rows && cc < cols) tile[k][tx] = in[(size_t)rr * cols + cc];
    }
    __syncthreads();
    for (int k = ty; k < 32; k += 8) {
        const long long cc = tc * 32 + k, rr = tr * 32 + tx;
        if (rr < rows && cc < cols) out[(size_t)cc * rows + rr] = tile[tx][k];
    }
}
