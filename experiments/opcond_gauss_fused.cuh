// K_PG: operating conditions at the Gauss points (K_P, props_kernels.cuh) fused with the
// Gauss-point matrices / element row records (K_G, osc2_gauss2_kernel), for small systems.
//
// One CTA = 32 conductors x a few warps working on ONE element.  The ~60 coefficient values of an
// element never go to HBM: phase 1 evaluates them component by component (the host balances the
// components over the warps by their pow() count), phase 1b the interface heat-transfer models, and phase 2 turns them
// into the row records K_F consumes, warp w building matrix rows w, w + n_warps, ....  Everything a
// thread shares with the other warps of its conductor sits in shared memory as [slot][32 lanes],
// conflict free; global loads and stores are 256-byte rows over the batch.
//
// Arithmetic: phase 1 is the code of osc2_opcond_kernel, phase 2 the statements of
// osc2_gauss2_kernel filtered by matrix row -- a row is owned by one warp, the statements that
// touch it keep their reference order (`=` vs `+=`, smc:440-668), and a warp only executes the
// statements of its own row (the branch is warp uniform).
#pragma once

#include "props_kernels.cuh"
#include "step_kernels_fast.cuh"

struct FusedSlots {
    int M, S, nff, nfs, nss, nes, d;
    __host__ __device__ int fg(int q, int j) const { return q * M + j; }                       // [9][M]
    __host__ __device__ int sg(int q, int l) const { return 9 * M + q * S + l; }               // [3][S]
    __host__ __device__ int sq(int side, int l, int col) const { return 9 * M + 3 * S + (side * S + l) * 2 + col; }   // [2][S][2]
    __host__ __device__ int hff(int oc, int k) const { return 9 * M + 7 * S + oc * nff + k; }
    __host__ __device__ int hfs(int k) const { return 9 * M + 7 * S + 2 * nff + k; }
    __host__ __device__ int hss(int k) const { return 9 * M + 7 * S + 2 * nff + nfs + k; }
    __host__ __device__ int hes(int k, int q) const { return 9 * M + 7 * S + 2 * nff + nfs + nss + k * 3 + q; }
    __host__ __device__ int geo() const { return 9 * M + 7 * S + 2 * nff + nfs + nss + 3 * nes; }
    __host__ __device__ int dz() const { return geo(); }
    __host__ __device__ int pff(int oc, int k) const { return geo() + 1 + oc * nff + k; }
    __host__ __device__ int pfs(int k) const { return geo() + 1 + 2 * nff + k; }
    __host__ __device__ int pss(int k) const { return geo() + 1 + 2 * nff + nfs + k; }
    __host__ __device__ int pes(int k) const { return geo() + 1 + 2 * nff + nfs + nss + k; }
    __host__ __device__ int qs(int side, int l) const { return geo() + 1 + 2 * nff + nfs + nss + nes + side * S + l; }
    __host__ __device__ int aux() const { return geo() + 1 + 2 * nff + nfs + nss + nes + 2 * S; }
    __host__ __device__ int tf(int j) const { return aux() + j; }
    __host__ __device__ int hst(int j) const { return aux() + M + j; }
    __host__ __device__ int rcp(int j) const { return aux() + 2 * M + j; }
    __host__ __device__ int ts(int l) const { return aux() + 3 * M + l; }
    __host__ __device__ int count() const { return aux() + 3 * M + S; }
    // shared-memory doubles per CTA: inputs + one S row per warp
    __host__ __device__ size_t smem_doubles(int n_warps) const { return ((size_t)count() + (size_t)n_warps * d) * 32; }
};

__global__ void __launch_bounds__(128, 6) osc2_opcond_gauss_kernel(const osc2_topology *__restrict__ tp, StepDev p, PropsDev q,
                                                                 int materialize)
{
    OSC2_DYN_SMEM(sm_dyn);
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
    const int tiles = (p.B + 31) / 32;
    const int e = (int)(blockIdx.x / tiles);
    const long long b_ll = (long long)(blockIdx.x % tiles) * 32 + lane;
    const bool act = b_ll < (long long)p.B;
    const size_t B = (size_t)p.B;
    const size_t b = act ? (size_t)b_ll : 0;
    const int M = p.M, S = p.S, d = p.d, E = p.E, N = p.N;
    const FusedSlots sl{M, S, p.nff, p.nfs, p.nss, p.nes, d};
    const osc2_model *md = q.model;
    double *in = sm_dyn + lane;                                   // IN(slot) = in[slot * 32]
    double *srow = sm_dyn + (size_t)sl.count() * 32 + (size_t)w * d * 32 + lane;   // SR(c) = srow[c * 32]
#define IN(slot) in[(size_t)(slot) * 32]
#define SR(c) srow[(size_t)(c) * 32]
    const double *x0 = p.sysvar + ((size_t)e * d) * B + b;        // node e
    const double *x1 = x0 + (size_t)d * B;                        // node e + 1

    // ---------------------------------------------------------------- phase 1: components
    for (int item = 0; item < M + S + 1; ++item) {
        if (q.item_owner[item] != w) continue;      // warp uniform
        if (item < M) {
            const int j = item;
            const FluidTableDev &tab = q.fluid[md->ch_fluid[j]];
            const double v = (x0[(size_t)j * B] + x1[(size_t)j * B]) / 2.0;
            const double pr = (x0[(size_t)(M + j) * B] + x1[(size_t)(M + j) * B]) / 2.0;
            const double T = (x0[(size_t)(2 * M + j) * B] + x1[(size_t)(2 * M + j) * B]) / 2.0;
            const osc2p::Cell c = osc2p::locate(tab, pr, T);
            const double rho = osc2p::lookup(tab, c, OSC2_FP_RHO), mu = osc2p::lookup(tab, c, OSC2_FP_MU);
            const double cv = osc2p::lookup(tab, c, OSC2_FP_CV), kk = osc2p::lookup(tab, c, OSC2_FP_K);
            const double re = fabs(tp->ch_dh[j] * v * rho / mu);
            const double grun = osc2p::lookup(tab, c, OSC2_FP_BETA) / osc2p::lookup(tab, c, OSC2_FP_KAPPA) / cv / rho;
            const int hcorr = md->ch_htc_corr[j];
            const double nu = (hcorr == 212) ? osc2p::nusselt_212(re)
                              : (hcorr == 1) ? osc2p::nusselt_1(re, osc2p::lookup(tab, c, OSC2_FP_PRANDTL))
                                             : osc2p::nusselt_more(hcorr, re, osc2p::lookup(tab, c, OSC2_FP_PRANDTL), tp->ch_dh[j]);
            const double ftot = osc2p::friction_total(md, tp, j, re, q.cb_iters[(size_t)j * B + b]);
            IN(sl.fg(0, j)) = v; IN(sl.fg(1, j)) = pr; IN(sl.fg(2, j)) = T; IN(sl.fg(3, j)) = rho;
            IN(sl.fg(4, j)) = osc2p::lookup(tab, c, OSC2_FP_SOUND);
            IN(sl.fg(5, j)) = grun;
            IN(sl.fg(6, j)) = osc2p::lookup(tab, c, OSC2_FP_H);
            IN(sl.fg(7, j)) = cv;
            IN(sl.fg(8, j)) = ftot;
            IN(sl.tf(j)) = T;
            IN(sl.hst(j)) = nu * kk / tp->ch_dh[j];
            IN(sl.rcp(j)) = kk * rho * osc2p::lookup(tab, c, OSC2_FP_CP);
            if (e == 0 && act) {   // nodal values the boundary rows look at
                double *fn = const_cast<double *>(p.fl_node);
                const double *xl = p.sysvar + ((size_t)(N - 1) * d) * B + b;
                fn[((size_t)0 * M + j) * B + b] = x0[(size_t)j * B];
                fn[((size_t)1 * M + j) * B + b] = xl[(size_t)j * B];
                const osc2p::Cell cf = osc2p::locate(tab, x0[(size_t)(M + j) * B], x0[(size_t)(2 * M + j) * B]);
                const osc2p::Cell cl = osc2p::locate(tab, xl[(size_t)(M + j) * B], xl[(size_t)(2 * M + j) * B]);
                fn[((size_t)2 * M + j) * B + b] = osc2p::lookup(tab, cf, OSC2_FP_RHO);
                fn[((size_t)3 * M + j) * B + b] = osc2p::lookup(tab, cl, OSC2_FP_RHO);
            }
        } else if (item < M + S) {
            const int l = item - M, is = 3 * M + l;
            const double T = fabs(x0[(size_t)is * B] + x1[(size_t)is * B]) / 2.0;
            const double biss = q.b_field[((size_t)l * 2 + 0) * B + b], boss = q.b_field[((size_t)l * 2 + 1) * B + b];
            const double bg = fabs(osc2p::linspace_at(biss, boss, N, e) + osc2p::linspace_at(biss, boss, N, e + 1)) / 2.0;
            const int nmat = md->sol_nmat[l];
            const int kind = md->sol_kind[l];
            double tc = 0.0, tcs = 0.0;
            if (kind == OSC2_SOLID_STRAND_MIXED) {
                const double eps = q.eps[(size_t)l * B + b];
                bool nbti = false;
                for (int i = 0; i < md->sol_nmat[l]; ++i) nbti = nbti || (md->sol_mat[l][i] == OSC2_MAT_NBTI);
                tc = nbti ? osc2p::tc_nbti(bg, md->sol_bc20m[l], md->sol_tc0m[l])
                          : osc2p::tc_nb3sn(bg, (eps + eps) / 2.0, md->sol_tc0m[l], md->sol_bc20m[l]);
                tcs = q.tcs ? q.tcs[((size_t)l * E + e) * B + b] : 0.0;
            }
            const double tmax = q.tmax[(size_t)l * B + b], tmin = q.tmin[(size_t)l * B + b];
            double nsum = 0.0, cps = 0.0, ksum = 0.0, rho1 = 0.0, cp1 = 0.0, k1 = 0.0;
            for (int i = 0; i < nmat; ++i) {
                const int mat = md->sol_mat[l][i];
                const double wgt = md->sol_weight[l][i];
                const double rho_i = osc2p::mat_density(mat);
                double cp_i, k_i;
                osc2p::material_cp_k(mat, T, bg, tcs, tc, tmin, tmax, md->sol_rrr[l], md->sol_tc0m[l], q.sol_const[2 * l],
                                     q.sol_const[2 * l + 1], cp_i, k_i);
                const double num = rho_i * wgt;
                if (i == 0) { nsum = num; cps = cp_i * num; ksum = k_i * wgt; rho1 = rho_i; cp1 = cp_i; k1 = k_i; }
                else { nsum = nsum + num; cps = cps + cp_i * num; ksum = ksum + k_i * wgt; }
            }
            double rho, cp, kk;
            if ((kind == OSC2_SOLID_JACKET && nmat == 1) || kind == OSC2_SOLID_STRAND_STAB) { rho = rho1; cp = cp1; kk = k1; }
            else { rho = nsum / md->sol_weight_sum[l]; cp = cps / nsum; kk = ksum / md->sol_weight_sum[l]; }
            IN(sl.sg(0, l)) = rho; IN(sl.sg(1, l)) = cp; IN(sl.sg(2, l)) = kk;
            IN(sl.ts(l)) = T;
            const double n0 = q.q_window[((size_t)l * 4 + 0) * B + b], n1 = q.q_window[((size_t)l * 4 + 1) * B + b];
            const double tb = q.q_window[((size_t)l * 4 + 2) * B + b], te = q.q_window[((size_t)l * 4 + 3) * B + b];
            const double q0 = q.q0[(size_t)l * B + b];
            const bool on_now = md->sol_heated[l] && q.time >= tb && q.time <= te;
            const bool on_t0 = md->sol_heated[l] && 0.0 >= tb && 0.0 <= te;
            for (int side = 0; side < 2; ++side) {
                const int node = e + side;
                const bool inw = node >= (int)n0 && node <= (int)n1;
                IN(sl.sq(side, l, 0)) = (on_now && inw) ? q0 : 0.0;
                IN(sl.sq(side, l, 1)) = (on_t0 && inw) ? q0 : 0.0;
                IN(sl.qs(side, l)) = p.qsource[((size_t)(e + side) * S + l) * B + b];
            }
        } else {
            // geometry of this element
            IN(sl.dz()) = p.dz[(size_t)e * B + b];
            for (int k = 0; k < p.nff; ++k) {
                IN(sl.pff(0, k)) = p.peri_ff[(((size_t)0 * p.nff + k) * E + e) * B + b];
                IN(sl.pff(1, k)) = p.peri_ff[(((size_t)1 * p.nff + k) * E + e) * B + b];
            }
            for (int k = 0; k < p.nfs; ++k) IN(sl.pfs(k)) = p.peri_fs[((size_t)k * E + e) * B + b];
            for (int k = 0; k < p.nss; ++k) IN(sl.pss(k)) = p.peri_ss[((size_t)k * E + e) * B + b];
            for (int k = 0; k < p.nes; ++k) IN(sl.pes(k)) = p.peri_es[((size_t)k * E + e) * B + b];
        }
    }
    __syncthreads();
    // ---------------------------------------------------------------- phase 1b: interface HTCs
    const int n_itf = p.nfs + p.nff + p.nss + p.nes;
    for (int item = w; item < n_itf; item += nw) {
        if (item < p.nfs) {
            const int k = item, j = tp->fs[k][0], l = tp->fs[k][1];
            double h;
            if (md->fs_choice[k] == 2) {
                const double Ts = IN(sl.ts(l)), Tf = IN(sl.tf(j));
                const double hK = 200.0 * (Ts + Tf) * (osc2p::sq(Ts) + osc2p::sq(Tf));
                const double tqbeg = q.q_window[((size_t)l * 4 + 2) * B + b];
                double full = 0.0;
                if (q.time > tqbeg) {
                    const double htr = sqrt(IN(sl.rcp(j)) / (3.141592653589793 * (q.time - tqbeg)));
                    full = (hK * htr) / (hK + htr);
                }
                h = fmax(IN(sl.hst(j)) * md->fs_mlt[k], full);
            } else {
                h = md->fs_htc[k] * md->fs_mlt[k];
            }
            IN(sl.hfs(k)) = h;
        } else if (item < p.nfs + p.nff) {
            const int k = item - p.nfs, ja = tp->ff[k][0], jb = tp->ff[k][1];
            double ho, hc;
            if (md->ff_choice[k] == 2) {
                const double h1 = IN(sl.hst(ja)), h2 = IN(sl.hst(jb));
                ho = md->ff_mlt[k] * h1 * h2 / (h1 + h2);
                const double cond = osc2p::k_ss((IN(sl.tf(ja)) + IN(sl.tf(jb))) / 2);
                const double rwall = md->ff_thick[k] / cond;
                hc = md->ff_mlt[k] / (1 / h1 + 1 / h2 + rwall);
            } else {
                ho = md->ff_htc[k] * md->ff_mlt[k];
                hc = ho;
            }
            IN(sl.hff(0, k)) = ho;
            IN(sl.hff(1, k)) = hc;
        } else if (item < p.nfs + p.nff + p.nss) {
            const int k = item - p.nfs - p.nff;
            double h;
            if (md->ss_choice[k] == 3 || md->ss_choice[k] == -3) {
                h = 0.0;
            } else if (md->ss_choice[k] == 1) {
                const double rr = md->ss_thick_rc[k] / IN(sl.sg(2, tp->ss[k][0]));
                const double rc = md->ss_thick_cr[k] / IN(sl.sg(2, tp->ss[k][1]));
                h = md->ss_mlt[k] * (1.0 / (rr + md->ss_rcontact[k] + rc));
            } else {
                h = md->ss_htc[k] * md->ss_mlt[k];
            }
            IN(sl.hss(k)) = h;
        } else {
            const int k = item - p.nfs - p.nff - p.nss;
            IN(sl.hes(k, 0)) = md->es_htc[k];
            IN(sl.hes(k, 1)) = 0.0;
            IN(sl.hes(k, 2)) = 0.0;
        }
    }
    __syncthreads();
    // radiative exchange between jackets: one warp walks the interfaces in order (two interfaces may feed
    // the same solid's source), see osc2_opcond_kernel
    if (w == 0) {
        for (int k = 0; k < p.nss; ++k) {
            const int ch = md->ss_choice[k];
            if (ch != 3 && ch != -3) continue;
            const int a = tp->ss[k][0], c = tp->ss[k][1];
            for (int side = 0; side < 2; ++side) {
                const double *xn = side ? x1 : x0;
                const double Ta = xn[(size_t)(3 * M + a) * B], Tc = xn[(size_t)(3 * M + c) * B];
                const double hr = osc2p::radiative_htc(md, k, Ta, Tc);
                const double ph = p.peri_ss_nodal[((size_t)k * N + (e + side)) * B + b] * hr;
                IN(sl.sq(side, a, 0)) += ph * (Tc - Ta);
                IN(sl.sq(side, c, 0)) += ph * (Ta - Tc);
            }
        }
    }
    __syncthreads();
    // optional: materialise the coefficient arrays (osc2_download_step_inputs / parity tests)
    if (materialize && act) {
        double *fg = const_cast<double *>(p.fl_gauss), *sg = const_cast<double *>(p.sol_gauss);
        double *sqv = const_cast<double *>(p.sol_q), *hff = const_cast<double *>(p.htc_ff);
        double *hfs = const_cast<double *>(p.htc_fs), *hss = const_cast<double *>(p.htc_ss), *hes = const_cast<double *>(p.htc_es);
        for (int s = w; s < sl.geo(); s += nw) {
            const double v = IN(s);
            if (s < 9 * M) fg[((size_t)s * E + e) * B + b] = v;                                         // [q][j] = s
            else if (s < 9 * M + 3 * S) sg[((size_t)(s - 9 * M) * E + e) * B + b] = v;
            else if (s < 9 * M + 7 * S) {
                const int t = s - 9 * M - 3 * S, col = t & 1, sl_ = t >> 1;                           // (side*S + l)*2 + col
                sqv[((((size_t)sl_) * E + e) * 2 + col) * B + b] = v;
            } else if (s < 9 * M + 7 * S + 2 * p.nff) hff[((size_t)(s - 9 * M - 7 * S) * E + e) * B + b] = v;
            else if (s < 9 * M + 7 * S + 2 * p.nff + p.nfs) hfs[((size_t)(s - 9 * M - 7 * S - 2 * p.nff) * E + e) * B + b] = v;
            else if (s < 9 * M + 7 * S + 2 * p.nff + p.nfs + p.nss) hss[((size_t)(s - 9 * M - 7 * S - 2 * p.nff - p.nfs) * E + e) * B + b] = v;
            else hes[((size_t)(s - 9 * M - 7 * S - 2 * p.nff - p.nfs - p.nss) * E + e) * B + b] = v;  // [k][3] = k*3 + q
        }
    }

    // ---------------------------------------------------------------- phase 2: row records
    const GmSlots2 gs{d, M, S};
    const double dz = IN(sl.dz());
    const double alpha = 0.0;
    const double cdiag = dz * (1. / 3. + alpha), coff = dz * (1. / 6. - alpha);
    const bool first = (p.num_step == 1);
    const double dz_6 = dz / 6.;
#define FG(qq, jj) IN(sl.fg(qq, jj))
    for (int r = w; r < d; r += nw) {
        double *g = p.gm + ((size_t)e * p.NS) * B + b;
#define GOUT(slot) g[(size_t)(slot) * B]
        for (int c = 0; c < d; ++c) SR(c) = 0.0;
        double kb, mdv, mov, ahd = 0.0, aho = 0.0, lod0 = 0.0, lod1 = 0.0, lod2 = 0.0, lod3 = 0.0;
        if (r < 3 * M) {
            const int kind = r / M, j = r % M;
            const int iv = j, ip = M + j, it = 2 * M + j;
            const double v = FG(FL_V, j), rho = FG(FL_RHO, j), c0 = FG(FL_C0, j), grun = FG(FL_GRUN, j);
            const double kd = dz * 1.0 * fabs(v) / 2.0;
            kb = kd / dz;
            mdv = cdiag;
            mov = coff;
            ahd = v / 2.;
            aho = (kind == 0) ? (1. / rho) / 2. : (kind == 1) ? (rho * (c0 * c0)) / 2. : (grun * FG(FL_T, j)) / 2.;
            const double svv = 2.0 * FG(FL_F, j) * fabs(v) / tp->ch_dh[j];
            if (kind == 0) SR(iv) = svv;
            else if (kind == 1) SR(iv) = -svv * grun * rho * v;
            else SR(iv) = -svv / FG(FL_CV, j) * v;
            for (int k = 0; k < p.nff; ++k) {
                const int ca = tp->ff[k][0], cb = tp->ff[k][1];
                if (ca != j && cb != j) continue;
                const double pa = FG(FL_P, ca), pb = FG(FL_P, cb);
                double delta_p = fabs(pa - pb);
                if (delta_p < tp->delta_p_min) delta_p = tp->delta_p_min;
                const int cu = (pb < pa) ? ca : cb;
                const double vel = FG(FL_V, cu);
                const double peri_o = IN(sl.pff(0, k)), peri_c = IN(sl.pff(1, k));
                const double K1 = peri_o * sqrt(2. * FG(FL_RHO, cu) / (tp->k_loc * delta_p));
                const double K2 = K1 * tp->lambda_v * vel;
                const double vl = vel * tp->lambda_v;
                const double K3 = K1 * (FG(FL_H, cu) + .5 * (vl * vl));
                if (kind == 0 && j == ca && act) {      // one owner per interface writes the transport coefficients
                    p.k123[(((size_t)0 * p.nff + k) * E + e) * B + b] = K1;
                    p.k123[(((size_t)1 * p.nff + k) * E + e) * B + b] = K2;
                    p.k123[(((size_t)2 * p.nff + k) * E + e) * B + b] = K3;
                }
                const double coef_htc = peri_o * IN(sl.hff(0, k)) + peri_c * IN(sl.hff(1, k));
                for (int side = 0; side < 2; ++side) {
                    const int c1 = side ? cb : ca, c2 = side ? ca : cb;
                    if (c1 != j) continue;
                    const int p1 = M + c1, t1 = 2 * M + c1, p2 = M + c2, t2 = 2 * M + c2;
                    const double A = tp->ch_area[c1];
                    const double phi = grun, h = FG(FL_H, c1), cv = FG(FL_CV, c1), T = FG(FL_T, c1);
                    if (kind == 0) {
                        const double s_vj_pj = (K1 * v - K2) / (A * rho);
                        SR(p1) -= s_vj_pj;
                        SR(p2) = s_vj_pj;
                    } else if (kind == 1) {
                        const double cga = phi / A;
                        const double s_pj_pj = cga * (K3 - v * K2 - (h - (v * v) / 2. - (c0 * c0) / phi) * K1);
                        SR(p1) += s_pj_pj;
                        SR(p2) = -s_pj_pj;
                        const double s_pj_tj = cga * coef_htc;
                        SR(t1) += s_pj_tj;
                        SR(t2) = -s_pj_tj;
                    } else {
                        const double crca = 1. / (rho * cv * A);
                        const double s_tj_pj = crca * (K3 - v * K2 - (h - (v * v) / 2. - phi * cv * T) * K1);
                        SR(p1) += s_tj_pj;
                        SR(p2) = -s_tj_pj;
                        const double s_tj_tj = crca * coef_htc;
                        SR(t1) += s_tj_tj;
                        SR(t2) = -s_tj_tj;
                    }
                }
            }
            if (kind > 0)
                for (int k = 0; k < p.nfs; ++k) {
                    if (tp->fs[k][0] != j) continue;
                    const int is = 3 * M + tp->fs[k][1];
                    const double A = tp->ch_area[j];
                    const double coef_htc = IN(sl.pfs(k)) * IN(sl.hfs(k));
                    if (kind == 1) {
                        const double s_pj_tj = (grun / A) * coef_htc;
                        SR(it) += s_pj_tj;
                        SR(is) = -s_pj_tj;
                    } else {
                        const double crca = 1. / (rho * FG(FL_CV, j) * A);
                        const double s_tj_tj = crca * coef_htc;
                        SR(it) += s_tj_tj;
                        SR(is) = -s_tj_tj;
                    }
                }
            (void)ip;
        } else {
            const int l = r - 3 * M, is = r;
            for (int k = 0; k < p.nfs; ++k) {
                if (tp->fs[k][1] != l) continue;
                const int it = 2 * M + tp->fs[k][0];
                const double coef_htc = IN(sl.pfs(k)) * IN(sl.hfs(k));
                SR(is) += coef_htc;
                SR(it) = -coef_htc;
            }
            for (int k = 0; k < p.nss; ++k) {
                const int a = 3 * M + tp->ss[k][0], c = 3 * M + tp->ss[k][1];
                if (a != is && c != is) continue;
                const double coef = IN(sl.pss(k)) * IN(sl.hss(k));
                if (a == is) { SR(a) += coef; SR(c) = -coef; }
                if (c == is) { SR(c) += coef; SR(a) = -coef; }
            }
            const double A = tp->sol_area[l], cs = tp->sol_costeta[l];
            const double mm = A * IN(sl.sg(0, l)) * IN(sl.sg(1, l)) / cs;
            mdv = cdiag * mm;
            mov = coff * mm;
            kb = (A * IN(sl.sg(2, l)) / cs) / dz;
            const double qs0 = IN(sl.qs(0, l)), qs1 = IN(sl.qs(1, l));
            double sv0 = IN(sl.sq(0, l, 0)) - qs0, sv1 = IN(sl.sq(1, l, 0)) - qs1;
            double sp0 = first ? IN(sl.sq(0, l, 1)) - qs0 : 0.0, sp1 = first ? IN(sl.sq(1, l, 1)) - qs1 : 0.0;
            for (int k = 0; k < p.nes; ++k) {
                if (tp->es[k] != l) continue;
                const double h0 = IN(sl.hes(k, 0));
                const double peri = IN(sl.pes(k));
                const bool rect = tp->es_choice2[k] && tp->is_rectangular;
                double coef_htc;
                if (rect) coef_htc = 2. * tp->height * h0 + tp->width * (IN(sl.hes(k, 1)) + IN(sl.hes(k, 2)));
                else coef_htc = h0 * peri;
                SR(is) += coef_htc;
                if (!rect) {
                    const double env_heat = (peri * h0) * p.t_env;
                    sv0 += env_heat; sv1 += env_heat;
                    if (first) { sp0 += env_heat; sp1 += env_heat; }
                }
            }
            lod0 = (2. * sv0 + sv1) * dz_6;
            lod1 = (sv0 + 2. * sv1) * dz_6;
            lod2 = (2. * sp0 + sp1) * dz_6;
            lod3 = (sp0 + 2. * sp1) * dz_6;
        }
        if (act) {
            for (int c = 0; c < d; ++c) GOUT(gs.sd(r, c)) = SR(c) * dz / 3.;
            GOUT(gs.kb(r)) = kb;
            GOUT(gs.md(r)) = mdv;
            GOUT(gs.mo(r)) = mov;
            GOUT(gs.ahd(r)) = ahd;
            GOUT(gs.aho(r)) = aho;
            GOUT(gs.lod(r, LOD_UP)) = lod0;
            GOUT(gs.lod(r, LOD_LO)) = lod1;
            GOUT(gs.lod(r, LOD_UP_PREV)) = lod2;
            GOUT(gs.lod(r, LOD_LO_PREV)) = lod3;
        }
#undef GOUT
    }
#undef FG
#undef IN
#undef SR
}
