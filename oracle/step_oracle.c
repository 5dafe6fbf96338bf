/* CPU ORACLE (test infrastructure, not product code).  See step_oracle.h.
 *
 * Every arithmetic expression keeps the reference's operand order and is
 * compiled with -ffp-contract=off, so results are bit-identical to the numpy
 * code wherever numpy itself is correctly rounded (`np.float64 ** 2.` goes
 * through libm pow and is not; see opensc2_b200/problem.py:quantize26).
 *
 * Band storage: the reference keeps band[k][i] = A[i][i + k - Main_diag] as a
 * (Full, Total) array; here the same entries are stored transposed,
 * B(i,k) = band[i*Full + k], which changes no arithmetic.
 */
#define _POSIX_C_SOURCE 200809L
#include "step_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <pthread.h>
#include <unistd.h>

/* numpy/core/src/umath/loops_utils.h.src: DOUBLE_pairwise_sum (PW_BLOCKSIZE 128). */
double osc2o_pairwise_sum(const double *a, long n, long stride)
{
    if (n < 8) {
        double res = 0.0;
        for (long i = 0; i < n; i++) res += a[i * stride];
        return res;
    } else if (n <= 128) {
        double r[8], res;
        long i;
        for (i = 0; i < 8; i++) r[i] = a[i * stride];
        for (i = 8; i < n - (n % 8); i += 8) {
            for (int k = 0; k < 8; k++) r[k] += a[(i + k) * stride];
        }
        res = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]));
        for (; i < n; i++) res += a[i * stride];
        return res;
    } else {
        long n2 = n / 2;
        n2 -= n2 % 8;
        return osc2o_pairwise_sum(a, n2, stride) + osc2o_pairwise_sum(a + n2 * stride, n - n2, stride);
    }
}

#define G(q, j) (fl_gauss[((size_t)(q) * M + (j)) * E + e])
enum { FL_V, FL_P, FL_T, FL_RHO, FL_C0, FL_GRUN, FL_H, FL_CV, FL_F };

/* step_matrix_construction.py:383-557 (__smat_fluid_interface): rows of c1. */
static void smat_fluid_interface(double *S, int d, int M, int E, int e, const double *fl_gauss,
                                 const double *ch_area, int c1, int c2, double K1, double K2,
                                 double K3, double coef_htc)
{
    const int v1 = c1, p1 = M + c1, t1 = 2 * M + c1, p2 = M + c2, t2 = 2 * M + c2;
    const double v = G(FL_V, c1), rho = G(FL_RHO, c1), A = ch_area[c1], phi = G(FL_GRUN, c1);
    const double h = G(FL_H, c1), cv = G(FL_CV, c1), c0 = G(FL_C0, c1), T = G(FL_T, c1);

    const double s_vj_pj = (K1 * v - K2) / (A * rho);
    S[v1 * d + p1] -= s_vj_pj;
    S[v1 * d + p2] = s_vj_pj;

    const double cga = phi / A;
    const double s_pj_pj = cga * (K3 - v * K2 - (h - (v * v) / 2. - (c0 * c0) / phi) * K1);
    S[p1 * d + p1] += s_pj_pj;
    S[p1 * d + p2] = -s_pj_pj;
    const double s_pj_tj = cga * coef_htc;
    S[p1 * d + t1] += s_pj_tj;
    S[p1 * d + t2] = -s_pj_tj;

    const double crca = 1. / (rho * cv * A);
    const double s_tj_pj = crca * (K3 - v * K2 - (h - (v * v) / 2. - phi * cv * T) * K1);
    S[t1 * d + p1] += s_tj_pj;
    S[t1 * d + p2] = -s_tj_pj;
    const double s_tj_tj = crca * coef_htc;
    S[t1 * d + t1] += s_tj_tj;
    S[t1 * d + t2] = -s_tj_tj;
}

int osc2o_step(const osc2o_topology *tp, osc2o_step_io *io)
{
    const int M = tp->n_ch, Sn = tp->n_sol, d = 3 * M + Sn, d2 = 2 * d;
    const int N = io->n_nodes, E = N - 1;
    const long Total = (long)d * N;
    const int Half = 2 * d, Main = 2 * d - 1, Full = 4 * d - 1;
    const double dt = io->dt, theta = tp->theta;
    const double *fl_gauss = io->fl_gauss;
    const int first = (io->num_step == 1);
    io->status = 0;

    double *MAS = calloc((size_t)Total * Full, sizeof(double));
    double *FLX = calloc((size_t)Total * Full, sizeof(double));
    double *DIF = calloc((size_t)Total * Full, sizeof(double));
    double *SOR = calloc((size_t)Total * Full, sizeof(double));
    double *SYS = calloc((size_t)Total * Full, sizeof(double));
    double *Known = calloc((size_t)Total, sizeof(double));
    double *work = calloc((size_t)Full + 8, sizeof(double));
    double *Mm = malloc(sizeof(double) * d * d), *Am = malloc(sizeof(double) * d * d);
    double *Km = malloc(sizeof(double) * d * d), *Sm = malloc(sizeof(double) * d * d);
    double *svec = malloc(sizeof(double) * d * 2), *svec_prev = malloc(sizeof(double) * d * 2);
    double *K123 = calloc((size_t)3 * (tp->nff > 0 ? tp->nff : 1) * (E > 0 ? E : 1), sizeof(double));
    double *EL[4];
    for (int q = 0; q < 4; q++) EL[q] = malloc(sizeof(double) * d2 * d2);
    double *FIN[4] = {MAS, FLX, DIF, SOR};
    double *elslod = malloc(sizeof(double) * d2), *elslod_prev = malloc(sizeof(double) * d2);
    double *sysvar = io->sysvar, *syslod = io->syslod;
#define BAND(Mx, i, k) ((Mx)[(size_t)(i) * Full + (k)])

    /* tsf:235-243 theta-method history shift */
    if (io->num_step > 1) {
        for (long i = 0; i < Total; i++) {
            syslod[2 * i + 1] = syslod[2 * i];
            syslod[2 * i] = 0.0;
        }
    }

    /* smc:103-218 transport coefficients K', K'', K''' */
    for (int k = 0; k < tp->nff; k++) {
        const int c1 = tp->ff[2 * k], c2 = tp->ff[2 * k + 1];
        for (int e = 0; e < E; e++) {
            const double p1 = G(FL_P, c1), p2 = G(FL_P, c2);
            double delta_p = fabs(p1 - p2);
            if (delta_p < tp->delta_p_min) delta_p = tp->delta_p_min;
            const int c = (p2 < p1) ? c1 : c2;
            const double vel = G(FL_V, c);
            const double peri_open = io->peri_ff[((size_t)0 * tp->nff + k) * E + e];
            const double k1 = peri_open * sqrt(2. * G(FL_RHO, c) / (tp->k_loc * delta_p));
            const double k2 = k1 * tp->lambda_v * vel;
            const double vl = vel * tp->lambda_v;
            const double k3 = k1 * (G(FL_H, c) + .5 * (vl * vl));
            K123[((size_t)0 * tp->nff + k) * E + e] = k1;
            K123[((size_t)1 * tp->nff + k) * E + e] = k2;
            K123[((size_t)2 * tp->nff + k) * E + e] = k3;
        }
    }

    /* tsf:281-491 element loop */
    for (int e = 0; e < E; e++) {
        const double dz = io->dz[e];
        memset(Mm, 0, sizeof(double) * d * d);
        memset(Am, 0, sizeof(double) * d * d);
        memset(Km, 0, sizeof(double) * d * d);
        memset(Sm, 0, sizeof(double) * d * d);
        memset(svec, 0, sizeof(double) * d * 2);
        memset(svec_prev, 0, sizeof(double) * d * 2);

        for (int i = 0; i < 3 * M; i++) Mm[i * d + i] = 1.0;            /* tsf:319-322 */
        for (int j = 0; j < M; j++) {
            const int iv = j, ip = M + j, it = 2 * M + j;
            const double v = G(FL_V, j), rho = G(FL_RHO, j), c0 = G(FL_C0, j);
            /* smc:62-101 build_amat */
            Am[iv * d + iv] = v; Am[ip * d + ip] = v; Am[it * d + it] = v;
            Am[iv * d + ip] = 1. / rho;
            Am[ip * d + iv] = rho * (c0 * c0);
            Am[it * d + iv] = G(FL_GRUN, j) * G(FL_T, j);
            /* smc:220-260 build_kmat_fluid (UPWEQT = 1 on fluid equations) */
            const double kd = dz * 1.0 * fabs(v) / 2.0;
            Km[iv * d + iv] = kd; Km[ip * d + ip] = kd; Km[it * d + it] = kd;
            /* smc:262-311 build_smat_fluid */
            const double svv = 2.0 * G(FL_F, j) * fabs(v) / tp->ch_dh[j];
            Sm[iv * d + iv] = svv;
            Sm[ip * d + iv] = -svv * G(FL_GRUN, j) * rho * v;
            Sm[it * d + iv] = -svv / G(FL_CV, j) * v;
        }
        /* smc:313-381 fluid-fluid interfaces */
        for (int k = 0; k < tp->nff; k++) {
            const int c1 = tp->ff[2 * k], c2 = tp->ff[2 * k + 1];
            const double K1 = K123[((size_t)0 * tp->nff + k) * E + e];
            const double K2 = K123[((size_t)1 * tp->nff + k) * E + e];
            const double K3 = K123[((size_t)2 * tp->nff + k) * E + e];
            const double coef_htc =
                io->peri_ff[((size_t)0 * tp->nff + k) * E + e] * io->htc_ff[((size_t)0 * tp->nff + k) * E + e]
                + io->peri_ff[((size_t)1 * tp->nff + k) * E + e] * io->htc_ff[((size_t)1 * tp->nff + k) * E + e];
            smat_fluid_interface(Sm, d, M, E, e, fl_gauss, tp->ch_area, c1, c2, K1, K2, K3, coef_htc);
            smat_fluid_interface(Sm, d, M, E, e, fl_gauss, tp->ch_area, c2, c1, K1, K2, K3, coef_htc);
        }
        /* smc:559-670 fluid-solid interfaces */
        for (int k = 0; k < tp->nfs; k++) {
            const int j = tp->fs[2 * k], l = tp->fs[2 * k + 1];
            const int ip = M + j, it = 2 * M + j, is = 3 * M + l;
            const double A = tp->ch_area[j];
            const double cga = G(FL_GRUN, j) / A;
            const double coef_htc = io->peri_fs[(size_t)k * E + e] * io->htc_fs[(size_t)k * E + e];
            const double s_pj_tj = cga * coef_htc;
            Sm[ip * d + it] += s_pj_tj;
            Sm[ip * d + is] = -s_pj_tj;
            const double crca = 1. / (G(FL_RHO, j) * G(FL_CV, j) * A);
            const double s_tj_tj = crca * coef_htc;
            Sm[it * d + it] += s_tj_tj;
            Sm[it * d + is] = -s_tj_tj;
            Sm[is * d + is] += coef_htc;
            Sm[is * d + it] = -coef_htc;
        }
        /* tsf:369-404 solids: smc:672-730, 888-948 */
        for (int l = 0; l < Sn; l++) {
            const int is = 3 * M + l;
            const double A = tp->sol_area[l], cs = tp->sol_costeta[l];
            const double rho = io->sol_gauss[((size_t)0 * Sn + l) * E + e];
            const double cp = io->sol_gauss[((size_t)1 * Sn + l) * E + e];
            const double kk = io->sol_gauss[((size_t)2 * Sn + l) * E + e];
            Mm[is * d + is] = A * rho * cp / cs;
            Km[is * d + is] = A * kk / cs;
            const double *Q1 = io->sol_q + (((size_t)0 * Sn + l) * E + e) * 2;
            const double *Q2 = io->sol_q + (((size_t)1 * Sn + l) * E + e) * 2;
            const double qs0 = io->qsource[(size_t)e * Sn + l], qs1 = io->qsource[(size_t)(e + 1) * Sn + l];
            svec[is * 2 + 0] = Q1[0] - qs0;
            svec[is * 2 + 1] = Q2[0] - qs1;
            if (first) {
                svec_prev[is * 2 + 0] = Q1[1] - qs0;
                svec_prev[is * 2 + 1] = Q2[1] - qs1;
            }
        }
        /* smc:732-828 solid-solid interfaces */
        for (int k = 0; k < tp->nss; k++) {
            const int a = 3 * M + tp->ss[2 * k], b = 3 * M + tp->ss[2 * k + 1];
            const double coef = io->peri_ss[(size_t)k * E + e] * io->htc_ss[(size_t)k * E + e];
            Sm[a * d + a] += coef;
            Sm[a * d + b] = -coef;
            Sm[b * d + b] += coef;
            Sm[b * d + a] = -coef;
        }
        /* smc:830-886, 950-1013 environment-solid interfaces */
        for (int k = 0; k < tp->nes; k++) {
            const int is = 3 * M + tp->es[k];
            const double *h = io->htc_es + (size_t)k * 3 * E;
            const double peri = io->peri_es[(size_t)k * E + e];
            const int rect = tp->es_choice2[k] && tp->is_rectangular;
            double coef_htc;
            if (rect) coef_htc = 2. * tp->height * h[0 * E + e] + tp->width * (h[1 * E + e] + h[2 * E + e]);
            else coef_htc = h[0 * E + e] * peri;
            Sm[is * d + is] += coef_htc;
            if (!rect) { /* the rectangular branch never adds env_heat (indentation, smc:987-1011) */
                const double coef = peri * h[0 * E + e];
                const double env_heat = coef * io->t_env;
                svec[is * 2 + 0] += env_heat;
                svec[is * 2 + 1] += env_heat;
                if (first) {
                    svec_prev[is * 2 + 0] += env_heat;
                    svec_prev[is * 2 + 1] += env_heat;
                }
            }
        }

        /* smc:1015-1163 element matrices (ALFA = 0, tsf:265) */
        const double alpha = 0.0;
        const double cdiag = dz * (1. / 3. + alpha), coff = dz * (1. / 6. - alpha);
        for (int r = 0; r < d; r++) {
            for (int c = 0; c < d; c++) {
                const double m = Mm[r * d + c], a = Am[r * d + c] / 2., kb = Km[r * d + c] / dz;
                const double sd = Sm[r * d + c] * dz / 3., so = sd / 2.;
                double *elm = EL[0], *ela = EL[1], *elk = EL[2], *els = EL[3];
                elm[r * d2 + c] = cdiag * m;       elm[r * d2 + d + c] = coff * m;
                elm[(d + r) * d2 + c] = coff * m;  elm[(d + r) * d2 + d + c] = cdiag * m;
                ela[r * d2 + c] = -a;              ela[r * d2 + d + c] = a;
                ela[(d + r) * d2 + c] = -a;        ela[(d + r) * d2 + d + c] = a;
                elk[r * d2 + c] = kb;              elk[r * d2 + d + c] = -kb;
                elk[(d + r) * d2 + c] = -kb;       elk[(d + r) * d2 + d + c] = kb;
                els[r * d2 + c] = sd;              els[r * d2 + d + c] = so;
                els[(d + r) * d2 + c] = so;        els[(d + r) * d2 + d + c] = sd;
            }
        }
        /* smc:1165-1216 element load */
        const double dz_6 = dz / 6.;
        for (int r = 0; r < d; r++) {
            elslod[r] = (2. * svec[r * 2] + svec[r * 2 + 1]) * dz_6;
            elslod[d + r] = (svec[r * 2] + 2. * svec[r * 2 + 1]) * dz_6;
            elslod_prev[r] = (2. * svec_prev[r * 2] + svec_prev[r * 2 + 1]) * dz_6;
            elslod_prev[d + r] = (svec_prev[r * 2] + 2. * svec_prev[r * 2 + 1]) * dz_6;
        }
        /* smc:1218-1262 assemble_matrix; :1264-1318 assemble_syslod */
        const long jump = (long)d * e;
        for (int r = 0; r < Half; r++) {
            const long i = jump + r;
            for (int q = 0; q < 4; q++)
                for (int c = 0; c < Half; c++) BAND(FIN[q], i, Half - 1 - r + c) += EL[q][r * d2 + c];
            syslod[2 * i] += elslod[r];
            if (first) syslod[2 * i + 1] += elslod_prev[r];
        }
    }

    /* smc:1320-1364 SYSMAT; :1366-1474 Known */
    for (size_t q = 0; q < (size_t)Total * Full; q++)
        SYS[q] = MAS[q] / dt + theta * (FLX[q] + DIF[q] + SOR[q]);
    for (long i = 0; i < Total; i++) {
        long lo, hi; /* sysvar index range [lo, hi) */
        if (i <= Half - 1) { lo = 0; hi = Half + i; }
        else if (i >= Total - (Half - 1)) { lo = i - (Half - 1); hi = Total; }
        else { lo = i - (Half - 1); hi = i + Half; }
        if (hi > Total) hi = Total;   /* numpy fancy index would raise; never hit for N >= 2 */
        const long n = hi - lo;
        for (long c = lo; c < hi; c++) {
            const int k = (int)(c - i + Half - 1);
            work[c - lo] = (BAND(MAS, i, k) / dt - (1.0 - theta) * (BAND(FLX, i, k) + BAND(DIF, i, k) + BAND(SOR, i, k))) * sysvar[c];
        }
        Known[i] = osc2o_pairwise_sum(work, n, 1);
    }
    for (long i = 0; i < Total; i++)
        Known[i] += (+theta * syslod[2 * i] + (1.0 - theta) * syslod[2 * i + 1]);

    /* tsf:524-533 boundary conditions, fluid_component.py:255-565 */
    for (int j = 0; j < M; j++) {
        const long first_node = 0, last_node = Total - d;
        const int fwd = tp->ch_forward[j];
        const long inl = fwd ? first_node : last_node, out = fwd ? last_node : first_node;
        const double p_inl = io->fl_bc[0 * M + j], p_out = io->fl_bc[1 * M + j];
        const double T_inl = io->fl_bc[2 * M + j], T_out = io->fl_bc[3 * M + j];
        const double mdt_in = io->fl_bc[4 * M + j], mdt_out = io->fl_bc[5 * M + j];
        const double v_first = io->fl_node[0 * M + j], v_last = io->fl_node[1 * M + j];
        const double rho_first = io->fl_node[2 * M + j], rho_last = io->fl_node[3 * M + j];
        const double area = tp->ch_area[j];
        long rows[4]; double vals[4]; int nb = 0;
        switch (tp->ch_intial[j]) {
        case 1: /* impose_pressure_drop */
            rows[nb] = inl + M + j; vals[nb++] = p_inl;
            rows[nb] = out + M + j; vals[nb++] = p_out;
            break;
        case 2: /* impose_inl_p_out_v */
            rows[nb] = out + j; vals[nb++] = mdt_out / (fwd ? rho_last : rho_first) / area;
            rows[nb] = inl + M + j; vals[nb++] = p_inl;
            break;
        default: /* 3: impose_inl_v_out_p */
            rows[nb] = inl + j; vals[nb++] = mdt_in / (fwd ? rho_first : rho_last) / area;
            rows[nb] = out + M + j; vals[nb++] = p_out;
            break;
        }
        /* temperature: inlet if inflow at the inlet end, outlet if backflow at the outlet end */
        if (fwd) {
            if (v_first > 0) { rows[nb] = first_node + 2 * M + j; vals[nb++] = T_inl; }
            if (v_last < 0) { rows[nb] = last_node + 2 * M + j; vals[nb++] = T_out; }
        } else {
            if (v_last < 0) { rows[nb] = last_node + 2 * M + j; vals[nb++] = T_inl; }
            if (v_first > 0) { rows[nb] = first_node + 2 * M + j; vals[nb++] = T_out; }
        }
        for (int b = 0; b < nb; b++) {
            for (int k = 0; k < Full; k++) BAND(SYS, rows[b], k) = 0.0;
            BAND(SYS, rows[b], Main) = 1.0;
            Known[rows[b]] = vals[b];
        }
    }

    /* tsf:538-551 row scaling */
    for (long i = 0; i < Total; i++) {
        double mx = 0.0;
        for (int k = 0; k < Full; k++) { const double a = fabs(BAND(SYS, i, k)); if (a > mx || a != a) mx = a; }
        for (int k = 0; k < Full; k++) BAND(SYS, i, k) = BAND(SYS, i, k) / mx;
        Known[i] = Known[i] / mx;
    }
    if (io->sysmat) for (long i = 0; i < Total; i++) for (int k = 0; k < Full; k++) io->sysmat[(size_t)k * Total + i] = BAND(SYS, i, k);
    if (io->known) memcpy(io->known, Known, sizeof(double) * Total);

    /* tsf:602-698 gredub */
    {
        const int K1 = Main, K21 = 2 * Main;
        int JJ = K21;
        for (long I = Total - Main; I < Total; I++) {
            if (I >= 0) for (int k = JJ; k <= K21; k++) BAND(SYS, I, k) = 0.0;
            JJ--;
        }
        for (long I = 1; I < Total && !io->status; I++) {
            long lo = I - Main; if (lo < 0) lo = 0;
            for (long r = lo; r < I; r++)
                if (fabs(BAND(SYS, r, K1)) <= 1.0e-20) { io->status = -(int)(r + 1); break; }
            if (io->status) break;
            for (long r = lo; r < I; r++) {
                const int kk = (int)(r - I + K1);          /* band slot of column r in row I */
                const double q = BAND(SYS, I, kk) / BAND(SYS, r, K1);
                for (int m = 1; m <= K1; m++)
                    BAND(SYS, I, kk + m) = BAND(SYS, I, kk + m) - BAND(SYS, r, K1 + m) * q;
                BAND(SYS, I, kk) = q;
            }
        }
    }
    /* tsf:701-808 gbacsb */
    if (!io->status) {
        const int K1 = Main;
        for (long I = 1; I < Total; I++) {
            long lo = I - Main; if (lo < 0) lo = 0;
            for (long r = lo; r < I; r++) Known[I] = Known[I] - Known[r] * BAND(SYS, I, (int)(r - I + K1));
        }
        for (long r = Total - 1; r >= 0; r--)
            if (fabs(BAND(SYS, r, K1)) <= 1.0e-20) { io->status = -(int)(r + 1); }
        if (!io->status) {
            Known[Total - 1] = Known[Total - 1] / BAND(SYS, Total - 1, K1);
            for (long i = Total - 2; i >= 0; i--) {
                long jj = i + Main; if (jj >= Total) jj = Total - 1;
                double q = Known[i];
                for (long c = i + 1; c <= jj; c++) q = q - BAND(SYS, i, (int)(c - i + K1)) * Known[c];
                Known[i] = q / BAND(SYS, i, K1);
            }
        }
    }

    if (!io->status) {
        /* tsf:567-591: SYSVAR <- x; norms with the in-place squaring quirks */
        memcpy(sysvar, Known, sizeof(double) * Total);
        double *sq = MAS;  /* reuse as scratch */
        for (long i = 0; i < Total; i++) sq[i] = Known[i] * Known[i];      /* Known **= 2 */
        double *chg = FLX, *eig = DIF;
        for (long i = 0; i < Total; i++) {
            chg[i] = sq[i] - sysvar[i];
            eig[i] = fabs(chg[i] / dt) / (fabs(sq[i]) + 1.0e-5);
        }
        for (int q = 0; q < d; q++) {
            if (io->norm_solution) io->norm_solution[q] = sqrt(osc2o_pairwise_sum(sq + q, N, d));
        }
        for (long i = 0; i < Total; i++) chg[i] = chg[i] * chg[i];
        for (int q = 0; q < d; q++) {
            if (io->norm_change) io->norm_change[q] = sqrt(osc2o_pairwise_sum(chg + q, N, d));
            if (io->eqteig) {
                int src = q;
                if (q >= M && q < 2 * M) src = q - M;   /* pressure <- velocity, tsf:886-888 */
                double mx = eig[src];
                for (long n = 1; n < N; n++) { const double x = eig[n * d + src]; if (x > mx) mx = x; }
                io->eqteig[q] = mx;
            }
        }
    }
    if (io->k123) memcpy(io->k123, K123, sizeof(double) * 3 * tp->nff * E);

    free(MAS); free(FLX); free(DIF); free(SOR); free(SYS); free(Known); free(work);
    free(Mm); free(Am); free(Km); free(Sm); free(svec); free(svec_prev); free(K123);
    for (int q = 0; q < 4; q++) free(EL[q]);
    free(elslod); free(elslod_prev);
    return io->status;
}

typedef struct {
    const osc2o_topology *tp;
    const osc2o_step_io *io0;
    int batch, *status, *next, *worst;
    pthread_mutex_t *mu;
} batch_job;

static void *batch_worker(void *arg)
{
    batch_job *jb = (batch_job *)arg;
    const osc2o_topology *tp = jb->tp;
    const osc2o_step_io *io0 = jb->io0;
    const int M = tp->n_ch, Sn = tp->n_sol, d = 3 * M + Sn;
    const int N = io0->n_nodes, E = N - 1;
    const size_t Total = (size_t)d * N;
    for (;;) {
        pthread_mutex_lock(jb->mu);
        const int b = (*jb->next)++;
        pthread_mutex_unlock(jb->mu);
        if (b >= jb->batch) break;
        osc2o_step_io io = *io0;
#define ADV(f, n) if (io0->f) io.f = io0->f + (size_t)b * (n)
        ADV(dz, E); ADV(fl_gauss, (size_t)9 * M * E); ADV(fl_node, 4 * M); ADV(fl_bc, 6 * M);
        ADV(sol_gauss, (size_t)3 * Sn * E); ADV(sol_q, (size_t)2 * Sn * E * 2);
        ADV(peri_ff, (size_t)2 * tp->nff * E); ADV(htc_ff, (size_t)2 * tp->nff * E);
        ADV(peri_fs, (size_t)tp->nfs * E); ADV(htc_fs, (size_t)tp->nfs * E);
        ADV(peri_ss, (size_t)tp->nss * E); ADV(htc_ss, (size_t)tp->nss * E);
        ADV(peri_es, (size_t)tp->nes * E); ADV(htc_es, (size_t)tp->nes * 3 * E);
        ADV(qsource, (size_t)N * Sn); ADV(sysvar, Total); ADV(syslod, Total * 2);
        ADV(k123, (size_t)3 * tp->nff * E); ADV(norm_solution, d); ADV(norm_change, d); ADV(eqteig, d);
        ADV(sysmat, Total * (4 * d - 1)); ADV(known, Total);
#undef ADV
        const int st = osc2o_step(tp, &io);
        if (jb->status) jb->status[b] = st;
        if (st) {
            pthread_mutex_lock(jb->mu);
            *jb->worst = st;
            pthread_mutex_unlock(jb->mu);
        }
    }
    return NULL;
}

int osc2o_step_batch(const osc2o_topology *tp, const osc2o_step_io *io0, int batch, int n_threads, int *status)
{
    if (n_threads <= 0) {
        long nc = sysconf(_SC_NPROCESSORS_ONLN);
        n_threads = nc > 0 ? (int)nc : 1;
    }
    if (n_threads > batch) n_threads = batch;
    if (n_threads > 256) n_threads = 256;
    int next = 0, worst = 0;
    pthread_mutex_t mu;
    pthread_mutex_init(&mu, NULL);
    batch_job jb = {tp, io0, batch, status, &next, &worst, &mu};
    pthread_t th[256];
    for (int t = 1; t < n_threads; t++) pthread_create(&th[t], NULL, batch_worker, &jb);
    batch_worker(&jb);
    for (int t = 1; t < n_threads; t++) pthread_join(th[t], NULL);
    pthread_mutex_destroy(&mu);
    return worst;
}
