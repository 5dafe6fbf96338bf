/* CPU ORACLE (test infrastructure) -- see props_oracle.h.  Every function cites the
 * reference lines it restates; paths are relative to /root/reference/source_code. */
#define _GNU_SOURCE
#include "props_oracle.h"

#include <math.h>
#include <pthread.h>
#include <stdlib.h>
#include <string.h>

static double sq(double x) { return x * x; }   /* numpy: x ** 2 -> square */

/* ---- copper: properties_of_materials/copper.py ---------------------------------- */
/* rhoecu0_nist, copper.py:221-275 */
static double rhoecu0(double t, double rrr)
{
    const double P0 = 1.553e-8, P1 = 1.171e-17, P2 = 4.49, P3 = 3.841e10, P4 = 1.14, P5 = 50, P6 = 6.428,
                 P7 = 0.4531;
    const double rho0 = P0 / rrr;
    const double rhoi = (P1 * pow(t, P2)) / (1.0 + P1 * P3 * pow(t, P2 - P4) * exp(-pow(P5 / t, P6)));
    const double rhoi0 = P7 * rhoi * rho0 / (rhoi + rho0);
    return rho0 + rhoi + rhoi0;
}
/* electrical_resistivity_cu_nist, copper.py:140-217 */
static double rhoecu(double t, double b, double rrr)
{
    static const double a[5] = {-2.662, 0.3168, 0.6229, -0.1839, 0.01827};
    const double T0 = 273.0;
    if (!(rrr > 1.0 && b >= 0 && t > 0)) return 0.0;
    const double rhoN = rhoecu0(t, rrr);
    double ccc = 0.0;
    if (b > 0) {
        const double xxx = rhoecu0(T0, rrr) / rhoecu0(t, rrr) * b;
        const double L = log10(xxx);
        double aaa = 0.0;
        aaa = aaa + a[0] * 1.0;
        aaa = aaa + a[1] * L;
        aaa = aaa + a[2] * sq(L);
        aaa = aaa + a[3] * pow(L, 3.0);
        aaa = aaa + a[4] * pow(L, 4.0);
        ccc = pow(10.0, aaa);
    }
    return rhoN * (1.0 + ccc);
}
/* thermal_conductivity_cu_nist, copper.py:5-71 */
static double k_cu(double t, double b, double rrr)
{
    const double beta = 0.634 / rrr, betat = beta / 0.0003;
    const double P1 = 1.754e-8, P2 = 2.763, P3 = 1102, P4 = -0.165, P5 = 70, P6 = 1.756;
    const double P7 = 0.838 / pow(betat, 0.1661);
    const double W0 = beta / t;
    const double Wi = (P1 * pow(t, P2)) / (1 + (P1 * P3 * pow(t, P2 + P4) * exp(-pow(P5 / t, P6))));
    const double Wi0 = P7 * Wi * W0 / (Wi + W0);
    return 1.0 / (Wi + W0 + Wi0) * rhoecu0(t, rrr) / rhoecu(t, b, rrr);
}
/* isobaric_specific_heat_cu_nist, copper.py:75-136 (np.piecewise: 0 outside [4, 300]) */
static double cp_cu(double t)
{
    const double A0 = -1.91844, A1 = -0.15973, A2 = 8.61013, A3 = -18.996, A4 = 21.9661, A5 = -12.7328,
                 A6 = 3.54322, A7 = -0.3797;
    if (!(t >= 4 && t <= 300)) return 0.0;
    const double L = log10(t);
    return pow(10.0, A0 + A1 * L + A2 * sq(L) + A3 * pow(L, 3.0) + A4 * pow(L, 4.0) + A5 * pow(L, 5.0)
                         + A6 * pow(L, 6.0) + A7 * pow(L, 7.0));
}

/* ---- Nb3Sn: properties_of_materials/niobium3_tin.py -------------------------------- */
/* SNBSN, niobium3_tin.py:371-434 */
static double snbsn(double eps)
{
    const double Ca1 = 47.02, Ca2 = 11.76, EPS0a = 2.31e-3;
    const double EPSsh = Ca2 * EPS0a / sqrt(pow(Ca1, 2.0) - pow(Ca2, 2.0));
    return 1.0 + 1.0 / (1.0 - Ca1 * EPS0a)
                     * (Ca1 * (sqrt(pow(EPSsh, 2.0) + pow(EPS0a, 2.0)) - sqrt(sq(eps - EPSsh) + pow(EPS0a, 2.0)))
                        - Ca2 * eps);
}
/* critical_temperature_nb3sn, niobium3_tin.py:437-490 */
static double tc_nb3sn(double b, double eps, double tc0m, double bc20m)
{
    const double blim = fmax(b, 0.01);
    const double s = snbsn(eps);
    const double blcase = blim / (bc20m * s);
    if (!(blcase < 1.0)) return 0.0;
    return tc0m * pow(s, 1.0 / 3.0) * pow(1.0 - blcase, 1.0 / 1.52);
}
/* thermal_conductivity_nb3sn, niobium3_tin.py:62-124 */
static double k_nb3sn(double tt)
{
    const double T0 = 21.9620369, A = 1.3905e14, B = 76.6127237, C = -7.1614565e-1, D = 5.40229689,
                 n = 3.66343961, m = 9.3220843;
    tt = fmin(tt, 750.0);
    tt = fmax(tt, 4.0);
    if (tt <= T0) return A * pow(tt, n) / pow(B + tt, m);
    if (tt > T0) return C * log(tt) + D;
    return 0.0;
}
/* isobaric_specific_heat_nb3sn, niobium3_tin.py:127-262 */
static double cp_nb3sn(double T, double TCS, double TC, double TC0)
{
    const double AA = 38.2226877, BB = -848.364226, CC = 1415.13808, DD = -346.837966;
    const double a = 6.80458608, b = 59.9209182, c = 25.8286334, d = 8.77918335;
    T = fmin(T, 1000.0);
    double CPN = 0.0;
    if (T <= 10.0) CPN = 7.5475e-3 * sq(T);
    else if (T > 10.0 && T <= 20.0) CPN = (-0.3 + 0.00375 * sq(T)) / 0.09937;
    else if (T > 20.0)
        CPN = AA * T / (a + T) + BB * sq(T) / sq(b + T) + CC * pow(T, 3.0) / pow(c + T, 3.0)
              + DD * pow(T, 4.0) / pow(d + T, 4.0);
    double CPNTC = 0.0;
    if (TC <= 10.0) CPNTC = 7.5475e-3 * sq(TC);
    else if (TC > 10.0 && TC <= 20.0) CPNTC = (-0.3 + 0.00375 * sq(TC)) / 0.09937;
    double CPS = 0.0;
    if (T <= TC) {
        const double TC_norm = TC / TC0;
        const double DBC2DT = -0.46306 - 0.067830 * TC;
        const double DELCP = 1500.0 * sq(DBC2DT) / (2.0 * sq(27.2 / (1.0 + (0.34 * TC_norm))) - 1.0);
        CPS = (CPNTC + DELCP) * pow(T / TC, 3.0);
    }
    double CP = 0.0;
    if (T <= TCS) CP = CPS;
    if (T > TCS && T <= TC) {
        double F = 1.0;
        if (TCS < TC) F = (T - TCS) / (TC - TCS);
        CP = F * CPN + (1.0 - F) * CPS;
    }
    if (T > TC) CP = CPN;
    return CP;
}

/* ---- stainless steel: properties_of_materials/stainless_steel.py ------------------- */
static double four_term(double T, const double C[4], const double s[4], const double n[4])
{
    return C[0] * T / pow(s[0] + T, n[0]) + C[1] * sq(T) / pow(s[1] + T, n[1])
           + C[2] * pow(T, 3.0) / pow(s[2] + T, n[2]) + C[3] * pow(T, 4.0) / pow(s[3] + T, n[3]);
}
/* thermal_conductivity_ss, stainless_steel.py:5-38 */
static double k_ss(double t)
{
    static const double C[4] = {21224.0426, 103.007623, 135.870511, 0.07942859};
    static const double s[4] = {10844822.8, 6800262.44, 34.4411414, 0.23194142};
    static const double n[4] = {1.07801291, 1.88806818, 3.44233222, 3.20629457};
    return four_term(t, C, s, n);
}
/* isobaric_specific_heat_ss, stainless_steel.py:41-88; the formula is chosen from max(T)
 * of the WHOLE array handed to the function */
static double cp_ss(double T, double tmax_of_array)
{
    static const double C[4] = {-468.772387, -81.2353649, 827.209072, 29.9811823};
    static const double s[4] = {422.117262, 1235598.51, 51.2732976, 0.44591988};
    static const double n[4] = {0.57149148, 2.0006633, 2.79384889, 3.25864797};
    const double T0 = 12.8973011, A1 = 0.46437387, A2 = 0.00035157, TMIN = 1.0, TMAX = 1253.0;
    if (tmax_of_array <= TMIN) return 0.498 * T + 3.71e-4 * pow(T, 3.0);
    if (tmax_of_array < T0) {
        T = fmin(T, TMAX);
        return A1 * T + A2 * pow(T, 3.0);
    }
    T = fmin(T, TMAX);
    return four_term(T, C, s, n);
}
/* ---- glass epoxy: properties_of_materials/glass_epoxy.py --------------------------- */
static double k_ge(double t)
{
    static const double C[4] = {0.106635834, 4.254937095, 0.369710568, -0.220612963};
    static const double s[4] = {4.374557267, 78.36483015, 0.896107679, 0.261916360};
    static const double n[4] = {0.551227580, 2.274402853, 2.979875183, 3.644905343};
    t = fmin(t, 1000.0);
    t = fmax(t, 1.0);
    return four_term(t, C, s, n);
}
static double cp_ge(double t)
{
    static const double C[4] = {0.309399374, -1387.509141, 283.0320589, 1.822153805};
    static const double s[4] = {113.1687618, 80.88333136, 26.27075203, 1.104229335};
    static const double n[4] = {0.254475819, 1.938523579, 2.603466143, 3.522809815};
    t = fmin(t, 1000.0);
    t = fmax(t, 1.0);
    return four_term(t, C, s, n);
}

/* ---- aluminium: properties_of_materials/aluminium.py --------------------------------- */
/* thermal_conductivity_al, aluminium.py:5-77: all zeros when ANY element of the array is outside
 * [2, 1000] K (the function prints an error and returns its zero initialisation) */
static double k_al(double T, double tmin, double tmax)
{
    static const double t0[12] = {2., 4., 6., 8., 10., 20., 40., 60., 80., 100., 200., 300.};
    static const double sl[12] = {4.2345, 4.2000, 4.3650, 4.8550, 4.3430, 3.8805, 2.0375, 0.5360, 0.1145, -0.0452, -0.0291, 0.0};
    static const double of[12] = {6.6510, 15.1200, 23.5200, 32.2500, 41.9600, 85.3900, 163.000, 203.75, 214.47, 216.76, 212.24, 209.3300};
    static const double hi[12] = {4., 6., 8., 10., 20., 40., 60., 80., 100., 200., 300., 1000.};
    if (tmin < 2.0 || tmax > 1000.0) return 0.0;
    for (int i = 0; i < 12; ++i) {
        const int in = (i == 0) ? (T >= 2 && T <= hi[0]) : (T > t0[i] && T <= hi[i]);
        if (in) return (i == 11) ? of[11] : sl[i] * (T - t0[i]) + of[i];
    }
    return 0.0;
}
/* isobaric_specific_heat_al, aluminium.py:80-164 */
static double cp_al(double T, double tmin, double tmax)
{
    const double T0 = 11.0578562, T1 = 53.2685055, a1 = 0.04257809, a2 = 0.00340313, a3 = 0.00063472;
    const double b0 = -2.50413265, b1 = 0.71747923, b2 = -0.06270275, b3 = 0.0031683, b4 = -2.0154e-05;
    static const double C[4] = {11458.0127, -438.603857, 331.057118, -10001.7006};
    static const double sh[4] = {-7.12302329, -34.6373149, -29.4649313, -3.30673113};
    static const double n[4] = {0.9531781, 2.00135991, 2.974805, 3.94384164};
    if (tmin < 2.0 || tmax > 1000.0) return 0.0;
    if (T >= 2.0 && T <= T0) return a1 * T + a2 * sq(T) + a3 * pow(T, 3.0);
    if (T > T0 && T <= T1) return b0 + b1 * T + b2 * sq(T) + b3 * pow(T, 3.0) + b4 * pow(T, 4.0);
    if (T > T1 && T <= 1000.0) return four_term(T, C, sh, n);
    return 0.0;
}
/* ---- RE123: properties_of_materials/rare_earth_123.py ---------------------------------- */
/* thermal_conductivity_re123, rare_earth_123.py:65-135 */
static double k_re123(double T)
{
    static const double A[2] = {-1.266103106492942e-5, -1.316704893540544e-9}, B[2] = {0.002670105477219, 2.636151594117440e-06};
    static const double C[2] = {-0.197035302542769, -0.001601689632073}, D[2] = {4.933962659604384, 0.428760641312538};
    static const double E[2] = {29.651536939501670, -53.306643333287260}, F[2] = {66.578192505447330, 3.682252343338599e03};
    T = fmin(T, 300.0);
    T = fmax(T, 4.0);
    const int i = (T <= 70.0) ? 0 : 1;
    return A[i] * pow(T, 5.0) + B[i] * pow(T, 4.0) + C[i] * pow(T, 3.0) + D[i] * sq(T) + E[i] * T + F[i];
}
/* isobaric_specific_heat_re123, rare_earth_123.py:138-190 */
static double cp_re123(double T)
{
    T = fmin(T, 300.0);
    T = fmax(T, 4.0);
    return -7.567485538209158e-10 * pow(T, 5.0) + 6.351452642016898e-07 * pow(T, 4.0) + -1.947975786547597e-04 * pow(T, 3.0)
           + 0.023616673974415 * sq(T) + 0.239331954284042 * T + -1.096191721280114;
}

/* magnesium_diboride.py:8-77 (the copper NIST form at the default rrr = 100: strand_mixed_component.py calls it
 * with temperature and field only), :79-139 */
static double k_mgb2(double t, double b) { return k_cu(t, b, 100.0); }
static double cp_mgb2(double t)
{
    static const double A[8] = {-1.91844, -0.15973, 8.61013, -18.996, 21.9661, -12.7328, 3.54322, -0.3797};
    if (!(t >= 4 && t <= 300)) return 0.0;
    const double L = log10(t);
    return pow(10.0, A[0] + A[1] * L + A[2] * sq(L) + A[3] * pow(L, 3.0) + A[4] * pow(L, 4.0) + A[5] * pow(L, 5.0)
                         + A[6] * pow(L, 6.0) + A[7] * pow(L, 7.0));
}
/* niobium_titanium.py:9-62 */
static double k_nbti(double T)
{
    const double A = 6.5, Tm = 25.0, Tnb = 50.0, K0 = 21.5;
    T = fmin(T, 750.0);
    T = fmax(T, 1.0);
    double k = A * T / (1.0 + pow(T / Tm, 3.0));
    if (T >= Tnb) k = k + K0 * (1.0 - exp(-(T - Tnb) / Tnb));
    return k;
}
/* niobium_titanium.py:274-286 */
static double tc_nbti(double b, double bc20, double tc0)
{
    return (b <= bc20) ? tc0 * pow(1 - b / bc20, 1 / 1.7) : 0.0;
}
/* niobium_titanium.py:64-209 */
static double cp_nbti(double T, double B, double TCS, double TC)
{
    static const double X[7] = {0.0, 2.0, 3.0, 4.0, 5.0, 6.0, 7.0};
    static const double AB[7] = {0.341000e-01, 0.282900e-01, 0.219200e-01, 0.158600e-01, 0.993000e-02, 0.491000e-02, 0.152000e-02};
    static const double E[7] = {0.233300e01, 0.244600e01, 0.259700e01, 0.280600e01, 0.310700e01, 0.356800e01, 0.434800e01};
    const double AA = 17.971311, BB = -3215.18087, CC = 4325.5737, DD = -711.296119;
    const double a = 88.768525, b = 3.66034399, c = 38.0453266, d = 14.67806;
    const double na = 1.0456476, nb = 51.7204943, nc = 4.20661339, nd = 6.01104574;
    if (T > 1.0) {
        T = fmin(T, 1000.0);
        return AA * pow(T / (a + T), na) + BB * pow(T / (b + T), nb) + CC * pow(T / (c + T), nc) + DD * pow(T / (d + T), nd);
    }
    int I = 0;
    for (int jj = 1; jj < 7; ++jj) if (fabs(B) >= X[jj]) I = jj;
    double CP = 0.0;
    if (T > TC) CP = 0.161 * T + 0.00279 * pow(T, 3.0);
    if (T < TCS) CP = AB[I] * pow(fabs(T), E[I]);
    if (T <= TC && T >= TCS) {
        const double CPS = AB[I] * pow(fabs(T), E[I]);
        const double CPN = 0.161 * T + 0.00279 * pow(T, 3.0);
        double F = 1.0;
        if (TCS < TC) F = (T - TCS) / (TC - TCS);
        CP = F * CPN + (1.0 - F) * CPS;
    }
    return CP;
}
/* silver.py:5-60 (np.piecewise: where two conditions hold the later one wins; none: 0) */
static double k_ag(double T)
{
    T = fmin(T, 1234.0);
    T = fmax(T, 1.0);
    if (T >= 1.0 && T < 13.0) return 3.5389 * pow(T, 4.0) - 8.2624e1 * pow(T, 3.0) + 2.7138e2 * sq(T) + 3.6897e3 * T;
    if (T >= 13.0 && T < 35.0) return -1.3578 * pow(T, 3.0) + 1.2791e2 * sq(T) - 4.1285e3 * T + 4.7377e4;
    if (T >= 35.0 && T < 100.0)
        return 1.8407e-4 * pow(T, 4.0) - 5.9684e-2 * pow(T, 3.0) + 7.241 * sq(T) - 3.9169e2 * T + 8.4877e3;
    if (T >= 100.0 && T < 273.0)
        return 1.4878e-7 * pow(T, 4.0) - 1.2548e-4 * pow(T, 3.0) + 3.9209e-2 * sq(T) - 5.4107 * T + 7.0959e2;
    if (T >= 273.0 && T < 1234.0) return 9.3528e-9 * pow(T, 3.0) - 2.6629e-5 * sq(T) - 5.4314e-2 * T + 4.4529e2;
    return 0.0;
}
/* silver.py:63-110 */
static double cp_ag(double T)
{
    T = fmin(T, 300.0);
    T = fmax(T, 1.0);
    double v;
    if (T >= 285.0) v = (2.343447902350777e2 - 2.343414867962295e2) / (285 - 284.9) * (T - 295) + 2.343447902350777e2;
    else if (T > 50.0)
        v = -0.000000104620955 * pow(T, 4.0) + 0.000091884831207 * pow(T, 3.0) - 0.030295300516950 * sq(T)
            + 4.598581892940776 * T - 52.331891977926269;
    else v = 0.047220784283425 * sq(T) - 0.018765676876719 * T - 1.946839691056091;
    return fmax(v, 1.0);
}
/* hastelloy_c276.py:5-53 */
static double k_hc276(double T)
{
    T = fmin(T, 350.0);
    T = fmax(T, 1.0);
    if (T >= 300.0) return (12.240619439579623 - 12.219768038241552) / (300 - 299) * (T - 300) + 12.240619439579623;
    if (T > 60.0) return 0.000000124577333 * pow(T, 3.0) - 0.000079993515434 * sq(T) + 0.035243632195403 * T + 5.503358179018721;
    return 0.000028975912976 * pow(T, 3.0) - 0.004655231435231 * sq(T) + 0.293992451992451 * T + 0.166939393939405;
}
/* hastelloy_c276.py:57-105 */
static double cp_hc276(double T)
{
    T = fmin(T, 350.0);
    T = fmax(T, 1.0);
    if (T >= 300.0) return (4.049021473239762e2 - 4.043151575124803e002) / (300 - 299) * (T - 300) + 4.049021473239762e2;
    if (T > 60.0) return 0.0000302086983 * pow(T, 3.0) - 0.0226456716880 * sq(T) + 6.0225562313795 * T - 179.3891242698739;
    return 0.000092969696970 * pow(T, 3.0) + 0.046381858141858 * sq(T) - 0.738681152181147 * T + 4.091909090909040;
}
/* solder_sn60_pb40.py:5-48 */
static double k_sn60pb40(double T)
{
    T = fmin(T, 350.0);
    T = fmax(T, 4.0);
    if (T >= 300.0) return 0.0902 * (T - 300.0) + 51.1010;
    return 2.667e-010 * pow(T, 5.0) - 2.483e-7 * pow(T, 4.0) + 8.769e-005 * pow(T, 3.0) - 0.01453 * sq(T) + 1.147 * T + 10.22;
}
/* solder_sn60_pb40.py:52-96 */
static double cp_sn60pb40(double T)
{
    T = fmin(T, 350.0);
    T = fmax(T, 1.0);
    double v;
    if (T >= 300.0) v = 0.2316 * (T - 300.0) + 181.0480;
    else v = 6.856e-010 * pow(T, 5.0) - 6.764e-7 * pow(T, 4.0) + 0.0002604 * pow(T, 3.0) - 0.0492 * sq(T) + 4.728 * T - 27.32;
    return fmax(v, 1.0);
}
static double mat_density(int mat)   /* density_cu / _nb3sn / _ss / _ge: constants */
{
    switch (mat) {
    case OSC2_MAT_CU_NIST: return 8900.0;
    case OSC2_MAT_NB3SN: return 8950.0;
    case OSC2_MAT_SS: return 7800.0;
    case OSC2_MAT_AL: return 2700.0;
    case OSC2_MAT_RE123: return 6380.0;
    case OSC2_MAT_MGB2: return 2570.0;                     /* magnesium_diboride.py density_mgb2 */
    case OSC2_MAT_NBTI: return 6160.0;                     /* niobium_titanium.py density_nbti */
    case OSC2_MAT_AG: return 10630.0;                       /* silver.py density_ag */
    case OSC2_MAT_HC276: return 8890.0;                    /* hastelloy_c276.py density_hc276 */
    case OSC2_MAT_SN60PB40: return 7310.0 * 0.6 + 11340.0 * 0.4;   /* solder_sn60_pb40.py density_sn60pb40 */
    default: return 2000.0;
    }
}

/* ---- channel correlations: channel.py --------------------------------------------- */
/* IFRICTION 124: blasius_formula channel.py:392-401 overwritten above Re 1e4 (channel.py:782-795);
 * laminar slot stays 0 so total = max(0, turbulent) (channel.py:1104-1115) */
static double friction_124(double re)
{
    re = fabs(re);
    double f = (0.316 * pow(re, -0.25)) / 4.0;
    if (re > 1e4) f = 2.21 * pow(re, -0.4) / 4.0;
    return fmax(0.0, f);
}
/* general_laminar_correlation channel.py:284-299: 16/Re where Re >= 1e-5, else the initial 0 */
static double friction_laminar(double re) { return re >= 1.0e-5 ? 16.0 / re : 0.0; }
/* _eval_total_friction_factor_L_lim_T (channel.py:1068-1101).  Observable quirk: eval_friction_factor (:1119-1138) calls
 * self.total_friction_factor_correlation(reynolds, nodal), so `nodal` binds to the L_lim_T parameter: llim = True = 1 in the
 * nodal pass, False = 0 at the Gauss points (never the default 4e3). */
static double total_llimt(double re, double lam, double turb, double llim)
{
    double tot = 0.0;
    if (re <= 2.0e3) tot = lam;
    if (re > 2.0e3 && re < llim) tot = (re - 2000.0) / (llim - 2000.0) * (turb - lam) + lam;
    if (re >= llim) tot = turb;
    return tot;
}
/* IFRICTION 102..109: turbulent_friction_with_newton_hole + _eval_ft_newton (channel.py:190-222, 349-390).  The loop runs
 * on the WHOLE array until the largest relative change is below 1e-2; re[] >= 0, tot[] out (total friction factor before
 * the multiplier: max(laminar, turbulent), 106: interpolation); returns the iteration count. */
static int newton_hole_array(const double *re, double *tot, long n, int ifr, double dh, double llim)
{
    const int fit64 = (ifr == 102 || ifr == 106 || ifr == 107);
    const double aa = fit64 ? 6.4 : 11.88, bb = fit64 ? 0.1717 : 0.039, cc = fit64 ? -0.3428 : -0.299;
    const double gg = ifr == 103 ? 3.0e-3 : ifr == 104 ? 1.5e-3 : ifr == 106 ? 2.0e-3 : ifr == 108 ? 2.40e-3 : ifr == 109 ? 5.30e-3 : 3.05e-3;
    const double tt = ifr == 104 ? 1.5e-3 : 1e-3;
    const double goverh = gg / tt, hd2 = 2.0 * tt / dh;
    for (long i = 0; i < n; ++i) tot[i] = 1e-2;
    int it = 0;
    double errmax = 10.0;
    while (errmax >= 1e-2 && it < 1000) {
        errmax = 0.0;
        for (long i = 0; i < n; ++i) {
            const double hp = tt / dh * re[i];
            const double hpiu = hp * sqrt(0.5 * tot[i]);
            const double rhpiu = aa * pow(hpiu, bb) * pow(goverh, cc);
            const double fn = 2.0 / pow(rhpiu - 2.5 * log10(hd2) - 3.75, 2.0);
            const double err = fabs((fn - tot[i]) / fn);
            if (err > errmax || err != err) errmax = err;
            tot[i] = fn;
        }
        ++it;
    }
    for (long i = 0; i < n; ++i) {
        const double lam = friction_laminar(re[i]);
        tot[i] = (ifr == 106) ? total_llimt(re[i], lam, tot[i], llim) : fmax(lam, tot[i]);
    }
    return it;
}
/* IFRICTION 110, 111, 112, 121 (channel.py:392-401, 403-413, 415-438, 700-717): total = max(laminar, turbulent)
 * (channel.py:1104-1115).  110-112: laminar slot = general_laminar_correlation.  121: the laminar slot calls
 * haaland_equation again, which writes the TURBULENT array, so laminar stays 0.  112 assigns by masks that
 * leave Re = 2e3 < . , Re = 4e3 exactly at the initial 0 of the turbulent array. */
static double friction_simple(int ifr, double re, double rough, double dh, double vf, double area, double llim, double lamc)
{
    re = fabs(re);
    double turb = 0.0, lam = 0.0;
    switch (ifr) {
    case 110: turb = (0.316 * pow(re, -0.25)) / 4.0; lam = friction_laminar(re); break;
    case 111: turb = 0.00128 + 0.1143 * pow(fmax(re, 4000.0), -0.311); lam = friction_laminar(re); break;
    case 112:
        if (re <= 2.0e3) turb = 64 / re / 4.0;
        if (re > 2.0e3 && re < 4.0e3) turb = 2.6471e-3 * pow(re, 0.3279) / 4.0;
        if (re > 4.0e3 && re < 3.0e4) turb = 0.3194 * pow(re, -0.25) / 4.0;
        if (re >= 3.0e4) turb = 1.0 / pow(1.8 * log10(re) / log10(10.0) - 1.64, 2.0) / 4.0;
        lam = friction_laminar(re);
        break;
    case 121: turb = pow(-1.8 * log10(pow(rough / dh / 3.7, 1.11) + (6.9 / re)), -2.0) / 4.0; break;
    /* katheder_correlation_bundle channel.py (204: 0.843, 0.0265; 205: 0.88, 0.051) */
    case 204: turb = pow(vf, 0.72) * (19.5 / pow(re, 0.843) + 0.0265) / 4.0; lam = friction_laminar(re); break;
    case 205: turb = pow(vf, 0.72) * (19.5 / pow(re, 0.88) + 0.051) / 4.0; lam = friction_laminar(re); break;
    /* darcy_forcheimer_porus_medium_correlation_bundle + _eval_jj_and_kk (206: 19.6e-9, 2.42, 5.80; 209: 20.9e-9, 19.1, 4.23) */
    case 206: case 209: {
        const double c0 = (ifr == 206) ? 19.6e-9 : 20.9e-9, c1 = (ifr == 206) ? 2.42 : 19.1, c2 = (ifr == 206) ? 5.80 : 4.23;
        const double jj = c1 / pow(vf, c2);
        const double kk = c0 * pow(vf, 3.0) / pow(1 - vf, 2.0);
        turb = pow(dh, 2.0) * vf / 2.0 / kk / re + (dh * pow(vf, 2.0)) / 2.0 * jj;
        lam = friction_laminar(re);
        break;
    }
    /* iter_conductor_hole (100, 101), channel.py:321-347 */
    case 100: case 101: {
        const double aa = ifr == 100 ? 0.45 : 0.36, bb = ifr == 100 ? -0.034 : -0.038, corr = (dh + 2 * 1e-3) / dh;
        turb = (aa * pow(re / corr, bb)) / 4.0 / pow(corr, 5.0);
        lam = friction_laminar(re);
        break;
    }
    /* 120: triangular duct, laminar 13.33 / min(Re, 2000) (:688-698), turbulent :650-662, total by interpolation (:1068-1101) */
    case 119:   /* rectangular duct (:664-686): laminar lamc / min(Re, 2000), lamc = 24 (1 - 1.3553 a + ...) from the host */
        return total_llimt(re, lamc / fmin(re, 2000.0), 0.00128 + 0.1143 * pow(fmax(re, 4000.0), -0.311), llim);
    case 120:
        return total_llimt(re, 13.33 / fmin(re, 2000.0), 0.00128 + 0.1143 * pow(fmax(re, 4000.0), -0.311), llim);
    /* hole laws 113 (memorandum_bessette_iter_cs_spiral_hole_y2015, channel.py:447-468: its closing statement
     * `turbulent / 4.0 / corr_diam ** 5` has no effect and is not applied), 114 (:470-480), 115 (:482-496), 118 (:636-648);
     * corr_diam of _eval_correction_diameter_hole (:305-319, tthick = 1e-3) */
    case 113: {
        const double corr = (dh + 2 * 1e-3) / dh, x = re / corr;
        if (x < 1.5e5) turb = 3.160 * pow(x, -0.249);
        if (x >= 1.5e5) turb = 0.216 * pow(x, -0.024);
        lam = friction_laminar(re);
        break;
    }
    case 114: turb = 0.0958 * pow(re, -0.181); lam = friction_laminar(re); break;
    case 115: { const double corr = 3.18e-3 / dh; turb = 0.25 * 0.3164 * pow(re / corr, -0.25) / corr; lam = friction_laminar(re); break; }
    case 118: turb = 0.1687 / 4.0 * pow(re, -0.1129); lam = friction_laminar(re); break;
    /* bundle laws 200 (tronza_iter_tf_correlation_bundle, :799-822), 201 (Nicollet, :824-840), 202 (Wanner, :842-853),
     * 203 (KSTAR, :855-866), 207 (LHC Poncet, :939-963: nothing assigned below Re 2e3), 210 (DEMO TF HTS, :1015-1037) */
    case 200: {
        const double doutct = dh + 2 * 1e-3;
        turb = 2.46 * pow(re / dh * 0.25e-3, -0.52) * dh / 0.25e-3
               * pow(area / (0.3 * (pow(39.8e-3, 2.0) - pow(doutct, 2.0)) * 0.25 * 3.141592653589793), 2.0) / 4.0;
        lam = friction_laminar(re);
        break;
    }
    case 201: turb = (0.0231 + 19.5 / pow(re / (6.0 / 5.0), 0.7953)) / pow(vf, 0.742) / 4.0; lam = friction_laminar(re); break;
    case 202: turb = 6.1 * pow(re, -0.51) / 4.0; lam = friction_laminar(re); break;
    case 203: turb = 0.4335 * pow(re, -0.263); lam = friction_laminar(re); break;
    case 207:
        if (re >= 2e3 && re < 4e3) turb = 2.6471e-3 * pow(re, 0.3279) / 4.0;
        if (re >= 4e3 && re < 3e4) turb = 0.3194 * pow(re, -0.25) / 4.0;
        if (re >= 3e4) turb = 1.0 / pow(1.8 * log10(re) / log10(10.0) - 1.64, 2.0) / 4.0;
        lam = friction_laminar(re);
        break;
    case 210:
        if (re < 1.5e3) turb = 47.65 * pow(re, -0.885) / 4.0;
        if (re >= 1.5e3 && re <= 2e5) turb = 1.093 * pow(re, -0.338) / 4.0;
        if (re > 2e5) turb = 0.0377 / 4.0;
        lam = friction_laminar(re);
        break;
    /* 211 (hts_cl_correlation_bundle, :1039-1066): the laminar slot calls the same method, which writes the
     * turbulent array (laminar stays 0), and the method sets FRICTION_MULTIPLIER to 1 (applied by the caller) */
    case 211:
        if (re < 1e3) turb = 194.002 * pow(re, -0.52) / 4.0;
        if (re >= 1e3 && re < 2e3) turb = 4.8563 + 4.87e-4 * re / 4.0;
        if (re >= 2e3) turb = 14.5148 * pow(re, -0.12) / 4.0;
        break;
    default: break;
    }
    return fmax(lam, turb);
}
/* IFRICTION 122: colebrook_formula channel.py:721-756 started from haaland_equation channel.py:700-717.
 * The loop runs on the WHOLE array until err.max() <= 1e-5 (or 100 iterations), so every element
 * performs the same, array-wide, number of iterations.  f[] in/out; returns that number. */
static int colebrook_array(const double *re, double *f, long n, double rough, double dh)
{
    for (long i = 0; i < n; ++i)
        f[i] = pow(-1.8 * log10(pow(rough / dh / 3.7, 1.11) + (6.9 / re[i])), -2.0) / 4.0;
    int it = 0;
    double errmax = 10.0;
    while (errmax > 1e-5 && it < 100) {
        errmax = 0.0;
        for (long i = 0; i < n; ++i) {
            const double fn = pow(-2.0 * log10((rough / dh / 3.7) + (2.51 / re[i] / sqrt(f[i]))), -2.0) / 4.0;
            const double err = fabs(f[i] - fn);
            if (err > errmax || err != err) errmax = err;
            f[i] = fn;
        }
        ++it;
    }
    return it;
}

/* dittus_boelter_nusselt_lower_limit (flag 1), channel.py:1156-1192 */
static double nusselt_1(double re, double pr)
{
    double nu = 0.023 * pow(re, 0.8) * pow(pr, 0.3);
    if (nu < 8.235) nu = 8.235;
    return nu;
}
/* rectangular_duct_enea_hts_cicc_htc (flag 212), channel.py:1245-1253.  Observable quirk:
 * eval_steady_state_htc (channel.py:1256-1273) forwards its own default nn = 0.3 positionally,
 * so the exponent the reference really uses is 0.3, not the correlation's default 0.71. */
/* Flag_htc_steady_corr 119 / 120 / 121: Dittus-Boelter with another lower limit (channel.py:1171-1243);
 * 211: correlation_211 (masks on Reynolds over an array of zeros) */
static double nusselt_more(int corr, double re, double pr, double dh)
{
    if (corr == 211) {
        double nu = 0.0;
        if (re < 1000.0) nu = 5.0969 * pow(re, 0.10);
        if (re >= 1000.0 && re < 2000.0) nu = 10.2029 - 3.3242e-5 * re;
        if (re >= 2000.0) nu = 0.0395 * pow(re, 0.73);
        return nu;
    }
    double lim = 4.01;                     /* 121 */
    if (corr == 120) lim = 2.181;
    if (corr == 119) {                     /* _eval_nusselt_lower_limit_119 */
        const double alpha = dh / (dh + 2 * 1e-3);
        const double NuTemp = 7.541 * (1.0 - 2.610 * alpha + 4.970 * pow(alpha, 2.0) - 5.119 * pow(alpha, 3.0) + 2.702 * pow(alpha, 4.0)
                                       - 0.548 * pow(alpha, 5.0));
        const double Nuflux = 8.235 * (1.0 - 10.6044 * alpha + 61.1755 * pow(alpha, 2.0) - 155.1803 * pow(alpha, 3.0)
                                       + 176.9203 * pow(alpha, 4.0) - 72.9236 * pow(alpha, 5.0));
        lim = (NuTemp + Nuflux) / 2.0;
    }
    double nu = 0.023 * pow(re, 0.8) * pow(pr, 0.3);
    if (nu < lim) nu = lim;
    return nu;
}
static double nusselt_212(double re) { return 0.42 * pow(re, 0.3); }

int osc2o_fit(int which, long n, const double *x, const double *y, const double *z, double p1, double p2, double *out)
{
    if (which == 32) {          /* friction_hole_newton: x = Reynolds (whole array), p1 = IFRICTION, p2 = Dh */
        double *ra = (double *)malloc(sizeof(double) * (size_t)(n > 0 ? n : 1));
        if (!ra) return -1;
        for (long i = 0; i < n; ++i) ra[i] = fabs(x[i]);
        newton_hole_array(ra, out, n, (int)p1, p2, 1.0);      /* nodal pass: L_lim_T = True */
        free(ra);
        (void)y; (void)z;
        return 0;
    }
    double tmax = 0.0, tmin = 0.0;
    if (which == 6 || which == 13 || which == 14) {
        tmax = tmin = x[0];
        for (long i = 1; i < n; ++i) { if (x[i] > tmax) tmax = x[i]; if (x[i] < tmin) tmin = x[i]; }
    }
    if (which == 17) {   /* Colebrook on the whole array: x = Reynolds (made positive like _quantities_initialization), p1 = roughness, p2 = Dh */
        double *re = malloc(sizeof(double) * (size_t)n);
        for (long i = 0; i < n; ++i) re[i] = fabs(x[i]);
        colebrook_array(re, out, n, p1, p2);
        for (long i = 0; i < n; ++i) out[i] = fmax(0.0, out[i]);
        free(re);
        return 0;
    }
    for (long i = 0; i < n; ++i) {
        switch (which) {
        case 0: out[i] = k_cu(x[i], y[i], p1); break;
        case 1: out[i] = cp_cu(x[i]); break;
        case 2: out[i] = k_nb3sn(x[i]); break;
        case 3: out[i] = cp_nb3sn(x[i], y[i], z[i], p1); break;
        case 4: out[i] = tc_nb3sn(x[i], y[i], p1, p2); break;
        case 5: out[i] = k_ss(x[i]); break;
        case 6: out[i] = cp_ss(x[i], tmax); break;
        case 7: out[i] = k_ge(x[i]); break;
        case 8: out[i] = cp_ge(x[i]); break;
        case 9: out[i] = friction_124(x[i]); break;
        case 10: out[i] = nusselt_1(x[i], y[i]); break;
        case 11: out[i] = nusselt_212(x[i]); break;
        case 28: out[i] = nusselt_more((int)p1, x[i], y[i], p2); break;   /* p1 = Flag_htc_steady_corr, x = Re, y = Pr, p2 = Dh */
        case 12: out[i] = rhoecu(x[i], y[i], p1); break;
        case 13: out[i] = k_al(x[i], tmin, tmax); break;
        case 14: out[i] = cp_al(x[i], tmin, tmax); break;
        case 15: out[i] = k_re123(x[i]); break;
        case 16: out[i] = cp_re123(x[i]); break;
        case 18: out[i] = k_ag(x[i]); break;
        case 19: out[i] = cp_ag(x[i]); break;
        case 20: out[i] = k_hc276(x[i]); break;
        case 21: out[i] = cp_hc276(x[i]); break;
        case 22: out[i] = k_sn60pb40(x[i]); break;
        case 23: out[i] = cp_sn60pb40(x[i]); break;
        case 25: out[i] = k_nbti(x[i]); break;
        case 29: out[i] = k_mgb2(x[i], y[i]); break;
        case 31: out[i] = mat_density((int)p1); break;
        case 30: out[i] = cp_mgb2(x[i]); break;
        case 26: out[i] = cp_nbti(x[i], y[i], z[i], out[i]); break;   /* x = T, y = B, z = Tcs; out[] holds Tc on entry */
        case 27: out[i] = tc_nbti(x[i], p1, p2); break;                            /* x = B, p1 = Bc20m, p2 = Tc0m */
        case 24: out[i] = friction_simple((int)p1, x[i], y[i], z[i], p2, y[i], 1.0, y[i]); break;   /* y: also the laminar numerator of 119 */   /* nodal pass: L_lim_T = True; */   /* p1 = IFRICTION, y = roughness (IFRICTION 200: CROSSECTION), z = Dh, p2 = void fraction */
        default: return -1;
        }
    }
    return 0;
}

/* ---- coolant tables ------------------------------------------------------------------ */
/* Bilinear interpolation on a regular (p, T) grid; this is the build's own definition
 * (the reference calls CoolProp.PropsSI, coolant.py:209-217), restated identically on
 * the device: index = clamp(floor((x - x_min) / step)), weights from the cell origin. */
double osc2o_fluid_lookup(const osc2o_fluid_table *tab, int prop, double p, double t)
{
    double fp = floor((p - tab->p_min) / tab->p_step), ft = floor((t - tab->t_min) / tab->t_step);
    if (!(fp >= 0.0)) fp = 0.0;
    if (fp > (double)(tab->n_p - 2)) fp = (double)(tab->n_p - 2);
    if (!(ft >= 0.0)) ft = 0.0;
    if (ft > (double)(tab->n_t - 2)) ft = (double)(tab->n_t - 2);
    const int ip = (int)fp, it = (int)ft;
    const double a = (p - (tab->p_min + fp * tab->p_step)) / tab->p_step;
    const double b = (t - (tab->t_min + ft * tab->t_step)) / tab->t_step;
    const double *f = tab->props + ((size_t)prop * tab->n_p + ip) * tab->n_t + it;
    const double f00 = f[0], f01 = f[1], f10 = f[tab->n_t], f11 = f[tab->n_t + 1];
    const double lo = f00 + b * (f01 - f00);
    const double hi = f10 + b * (f11 - f10);
    return lo + a * (hi - lo);
}

/* ---- operating conditions at the Gauss points ------------------------------------------ */
typedef struct {
    const osc2_topology *tp;
    const osc2_model *md;
    osc2o_opcond_io *io;
    int b0, b1;
} job_t;

/* np.linspace(start, stop, N)[i] */
static double linspace_at(double start, double stop, int N, int i)
{
    if (i == N - 1 && N > 1) return stop;
    const double div = (double)(N - 1), delta = stop - start;
    const double step = delta / div;
    if (step == 0.0) return ((double)i / div) * delta + start;
    return (double)i * step + start;
}

static void one_conductor(const osc2_topology *tp, const osc2_model *md, osc2o_opcond_io *io, int bb)
{
    const int M = tp->n_ch, S = tp->n_sol, d = 3 * M + S, N = io->n_nodes, E = N - 1;
    const double *x = io->sysvar + (size_t)bb * N * d;
    double *fg = io->fl_gauss + (size_t)bb * 9 * M * E;
    double *fn = io->fl_node + (size_t)bb * 4 * M;
    double *sg = io->sol_gauss + (size_t)bb * 3 * S * E;
    double *sqv = io->sol_q + (size_t)bb * 2 * S * E * 2;
    double *hff = io->htc_ff + (size_t)bb * 2 * tp->nff * E;
    double *hfs = io->htc_fs + (size_t)bb * tp->nfs * E;
    double *hss = io->htc_ss + (size_t)bb * tp->nss * E;
    double *hes = io->htc_es + (size_t)bb * tp->nes * 3 * E;
    const double time = io->time;
    /* per-channel Gauss quantities the interface models need */
    double *tf = malloc(sizeof(double) * 5 * (size_t)(M > 0 ? M : 1) * E);
    double *hst = tf + (size_t)M * E, *kf = hst + (size_t)M * E, *rf = kf + (size_t)M * E, *cpf = rf + (size_t)M * E;
    double *ts = malloc(sizeof(double) * 2 * (size_t)(S > 0 ? S : 1) * E);
    double *ks = ts + (size_t)S * E;

    double *rey = malloc(sizeof(double) * 2 * (size_t)E);
    for (int j = 0; j < M; ++j) {
        const osc2o_fluid_table *tab = &io->fluid[md->ch_fluid[j]];
        for (int e = 0; e < E; ++e) {
            /* coolant.py:134-155: arithmetic mean of the two nodes */
            const double v = (x[(size_t)e * d + j] + x[(size_t)(e + 1) * d + j]) / 2.0;
            const double p = (x[(size_t)e * d + M + j] + x[(size_t)(e + 1) * d + M + j]) / 2.0;
            const double T = (x[(size_t)e * d + 2 * M + j] + x[(size_t)(e + 1) * d + 2 * M + j]) / 2.0;
            double f[OSC2_N_FLUID_PROPS];
            for (int q = 0; q < OSC2_N_FLUID_PROPS; ++q) f[q] = osc2o_fluid_lookup(tab, q, p, T);
            /* coolant.py:165-180 */
            const double re = fabs(tp->ch_dh[j] * v * f[OSC2_FP_RHO] / f[OSC2_FP_MU]);
            const double grun = f[OSC2_FP_BETA] / f[OSC2_FP_KAPPA] / f[OSC2_FP_CV] / f[OSC2_FP_RHO];
            /* channel.py:1256-1273 */
            const double nu = (md->ch_htc_corr[j] == 212) ? nusselt_212(re) : (md->ch_htc_corr[j] == 1) ? nusselt_1(re, f[OSC2_FP_PRANDTL])
                              : nusselt_more(md->ch_htc_corr[j], re, f[OSC2_FP_PRANDTL], tp->ch_dh[j]);
            const double hsteady = nu * f[OSC2_FP_K] / tp->ch_dh[j];
            /* channel.py:1119-1138 (IFRICTION 122 is array-wide: filled in after this loop) */
            const int ifr = md->ch_ifriction[j];
            const double ftot = ((ifr == 124) ? friction_124(re)
                                 : (ifr != -99 && ifr != 122 && !(ifr >= 102 && ifr <= 109))
                                     ? friction_simple(ifr, re, md->ch_roughness[j], tp->ch_dh[j], md->ch_void_fraction[j], tp->ch_area[j], 0.0, md->ch_lam_coef[j])
                                 : 1.0) * (ifr == 211 ? 1.0 : md->ch_friction_mult[j]);
            rey[e] = re;
            fg[((size_t)0 * M + j) * E + e] = v;
            fg[((size_t)1 * M + j) * E + e] = p;
            fg[((size_t)2 * M + j) * E + e] = T;
            fg[((size_t)3 * M + j) * E + e] = f[OSC2_FP_RHO];
            fg[((size_t)4 * M + j) * E + e] = f[OSC2_FP_SOUND];
            fg[((size_t)5 * M + j) * E + e] = grun;
            fg[((size_t)6 * M + j) * E + e] = f[OSC2_FP_H];
            fg[((size_t)7 * M + j) * E + e] = f[OSC2_FP_CV];
            fg[((size_t)8 * M + j) * E + e] = ftot;
            tf[(size_t)j * E + e] = T;
            hst[(size_t)j * E + e] = hsteady;
            kf[(size_t)j * E + e] = f[OSC2_FP_K];
            rf[(size_t)j * E + e] = f[OSC2_FP_RHO];
            cpf[(size_t)j * E + e] = f[OSC2_FP_CP];
        }
        if (md->ch_ifriction[j] >= 102 && md->ch_ifriction[j] <= 109) {
            newton_hole_array(rey, rey + E, E, md->ch_ifriction[j], tp->ch_dh[j], 0.0);      /* Gauss points: L_lim_T = False */
            for (int e = 0; e < E; ++e) fg[((size_t)8 * M + j) * E + e] = rey[E + e] * md->ch_friction_mult[j];
        }
        if (md->ch_ifriction[j] == 122) {      /* total = max(laminar = 0, turbulent) * multiplier */
            colebrook_array(rey, rey + E, E, md->ch_roughness[j], tp->ch_dh[j]);
            for (int e = 0; e < E; ++e) fg[((size_t)8 * M + j) * E + e] = fmax(0.0, rey[E + e]) * md->ch_friction_mult[j];
        }
        /* nodal values the boundary conditions look at (fluid_component.py:255-565) */
        fn[0 * M + j] = x[j];
        fn[1 * M + j] = x[(size_t)(N - 1) * d + j];
        fn[2 * M + j] = osc2o_fluid_lookup(tab, OSC2_FP_RHO, x[M + j], x[2 * M + j]);
        fn[3 * M + j] = osc2o_fluid_lookup(tab, OSC2_FP_RHO, x[(size_t)(N - 1) * d + M + j], x[(size_t)(N - 1) * d + 2 * M + j]);
    }

    for (int l = 0; l < S; ++l) {
        const int is = 3 * M + l;
        const double biss = io->b_field[((size_t)bb * S + l) * 2], boss = io->b_field[((size_t)bb * S + l) * 2 + 1];
        const double eps = io->eps[(size_t)bb * S + l];
        const double *win = io->q_window + ((size_t)bb * S + l) * 4;
        const double q0 = io->q0[(size_t)bb * S + l];
        const int nmat = md->sol_nmat[l];
        /* conductor.py:4732-4744 */
        for (int e = 0; e < E; ++e) ts[(size_t)l * E + e] = fabs(x[(size_t)e * d + is] + x[(size_t)(e + 1) * d + is]) / 2.0;
        double tmax = ts[(size_t)l * E], tmin = ts[(size_t)l * E];
        for (int e = 1; e < E; ++e) {
            if (ts[(size_t)l * E + e] > tmax) tmax = ts[(size_t)l * E + e];
            if (ts[(size_t)l * E + e] < tmin) tmin = ts[(size_t)l * E + e];
        }
        for (int e = 0; e < E; ++e) {
            const double T = ts[(size_t)l * E + e];
            /* solid_component.py:383-408 */
            const double bg = fabs(linspace_at(biss, boss, N, e) + linspace_at(biss, boss, N, e + 1)) / 2.0;
            double tc = 0.0, tcs = 0.0;
            if (md->sol_kind[l] == OSC2_SOLID_STRAND_MIXED) {
                /* eps at Gauss = (eps + eps) / 2 (strand_component.py:392-399) */
                int nbti = 0;      /* strand_component.py:191-246: Tc by superconducting_material */
                for (int i = 0; i < md->sol_nmat[l]; ++i) nbti |= (md->sol_mat[l][i] == OSC2_MAT_NBTI);
                tc = nbti ? tc_nbti(bg, md->sol_bc20m[l], md->sol_tc0m[l])
                          : tc_nb3sn(bg, (eps + eps) / 2.0, md->sol_tc0m[l], md->sol_bc20m[l]);
                tcs = io->tcs ? io->tcs[((size_t)bb * S + l) * E + e] : 0.0;
            }
            double rho_i[OSC2_MAX_MATERIALS], cp_i[OSC2_MAX_MATERIALS], k_i[OSC2_MAX_MATERIALS];
            for (int i = 0; i < nmat; ++i) {
                const int mat = md->sol_mat[l][i];
                rho_i[i] = mat_density(mat);
                switch (mat) {
                case OSC2_MAT_CU_NIST: cp_i[i] = cp_cu(T); k_i[i] = k_cu(T, bg, md->sol_rrr[l]); break;
                case OSC2_MAT_NB3SN: cp_i[i] = cp_nb3sn(T, tcs, tc, md->sol_tc0m[l]); k_i[i] = k_nb3sn(T); break;
                case OSC2_MAT_SS: cp_i[i] = cp_ss(T, tmax); k_i[i] = k_ss(T); break;
                case OSC2_MAT_AL: cp_i[i] = cp_al(T, tmin, tmax); k_i[i] = k_al(T, tmin, tmax); break;
                case OSC2_MAT_RE123: cp_i[i] = cp_re123(T); k_i[i] = k_re123(T); break;
                case OSC2_MAT_AG: cp_i[i] = cp_ag(T); k_i[i] = k_ag(T); break;
                case OSC2_MAT_NBTI: cp_i[i] = cp_nbti(T, bg, tcs, tc); k_i[i] = k_nbti(T); break;
                case OSC2_MAT_MGB2: cp_i[i] = cp_mgb2(T); k_i[i] = k_mgb2(T, bg); break;
                case OSC2_MAT_HC276: cp_i[i] = cp_hc276(T); k_i[i] = k_hc276(T); break;
                case OSC2_MAT_SN60PB40: cp_i[i] = cp_sn60pb40(T); k_i[i] = k_sn60pb40(T); break;
                default: cp_i[i] = cp_ge(T); k_i[i] = k_ge(T); break;
                }
            }
            double rho, cp, kk;
            if ((md->sol_kind[l] == OSC2_SOLID_JACKET && nmat == 1) || md->sol_kind[l] == OSC2_SOLID_STRAND_STAB) {
                /* jacket_component.py:376-377, 414-415, 437-438; strand_stabilizer_component.py:156-198:
                 * single-material components return the fit itself */
                rho = rho_i[0]; cp = cp_i[0]; kk = k_i[0];
            } else {
                /* strand_mixed_component.py:415-514 / jacket_component.py:358-440 / stack_component.py:398-476
                 * (weights = layer thicknesses, sum = tape thickness):
                 * numerator = rho_i * w_i; sums over the material axis are sequential */
                double num[OSC2_MAX_MATERIALS], nsum = 0.0, cps = 0.0, ksum = 0.0;
                for (int i = 0; i < nmat; ++i) num[i] = rho_i[i] * md->sol_weight[l][i];
                nsum = num[0];
                for (int i = 1; i < nmat; ++i) nsum = nsum + num[i];
                cps = cp_i[0] * num[0];
                for (int i = 1; i < nmat; ++i) cps = cps + cp_i[i] * num[i];
                ksum = k_i[0] * md->sol_weight[l][0];
                for (int i = 1; i < nmat; ++i) ksum = ksum + k_i[i] * md->sol_weight[l][i];
                rho = nsum / md->sol_weight_sum[l];
                cp = cps / nsum;
                kk = ksum / md->sol_weight_sum[l];
            }
            sg[((size_t)0 * S + l) * E + e] = rho;
            sg[((size_t)1 * S + l) * E + e] = cp;
            sg[((size_t)2 * S + l) * E + e] = kk;
            ks[(size_t)l * E + e] = kk;
            /* conductor.py:4519-4556 with zero Joule terms + solid_component.py:415-576:
             * EXTFLX = Q0 on nodes [first, last] while TQBEG <= t <= TQEND; the previous
             * column is the t = 0 value (copied at step 1) */
            const int n0 = (int)win[0], n1 = (int)win[1];
            const int on_now = md->sol_heated[l] && time >= win[2] && time <= win[3];
            const int on_t0 = md->sol_heated[l] && 0.0 >= win[2] && 0.0 <= win[3];
            for (int side = 0; side < 2; ++side) {
                const int node = e + side;
                const int in = node >= n0 && node <= n1;
                sqv[((((size_t)side * S + l) * E + e) * 2) + 0] = (on_now && in) ? q0 : 0.0;
                sqv[((((size_t)side * S + l) * E + e) * 2) + 1] = (on_t0 && in) ? q0 : 0.0;
            }
        }
    }
    /* radiative exchange between jackets (HTC_choice +-3): nodal HTC (conductor.py:5161-5228 with
     * _inner_radiative_htc :5428-5443), nodal heat P * h * (T_other - T_self) (jacket_component.py:174-232),
     * added to Q1 / Q2 of both jackets pair by pair (conductor.py:4558-4579); present column only (the
     * previous column holds the zero of t = 0) */
    for (int k = 0; k < tp->nss; ++k) {
        const int ch = md->ss_choice[k];
        if (ch != 3 && ch != -3) continue;
        const int a = tp->ss[k][0], c = tp->ss[k][1];
        const double *pn = io->peri_ss_nodal + ((size_t)bb * tp->nss + k) * N;
        for (int e = 0; e < E; ++e)
            for (int side = 0; side < 2; ++side) {
                const int node = e + side;
                const double Ta = x[(size_t)node * d + 3 * M + a], Tc = x[(size_t)node * d + 3 * M + c];
                double h;
                if (ch == -3) h = md->ss_htc[k] * md->ss_mlt[k];
                else if (!(md->ss_rad_den[k] < 1.0e300)) h = 0.0;
                else h = 5.670374419e-08 * (sq(Ta) + sq(Tc)) * (Ta + Tc) / md->ss_rad_den[k] * md->ss_rad_mlt[k];
                const double ph = pn[node] * h;
                sqv[((((size_t)side * S + a) * E + e) * 2) + 0] += ph * (Tc - Ta);
                sqv[((((size_t)side * S + c) * E + e) * 2) + 0] += ph * (Ta - Tc);
            }
    }
    /* interface heat transfer coefficients, conductor.py:4867-5160 */
    for (int k = 0; k < tp->nfs; ++k) {
        const int j = tp->fs[k][0], l = tp->fs[k][1];
        const double tqbeg = io->q_window[((size_t)bb * S + l) * 4 + 2];
        for (int e = 0; e < E; ++e) {
            double h;
            if (md->fs_choice[k] == 2) {
                const double Ts = ts[(size_t)l * E + e], Tf = tf[(size_t)j * E + e];
                const double hK = 200.0 * (Ts + Tf) * (sq(Ts) + sq(Tf));
                double full = 0.0;
                if (time > tqbeg) {
                    const double htr = sqrt((kf[(size_t)j * E + e] * rf[(size_t)j * E + e] * cpf[(size_t)j * E + e])
                                            / (M_PI * (time - tqbeg)));
                    full = (hK * htr) / (hK + htr);
                }
                h = fmax(hst[(size_t)j * E + e] * md->fs_mlt[k], full);
            } else {
                h = md->fs_htc[k] * md->fs_mlt[k];
            }
            hfs[(size_t)k * E + e] = h;
        }
    }
    for (int k = 0; k < tp->nff; ++k) {
        const int ja = tp->ff[k][0], jb = tp->ff[k][1];
        for (int e = 0; e < E; ++e) {
            double ho, hc;
            if (md->ff_choice[k] == 2) {
                const double h1 = hst[(size_t)ja * E + e], h2 = hst[(size_t)jb * E + e];
                ho = md->ff_mlt[k] * h1 * h2 / (h1 + h2);
                const double cond = k_ss((tf[(size_t)ja * E + e] + tf[(size_t)jb * E + e]) / 2);
                const double rwall = md->ff_thick[k] / cond;
                hc = md->ff_mlt[k] / (1 / h1 + 1 / h2 + rwall);
            } else {
                ho = md->ff_htc[k] * md->ff_mlt[k];
                hc = ho;
            }
            hff[((size_t)0 * tp->nff + k) * E + e] = ho;
            hff[((size_t)1 * tp->nff + k) * E + e] = hc;
        }
    }
    for (int k = 0; k < tp->nss; ++k) {
        const int la = tp->ss[k][0], lb = tp->ss[k][1];
        for (int e = 0; e < E; ++e) {
            double h;
            if (md->ss_choice[k] == 3 || md->ss_choice[k] == -3) {
                h = 0.0;      /* "cond" keeps its zero initialisation for radiative pairs */
            } else if (md->ss_choice[k] == 1) {
                const double rr = md->ss_thick_rc[k] / ks[(size_t)la * E + e];
                const double rc = md->ss_thick_cr[k] / ks[(size_t)lb * E + e];
                h = md->ss_mlt[k] * (1.0 / (rr + md->ss_rcontact[k] + rc));
            } else {
                h = md->ss_htc[k] * md->ss_mlt[k];
            }
            hss[(size_t)k * E + e] = h;
        }
    }
    for (int k = 0; k < tp->nes; ++k)
        for (int e = 0; e < E; ++e) {
            hes[((size_t)k * 3 + 0) * E + e] = md->es_htc[k];
            hes[((size_t)k * 3 + 1) * E + e] = 0.0;
            hes[((size_t)k * 3 + 2) * E + e] = 0.0;
        }
    free(tf);
    free(ts);
    free(rey);
}

static void *worker(void *arg)
{
    job_t *j = (job_t *)arg;
    for (int b = j->b0; b < j->b1; ++b) one_conductor(j->tp, j->md, j->io, b);
    return NULL;
}

int osc2o_operating_conditions(const osc2_topology *topo, const osc2_model *model, osc2o_opcond_io *io, int n_threads)
{
    const int B = io->batch;
    if (n_threads < 1) n_threads = 1;
    if (n_threads > B) n_threads = B;
    if (n_threads == 1) {
        job_t j = {topo, model, io, 0, B};
        worker(&j);
        return 0;
    }
    pthread_t *th = malloc(sizeof(pthread_t) * (size_t)n_threads);
    job_t *jobs = malloc(sizeof(job_t) * (size_t)n_threads);
    for (int t = 0; t < n_threads; ++t) {
        jobs[t].tp = topo; jobs[t].md = model; jobs[t].io = io;
        jobs[t].b0 = (int)((long long)B * t / n_threads);
        jobs[t].b1 = (int)((long long)B * (t + 1) / n_threads);
        pthread_create(&th[t], NULL, worker, &jobs[t]);
    }
    for (int t = 0; t < n_threads; ++t) pthread_join(th[t], NULL);
    free(th);
    free(jobs);
    return 0;
}
