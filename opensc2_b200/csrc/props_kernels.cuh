// K_P: operating conditions at the Gauss points, evaluated on the device from the
// device-resident state -- what the reference does on the host before every step():
//   Conductor.operating_conditions_em / _th   conductor.py:4591-4744
//   Conductor.eval_transp_coeff               conductor.py:4822-5406 (choices +-2, +-1, env constants)
//   Conductor.build_heat_source               conductor.py:4423-4581, solid_component.py:415-576
//   Coolant Gauss p, T, v / Reynolds / Gruneisen  coolant.py:134-180
//   Channel friction (-99, 124) and Nusselt (1, 212)  channel.py:392-401, 782-795, 1104-1273
//   material fits  properties_of_materials/{copper,niobium3_tin,stainless_steel,glass_epoxy}.py
// Coolant properties come from (p, T) tables by bilinear interpolation (the reference calls
// CoolProp.PropsSI, coolant.py:209-217).  One thread per (element, conductor); every array is
// batch-innermost, so a warp reads and writes 256 contiguous bytes per value.
#pragma once

#include "osc2_dev.h"
#include "osc2_div.cuh"

struct FluidTableDev {
    int n_p, n_t;
    double p_min, p_step, t_min, t_step;
    const double *props;   // [10][n_p][n_t]
};

struct PropsDev {
    const osc2_model *model;
    FluidTableDev fluid[OSC2_MAX_FLUIDS];
    // sweep, batch-innermost
    const double *b_field;   // [S][2][B]
    const double *eps;       // [S][B]
    const double *q0;        // [S][B]
    const double *q_window;  // [S][4][B]
    const double *tcs;       // [S][E][B] or nullptr
    const double *jop;       // [S][B] operating current density (osc2_eval_tcs) or nullptr
    unsigned long long *margin_key;   // [S][B] ordered key of min over the Gauss points of (Tcs - T), or nullptr
    double *tmax;            // [S][B] max Gauss temperature of each solid (whole-array branches)
    double *tmin;            // [S][B] min Gauss temperature (aluminium's whole-array range check)
    int *cb_iters;           // [M][B] array-wide Colebrook iteration count (IFRICTION 122)
    double *sol_const;       // [S][2] RRR-only constants of the copper fit: p7, rhoecu0(273 K)
    double *ch_scratch;      // [2][M][E][B] coolant kernel -> interface kernel: Nu k / Dh, k rho cp
    double time;
};

namespace osc2p {

__device__ __forceinline__ double sq(double x) { return x * x; }
// x ** 3, x ** 4 ...: numpy calls pw(); repeated multiplication differs from the correctly
// rounded power by at most a few 1e-16 relative -- inside the stated tolerance of the fits
__device__ __forceinline__ double cube(double x) { return x * x * x; }
__device__ __forceinline__ double quart(double x) { const double y = x * x; return y * y; }
// x ** y for x > 0 as exp(y log x): |y log x| stays below ~40 in every fit here, so the result is within
// a few 1e-15 of the correctly rounded power (the fits' stated tolerance is 2e-12) at a third of pw()'s cost
// (ONE copy of each: inlined, the ~45 expansions of exp/log made the kernel 150 KB of code, and ncu
// showed 3 stall cycles per issued instruction waiting for the instruction cache)
//
// exp and log here are table-driven (ncu, round 2: the 42 pw() calls of a CASE_1 element were 40 % of K_P's
// instructions, ~80 each with the CUDA library's exp and log):
//   log x, x = 2^e m:  m / c_i = 1 + r with c_i the centre of the 1/128-wide interval of m, |r| < 2^-8;
//                      log x = e ln2 + log c_i + (r - r^2/2 + ... - r^6/6)           table: 1/c_i, log c_i
//   exp z:             z = k ln2/64 + r, |r| <= ln2/128; exp z = 2^(k/64) (1 + r + ... + r^5/120)   table: 2^(k/64)
// Measured against long double over x in [1e-8, 1e12], y in [-10, 52] (tools/pw_accuracy.c): x^y within 2.5e-13
// relative where exp(y log x) with the library functions is within 1.3e-13 -- both errors are the rounding of
// y log x amplified by |y log x| (up to 600 in that sweep; below 40 in every fit here, i.e. < 2e-14), inside the
// stated 2e-12 of the fits.  Arguments outside the normal positive range (Re = 0 ...) take the library functions.
// The tables are written once per device by osc2_upload_math_tables().
__device__ double2 osc2_log_table[128];      // 1 / c_i (rounded), -log of that rounded value
__device__ double osc2_exp_table[64];        // 2^(k / 64)
__device__ __noinline__ double slow_log(double x) { return log(x); }
__device__ __noinline__ double slow_exp(double z) { return exp(z); }
__device__ __forceinline__ double flog(double x)
{
    const int hi = __double2hiint(x), lo = __double2loint(x);
    if ((unsigned)(hi - 0x00100000) >= 0x7fe00000u) return slow_log(x);          // zero, negative, subnormal, inf, NaN
    const int e = (hi >> 20) - 1023, i = (hi >> 13) & 127;
    const double m = __hiloint2double((hi & 0x000fffff) | 0x3ff00000, lo);
    const double2 t = osc2_log_table[i];
    const double r = __fma_rn(m, t.x, -1.0);
    double pl = __fma_rn(r, -1.0 / 6, 1.0 / 5);
    pl = __fma_rn(r, pl, -0.25);
    pl = __fma_rn(r, pl, 1.0 / 3);
    pl = __fma_rn(r, pl, -0.5);
    const double l1 = __fma_rn(__dmul_rn(r, r), pl, r);
    const double ed = (double)e;
    const double sum = __fma_rn(ed, 0x1.62e42fefa38p-1, t.y);                // ln2 high part: 42 bits, e * high is exact
    return __dadd_rn(__dadd_rn(sum, l1), __dmul_rn(ed, 0x1.ef35793c7673p-45));
}
__device__ __forceinline__ double fexp(double z)
{
    if (!(fabs(z) < 700.0)) return slow_exp(z);
    const double tk = __fma_rn(z, 64.0 / 0.6931471805599453094172321, 0x1.8p52);   // round to nearest integer
    const double kd = __dadd_rn(tk, -0x1.8p52);
    const int k = __double2loint(tk);
    double r = __fma_rn(-kd, 0x1.62e42fefa38p-7, z);                         // ln2 / 64, high and low part
    r = __fma_rn(-kd, 0x1.ef35793c7673p-51, r);
    double pe = __fma_rn(r, 1.0 / 120, 1.0 / 24);
    pe = __fma_rn(r, pe, 1.0 / 6);
    pe = __fma_rn(r, pe, 0.5);
    const double q = __fma_rn(__dmul_rn(r, r), pe, r);
    const double T = osc2_exp_table[k & 63];
    const double v = __fma_rn(T, q, T);
    return __hiloint2double(__double2hiint(v) + ((k >> 6) << 20), __double2loint(v));
}
__device__ __noinline__ double pw(double x, double y) { return fexp(y * flog(x)); }
__device__ __noinline__ double pw10(double y) { return fexp(y * 2.302585092994046); }
// four independent powers in one body: the instruction streams interleave (a single pw() is one dependent chain of
// ~25 FP64 operations, and K_P's solid kernel has only 4 warps per scheduler to hide it with)
struct Pw4 { double a, b, c, d; };
__device__ __noinline__ Pw4 pw4(double x0, double y0, double x1, double y1, double x2, double y2, double x3, double y3)
{
    Pw4 o;
    const double l0 = flog(x0), l1 = flog(x1), l2 = flog(x2), l3 = flog(x3);
    o.a = fexp(y0 * l0); o.b = fexp(y1 * l1); o.c = fexp(y2 * l2); o.d = fexp(y3 * l3);
    return o;
}
// exp(y0 L), exp(y1 L), exp(y2 L2) for logarithms already taken (the copper fits raise t to three powers)
struct Pw3 { double a, b, c; };
__device__ __noinline__ Pw3 exp3(double z0, double z1, double z2)
{
    Pw3 o;
    o.a = fexp(z0); o.b = fexp(z1); o.c = fexp(z2);
    return o;
}

// copper.py:221-275
// lt = log(t) (shared with k_cu: one logarithm feeds every power of t; log(50 / t) = log 50 - log t)
__device__ __forceinline__ double rhoecu0_l(double lt, double rrr)
{
    const double P0 = 1.553e-8, P1 = 1.171e-17, P2 = 4.49, P3 = 3.841e10, P4 = 1.14, P6 = 6.428, P7 = 0.4531;
    const double rho0 = P0 / rrr;
    const Pw3 e = exp3(P2 * lt, (P2 - P4) * lt, P6 * (3.9120230054281460586 - lt));
    const double rhoi = (P1 * e.a) / (1.0 + P1 * P3 * e.b * fexp(-e.c));
    const double rhoi0 = P7 * rhoi * rho0 / (rhoi + rho0);
    return rho0 + rhoi + rhoi0;
}
__device__ __noinline__ double rhoecu0(double t, double rrr) { return rhoecu0_l(flog(t), rrr); }
// copper.py:140-217, with rhoecu0(t, rrr) and rhoecu0(273, rrr) passed in (evaluated once)
__device__ inline double rhoecu(double t, double b, double rrr, double rho_t, double rho_273)
{
    if (!(rrr > 1.0 && b >= 0 && t > 0)) return 0.0;
    double ccc = 0.0;
    if (b > 0) {
        const double xxx = rho_273 / rho_t * b;
        const double L = flog(xxx) * 0.43429448190325182765;
        double aaa = 0.0;
        aaa = aaa + -2.662 * 1.0;
        aaa = aaa + 0.3168 * L;
        aaa = aaa + 0.6229 * sq(L);
        aaa = aaa + -0.1839 * cube(L);
        aaa = aaa + 0.01827 * quart(L);
        ccc = pw10(aaa);
    }
    return rho_t * (1.0 + ccc);
}
// copper.py:5-71.  p7 = 0.838 / (beta / 0.0003) ** 0.1661 and rho_273 = rhoecu0(273, rrr) depend on
// RRR only and are computed once per model (osc2_model_consts_kernel)
__device__ __noinline__ double k_cu(double t, double b, double rrr, double p7, double rho_273)
{
    const double beta = 0.634 / rrr;
    const double P1 = 1.754e-8, P2 = 2.763, P3 = 1102, P4 = -0.165, P6 = 1.756;     // P5 = 70
    const double W0 = beta / t;
    const double lt = flog(t);
    const Pw3 e = exp3(P2 * lt, (P2 + P4) * lt, P6 * (4.2484952420493593784 - lt));     // log(P5 / t) = log 70 - log t
    const double Wi = (P1 * e.a) / (1 + (P1 * P3 * e.b * fexp(-e.c)));
    const double Wi0 = p7 * Wi * W0 / (Wi + W0);
    const double rho_t = rhoecu0_l(lt, rrr);
    return 1.0 / (Wi + W0 + Wi0) * rho_t / rhoecu(t, b, rrr, rho_t, rho_273);
}
// copper.py:75-136
__device__ __noinline__ double cp_cu(double t)
{
    if (!(t >= 4 && t <= 300)) return 0.0;
    const double L = flog(t) * 0.43429448190325182765;
    const double L2 = sq(L), L3 = L2 * L, L4 = L2 * L2;
    return pw10(-1.91844 + -0.15973 * L + 8.61013 * L2 + -18.996 * L3 + 21.9661 * L4
                         + -12.7328 * (L4 * L) + 3.54322 * (L3 * L3) + -0.3797 * (L4 * L3));
}
// niobium3_tin.py:371-434
__device__ inline double snbsn(double eps)
{
    const double Ca1 = 47.02, Ca2 = 11.76, EPS0a = 2.31e-3;
    const double EPSsh = Ca2 * EPS0a / sqrt(Ca1 * Ca1 - Ca2 * Ca2);
    return 1.0 + 1.0 / (1.0 - Ca1 * EPS0a)
                     * (Ca1 * (sqrt(EPSsh * EPSsh + EPS0a * EPS0a) - sqrt(sq(eps - EPSsh) + EPS0a * EPS0a)) - Ca2 * eps);
}
// niobium3_tin.py:437-490
__device__ __noinline__ double tc_nb3sn(double b, double eps, double tc0m, double bc20m)
{
    const double blim = fmax(b, 0.01);
    const double s = snbsn(eps);
    const double blcase = blim / (bc20m * s);
    if (!(blcase < 1.0)) return 0.0;
    return tc0m * pw(s, 1.0 / 3.0) * pw(1.0 - blcase, 1.0 / 1.52);
}
// niobium3_tin.py:62-124
__device__ __noinline__ double k_nb3sn(double tt)
{
    const double T0 = 21.9620369;
    tt = fmin(tt, 750.0);
    tt = fmax(tt, 4.0);
    if (tt <= T0) return 1.3905e14 * pw(tt, 3.66343961) / pw(76.6127237 + tt, 9.3220843);
    if (tt > T0) return -7.1614565e-1 * log(tt) + 5.40229689;
    return 0.0;
}
// niobium3_tin.py:127-262
__device__ __noinline__ double cp_nb3sn(double T, double TCS, double TC, double TC0)
{
    const double AA = 38.2226877, BB = -848.364226, CC = 1415.13808, DD = -346.837966;
    const double a = 6.80458608, b = 59.9209182, c = 25.8286334, d = 8.77918335;
    T = fmin(T, 1000.0);
    double CPN = 0.0;
    if (T <= 10.0) CPN = 7.5475e-3 * sq(T);
    else if (T > 10.0 && T <= 20.0) CPN = (-0.3 + 0.00375 * sq(T)) / 0.09937;
    else if (T > 20.0)
        CPN = AA * T / (a + T) + BB * sq(T) / sq(b + T) + CC * cube(T) / cube(c + T)
              + DD * quart(T) / quart(d + T);
    double CPNTC = 0.0;
    if (TC <= 10.0) CPNTC = 7.5475e-3 * sq(TC);
    else if (TC > 10.0 && TC <= 20.0) CPNTC = (-0.3 + 0.00375 * sq(TC)) / 0.09937;
    double CPS = 0.0;
    if (T <= TC) {
        const double TC_norm = TC / TC0;
        const double DBC2DT = -0.46306 - 0.067830 * TC;
        const double DELCP = 1500.0 * sq(DBC2DT) / (2.0 * sq(27.2 / (1.0 + (0.34 * TC_norm))) - 1.0);
        CPS = (CPNTC + DELCP) * cube(T / TC);
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
// the four-term rational form shared by the SS and glass-epoxy fits
__device__ __noinline__ double four_term(double T, double C0, double C1, double C2, double C3, double s0, double s1,
                                   double s2, double s3, double n0, double n1, double n2, double n3)
{
    // x / (s + T)^n as x (s + T)^-n: four FP64 divisions (~40 instructions each) less, one rounding apart
    const Pw4 w = pw4(s0 + T, -n0, s1 + T, -n1, s2 + T, -n2, s3 + T, -n3);
    return C0 * T * w.a + C1 * sq(T) * w.b + C2 * cube(T) * w.c + C3 * quart(T) * w.d;
}
// stainless_steel.py:5-38
__device__ inline double k_ss(double t)
{
    return four_term(t, 21224.0426, 103.007623, 135.870511, 0.07942859, 10844822.8, 6800262.44, 34.4411414,
                     0.23194142, 1.07801291, 1.88806818, 3.44233222, 3.20629457);
}
// stainless_steel.py:41-88 -- branch on max(T) of the whole array
__device__ inline double cp_ss(double T, double tmax)
{
    if (tmax <= 1.0) return 0.498 * T + 3.71e-4 * cube(T);
    if (tmax < 12.8973011) {
        T = fmin(T, 1253.0);
        return 0.46437387 * T + 0.00035157 * cube(T);
    }
    T = fmin(T, 1253.0);
    return four_term(T, -468.772387, -81.2353649, 827.209072, 29.9811823, 422.117262, 1235598.51, 51.2732976,
                     0.44591988, 0.57149148, 2.0006633, 2.79384889, 3.25864797);
}
// glass_epoxy.py:5-42, 45-84
__device__ inline double k_ge(double t)
{
    t = fmin(t, 1000.0);
    t = fmax(t, 1.0);
    return four_term(t, 0.106635834, 4.254937095, 0.369710568, -0.220612963, 4.374557267, 78.36483015,
                     0.896107679, 0.261916360, 0.551227580, 2.274402853, 2.979875183, 3.644905343);
}
__device__ inline double cp_ge(double t)
{
    t = fmin(t, 1000.0);
    t = fmax(t, 1.0);
    return four_term(t, 0.309399374, -1387.509141, 283.0320589, 1.822153805, 113.1687618, 80.88333136,
                     26.27075203, 1.104229335, 0.254475819, 1.938523579, 2.603466143, 3.522809815);
}
// aluminium.py:5-77 -- zeros when any element of the array is outside [2, 1000] K
__device__ __noinline__ double k_al(double T, double tmin, double tmax)
{
    if (tmin < 2.0 || tmax > 1000.0) return 0.0;
    if (T >= 2.0 && T <= 4.0) return 4.2345 * (T - 2.0) + 6.6510;
    if (T > 4.0 && T <= 6.0) return 4.2000 * (T - 4.0) + 15.1200;
    if (T > 6.0 && T <= 8.0) return 4.3650 * (T - 6.0) + 23.5200;
    if (T > 8.0 && T <= 10.0) return 4.8550 * (T - 8.0) + 32.2500;
    if (T > 10.0 && T <= 20.0) return 4.3430 * (T - 10.0) + 41.9600;
    if (T > 20.0 && T <= 40.0) return 3.8805 * (T - 20.0) + 85.3900;
    if (T > 40.0 && T <= 60.0) return 2.0375 * (T - 40.0) + 163.000;
    if (T > 60.0 && T <= 80.0) return 0.5360 * (T - 60.0) + 203.75;
    if (T > 80.0 && T <= 100.0) return 0.1145 * (T - 80.0) + 214.47;
    if (T > 100.0 && T <= 200.0) return -0.0452 * (T - 100.0) + 216.76;
    if (T > 200.0 && T <= 300.0) return -0.0291 * (T - 200.0) + 212.24;
    if (T > 300.0 && T <= 1000.0) return 209.3300;
    return 0.0;
}
// aluminium.py:80-164
__device__ __noinline__ double cp_al(double T, double tmin, double tmax)
{
    if (tmin < 2.0 || tmax > 1000.0) return 0.0;
    if (T >= 2.0 && T <= 11.0578562) return 0.04257809 * T + 0.00340313 * sq(T) + 0.00063472 * cube(T);
    if (T > 11.0578562 && T <= 53.2685055)
        return -2.50413265 + 0.71747923 * T + -0.06270275 * sq(T) + 0.0031683 * cube(T) + -2.0154e-05 * quart(T);
    if (T > 53.2685055 && T <= 1000.0)
        return four_term(T, 11458.0127, -438.603857, 331.057118, -10001.7006, -7.12302329, -34.6373149, -29.4649313,
                         -3.30673113, 0.9531781, 2.00135991, 2.974805, 3.94384164);
    return 0.0;
}
// rare_earth_123.py:65-135, 138-190
__device__ inline double k_re123(double T)
{
    T = fmin(T, 300.0);
    T = fmax(T, 4.0);
    const double T2 = sq(T), T3 = T2 * T, T4 = T2 * T2, T5 = T4 * T;
    if (T <= 70.0)
        return -1.266103106492942e-5 * T5 + 0.002670105477219 * T4 + -0.197035302542769 * T3 + 4.933962659604384 * T2
               + 29.651536939501670 * T + 66.578192505447330;
    return -1.316704893540544e-9 * T5 + 2.636151594117440e-06 * T4 + -0.001601689632073 * T3 + 0.428760641312538 * T2
           + -53.306643333287260 * T + 3.682252343338599e03;
}
__device__ inline double cp_re123(double T)
{
    T = fmin(T, 300.0);
    T = fmax(T, 4.0);
    const double T2 = sq(T), T3 = T2 * T, T4 = T2 * T2, T5 = T4 * T;
    return -7.567485538209158e-10 * T5 + 6.351452642016898e-07 * T4 + -1.947975786547597e-04 * T3 + 0.023616673974415 * T2
           + 0.239331954284042 * T + -1.096191721280114;
}
// magnesium_diboride.py:8-77 (the copper NIST form at its default rrr = 100), :79-139
__device__ __noinline__ double k_mgb2(double t, double b)
{
    const double beta = 0.634 / 100.0, betat = beta / 0.0003;
    return k_cu(t, b, 100.0, 0.838 / pw(betat, 0.1661), rhoecu0(273.0, 100.0));
}
__device__ __noinline__ double cp_mgb2(double t)
{
    if (!(t >= 4 && t <= 300)) return 0.0;
    const double L = log10(t), L2 = L * L, L3 = L2 * L, L4 = L2 * L2, L5 = L4 * L, L6 = L3 * L3, L7 = L6 * L;
    return pw10(-1.91844 + -0.15973 * L + 8.61013 * L2 + -18.996 * L3 + 21.9661 * L4 + -12.7328 * L5 + 3.54322 * L6 + -0.3797 * L7);
}
// niobium_titanium.py:9-62, 64-209, 274-286 (superconductor of a StrandMixedComponent)
__device__ __noinline__ double k_nbti(double T)
{
    T = fmin(T, 750.0);
    T = fmax(T, 1.0);
    double k = 6.5 * T / (1.0 + cube(T / 25.0));
    if (T >= 50.0) k = k + 21.5 * (1.0 - exp(-(T - 50.0) / 50.0));
    return k;
}
__device__ __noinline__ double tc_nbti(double b, double bc20, double tc0)
{
    return (b <= bc20) ? tc0 * pw(1 - b / bc20, 1 / 1.7) : 0.0;
}
__device__ __noinline__ double cp_nbti(double T, double B, double TCS, double TC)
{
    if (T > 1.0) {
        T = fmin(T, 1000.0);
        return 17.971311 * pw(T / (88.768525 + T), 1.0456476) + -3215.18087 * pw(T / (3.66034399 + T), 51.7204943)
               + 4325.5737 * pw(T / (38.0453266 + T), 4.20661339) + -711.296119 * pw(T / (14.67806 + T), 6.01104574);
    }
    // T <= 1 K: field-dependent superconducting-state table, mixed between Tcs and Tc
    const double ab[7] = {0.341000e-01, 0.282900e-01, 0.219200e-01, 0.158600e-01, 0.993000e-02, 0.491000e-02, 0.152000e-02};
    const double ee[7] = {0.233300e01, 0.244600e01, 0.259700e01, 0.280600e01, 0.310700e01, 0.356800e01, 0.434800e01};
    const double fb = fabs(B);
    const int I = fb >= 7.0 ? 6 : fb >= 6.0 ? 5 : fb >= 5.0 ? 4 : fb >= 4.0 ? 3 : fb >= 3.0 ? 2 : fb >= 2.0 ? 1 : 0;
    double CP = 0.0;
    if (T > TC) CP = 0.161 * T + 0.00279 * cube(T);
    if (T < TCS) CP = ab[I] * pw(fabs(T), ee[I]);
    if (T <= TC && T >= TCS) {
        const double CPS = ab[I] * pw(fabs(T), ee[I]);
        const double CPN = 0.161 * T + 0.00279 * cube(T);
        double F = 1.0;
        if (TCS < TC) F = (T - TCS) / (TC - TCS);
        CP = F * CPN + (1.0 - F) * CPS;
    }
    return CP;
}
// silver.py:5-110, hastelloy_c276.py:5-105, solder_sn60_pb40.py:5-96 (StackComponent layers).  np.piecewise:
// where two conditions hold the later one wins, where none holds the value is 0.
__device__ __noinline__ double k_ag(double T)
{
    T = fmin(T, 1234.0);
    T = fmax(T, 1.0);
    const double T2 = sq(T), T3 = T2 * T, T4 = T2 * T2;
    if (T < 13.0) return 3.5389 * T4 - 8.2624e1 * T3 + 2.7138e2 * T2 + 3.6897e3 * T;
    if (T < 35.0) return -1.3578 * T3 + 1.2791e2 * T2 - 4.1285e3 * T + 4.7377e4;
    if (T < 100.0) return 1.8407e-4 * T4 - 5.9684e-2 * T3 + 7.241 * T2 - 3.9169e2 * T + 8.4877e3;
    if (T < 273.0) return 1.4878e-7 * T4 - 1.2548e-4 * T3 + 3.9209e-2 * T2 - 5.4107 * T + 7.0959e2;
    if (T < 1234.0) return 9.3528e-9 * T3 - 2.6629e-5 * T2 - 5.4314e-2 * T + 4.4529e2;
    return 0.0;
}
__device__ __noinline__ double cp_ag(double T)
{
    T = fmin(T, 300.0);
    T = fmax(T, 1.0);
    const double T2 = sq(T), T3 = T2 * T, T4 = T2 * T2;
    double v;
    if (T >= 285.0) v = (2.343447902350777e2 - 2.343414867962295e2) / (285 - 284.9) * (T - 295) + 2.343447902350777e2;
    else if (T > 50.0)
        v = -0.000000104620955 * T4 + 0.000091884831207 * T3 - 0.030295300516950 * T2 + 4.598581892940776 * T - 52.331891977926269;
    else v = 0.047220784283425 * T2 - 0.018765676876719 * T - 1.946839691056091;
    return fmax(v, 1.0);
}
__device__ __noinline__ double k_hc276(double T)
{
    T = fmin(T, 350.0);
    T = fmax(T, 1.0);
    const double T2 = sq(T), T3 = T2 * T;
    if (T >= 300.0) return (12.240619439579623 - 12.219768038241552) / (300 - 299) * (T - 300) + 12.240619439579623;
    if (T > 60.0) return 0.000000124577333 * T3 - 0.000079993515434 * T2 + 0.035243632195403 * T + 5.503358179018721;
    return 0.000028975912976 * T3 - 0.004655231435231 * T2 + 0.293992451992451 * T + 0.166939393939405;
}
__device__ __noinline__ double cp_hc276(double T)
{
    T = fmin(T, 350.0);
    T = fmax(T, 1.0);
    const double T2 = sq(T), T3 = T2 * T;
    if (T >= 300.0) return (4.049021473239762e2 - 4.043151575124803e002) / (300 - 299) * (T - 300) + 4.049021473239762e2;
    if (T > 60.0) return 0.0000302086983 * T3 - 0.0226456716880 * T2 + 6.0225562313795 * T - 179.3891242698739;
    return 0.000092969696970 * T3 + 0.046381858141858 * T2 - 0.738681152181147 * T + 4.091909090909040;
}
__device__ __noinline__ double k_sn60pb40(double T)
{
    T = fmin(T, 350.0);
    T = fmax(T, 4.0);
    if (T >= 300.0) return 0.0902 * (T - 300.0) + 51.1010;
    const double T2 = sq(T), T3 = T2 * T, T4 = T2 * T2, T5 = T4 * T;
    return 2.667e-010 * T5 - 2.483e-7 * T4 + 8.769e-005 * T3 - 0.01453 * T2 + 1.147 * T + 10.22;
}
__device__ __noinline__ double cp_sn60pb40(double T)
{
    T = fmin(T, 350.0);
    T = fmax(T, 1.0);
    double v;
    if (T >= 300.0) v = 0.2316 * (T - 300.0) + 181.0480;
    else {
        const double T2 = sq(T), T3 = T2 * T, T4 = T2 * T2, T5 = T4 * T;
        v = 6.856e-010 * T5 - 6.764e-7 * T4 + 0.0002604 * T3 - 0.0492 * T2 + 4.728 * T - 27.32;
    }
    return fmax(v, 1.0);
}
__device__ __noinline__ double mat_density_more(int mat)
{
    return mat == OSC2_MAT_AG ? 10630.0 : mat == OSC2_MAT_NBTI ? 6160.0 : mat == OSC2_MAT_HC276 ? 8890.0 : mat == OSC2_MAT_MGB2 ? 2570.0
           : (7310.0 * 0.6 + 11340.0 * 0.4);
}
__device__ inline double mat_density(int mat)
{
    if (mat > OSC2_MAT_RE123) return mat_density_more(mat);
    return mat == OSC2_MAT_CU_NIST ? 8900.0 : mat == OSC2_MAT_NB3SN ? 8950.0 : mat == OSC2_MAT_SS ? 7800.0
           : mat == OSC2_MAT_AL ? 2700.0 : mat == OSC2_MAT_RE123 ? 6380.0 : 2000.0;
}
__device__ inline double friction_124(double re);
// IFRICTION 122: haaland_equation channel.py:700-717 as the start, colebrook_formula channel.py:721-756
__device__ inline double haaland(double re, double rough, double dh)
{
    const double x = -1.8 * log10(pw(rough / dh / 3.7, 1.11) + (6.9 / re));
    return (1.0 / (x * x)) / 4.0;
}
__device__ inline double colebrook_step(double re, double f, double rough, double dh)
{
    const double x = -2.0 * log10((rough / dh / 3.7) + (2.51 / re / sqrt(f)));
    return (1.0 / (x * x)) / 4.0;
}
// number of iterations THIS element needs (the reference iterates the whole array until the largest
// error is <= 1e-5; each element's fixed-point iteration contracts, so the array-wide count is the max)
__device__ inline int colebrook_count(double re, double rough, double dh)
{
    double f = haaland(re, rough, dh);
    int it = 0;
    double err = 10.0;
    while (err > 1e-5 && it < 100) {
        const double fn = colebrook_step(re, f, rough, dh);
        err = fabs(f - fn);
        if (err != err) err = 10.0;
        f = fn;
        ++it;
    }
    return it;
}
// IFRICTION 102..109: turbulent_friction_with_newton_hole + _eval_ft_newton (channel.py:190-222, 349-390): the fixed-point
// iteration f <- 2 / (a (hp sqrt(f/2))^b (g/h)^c - 2.5 log10(2 h / Dh) - 3.75)^2 from f = 1e-2, run on the WHOLE array
// until the largest relative change is below 1e-2 -- like Colebrook's, an array-wide iteration count
struct HoleCoeff { double aa, bb, cc, gg, tt; };
__device__ inline HoleCoeff hole_coeff(int ifr)
{
    const bool fit64 = (ifr == 102 || ifr == 106 || ifr == 107);      // "coefficient for 7/9 mm spiral"
    HoleCoeff c;
    c.aa = fit64 ? 6.4 : 11.88; c.bb = fit64 ? 0.1717 : 0.039; c.cc = fit64 ? -0.3428 : -0.299;
    c.gg = ifr == 103 ? 3.0e-3 : ifr == 104 ? 1.5e-3 : ifr == 106 ? 2.0e-3 : ifr == 108 ? 2.40e-3 : ifr == 109 ? 5.30e-3 : 3.05e-3;
    c.tt = ifr == 104 ? 1.5e-3 : 1e-3;
    return c;
}
__device__ inline double hole_step(const HoleCoeff &c, double re, double f, double dh)
{
    const double goverh = c.gg / c.tt, hp = c.tt / dh * re, hd2 = 2.0 * c.tt / dh;
    const double hpiu = hp * sqrt(0.5 * f);
    const double rhpiu = c.aa * pw(hpiu, c.bb) * pw(goverh, c.cc);
    const double x = rhpiu - 2.5 * log10(hd2) - 3.75;
    return 2.0 / (x * x);
}
__device__ __noinline__ int hole_count(int ifr, double re, double dh)
{
    const HoleCoeff c = hole_coeff(ifr);
    double f = 1e-2, err = 10.0;
    int it = 0;
    while (err >= 1e-2 && it < 1000) {
        const double fn = hole_step(c, re, f, dh);
        err = fabs((fn - f) / fn);
        f = fn;
        ++it;
    }
    return it;
}
__device__ __noinline__ double hole_friction(int ifr, double re, double dh, int iters)
{
    const HoleCoeff c = hole_coeff(ifr);
    double f = 1e-2;
    for (int it = 0; it < iters; ++it) f = hole_step(c, re, f, dh);
    return f;
}
// _eval_total_friction_factor_L_lim_T (channel.py:1068-1101): laminar up to Re 2000, turbulent from L_lim_T, linear in Re in
// between (a NaN Reynolds number leaves the initial 0).  Observable quirk: eval_friction_factor (:1119-1138) calls it as
// total_friction_factor_correlation(reynolds, nodal), so `nodal` lands in the L_lim_T parameter -- L_lim_T is True = 1 in
// the nodal pass and False = 0 at the Gauss points, never its default 4e3: the total is the turbulent factor from Re = 1
// (nodes) / for every Re (Gauss points), and the laminar one only below
__device__ inline double total_llimt(double re, double lam, double turb, double llim)
{
    double tot = 0.0;
    if (re <= 2.0e3) tot = lam;
    if (re > 2.0e3 && re < llim) tot = (re - 2000.0) / (llim - 2000.0) * (turb - lam) + lam;
    if (re >= llim) tot = turb;
    return tot;
}
// array-wide iteration count of channel j's friction law, this element's contribution (0: the law has none)
__device__ inline int friction_iter_count(const osc2_model *md, const osc2_topology *tp, int j, double re)
{
    const int ifr = md->ch_ifriction[j];
    if (ifr == 122) return colebrook_count(re, md->ch_roughness[j], tp->ch_dh[j]);
    if (ifr >= 102 && ifr <= 109) return hole_count(ifr, re, tp->ch_dh[j]);
    return 0;
}
__device__ inline bool friction_iterates(int ifr) { return ifr == 122 || (ifr >= 102 && ifr <= 109); }
// turbulent slot of the hole laws 113 / 114 / 115 / 118 (channel.py:447-496, 636-648; the closing statement of 113 has
// no effect and is not applied) and of the bundle laws 200 / 201 / 202 / 203 / 207 / 210 / 211 (:799-866, 939-963,
// 1015-1066); re > 0 on entry.  211: the laminar slot stays 0 and the multiplier is forced to 1 (friction_total)
__device__ __noinline__ double friction_more(int ifr, double re, double dh, double vf, double area)
{
    double turb = 0.0;
    if (ifr == 100 || ifr == 101) {          // iter_conductor_hole (channel.py:321-347)
        const double aa = ifr == 100 ? 0.45 : 0.36, bb = ifr == 100 ? -0.034 : -0.038, corr = (dh + 2 * 1e-3) / dh;
        turb = (aa * pw(re / corr, bb)) / 4.0 / pw(corr, 5.0);
    } else if (ifr == 120) {                 // duct_demo_common_memo_turbulent_flow_hole (:650-662)
        turb = 0.00128 + 0.1143 * pw(fmax(re, 4000.0), -0.311);
    } else if (ifr == 113) {
        const double corr = (dh + 2 * 1e-3) / dh, x = re / corr;
        if (x < 1.5e5) turb = 3.160 * pw(x, -0.249);
        if (x >= 1.5e5) turb = 0.216 * pw(x, -0.024);
    } else if (ifr == 114) turb = 0.0958 * pw(re, -0.181);
    else if (ifr == 115) { const double corr = 3.18e-3 / dh; turb = 0.25 * 0.3164 * pw(re / corr, -0.25) / corr; }
    else if (ifr == 118) turb = 0.1687 / 4.0 * pw(re, -0.1129);
    else if (ifr == 200) {
        const double doutct = dh + 2 * 1e-3;
        turb = 2.46 * pw(re / dh * 0.25e-3, -0.52) * dh / 0.25e-3 * sq(area / (0.3 * (sq(39.8e-3) - sq(doutct)) * 0.25 * 3.141592653589793)) / 4.0;
    } else if (ifr == 201) turb = (0.0231 + 19.5 / pw(re / (6.0 / 5.0), 0.7953)) / pw(vf, 0.742) / 4.0;
    else if (ifr == 202) turb = 6.1 * pw(re, -0.51) / 4.0;
    else if (ifr == 203) turb = 0.4335 * pw(re, -0.263);
    else if (ifr == 207) {
        if (re >= 2e3 && re < 4e3) turb = 2.6471e-3 * pw(re, 0.3279) / 4.0;
        if (re >= 4e3 && re < 3e4) turb = 0.3194 * pw(re, -0.25) / 4.0;
        if (re >= 3e4) { const double x = 1.8 * log10(re) / 1.0 - 1.64; turb = 1.0 / (x * x) / 4.0; }
    } else if (ifr == 210) {
        if (re < 1.5e3) turb = 47.65 * pw(re, -0.885) / 4.0;
        if (re >= 1.5e3 && re <= 2e5) turb = 1.093 * pw(re, -0.338) / 4.0;
        if (re > 2e5) turb = 0.0377 / 4.0;
    } else if (ifr == 211) {
        if (re < 1e3) turb = 194.002 * pw(re, -0.52) / 4.0;
        if (re >= 1e3 && re < 2e3) turb = 4.8563 + 4.87e-4 * re / 4.0;
        if (re >= 2e3) turb = 14.5148 * pw(re, -0.12) / 4.0;
    }
    return turb;
}
// total friction factor of channel j (channel.py:1119-1138)
// IFRICTION 110 (Blasius), 111 (Bhatti-Shah), 112 (LHC magnets recipe), 121 (Haaland), 204 / 205 (Katheder), 206 / 209
// (Darcy-Forchheimer porous medium): channel.py:392-438, 700-717, katheder_correlation_bundle, _eval_jj_and_kk;
// total = max(laminar, turbulent) with the general laminar correlation 16/Re (channel.py:284-299, 1104-1115);
// for 121 the laminar slot calls the Haaland method again, which writes the turbulent array: laminar stays 0
__device__ __noinline__ double friction_simple(int ifr, double re, double rough, double dh, double vf, double area, double llim, double lamc)
{
    re = fabs(re);
    double turb = 0.0, lam = 0.0;
    if (ifr == 121) return fmax(0.0, haaland(re, rough, dh));
    lam = re >= 1.0e-5 ? 16.0 / re : 0.0;
    if (ifr == 204) turb = pw(vf, 0.72) * (19.5 / pw(re, 0.843) + 0.0265) / 4.0;        // Katheder (bundle)
    else if (ifr == 205) turb = pw(vf, 0.72) * (19.5 / pw(re, 0.88) + 0.051) / 4.0;
    else if (ifr == 206 || ifr == 209) {                                                // Darcy-Forchheimer (bundle)
        const double c0 = (ifr == 206) ? 19.6e-9 : 20.9e-9, c1 = (ifr == 206) ? 2.42 : 19.1, c2 = (ifr == 206) ? 5.80 : 4.23;
        const double jj = c1 / pw(vf, c2);
        const double kk = c0 * cube(vf) / sq(1 - vf);
        turb = sq(dh) * vf / 2.0 / kk / re + (dh * sq(vf)) / 2.0 * jj;
    }
    else if (ifr == 110) turb = (0.316 * pw(re, -0.25)) / 4.0;
    else if (ifr == 111) turb = 0.00128 + 0.1143 * pw(fmax(re, 4000.0), -0.311);
    else if (ifr == 119 || ifr == 120)      // rectangular (:664-686) / triangular (:688-698) duct: laminar c / min(Re, 2000), turbulent
        return total_llimt(re, (ifr == 119 ? lamc : 13.33) / fmin(re, 2000.0), friction_more(120, re, dh, vf, area), llim);   // :650-662
    else if (ifr != 112) return fmax(friction_more(ifr, re, dh, vf, area), ifr == 211 ? 0.0 : lam);
    else {
        if (re <= 2.0e3) turb = 64 / re / 4.0;
        if (re > 2.0e3 && re < 4.0e3) turb = 2.6471e-3 * pw(re, 0.3279) / 4.0;
        if (re > 4.0e3 && re < 3.0e4) turb = 0.3194 * pw(re, -0.25) / 4.0;
        if (re >= 3.0e4) { const double x = 1.8 * log10(re) / 1.0 - 1.64; turb = 1.0 / (x * x) / 4.0; }
    }
    return fmax(lam, turb);
}
// nodal: evaluated in the nodal pass (true) or at the Gauss points (false) -- only the interpolated totals care (total_llimt)
__device__ inline double friction_total(const osc2_model *md, const osc2_topology *tp, int j, double re, int cb_iters, bool nodal)
{
    const double llim = nodal ? 1.0 : 0.0;
    double f = 1.0;
    const int ifr = md->ch_ifriction[j];
    if (ifr == 124) f = friction_124(re);
    else if (ifr >= 102 && ifr <= 109) {    // Newton hole laws: 106 interpolates, the others take the maximum
        const double ra = fabs(re), lam = ra >= 1.0e-5 ? 16.0 / ra : 0.0;
        const double turb = hole_friction(ifr, ra, tp->ch_dh[j], cb_iters);
        f = (ifr == 106) ? total_llimt(ra, lam, turb, llim) : fmax(lam, turb);
    }
    else if (ifr != -99 && ifr != 122)      // the closed-form laws (validated by osc2_set_model)
        f = friction_simple(ifr, re, md->ch_roughness[j], tp->ch_dh[j], md->ch_void_fraction[j], tp->ch_area[j], llim, md->ch_lam_coef[j]);
    else if (md->ch_ifriction[j] == 122) {
        const double rough = md->ch_roughness[j], dh = tp->ch_dh[j];
        f = haaland(fabs(re), rough, dh);
        for (int it = 0; it < cb_iters; ++it) f = colebrook_step(fabs(re), f, rough, dh);
        f = fmax(0.0, f);
    }
    return f * (ifr == 211 ? 1.0 : md->ch_friction_mult[j]);      // hts_cl_correlation_bundle sets FRICTION_MULTIPLIER = 1
}
// HTC["sol_sol"]["rad"] of interface k at one node: choice 3 = _inner_radiative_htc (conductor.py:5428-5443,
// symmetric in the two temperatures; the denominator is prepared on the host), -3 = constant
__device__ inline double radiative_htc(const osc2_model *md, int k, double Ta, double Tc)
{
    if (md->ss_choice[k] == -3) return md->ss_htc[k] * md->ss_mlt[k];
    const double den = md->ss_rad_den[k];
    if (!(den < 1.0e300)) return 0.0;          // an emissivity or the view factor is 0: the reference leaves zeros
    return 5.670374419e-08 * (sq(Ta) + sq(Tc)) * (Ta + Tc) / den * md->ss_rad_mlt[k];
}
// the materials outside the shipped decks (StackComponent layers, NbTi) sit behind two calls that RETURN their
// value (no reference arguments: those would pin cp and k of the hot path to local memory), so that the
// straight-line code of K_P, which is instruction-cache bound, does not grow with them
__device__ __noinline__ double material_cp_more(int mat, double T, double bg, double tcs, double tc)
{
    return mat == OSC2_MAT_AG ? cp_ag(T) : mat == OSC2_MAT_HC276 ? cp_hc276(T) : mat == OSC2_MAT_SN60PB40 ? cp_sn60pb40(T)
           : mat == OSC2_MAT_MGB2 ? cp_mgb2(T) : cp_nbti(T, bg, tcs, tc);
}
__device__ __noinline__ double material_k_more(int mat, double T, double bg)
{
    return mat == OSC2_MAT_AG ? k_ag(T) : mat == OSC2_MAT_HC276 ? k_hc276(T) : mat == OSC2_MAT_SN60PB40 ? k_sn60pb40(T)
           : mat == OSC2_MAT_MGB2 ? k_mgb2(T, bg) : k_nbti(T);
}
// cp and k of one constituent material at a Gauss point
__device__ inline void material_cp_k(int mat, double T, double bg, double tcs, double tc, double tmin, double tmax,
                                     double rrr, double tc0m, double p7, double rho273, double &cp, double &k)
{
    if (mat == OSC2_MAT_CU_NIST) { cp = cp_cu(T); k = k_cu(T, bg, rrr, p7, rho273); }
    else if (mat == OSC2_MAT_NB3SN) { cp = cp_nb3sn(T, tcs, tc, tc0m); k = k_nb3sn(T); }
    else if (mat == OSC2_MAT_SS) { cp = cp_ss(T, tmax); k = k_ss(T); }
    else if (mat == OSC2_MAT_AL) { cp = cp_al(T, tmin, tmax); k = k_al(T, tmin, tmax); }
    else if (mat == OSC2_MAT_RE123) { cp = cp_re123(T); k = k_re123(T); }
    else if (mat == OSC2_MAT_GE) { cp = cp_ge(T); k = k_ge(T); }
    else { cp = material_cp_more(mat, T, bg, tcs, tc); k = material_k_more(mat, T, bg); }
}
// channel.py:392-401, 782-795, 1104-1115
__device__ inline double friction_124(double re)
{
    re = fabs(re);
    double f = (0.316 * pw(re, -0.25)) / 4.0;
    if (re > 1e4) f = 2.21 * pw(re, -0.4) / 4.0;
    return fmax(0.0, f);
}
// channel.py:1156-1192
__device__ inline double nusselt_1(double re, double pr)
{
    double nu = 0.023 * pw(re, 0.8) * pw(pr, 0.3);
    if (nu < 8.235) nu = 8.235;
    return nu;
}
// channel.py:1245-1273: the exponent really used is eval_steady_state_htc's nn = 0.3
__device__ inline double nusselt_212(double re) { return 0.42 * pw(re, 0.3); }
// Flag_htc_steady_corr 119 / 120 / 121 (Dittus-Boelter with another lower limit, channel.py:1171-1243) and 211
// (correlation_211: masks on Reynolds over an array of zeros)
__device__ __noinline__ double nusselt_more(int corr, double re, double pr, double dh)
{
    if (corr == 211) {
        double nu = 0.0;
        if (re < 1000.0) nu = 5.0969 * pw(re, 0.10);
        if (re >= 1000.0 && re < 2000.0) nu = 10.2029 - 3.3242e-5 * re;
        if (re >= 2000.0) nu = 0.0395 * pw(re, 0.73);
        return nu;
    }
    double lim = 4.01;
    if (corr == 120) lim = 2.181;
    if (corr == 119) {
        const double a = dh / (dh + 2 * 1e-3), a2 = a * a, a3 = a2 * a, a4 = a2 * a2, a5 = a4 * a;
        const double NuTemp = 7.541 * (1.0 - 2.610 * a + 4.970 * a2 - 5.119 * a3 + 2.702 * a4 - 0.548 * a5);
        const double Nuflux = 8.235 * (1.0 - 10.6044 * a + 61.1755 * a2 - 155.1803 * a3 + 176.9203 * a4 - 72.9236 * a5);
        lim = (NuTemp + Nuflux) / 2.0;
    }
    double nu = 0.023 * pw(re, 0.8) * pw(pr, 0.3);
    if (nu < lim) nu = lim;
    return nu;
}


// bilinear (p, T) table lookup: cell index and weights once per point, then one
// 4-value stencil per property
struct Cell { int off; double a, b; int n_t; };
__device__ inline Cell locate(const FluidTableDev &tab, double p, double t)
{
    // two quotients per axis share their divisor: exact division through one refined reciprocal (osc2_div.cuh)
    const Osc2Rcp rps = osc2_rcp_prepare(tab.p_step), rts = osc2_rcp_prepare(tab.t_step);
    double fp = floor(osc2_div(p - tab.p_min, rps)), ft = floor(osc2_div(t - tab.t_min, rts));
    if (!(fp >= 0.0)) fp = 0.0;
    if (fp > (double)(tab.n_p - 2)) fp = (double)(tab.n_p - 2);
    if (!(ft >= 0.0)) ft = 0.0;
    if (ft > (double)(tab.n_t - 2)) ft = (double)(tab.n_t - 2);
    Cell c;
    c.off = (int)fp * tab.n_t + (int)ft;
    c.a = osc2_div(p - (tab.p_min + fp * tab.p_step), rps);
    c.b = osc2_div(t - (tab.t_min + ft * tab.t_step), rts);
    c.n_t = tab.n_t;
    return c;
}
__device__ inline double lookup(const FluidTableDev &tab, const Cell &c, int prop)
{
    const double *f = tab.props + (size_t)prop * tab.n_p * tab.n_t + c.off;
    const double f00 = __ldg(f), f01 = __ldg(f + 1), f10 = __ldg(f + c.n_t), f11 = __ldg(f + c.n_t + 1);
    const double lo = f00 + c.b * (f01 - f00);
    const double hi = f10 + c.b * (f11 - f10);
    return lo + c.a * (hi - lo);
}
// np.linspace(start, stop, N)[i]
__device__ inline double linspace_at(double start, double stop, int N, int i)
{
    if (i == N - 1 && N > 1) return stop;
    const double div = (double)(N - 1), delta = stop - start;
    const double step = delta / div;
    if (step == 0.0) return ((double)i / div) * delta + start;
    return (double)i * step + start;
}

// |linspace[i] + linspace[i + 1]| / 2 with the step's division done once
__device__ inline double linspace_mid(double start, double stop, int N, int i)
{
    const double div = (double)(N - 1), delta = stop - start;
    const double step = delta / div;
    double a, b;
    if (step == 0.0) { a = ((double)i / div) * delta + start; b = ((double)(i + 1) / div) * delta + start; }
    else { a = (double)i * step + start; b = (double)(i + 1) * step + start; }
    if (N > 1 && i == N - 1) a = stop;
    if (N > 1 && i + 1 == N - 1) b = stop;
    return fabs(a + b) / 2.0;
}

}  // namespace osc2p

__global__ void osc2_model_consts_kernel(const osc2_model *md, int S, double *sol_const)
{
    const int l = blockIdx.x * blockDim.x + threadIdx.x;
    if (l >= S) return;
    const double rrr = md->sol_rrr[l];
    const double beta = 0.634 / rrr, betat = beta / 0.0003;
    sol_const[2 * l] = 0.838 / osc2p::pw(betat, 0.1661);
    sol_const[2 * l + 1] = osc2p::rhoecu0(273.0, rrr);
}

// Array-wide quantities the element-wise fits branch on, one thread per (chunk of 64 elements, conductor):
//   max / min over the Gauss points of |T_i + T_i+1| / 2 for every solid (isobaric_specific_heat_ss,
//   the aluminium fits); temperatures are >= 0 so the IEEE bit pattern orders like an unsigned integer;
//   the iteration count of the Colebrook loop for channels with IFRICTION 122.
__global__ void osc2_solid_tmax_kernel(const osc2_topology *__restrict__ tp, StepDev p, PropsDev q)
{
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const int chunks = (p.E + 63) / 64;
    if (idx >= (long long)chunks * p.B) return;
    const size_t B = (size_t)p.B;
    const size_t b = (size_t)(idx % p.B);
    const int c = (int)(idx / p.B);
    const int e0 = c * 64, e1 = min(p.E, e0 + 64);
    const int M = p.M, d = p.d;
    for (int l = 0; l < p.S; ++l) {
        const int is = 3 * M + l;
        double prev = p.sysvar[((size_t)e0 * d + is) * B + b];
        double mx = 0.0, mn = __longlong_as_double(0x7ff0000000000000ll);
        // temperature margin of the superconducting strands: min (Tcs - T) over the Gauss points
        const bool want_margin = q.tcs && q.margin_key && q.model->sol_kind[l] == OSC2_SOLID_STRAND_MIXED;
        double mg = __longlong_as_double(0x7ff0000000000000ll);
        for (int e = e0; e < e1; ++e) {
            const double nxt = p.sysvar[((size_t)(e + 1) * d + is) * B + b];
            const double tg = fabs(prev + nxt) / 2.0;
            if (tg > mx) mx = tg;
            if (tg < mn) mn = tg;
            if (want_margin) { const double m = q.tcs[((size_t)l * p.E + e) * B + b] - tg; if (m < mg) mg = m; }
            prev = nxt;
        }
        atomicMax((unsigned long long *)(q.tmax + (size_t)l * B + b), (unsigned long long)__double_as_longlong(mx));
        atomicMin((unsigned long long *)(q.tmin + (size_t)l * B + b), (unsigned long long)__double_as_longlong(mn));
        if (want_margin) {
            const unsigned long long u = (unsigned long long)__double_as_longlong(mg);
            atomicMin(q.margin_key + (size_t)l * B + b, (u >> 63) ? ~u : (u | 0x8000000000000000ull));
        }
    }
    const osc2_model *md = q.model;
    for (int j = 0; j < M; ++j) {
        if (!osc2p::friction_iterates(md->ch_ifriction[j])) continue;
        const FluidTableDev &tab = q.fluid[md->ch_fluid[j]];
        int worst = 0;
        for (int e = e0; e < e1; ++e) {
            const double *x0 = p.sysvar + ((size_t)e * d) * B + b, *x1 = x0 + (size_t)d * B;
            const double v = (x0[(size_t)j * B] + x1[(size_t)j * B]) / 2.0;
            const double pr = (x0[(size_t)(M + j) * B] + x1[(size_t)(M + j) * B]) / 2.0;
            const double T = (x0[(size_t)(2 * M + j) * B] + x1[(size_t)(2 * M + j) * B]) / 2.0;
            const osc2p::Cell cl = osc2p::locate(tab, pr, T);
            const double re = fabs(tp->ch_dh[j] * v * osc2p::lookup(tab, cl, OSC2_FP_RHO) / osc2p::lookup(tab, cl, OSC2_FP_MU));
            const int it = osc2p::friction_iter_count(md, tp, j, re);
            if (it > worst) worst = it;
        }
        atomicMax(q.cb_iters + (size_t)j * B + b, worst);
    }
}

// ---------------------------------------------------------------------------
// K_P as three kernels:
//   osc2_opcond_coolant_kernel     thread per (channel, element, conductor)   tables, Re, Nu, friction
//   osc2_opcond_solid_kernel       thread per (solid, element, conductor)     Tc, material fits, heat source
//   osc2_opcond_interface_kernel   thread per (element, conductor)            radiative exchange, HTCs
// The steady-state coefficient Nu k / Dh and k rho cp of a channel travel to the interface kernel through
// `ch_scratch` [2][M][E][B].  (Round 1 had one kernel; the split by itself changed nothing -- 14.57 vs 14.59 ms --
// but the solid part now takes its own register budget: 80 registers / 6 CTAs per SM, 12.7 -> 11.7 ms.)
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(128, 5) osc2_opcond_coolant_kernel(const osc2_topology *__restrict__ tp, StepDev p, PropsDev q)
{
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long EB = (long long)p.E * p.B;
    if (idx >= EB * p.M) return;
    const int j = (int)(idx / EB);
    const long long rem = idx - (long long)j * EB;
    const int e = (int)(rem / p.B);
    const size_t b = (size_t)(rem % p.B);
    const size_t B = (size_t)p.B;
    const int M = p.M, d = p.d, E = p.E, N = p.N;
    const osc2_model *md = q.model;
    const double *x0 = p.sysvar + ((size_t)e * d) * B + b;
    const double *x1 = x0 + (size_t)d * B;
    double *fg = const_cast<double *>(p.fl_gauss);
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
    const double ftot = osc2p::friction_total(md, tp, j, re, q.cb_iters[(size_t)j * B + b], false);
#define FGW(qq) fg[((((size_t)(qq) * M + j) * E) + e) * B + b]
    FGW(0) = v; FGW(1) = pr; FGW(2) = T; FGW(3) = rho;
    FGW(4) = osc2p::lookup(tab, c, OSC2_FP_SOUND);
    FGW(5) = grun;
    FGW(6) = osc2p::lookup(tab, c, OSC2_FP_H);
    FGW(7) = cv;
    FGW(8) = ftot;
#undef FGW
    q.ch_scratch[(((size_t)0 * M + j) * E + e) * B + b] = nu * kk / tp->ch_dh[j];
    q.ch_scratch[(((size_t)1 * M + j) * E + e) * B + b] = kk * rho * osc2p::lookup(tab, c, OSC2_FP_CP);   // k rho cp of the transient HTC
    if (e == 0) {   // nodal values the boundary rows look at
        double *fn = const_cast<double *>(p.fl_node);
        const double *xl = p.sysvar + ((size_t)(N - 1) * d) * B + b;
        fn[((size_t)0 * M + j) * B + b] = x0[(size_t)j * B];
        fn[((size_t)1 * M + j) * B + b] = xl[(size_t)j * B];
        const osc2p::Cell cf = osc2p::locate(tab, x0[(size_t)(M + j) * B], x0[(size_t)(2 * M + j) * B]);
        const osc2p::Cell cl = osc2p::locate(tab, xl[(size_t)(M + j) * B], xl[(size_t)(2 * M + j) * B]);
        fn[((size_t)2 * M + j) * B + b] = osc2p::lookup(tab, cf, OSC2_FP_RHO);
        fn[((size_t)3 * M + j) * B + b] = osc2p::lookup(tab, cl, OSC2_FP_RHO);
    }
}

__global__ void __launch_bounds__(128, 6) osc2_opcond_solid_kernel(StepDev p, PropsDev q)
{
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long EB = (long long)p.E * p.B;
    if (idx >= EB * p.S) return;
    const int l = (int)(idx / EB);
    const long long rem = idx - (long long)l * EB;
    const int e = (int)(rem / p.B);
    const size_t b = (size_t)(rem % p.B);
    const size_t B = (size_t)p.B;
    const int M = p.M, S = p.S, d = p.d, E = p.E, N = p.N;
    const osc2_model *md = q.model;
    const double *x0 = p.sysvar + ((size_t)e * d) * B + b;
    const double *x1 = x0 + (size_t)d * B;
    double *sg = const_cast<double *>(p.sol_gauss);
    double *sqv = const_cast<double *>(p.sol_q);
    const int is = 3 * M + l;
    const double T = fabs(x0[(size_t)is * B] + x1[(size_t)is * B]) / 2.0;
    const double biss = q.b_field[((size_t)l * 2 + 0) * B + b], boss = q.b_field[((size_t)l * 2 + 1) * B + b];
    const double bg = osc2p::linspace_mid(biss, boss, N, e);
    const int nmat = md->sol_nmat[l];
    const int kind = md->sol_kind[l];
    double tc = 0.0, tcs = 0.0;
    if (kind == OSC2_SOLID_STRAND_MIXED) {
        const double eps = q.eps[(size_t)l * B + b];
        tc = (md->sol_mat[l][0] == OSC2_MAT_NBTI || md->sol_mat[l][nmat - 1] == OSC2_MAT_NBTI)
                 ? osc2p::tc_nbti(bg, md->sol_bc20m[l], md->sol_tc0m[l])
             : osc2p::tc_nb3sn(bg, (eps + eps) / 2.0, md->sol_tc0m[l], md->sol_bc20m[l]);
        tcs = q.tcs ? q.tcs[((size_t)l * E + e) * B + b] : 0.0;
    }
    const double tmax = q.tmax[(size_t)l * B + b], tmin = q.tmin[(size_t)l * B + b];
    double nsum = 0.0, cps = 0.0, ksum = 0.0, rho1 = 0.0, cp1 = 0.0, k1 = 0.0;
    for (int i = 0; i < nmat; ++i) {
        const int mat = md->sol_mat[l][i];
        const double w = md->sol_weight[l][i];
        const double rho_i = osc2p::mat_density(mat);
        double cp_i, k_i;
        osc2p::material_cp_k(mat, T, bg, tcs, tc, tmin, tmax, md->sol_rrr[l], md->sol_tc0m[l], q.sol_const[2 * l],
                             q.sol_const[2 * l + 1], cp_i, k_i);
        const double num = rho_i * w;
        if (i == 0) { nsum = num; cps = cp_i * num; ksum = k_i * w; rho1 = rho_i; cp1 = cp_i; k1 = k_i; }
        else { nsum = nsum + num; cps = cps + cp_i * num; ksum = ksum + k_i * w; }
    }
    double rho, cp, kk;
    if ((kind == OSC2_SOLID_JACKET && nmat == 1) || kind == OSC2_SOLID_STRAND_STAB) { rho = rho1; cp = cp1; kk = k1; }
    else { rho = nsum / md->sol_weight_sum[l]; cp = cps / nsum; kk = ksum / md->sol_weight_sum[l]; }
    sg[(((size_t)0 * S + l) * E + e) * B + b] = rho;
    sg[(((size_t)1 * S + l) * E + e) * B + b] = cp;
    sg[(((size_t)2 * S + l) * E + e) * B + b] = kk;
    // square-wave external heating
    const double n0 = q.q_window[((size_t)l * 4 + 0) * B + b], n1 = q.q_window[((size_t)l * 4 + 1) * B + b];
    const double tb = q.q_window[((size_t)l * 4 + 2) * B + b], te = q.q_window[((size_t)l * 4 + 3) * B + b];
    const double q0 = q.q0[(size_t)l * B + b];
    const bool on_now = md->sol_heated[l] && q.time >= tb && q.time <= te;
    const bool on_t0 = md->sol_heated[l] && 0.0 >= tb && 0.0 <= te;
    for (int side = 0; side < 2; ++side) {
        const int node = e + side;
        const bool in = node >= (int)n0 && node <= (int)n1;
        sqv[(((((size_t)side * S + l) * E + e) * 2) + 0) * B + b] = (on_now && in) ? q0 : 0.0;
        sqv[(((((size_t)side * S + l) * E + e) * 2) + 1) * B + b] = (on_t0 && in) ? q0 : 0.0;
    }
}

__global__ void __launch_bounds__(128) osc2_opcond_interface_kernel(const osc2_topology *__restrict__ tp, StepDev p, PropsDev q)
{
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (long long)p.E * p.B) return;
    const int e = (int)(idx / p.B);
    const size_t b = (size_t)(idx % p.B);
    const size_t B = (size_t)p.B;
    const int M = p.M, S = p.S, d = p.d, E = p.E, N = p.N;
    const osc2_model *md = q.model;
    const double *x0 = p.sysvar + ((size_t)e * d) * B + b;
    const double *x1 = x0 + (size_t)d * B;
    const double *sg = p.sol_gauss;
    double *sqv = const_cast<double *>(p.sol_q);
#define TF(j) ((x0[(size_t)(2 * M + (j)) * B] + x1[(size_t)(2 * M + (j)) * B]) / 2.0)
#define HST(j) q.ch_scratch[(((size_t)0 * M + (j)) * E + e) * B + b]
#define RCP(j) q.ch_scratch[(((size_t)1 * M + (j)) * E + e) * B + b]
    // radiative exchange between jackets (HTC_choice +-3): explicit, evaluated at the two NODES of the element
    // from the nodal temperatures and added to the heat sources of both solids, interface by interface
    // (jacket_component.py:174-232, conductor.py:5161-5228, 4558-4579)
    for (int k = 0; k < p.nss; ++k) {
        const int ch = md->ss_choice[k];
        if (ch != 3 && ch != -3) continue;
        const int a = tp->ss[k][0], c = tp->ss[k][1];
        for (int side = 0; side < 2; ++side) {
            const double *xn = side ? x1 : x0;
            const double Ta = xn[(size_t)(3 * M + a) * B], Tc = xn[(size_t)(3 * M + c) * B];
            const double h = osc2p::radiative_htc(md, k, Ta, Tc);
            const double ph = p.peri_ss_nodal[((size_t)k * N + (e + side)) * B + b] * h;
            sqv[(((((size_t)side * S + a) * E + e) * 2) + 0) * B + b] += ph * (Tc - Ta);
            sqv[(((((size_t)side * S + c) * E + e) * 2) + 0) * B + b] += ph * (Ta - Tc);
        }
    }
    double *hfs = const_cast<double *>(p.htc_fs);
    for (int k = 0; k < p.nfs; ++k) {
        const int j = tp->fs[k][0], l = tp->fs[k][1];
        double h;
        if (md->fs_choice[k] == 2) {
            const int is = 3 * M + l;
            const double Ts = fabs(x0[(size_t)is * B] + x1[(size_t)is * B]) / 2.0, Tf = TF(j);
            const double hK = 200.0 * (Ts + Tf) * (osc2p::sq(Ts) + osc2p::sq(Tf));
            const double tqbeg = q.q_window[((size_t)l * 4 + 2) * B + b];
            double full = 0.0;
            if (q.time > tqbeg) {
                const double htr = sqrt(RCP(j) / (3.141592653589793 * (q.time - tqbeg)));
                full = (hK * htr) / (hK + htr);
            }
            h = fmax(HST(j) * md->fs_mlt[k], full);
        } else {
            h = md->fs_htc[k] * md->fs_mlt[k];
        }
        hfs[((size_t)k * E + e) * B + b] = h;
    }
    double *hff = const_cast<double *>(p.htc_ff);
    for (int k = 0; k < p.nff; ++k) {
        const int ja = tp->ff[k][0], jb = tp->ff[k][1];
        double ho, hc;
        if (md->ff_choice[k] == 2) {
            const double h1 = HST(ja), h2 = HST(jb);
            ho = md->ff_mlt[k] * h1 * h2 / (h1 + h2);
            const double cond = osc2p::k_ss((TF(ja) + TF(jb)) / 2);
            const double rwall = md->ff_thick[k] / cond;
            hc = md->ff_mlt[k] / (1 / h1 + 1 / h2 + rwall);
        } else {
            ho = md->ff_htc[k] * md->ff_mlt[k];
            hc = ho;
        }
        hff[(((size_t)0 * p.nff + k) * E + e) * B + b] = ho;
        hff[(((size_t)1 * p.nff + k) * E + e) * B + b] = hc;
    }
    double *hss = const_cast<double *>(p.htc_ss);
    for (int k = 0; k < p.nss; ++k) {
        double h;
        if (md->ss_choice[k] == 3 || md->ss_choice[k] == -3) {
            h = 0.0;                   // HTC["sol_sol"]["cond"] keeps its zero initialisation (conductor.py:5109-5113)
        } else if (md->ss_choice[k] == 1) {
            const double ka = sg[(((size_t)2 * S + tp->ss[k][0]) * E + e) * B + b];
            const double kb = sg[(((size_t)2 * S + tp->ss[k][1]) * E + e) * B + b];
            const double rr = md->ss_thick_rc[k] / ka;
            const double rc = md->ss_thick_cr[k] / kb;
            h = md->ss_mlt[k] * (1.0 / (rr + md->ss_rcontact[k] + rc));
        } else {
            h = md->ss_htc[k] * md->ss_mlt[k];
        }
        hss[((size_t)k * E + e) * B + b] = h;
    }
    double *hes = const_cast<double *>(p.htc_es);
    for (int k = 0; k < p.nes; ++k) {
        hes[(((size_t)k * 3 + 0) * E + e) * B + b] = md->es_htc[k];
        hes[(((size_t)k * 3 + 1) * E + e) * B + b] = 0.0;
        hes[(((size_t)k * 3 + 2) * E + e) * B + b] = 0.0;
    }
#undef TF
#undef HST
#undef RCP
}

// ---------------------------------------------------------------------------
// Nodal coolant pass after a step (what the reference does for its outputs: FluidComponent
// _eval_properties_nodal_gauss(nodal = True) + _compute_density_and_mass_flow_rates, coolant.py:134-263;
// Channel.eval_friction_factor(nodal = True), channel.py:1119-1138).  Thread per (channel, node, conductor);
// `out` is [OSC2_N_NODAL][M][N][B]: total_density, the nine other table properties in the order of the alias
// table (simulation.py:89-100), Reynolds, Gruneisen, mass_flow_rate, friction_factor -- columns 4 and 6..18 of
// save_properties' channel file.  Pass 0 writes everything but the friction factor and, for IFRICTION 122,
// takes the array-wide Colebrook iteration count; pass 1 writes the friction factor.
// ---------------------------------------------------------------------------
#define OSC2_N_NODAL 14
__global__ void __launch_bounds__(128) osc2_nodal_coolant_kernel(const osc2_topology *__restrict__ tp, StepDev p, PropsDev q,
                                                                 double *__restrict__ out, int *__restrict__ cb_nodal, const int pass)
{
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long NB = (long long)p.N * p.B;
    if (idx >= NB * p.M) return;
    const int j = (int)(idx / NB);
    const long long rem = idx - (long long)j * NB;
    const int n = (int)(rem / p.B);
    const size_t b = (size_t)(rem % p.B);
    const size_t B = (size_t)p.B;
    const int M = p.M, d = p.d, N = p.N;
    const osc2_model *md = q.model;
    const size_t plane = (size_t)M * N * B, at = ((size_t)j * N + n) * B + b;
    if (pass == 1) {
        out[13 * plane + at] = osc2p::friction_total(md, tp, j, out[10 * plane + at], cb_nodal[(size_t)j * B + b], true);
        return;
    }
    const double *x = p.sysvar + ((size_t)n * d) * B + b;
    const FluidTableDev &tab = q.fluid[md->ch_fluid[j]];
    const double v = x[(size_t)j * B], pr = x[(size_t)(M + j) * B], T = x[(size_t)(2 * M + j) * B];
    const osc2p::Cell c = osc2p::locate(tab, pr, T);
    const double rho = osc2p::lookup(tab, c, OSC2_FP_RHO), mu = osc2p::lookup(tab, c, OSC2_FP_MU);
    const double beta = osc2p::lookup(tab, c, OSC2_FP_BETA), kappa = osc2p::lookup(tab, c, OSC2_FP_KAPPA);
    const double cv = osc2p::lookup(tab, c, OSC2_FP_CV);
    const double re = fabs(tp->ch_dh[j] * v * rho / mu);
    out[0 * plane + at] = rho;
    out[1 * plane + at] = beta;
    out[2 * plane + at] = kappa;
    out[3 * plane + at] = osc2p::lookup(tab, c, OSC2_FP_PRANDTL);
    out[4 * plane + at] = mu;
    out[5 * plane + at] = osc2p::lookup(tab, c, OSC2_FP_H);
    out[6 * plane + at] = osc2p::lookup(tab, c, OSC2_FP_CP);
    out[7 * plane + at] = cv;
    out[8 * plane + at] = osc2p::lookup(tab, c, OSC2_FP_SOUND);
    out[9 * plane + at] = osc2p::lookup(tab, c, OSC2_FP_K);
    out[10 * plane + at] = re;
    out[11 * plane + at] = beta / kappa / cv / rho;
    out[12 * plane + at] = tp->ch_area[j] * v * rho;
    if (osc2p::friction_iterates(md->ch_ifriction[j]))
        atomicMax(cb_nodal + (size_t)j * B + b, osc2p::friction_iter_count(md, tp, j, re));
}

