// Current sharing temperature and temperature margin on the device.
//
//   StrandComponent.eval_tcs                      strand_component.py:266-349  (jop = |I_op| / A_sc)
//   current_sharing_temperature_nb3sn             properties_of_materials/niobium3_tin.py:493-623
//   critical_current_density_nb3sn                niobium3_tin.py:265-368 (Bottura's ITER-2008 parametrisation)
//   BCNBSN / SNBSN / critical_temperature_nb3sn   niobium3_tin.py:7-59, 371-434, 437-490
//
// The reference inverts Jc(T, B, eps) = Jop with scipy.optimize.bisect(xtol = 1e-5) between 3.5 K and Tc(B, eps), one
// Python call per node.  The kernel below runs the same bisection -- same start interval, same midpoint update
// xm = xa + dm, same stopping test |dm| < xtol + rtol |xm| with scipy's default rtol = 4 eps, same sign rule
// (fm * fa >= 0 moves the left end) -- one thread per (point, conductor), so the iterates are the reference's
// unless a 1-ulp difference of pow() flips the sign of an f that is zero to rounding (then the answers still agree
// to xtol).  Run ONCE at step 0 (TCS_EVALUATION False, conductor.py:4615-4619); full-precision pow() is affordable.
#pragma once

#include "props_kernels.cuh"

namespace osc2p {

// niobium3_tin.py:7-59
__device__ inline double bc_nb3sn(double T, double s, double tc0m, double bc20m)
{
    const double tl = T / (tc0m * pow(s, 1.0 / 3.0));
    if (!(tl < 1.0)) return 0.0;
    return bc20m * s * (1.0 - pow(tl, 1.52));
}
// niobium3_tin.py:265-368; zero outside T/Tc0 < 1 and B/Bc2 < 1, negative values clipped to zero
__device__ inline double jc_nb3sn(double T, double B, double eps, double tc0m, double bc20m, double c0)
{
    const double ppp = 0.84, qqq = 2.57;
    const double blim = fmax(B, 0.01);
    const double s = snbsn(eps);
    const double tl = T / (tc0m * pow(s, 1.0 / 3.0));
    if (!(tl < 1.0)) return 0.0;
    const double bl = blim / bc_nb3sn(T, s, tc0m, bc20m);
    if (!(bl < 1.0)) return 0.0;
    const double jc = c0 / blim * s * (1.0 - pow(tl, 1.52)) * (1.0 - tl * tl) * pow(bl, ppp) * pow(1 - bl, qqq);
    return jc < 0.0 ? 0.0 : jc;
}
// full-precision twin of tc_nb3sn (the per-step kernel uses exp(y log x))
__device__ inline double tc_nb3sn_exact(double b, double eps, double tc0m, double bc20m)
{
    const double blim = fmax(b, 0.01);
    const double s = snbsn(eps);
    const double blcase = blim / (bc20m * s);
    if (!(blcase < 1.0)) return 0.0;
    return tc0m * pow(s, 1.0 / 3.0) * pow(1.0 - blcase, 1.0 / 1.52);
}
// niobium3_tin.py:493-623 for one point.  status: 0 ok, 1 = bisect's "f(a) and f(b) must have different signs"
__device__ inline double tcs_nb3sn(double B, double eps, double jop, double tc0m, double bc20m, double c0, int &bad)
{
    const double blim = fmax(B, 0.01);
    const double bstar = blim / bc20m / snbsn(eps);
    if (!(bstar < 1.0)) return 0.0;
    const double jc0 = jc_nb3sn(0.0, B, eps, tc0m, bc20m, c0);
    if (!(jop < jc0 && jop >= 0.0)) return 0.0;
    const double t_upper = tc_nb3sn_exact(B, eps, tc0m, bc20m);
    if (jop <= 0.0) return t_upper;          // by definition (niobium3_tin.py:577-596)
    // scipy/optimize/Zeros/bisect.c
    double xa = 3.5, xb = t_upper;
    const double fa = jc_nb3sn(xa, B, eps, tc0m, bc20m, c0) - jop;
    const double fb = jc_nb3sn(xb, B, eps, tc0m, bc20m, c0) - jop;
    if (fa == 0.0) return xa;
    if (fb == 0.0) return xb;
    if ((fa < 0.0) == (fb < 0.0)) { bad = 1; return 0.0; }
    const double xtol = 1e-5, rtol = 8.881784197001252e-16;
    double dm = xb - xa, xm = xa;
    for (int i = 0; i < 100; ++i) {
        dm *= .5;
        xm = xa + dm;
        const double fm = jc_nb3sn(xm, B, eps, tc0m, bc20m, c0) - jop;
        if (fm * fa >= 0) xa = xm;
        if (fm == 0 || fabs(dm) < xtol + rtol * fabs(xm)) return xm;
    }
    return xm;
}

// doubles ordered as unsigned integers (negative values included), for atomicMin
__device__ inline unsigned long long ordered_key(double x)
{
    const unsigned long long u = (unsigned long long)__double_as_longlong(x);
    return (u >> 63) ? ~u : (u | 0x8000000000000000ull);
}

}  // namespace osc2p

// points 0..N-1: nodes (B = linspace, solid_component.py:383-388); N..N+E-1: Gauss points
// (B = |B_i + B_i+1| / 2, eps = (eps + eps) / 2, conductor.py:4667-4730)
__global__ void osc2_tcs_kernel(StepDev p, PropsDev q, double *tcs_gauss, double *tcs_nodal)
{
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const int P = p.N + p.E;
    if (idx >= (long long)P * p.B) return;
    const size_t B = (size_t)p.B, b = (size_t)(idx % p.B);
    const int pt = (int)(idx / p.B);
    const osc2_model *md = q.model;
    for (int l = 0; l < p.S; ++l) {
        const int nmat = md->sol_nmat[l];
        const bool sc = md->sol_kind[l] == OSC2_SOLID_STRAND_MIXED && md->sol_mat[l][nmat - 1] == OSC2_MAT_NB3SN;
        double v = 0.0;
        if (sc) {
            const double biss = q.b_field[((size_t)l * 2 + 0) * B + b], boss = q.b_field[((size_t)l * 2 + 1) * B + b];
            const double eps0 = q.eps[(size_t)l * B + b];
            double bf, eps;
            if (pt < p.N) { bf = osc2p::linspace_at(biss, boss, p.N, pt); eps = eps0; }
            else {
                const int e = pt - p.N;
                bf = fabs(osc2p::linspace_at(biss, boss, p.N, e) + osc2p::linspace_at(biss, boss, p.N, e + 1)) / 2.0;
                eps = (eps0 + eps0) / 2.0;
            }
            int bad = 0;
            v = osc2p::tcs_nb3sn(bf, eps, q.jop[(size_t)l * B + b], md->sol_tc0m[l], md->sol_bc20m[l], md->sol_c0[l], bad);
            if (bad) atomicCAS(p.status + b, 0, OSC2_STATUS_TCS);
        }
        if (pt < p.N) tcs_nodal[((size_t)l * p.N + pt) * B + b] = v;
        else tcs_gauss[((size_t)l * p.E + (pt - p.N)) * B + b] = v;
    }
}
