// K_F64 / K_B64: the two sweeps of step() for systems with NODOFS = 64 (CASE_2: 13 channels,
// 25 solids), register tiled.
//
// Reference being reproduced, operation for operation (same operand order, no FMA contraction):
//   SYSMAT / Known   step_matrix_construction.py:1320-1474
//   boundary rows    fluid_component.py:117-565
//   row scaling      transient_solution_functions.py:538-551
//   gredub           transient_solution_functions.py:602-698
//   gbacsb           transient_solution_functions.py:701-808
//
// The generic sweep (step_kernels.cuh) keeps the block row [L D U | rhs] of a node in shared
// memory, one thread per matrix row: 64 threads per conductor and two shared-memory loads + one
// store per multiply-subtract.  Here one CTA of 256 threads owns a conductor: matrix row r lives in
// the REGISTERS of four neighbouring lanes (3 x 16 entries each, columns dealt in pairs so that a
// pivot row is read with 128-bit loads), the elimination multiplier travels by warp shuffle, and only
// the pivot rows go through shared memory (one row of 130 doubles per pivot, read as a broadcast).
// A d = 64 node costs 2.3 d^3 = 6.1e5 multiply-subtract pairs, i.e. the path is bound by the FP64
// pipe and by the serial chain of 128 pivots per node, not by bytes (SURVEY.md 8d flags CASE_2).
//
// Row assembly uses the sparsity of the Gauss-point matrices: an entry (r, c) that is neither the
// diagonal, nor the AMAT partner column, nor a position the topology's interface statements can
// write (the same enumeration as osc2_build_smap) has SYSMAT value +0 and Known coefficient +0 in
// the reference, so only its product 0 * x_old enters the (order-preserving) pairwise sum.
//
// Factor rows are spilled per conductor as [node][row][130] = D' (64) | U (64) | rhs | 1/pivot
// refinement, contiguous, so that the forward sweep writes and the backward sweep reads full lines.
#pragma once

#include "step_kernels.cuh"
#include "step_kernels_fast.cuh"   // osc2_bcast
#include "osc2_div.cuh"

#define OSC2_D64_SW 200   /* strip row stride (doubles): even, 64-byte shift between neighbouring rows */
#define OSC2_D64_PW 130   /* pivot / factor row: D'(64) U(64) rhs rcp */

struct alignas(16) Osc2Pair { double x, y; };

// shared memory of K_F64 (doubles): strip | pivot rows | old solution at three nodes | touched SMAT
// entries of two elements, eight per thread | row scalars of two elements, ten per row | SYSLOD of two nodes
#define OSC2_D64_TS 8     /* prefetched touched entries per thread; further ones are read from global memory */
#define OSC2_D64_RS 10    /* dz, K, Msol, A diag, A partner, SV0, SV1, SVP0, SVP1, pad */
__host__ __device__ inline size_t osc2_d64_fwd_smem_doubles()
{
    return (size_t)64 * OSC2_D64_SW + 64 * OSC2_D64_PW + 3 * 64 + 2 * 256 * OSC2_D64_TS + 2 * 64 * OSC2_D64_RS + 2 * 64 * 2 + 64;
}

// 8-byte asynchronous copy global -> shared (LDGSTS): the element record of node n+1 travels while
// node n is eliminated, no registers in flight
#ifdef OSC2_EMU
inline void osc2_cp8(double *dst, const double *src) { *dst = *src; }
inline void osc2_cp_wait_all() {}
#else
__device__ __forceinline__ void osc2_cp8(double *dst, const double *src)
{
    const unsigned sa = (unsigned)__cvta_generic_to_shared(dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(sa), "l"(src) : "memory");
}
__device__ __forceinline__ void osc2_cp_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }
#endif

__device__ __forceinline__ Osc2Pair osc2_ld2(const double *q) { return *reinterpret_cast<const Osc2Pair *>(q); }
__device__ __forceinline__ void osc2_st2(double *q, double a, double b)
{
    Osc2Pair v; v.x = a; v.y = b;
    *reinterpret_cast<Osc2Pair *>(q) = v;
}

// numpy's pairwise sum of n <= 255 values splits into at most three leaves of <= 128 values:
// sum = leaf0 [+ leaf1 | + (leaf1 + leaf2)]   (loops_utils.h.src, PW_BLOCKSIZE 128)
__device__ inline int osc2_leaf_split(int n, int *start, int *len)
{
    if (n <= 128) { start[0] = 0; len[0] = n; return 1; }
    int n2 = n / 2;
    n2 -= n2 % 8;
    start[0] = 0; len[0] = n2;
    const int m = n - n2;
    if (m <= 128) { start[1] = n2; len[1] = m; return 2; }
    int m2 = m / 2;
    m2 -= m2 % 8;
    start[1] = n2; len[1] = m2;
    start[2] = n2 + m2; len[2] = m - m2;
    return 3;
}

__device__ __forceinline__ int osc2_ctz(unsigned m)
{
#ifdef OSC2_EMU
    return __builtin_ctz(m);
#else
    return __ffs((int)m) - 1;
#endif
}

// publish / wait on a shared-memory flag (in-node pivot rounds without a CTA barrier)
#ifdef OSC2_EMU
#include <thread>
inline void osc2_fence_block() { __atomic_thread_fence(__ATOMIC_SEQ_CST); }
inline double osc2_vload(const double *q) { double v; __atomic_load((const long long *)q, (long long *)&v, __ATOMIC_ACQUIRE); return v; }
inline void osc2_vstore(double *q, double v) { __atomic_store((long long *)q, (long long *)&v, __ATOMIC_RELEASE); }
inline void osc2_pause() { std::this_thread::yield(); }
#else
__device__ __forceinline__ void osc2_fence_block() { __threadfence_block(); }
__device__ __forceinline__ double osc2_vload(const double *q) { return *(const volatile double *)q; }
__device__ __forceinline__ void osc2_vstore(double *q, double v) { *(volatile double *)q = v; }
__device__ __forceinline__ void osc2_pause() {}
#endif

#ifdef OSC2_EMU
inline bool osc2_any(bool pred) { return emu_any(pred); }
#else
__device__ __forceinline__ bool osc2_any(bool pred) { return __any_sync(0xffffffffu, pred) != 0; }
#endif

// numpy pairwise sum (n <= 128: one block, loops_utils.h.src) of window positions j0 .. j0+n-1; the
// window position j is strip[cbase + (j - j0)] when that index lies in [0, 192) and +0 otherwise
// (band positions outside the three blocks).
__device__ __forceinline__ double osc2_leaf_sum(const double *strip, int cbase, int n)
{
    auto at = [&](int i) { const unsigned c = (unsigned)(cbase + i); return c < 192u ? strip[c] : 0.0; };
    if (n < 8) {
        double res = 0.0;
        for (int i = 0; i < n; ++i) res += at(i);
        return res;
    }
    double rr[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) rr[k] = at(k);
    int i;
    for (i = 8; i < n - (n % 8); i += 8) {
#pragma unroll
        for (int k = 0; k < 8; ++k) rr[k] += at(i + k);
    }
    double res = ((rr[0] + rr[1]) + (rr[2] + rr[3])) + ((rr[4] + rr[5]) + (rr[6] + rr[7]));
    for (; i < n; ++i) res += at(i);
    return res;
}

// ---------------------------------------------------------------------------
// K_F64
// ---------------------------------------------------------------------------
template <bool CAP, bool FLAGS>
__global__ void __launch_bounds__(256, 1) osc2_forward64_kernel(const osc2_topology *__restrict__ tp, StepDev p)
{
    OSC2_DYN_SMEM(sm_dyn);
    constexpr int d = 64, SW = OSC2_D64_SW, PW = OSC2_D64_PW;
    const int M = p.M, N = p.N;
    const size_t B = (size_t)p.B;
    const size_t b = (size_t)blockIdx.x;
    const int tid = (int)threadIdx.x;
    const int r = tid >> 2, h = tid & 3, lane = tid & 31, lane0 = lane & ~3, warp = tid >> 5;
    const GmSlots sl{d, M, p.S};
    const long long Total = (long long)d * N;
    constexpr int Half = 2 * d, Main = 2 * d - 1;
    const double dt = p.dt, theta = tp->theta;
    const bool first = (p.num_step == 1);

    double *strip = sm_dyn;                 // [64][SW]  products / unscaled row entries of the node
    double *piv = strip + 64 * SW;          // [64][PW]  final rows of node n-1, then of node n
    double *xs = piv + 64 * PW;             // [3][64]   old solution at nodes n-1, n, n+1
    double *sS = xs + 3 * 64;               // [2][256][TS]  touched SMAT entries of elements e, e+1 (by parity)
    double *sR = sS + 2 * 256 * OSC2_D64_TS;   // [2][64][RS] row scalars of those elements
    double *sSL = sR + 2 * 64 * OSC2_D64_RS;   // [2][64][2]  SYSLOD of the node (by parity)
    double *rdy = sSL + 2 * 64 * 2;            // [64] FLAGS: node count up to which pivot row k has been published
    double *srow = strip + r * SW;
    double *fac_out = p.spill + b * ((size_t)N * 64 * PW);

    // ---- row kind, AMAT pattern, boundary rows (as the generic sweep) ------
    int kind = 3, j = 0, aoffcol = -1, aoffq = 0;
    if (r < M) { kind = 0; j = r; aoffcol = M + j; aoffq = 1; }
    else if (r < 2 * M) { kind = 1; j = r - M; aoffcol = j; aoffq = 2; }
    else if (r < 3 * M) { kind = 2; j = r - 2 * M; aoffcol = j; aoffq = 3; }
    else { kind = 3; j = r - 3 * M; }
    bool bc_first = false, bc_last = false;
    double val_first = 0.0, val_last = 0.0;
    if (kind < 3) {
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

    // ---- columns of row r the Gauss-point matrices can touch ----------------
    unsigned long long rowbits = 1ull << r;
    if (aoffcol >= 0) rowbits |= 1ull << aoffcol;
    {
        auto add = [&](int rr, int cc) { if (rr == r) rowbits |= 1ull << cc; };
        for (int q = 0; q < M; ++q) { add(q, q); add(M + q, q); add(2 * M + q, q); }
        for (int k = 0; k < p.nff; ++k)
            for (int side = 0; side < 2; ++side) {
                const int c1 = tp->ff[k][side], c2 = tp->ff[k][1 - side];
                const int v1 = c1, p1 = M + c1, t1 = 2 * M + c1, p2 = M + c2, t2 = 2 * M + c2;
                add(v1, p1); add(v1, p2); add(p1, p1); add(p1, p2); add(p1, t1); add(p1, t2);
                add(t1, p1); add(t1, p2); add(t1, t1); add(t1, t2);
            }
        for (int k = 0; k < p.nfs; ++k) {
            const int jj = tp->fs[k][0], ip = M + jj, it = 2 * M + jj, is = 3 * M + tp->fs[k][1];
            add(ip, it); add(ip, is); add(it, it); add(it, is); add(is, is); add(is, it);
        }
        for (int k = 0; k < p.nss; ++k) {
            const int a = 3 * M + tp->ss[k][0], c = 3 * M + tp->ss[k][1];
            add(a, a); add(a, c); add(c, c); add(c, a);
        }
        for (int k = 0; k < p.nes; ++k) add(3 * M + tp->es[k], 3 * M + tp->es[k]);
    }
    // bit i of `mask`: local column i (pair i/2, element i%2) = column 8 (i/2) + 2 h + i%2 is touched
    unsigned mask = 0, diagbit = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) {
        const int c = ((i >> 1) << 3) + 2 * h + (i & 1);
        if ((rowbits >> c) & 1ull) mask |= 1u << i;
        if (c == r) diagbit = 1u << i;
    }

    const Osc2Rcp third = osc2_rcp_prepare(3.0);     // x / 3. with the shared-divisor division (exact, osc2_div.cuh)
    long long bad_row = -1;
    bool hung = false;
    if (tid < 64) {
        rdy[tid] = 0.0;
        xs[0 * d + tid] = p.sysvar[((size_t)0 * d + tid) * B + b];
        xs[1 * d + tid] = p.sysvar[((size_t)1 * d + tid) * B + b];
    }

    double Lr[16], Dr[16], Ur[16];

    // element e (between nodes e and e+1) and SYSLOD of node e
    auto prefetch = [&](int e) {
        if (e < N - 1) {
            const double *g = p.gm + ((size_t)e * p.NS) * B + b;
            double *ds = sS + ((e & 1) * 256 + tid) * OSC2_D64_TS;
            int t = 0;
            for (unsigned m = mask; m && t < OSC2_D64_TS; m &= m - 1, ++t) {
                const int i = osc2_ctz(m);
                const int c = ((i >> 1) << 3) + 2 * h + (i & 1);
                osc2_cp8(ds + t, g + (size_t)sl.s(r, c) * B);
            }
            if (h == 0) {
                double *dr = sR + ((e & 1) * 64 + r) * OSC2_D64_RS;
                osc2_cp8(dr + 0, p.dz + (size_t)e * B + b);
                osc2_cp8(dr + 1, g + (size_t)sl.k(r) * B);
                if (kind == 3) {
                    osc2_cp8(dr + 2, g + (size_t)sl.msol(j) * B);
                    osc2_cp8(dr + 5, g + (size_t)sl.sv(j, 0) * B);
                    osc2_cp8(dr + 6, g + (size_t)sl.sv(j, 1) * B);
                    if (first) {
                        osc2_cp8(dr + 7, g + (size_t)sl.svp(j, 0) * B);
                        osc2_cp8(dr + 8, g + (size_t)sl.svp(j, 1) * B);
                    }
                } else {
                    osc2_cp8(dr + 3, g + (size_t)sl.a(j, 0) * B);
                    osc2_cp8(dr + 4, g + (size_t)sl.a(j, aoffq) * B);
                }
            }
        }
        if (e < N && h == 0) {
            const size_t i_row = (size_t)e * d + r;
            osc2_cp8(sSL + ((e & 1) * 64 + r) * 2 + 0, p.syslod + (i_row * 2 + 0) * B + b);
            osc2_cp8(sSL + ((e & 1) * 64 + r) * 2 + 1, p.syslod + (i_row * 2 + 1) * B + b);
        }
    };
    prefetch(0);

    for (int n = 0; n < N; ++n) {
        osc2_cp_wait_all();
        __syncthreads();

        const bool hasL = n > 0, hasR = n < N - 1;
        const long long i_row = (long long)n * d + r;
        const double *gL = p.gm + ((size_t)(hasL ? n - 1 : 0) * p.NS) * B + b;   // entries beyond the prefetched eight
        const double *gR = p.gm + ((size_t)(hasR ? n : 0) * p.NS) * B + b;
#define GL(slot) gL[(size_t)(slot) * B]
#define GR(slot) gR[(size_t)(slot) * B]
        const double *tL = sS + (((n + 1) & 1) * 256 + tid) * OSC2_D64_TS, *tR = sS + ((n & 1) * 256 + tid) * OSC2_D64_TS;
        const double *rL = sR + (((n + 1) & 1) * 64 + r) * OSC2_D64_RS, *rR = sR + ((n & 1) * 64 + r) * OSC2_D64_RS;
        double dzL = 1.0, dzR = 1.0, mL = 0.0, mR = 0.0, kbL = 0.0, kbR = 0.0;
        double adL = 0.0, aoL = 0.0, adR = 0.0, aoR = 0.0;
        if (hasL) {
            dzL = rL[0];
            mL = (kind < 3) ? 1.0 : rL[2];
            kbL = rL[1] / dzL;
            if (kind < 3) { adL = rL[3]; aoL = rL[4]; }
        }
        if (hasR) {
            dzR = rR[0];
            mR = (kind < 3) ? 1.0 : rR[2];
            kbR = rR[1] / dzR;
            if (kind < 3) { adR = rR[3]; aoR = rR[4]; }
        }
        const double alpha = 0.0;   // ALFA (tsf:265): consistent mass
        const double cdL = dzL * (1. / 3. + alpha), coL = dzL * (1. / 6. - alpha);
        const double cdR = dzR * (1. / 3. + alpha), coR = dzR * (1. / 6. - alpha);

        // SYSLOD (smc:1165-1318): lane h = 0 of the row owns it
        double slterm = 0.0;
        if (h == 0) {
            double sl0 = sSL[((n & 1) * 64 + r) * 2 + 0];
            double sl1 = sSL[((n & 1) * 64 + r) * 2 + 1];
            if (p.num_step > 1) { sl1 = sl0; sl0 = 0.0; }
            if (kind == 3) {
                if (hasL) {
                    const double dz_6 = dzL / 6.;
                    sl0 += (rL[5] + 2. * rL[6]) * dz_6;
                    if (first) sl1 += (rL[7] + 2. * rL[8]) * dz_6;
                }
                if (hasR) {
                    const double dz_6 = dzR / 6.;
                    sl0 += (2. * rR[5] + rR[6]) * dz_6;
                    if (first) sl1 += (2. * rR[7] + rR[8]) * dz_6;
                }
            }
            p.syslod[((size_t)i_row * 2 + 0) * B + b] = sl0;
            p.syslod[((size_t)i_row * 2 + 1) * B + b] = sl1;
            slterm = theta * sl0 + (1.0 - theta) * sl1;
        }

        // one band entry: SYSMAT value and the Known coefficient (smc:1343-1346, 1432-1441)
        auto entry = [&](int part, int c, int t, double &sys, double &ck) {
            const double sLv = hasL ? (t < OSC2_D64_TS ? tL[t] : GL(sl.s(r, c))) : 0.0;
            const double sRv = hasR ? (t < OSC2_D64_TS ? tR[t] : GR(sl.s(r, c))) : 0.0;
            const bool dg = (c == r);
            double mas, flx, dif, sor;
            if (part == 0) {
                const double aval = dg ? adL : (c == aoffcol ? aoL : 0.0);
                mas = dg ? coL * mL : 0.0;
                flx = -(aval / 2.);
                dif = dg ? -kbL : 0.0;
                sor = osc2_div(sLv * dzL, third) / 2.;
            } else if (part == 2) {
                const double aval = dg ? adR : (c == aoffcol ? aoR : 0.0);
                mas = dg ? coR * mR : 0.0;
                flx = aval / 2.;
                dif = dg ? -kbR : 0.0;
                sor = osc2_div(sRv * dzR, third) / 2.;
            } else {
                mas = 0.0; flx = 0.0; dif = 0.0; sor = 0.0;
                if (hasL) {
                    const double aval = dg ? adL : (c == aoffcol ? aoL : 0.0);
                    if (dg) { mas += cdL * mL; dif += kbL; }
                    flx += aval / 2.;
                    sor += osc2_div(sLv * dzL, third);
                }
                if (hasR) {
                    const double aval = dg ? adR : (c == aoffcol ? aoR : 0.0);
                    if (dg) { mas += cdR * mR; dif += kbR; }
                    flx += -(aval / 2.);
                    sor += osc2_div(sRv * dzR, third);
                }
            }
            const double fds = (flx + dif) + sor;
            // off the diagonal the mass term is +0: 0 / dt is +0 for a positive time step (nvcc's own
            // division sends a zero numerator to its slow subroutine)
            const double mdt = (dg || !(dt > 0.0)) ? mas / dt : 0.0;
            sys = mdt + theta * fds;
            ck = mdt - (1.0 - theta) * fds;
        };

        const bool is_bc = (n == 0 && bc_first) || (n == N - 1 && bc_last);
        double rhs, zero = 0.0;
        unsigned pm0 = 0, pm1 = 0, pm2 = 0;    // which of my 16 columns of L / D / U come from the strip
        // Shuffles and warp barriers stay outside the boundary-row branch: the eight rows of a warp
        // may take different sides of it.
        // ---- Known (smc:1366-1443): products into the strip, numpy pairwise sum ----
        if (!is_bc) {
#pragma unroll
            for (int part = 0; part < 3; ++part) {
                const bool present = (part == 0) ? hasL : (part == 2 ? hasR : true);
                const double *xo = xs + ((n - 1 + part + 3) % 3) * d;
#pragma unroll
                for (int pp = 0; pp < 8; ++pp) {
                    const int c = 8 * pp + 2 * h;
                    // untouched entry: Known coefficient +0 (see the header)
                    osc2_st2(srow + part * d + c, present ? 0.0 * xo[c] : 0.0, present ? 0.0 * xo[c + 1] : 0.0);
                }
            }
            int t = 0;
            for (unsigned m = mask; m; m &= m - 1, ++t) {
                const int i = osc2_ctz(m);
                const int c = ((i >> 1) << 3) + 2 * h + (i & 1);
                for (int part = 0; part < 3; ++part) {
                    const bool present = (part == 0) ? hasL : (part == 2 ? hasR : true);
                    if (!present) continue;
                    const double *xo = xs + ((n - 1 + part + 3) % 3) * d;
                    double sys, ck;
                    entry(part, c, t, sys, ck);
                    srow[part * d + c] = ck * xo[c];
                }
            }
        }
        __syncwarp();
        double part_sum = 0.0;
        int nleaf = 1;
        if (!is_bc) {
            long long lo, hi;
            if (i_row <= Half - 1) { lo = 0; hi = Half + i_row; }
            else if (i_row >= Total - (Half - 1)) { lo = i_row - (Half - 1); hi = Total; }
            else { lo = i_row - (Half - 1); hi = i_row + Half; }
            const int Lfull = (int)(hi - lo);
            const long long off = (long long)(n - 1) * d - lo;
            int lstart[3], llen[3];
            nleaf = osc2_leaf_split(Lfull, lstart, llen);
            if (h < nleaf) part_sum = osc2_leaf_sum(srow, lstart[h] - (int)off, llen[h]);
        }
        const double s0 = osc2_bcast(part_sum, lane0), s1 = osc2_bcast(part_sum, lane0 + 1), s2 = osc2_bcast(part_sum, lane0 + 2);
        double known = (nleaf == 1) ? s0 : (nleaf == 2 ? s0 + s1 : s0 + (s1 + s2));
        known += osc2_bcast(slterm, lane0);
        __syncwarp();

        // ---- SYSMAT entries, row scaling (tsf:538-551) ----
        double mx = 0.0;
        if (!is_bc) {
            int t = 0;
            for (unsigned m = mask; m; m &= m - 1, ++t) {
                const int i = osc2_ctz(m);
                const int c = ((i >> 1) << 3) + 2 * h + (i & 1);
                for (int part = 0; part < 3; ++part) {
                    const bool present = (part == 0) ? hasL : (part == 2 ? hasR : true);
                    if (!present) continue;
                    double sys, ck;
                    entry(part, c, t, sys, ck);
                    srow[part * d + c] = sys;
                    const double a = fabs(sys);
                    if (a > mx) mx = a;
                }
            }
        }
        { const double o = osc2_bcast(mx, lane ^ 1); if (o > mx) mx = o; }
        { const double o = osc2_bcast(mx, lane ^ 2); if (o > mx) mx = o; }
        if (!is_bc) {
            const Osc2Rcp rmx = osc2_rcp_prepare(mx);
            for (unsigned m = mask; m; m &= m - 1) {
                const int i = osc2_ctz(m);
                const int c = ((i >> 1) << 3) + 2 * h + (i & 1);
                for (int part = 0; part < 3; ++part) {
                    const bool present = (part == 0) ? hasL : (part == 2 ? hasR : true);
                    if (!present) continue;
                    srow[part * d + c] = osc2_div(srow[part * d + c], rmx);
                }
            }
            zero = (mx > 0.0) ? 0.0 : 0.0 / mx;       // every other entry of the row: 0 / max|row|
            rhs = osc2_div(known, rmx);
            pm0 = hasL ? mask : 0u; pm1 = mask; pm2 = hasR ? mask : 0u;
        } else {
            // boundary row (fc:117-565): identity row, Known = imposed value
            if (diagbit) srow[d + r] = 1.0;
            pm1 = diagbit;
            rhs = (n == 0) ? val_first : val_last;
        }
#undef GL
#undef GR
#pragma unroll
        for (int pp = 0; pp < 8; ++pp) {
            const int c = 8 * pp + 2 * h;
            const Osc2Pair l2 = osc2_ld2(srow + c), d2 = osc2_ld2(srow + d + c), u2 = osc2_ld2(srow + 2 * d + c);
            Lr[2 * pp] = ((pm0 >> (2 * pp)) & 1u) ? l2.x : zero;
            Lr[2 * pp + 1] = ((pm0 >> (2 * pp + 1)) & 1u) ? l2.y : zero;
            Dr[2 * pp] = ((pm1 >> (2 * pp)) & 1u) ? d2.x : zero;
            Dr[2 * pp + 1] = ((pm1 >> (2 * pp + 1)) & 1u) ? d2.y : zero;
            Ur[2 * pp] = ((pm2 >> (2 * pp)) & 1u) ? u2.x : zero;
            Ur[2 * pp + 1] = ((pm2 >> (2 * pp + 1)) & 1u) ? u2.y : zero;
        }
        __syncwarp();          // the four lanes of a row are done with the row scalars of element n-1
        prefetch(n + 1);
        if (CAP) {
            if (p.cap_sysmat) {
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                    const int c = ((i >> 1) << 3) + 2 * h + (i & 1);
                    const double vals[3] = {Lr[i], Dr[i], Ur[i]};
#pragma unroll
                    for (int part = 0; part < 3; ++part) {
                        const long long col = (long long)(n - 1 + part) * d + c;
                        if (col < 0 || col >= Total) continue;
                        const long long kb = Main + (col - i_row);
                        p.cap_sysmat[((size_t)kb * Total + i_row) * B + b] = vals[part];
                    }
                }
                if (h == 0) p.cap_known[(size_t)i_row * B + b] = rhs;
            }
        }

        // ---- gredub + forward part of gbacsb against the final rows of node n-1 ----
        if (hasL) {
#pragma unroll
            for (int k8 = 0; k8 < 8; ++k8) {
#pragma unroll 1
                for (int kk = 0; kk < 8; ++kk) {
                    const int k = 8 * k8 + kk;
                    const double *pk = piv + k * PW;
                    const double num = (kk & 1) ? Lr[2 * k8 + 1] : Lr[2 * k8];
                    const double v = osc2_bcast(num, lane0 | (kk >> 1));
                    // a multiplier of zero leaves the row as it is (up to the sign of an exact zero): skipped
                    // when that holds for all eight rows of the warp -- the blocks of CASE_2 are sparse
                    if (!osc2_any(v != 0.0)) continue;
                    const double q = osc2_div(v, osc2_rcp_make(pk[k], pk[2 * d + 1]));
                    {
                        const Osc2Pair x = osc2_ld2(pk + 8 * k8 + 2 * h);
                        if (2 * h > kk) Lr[2 * k8] = Lr[2 * k8] - x.x * q;
                        if (2 * h + 1 > kk) Lr[2 * k8 + 1] = Lr[2 * k8 + 1] - x.y * q;
                    }
#pragma unroll
                    for (int pp = k8 + 1; pp < 8; ++pp) {
                        const Osc2Pair x = osc2_ld2(pk + 8 * pp + 2 * h);
                        Lr[2 * pp] = Lr[2 * pp] - x.x * q;
                        Lr[2 * pp + 1] = Lr[2 * pp + 1] - x.y * q;
                    }
#pragma unroll
                    for (int pp = 0; pp < 8; ++pp) {
                        const Osc2Pair x = osc2_ld2(pk + d + 8 * pp + 2 * h);
                        Dr[2 * pp] = Dr[2 * pp] - x.x * q;
                        Dr[2 * pp + 1] = Dr[2 * pp + 1] - x.y * q;
                    }
                    rhs = rhs - pk[2 * d] * q;
                }
            }
        }
        __syncthreads();      // every row is done with the rows of node n-1: `piv` may be rewritten
        // ... and with the old solution of node n-1: its slot takes the one of node n+2
        if (tid < 64 && n + 2 < N) osc2_cp8(xs + ((n + 2) % 3) * d + tid, p.sysvar + ((size_t)(n + 2) * d + tid) * B + b);

        // ---- pivots inside node n: row k is final once it has seen rows < k ----
#pragma unroll
        for (int k8 = 0; k8 < 8; ++k8) {
#pragma unroll 1
            for (int kk = 0; kk < 8; ++kk) {
                const int k = 8 * k8 + kk;
                double *pk = piv + k * PW;
                if (r == k) {
#pragma unroll
                    for (int pp = 0; pp < 8; ++pp) {
                        osc2_st2(pk + 8 * pp + 2 * h, Dr[2 * pp], Dr[2 * pp + 1]);
                        osc2_st2(pk + d + 8 * pp + 2 * h, Ur[2 * pp], Ur[2 * pp + 1]);
                    }
                    if (h == (kk >> 1)) {
                        const double pv = (kk & 1) ? Dr[2 * k8 + 1] : Dr[2 * k8];
                        pk[2 * d + 1] = osc2_rcp_prepare(pv).y;
                        if (fabs(pv) <= OSC2_TINY_PIVOT && bad_row < 0) bad_row = (long long)n * d + k;
                    }
                    if (h == 0) pk[2 * d] = rhs;
                }
                if (FLAGS) {
                    // no CTA barrier per round: the owner warp raises the row's flag, the warps below it
                    // wait for that flag only, so a warp that is late does not hold up the pivot chain
                    const double epoch = (double)(n + 1);
                    if (warp == k8) {
                        __syncwarp();
                        if (lane == 0) { osc2_fence_block(); osc2_vstore(rdy + k, epoch); }
                    } else if (warp > k8) {
                        if (!hung) {
                            int it = 0;
                            while (osc2_vload(rdy + k) < epoch) {
                                if (++it > (1 << 22)) { hung = true; break; }
                                osc2_pause();
                            }
                        }
                        osc2_fence_block();
                    }
                } else {
                    __syncthreads();
                }
                if (8 * warp + 7 > k) {
                    const double num = (kk & 1) ? Dr[2 * k8 + 1] : Dr[2 * k8];
                    const double v = osc2_bcast(num, lane0 | (kk >> 1));
                    if (osc2_any(r > k && v != 0.0) && r > k) {
                        const double q = osc2_div(v, osc2_rcp_make(pk[k], pk[2 * d + 1]));
                        {
                            const Osc2Pair x = osc2_ld2(pk + 8 * k8 + 2 * h);
                            if (2 * h > kk) Dr[2 * k8] = Dr[2 * k8] - x.x * q;
                            if (2 * h + 1 > kk) Dr[2 * k8 + 1] = Dr[2 * k8 + 1] - x.y * q;
                        }
#pragma unroll
                        for (int pp = k8 + 1; pp < 8; ++pp) {
                            const Osc2Pair x = osc2_ld2(pk + 8 * pp + 2 * h);
                            Dr[2 * pp] = Dr[2 * pp] - x.x * q;
                            Dr[2 * pp + 1] = Dr[2 * pp + 1] - x.y * q;
                        }
#pragma unroll
                        for (int pp = 0; pp < 8; ++pp) {
                            const Osc2Pair x = osc2_ld2(pk + d + 8 * pp + 2 * h);
                            Ur[2 * pp] = Ur[2 * pp] - x.x * q;
                            Ur[2 * pp + 1] = Ur[2 * pp + 1] - x.y * q;
                        }
                        rhs = rhs - pk[2 * d] * q;
                    }
                }
            }
        }
        // all 64 final rows are in `piv` (the last publish was followed by a barrier): spill them as they lie
        if (FLAGS) __syncthreads();
        {
            double *out = fac_out + (size_t)n * (64 * PW);
            for (int q = tid; q < 64 * PW / 2; q += 256) {
                const Osc2Pair x = osc2_ld2(piv + 2 * q);
                osc2_st2(out + 2 * q, x.x, x.y);
            }
        }
    }
    // first singular row of the conductor (gredub/gbacsb raise on abs(pivot) <= 1e-20)
    __syncthreads();
    if (hung) bad_row = 0;          // a flag never came (cannot happen; guards against a hang): conductor reported failed
    strip[tid] = (double)bad_row;
    __syncthreads();
    if (tid == 0) {
        long long worst = -1;
        for (int q = 0; q < 256; ++q) {
            const long long v = (long long)strip[q];
            if (v >= 0 && (worst < 0 || v < worst)) worst = v;
        }
        if (worst >= 0 && p.status[b] == 0) p.status[b] = -(int)(worst + 1);   // sticky: the first failing step is kept
    }
}

// ---------------------------------------------------------------------------
// K_B64: back substitution (gbacsb, tsf:772-802).  The subtractions of one row must follow the
// reference's column order (in-node columns r+1.., then the 64 columns of node n+1) and start with
// the unknown solved last, so a node is one serial chain of sum_r (127 - r) = 6112 dependent
// subtractions whatever the thread mapping.  One warp per conductor: the lanes form the products
// of a row in parallel (two in-node columns, two next-node columns each), hand them over through
// shared memory and every lane walks the chain, so the solved unknown needs no broadcast.  Seven
// conductors per SM are resident at a 1024-conductor batch; the factor rows stream from HBM one row
// ahead of the chain.
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(32) osc2_backward64_kernel(StepDev p)
{
    constexpr int d = 64, PW = OSC2_D64_PW;
    OSC2_DYN_SMEM(sm_dyn);
    double *prod = sm_dyn;     // [2][128]
    const int N = p.N, lane = (int)threadIdx.x;
    const size_t B = (size_t)p.B, b = (size_t)blockIdx.x;
    const double *fac = p.spill + b * ((size_t)N * 64 * PW);
    double xn0 = 0.0, xn1 = 0.0;
    for (int n = N - 1; n >= 0; --n) {
        const double *nd = fac + (size_t)n * (64 * PW);
        const bool hasR = n < N - 1;
        double x0 = 0.0, x1 = 0.0;
        // row 63 in flight
        const double *row = nd + 63 * PW;
        double dlo = row[lane], dhi = row[lane + 32], ulo = row[d + lane], uhi = row[d + 32 + lane];
        double rhs = row[2 * d], ry = row[2 * d + 1], pv = row[63];
        for (int r = d - 1; r >= 0; --r) {
            double *pr = prod + (r & 1) * 128;
            pr[lane] = dlo * x0;
            pr[lane + 32] = dhi * x1;
            pr[d + lane] = ulo * xn0;
            pr[d + 32 + lane] = uhi * xn1;
            double q = rhs;
            const Osc2Rcp rc = osc2_rcp_make(pv, ry);
            if (r > 0) {      // next row's factors
                row = nd + (r - 1) * PW;
                dlo = row[lane]; dhi = row[lane + 32]; ulo = row[d + lane]; uhi = row[d + 32 + lane];
                rhs = row[2 * d]; ry = row[2 * d + 1]; pv = row[r - 1];
            }
            __syncwarp();
#pragma unroll 8
            for (int c = r + 1; c < d; ++c) q = q - pr[c];
            if (hasR) {
#pragma unroll 8
                for (int c = 0; c < d; ++c) q = q - pr[d + c];
            }
            const double x = osc2_div(q, rc);
            if (lane == (r & 31)) { if (r < 32) x0 = x; else x1 = x; }
        }
        p.sysvar_new[((size_t)n * d + lane) * B + b] = x0;
        p.sysvar_new[((size_t)n * d + 32 + lane) * B + b] = x1;
        xn0 = x0; xn1 = x1;
    }
}
