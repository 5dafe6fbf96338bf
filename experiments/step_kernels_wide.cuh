// K_F with 16 lanes per conductor (Backward Euler, D a multiple of 8): every matrix row is split by
// columns over two lanes -- lane (r, h) holds columns [D/2 * h, D/2 * h + D/2) of the L, D and U
// blocks of row r -- so a warp carries 2 conductors instead of 4, the instruction stream per warp
// shrinks to ~60 %, the registers per thread to about half, and a batch of 4096 offers 2048 warps
// instead of 1024 (the single-warp sweep is issue/latency bound at 1.7 warps per scheduler,
// profiles/r1_final_ncu_full_summary.txt).
//
// Arithmetic is that of osc2_forward_fast_kernel entry by entry: each entry sees the same operations
// in the same order; the elimination multiplier of a (row, pivot) pair is computed redundantly, and
// identically, by both lanes of the row after one shuffle of the numerator.
#pragma once

#include "step_kernels_fast.cuh"

#ifdef OSC2_EMU
inline double osc2_shfl_xor(double v, int mask) { return emu_shfl(v, (int)(threadIdx.x % 32) ^ mask); }
#else
__device__ __forceinline__ double osc2_shfl_xor(double v, int mask) { return __shfl_xor_sync(0xffffffffu, v, mask); }
#endif

template <int D>
struct WideFwdSmem {
    static constexpr int PW = 2 * D + 2;                  // published pivot row: D' | U | rhs | 1/pivot
    static constexpr int PUB = 2 * D * PW;                // two nodes (double buffer)
    static constexpr int PER_COND = PUB + ((18 - (PUB % 16)) % 16);   // == 2 (mod 16): bank spread
};

template <int D, bool CAPTURE>
__global__ void __maxnreg__(128) osc2_forward_wide_kernel(const osc2_topology *__restrict__ tp, StepDev p)
{
    static_assert(D % 8 == 0, "Backward Euler closed-form Known needs D % 8 == 0");
    OSC2_DYN_SMEM(sm_dyn);
    using SM = WideFwdSmem<D>;
    constexpr int PW = SM::PW, RW = D + 9, H = D / 2, LPC = 2 * D, cpb = 32 / LPC;
    const int M = p.M, N = p.N;
    const size_t B = (size_t)p.B;
    const int grp = threadIdx.x / LPC, l16 = threadIdx.x % LPC;
    const int h = l16 / D, r = l16 % D;                   // column half, matrix row
    const int c0 = h * H;                                 // first column of this lane
    const long long b_ll = (long long)blockIdx.x * cpb + grp;
    const bool act = b_ll < (long long)p.B;
    const size_t b = act ? (size_t)b_ll : 0;
    const long long Total = (long long)D * N;
    const Osc2Rcp rdt = osc2_rcp_prepare(p.dt);
    const bool first = (p.num_step == 1);
    const bool shift = (p.num_step > 1);
    int range_flag = 0;
    osc2_rcp_check(rdt, range_flag);
    double *pub = sm_dyn + (size_t)grp * SM::PER_COND;    // [2][D][PW]

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

    // this lane's half of the row record of one element
    struct Elem {
        double sd[H];
        double kb, md, mo, ahd, aho, lodu, lodl, lodup, lodlp;
    };
    const size_t gstep = (size_t)p.NS * B;
    const double *gptr = p.gm + (size_t)r * RW * B + b;
    auto load_elem = [&](const double *g, Elem &el) {
#pragma unroll
        for (int cc = 0; cc < H; ++cc) el.sd[cc] = g[(size_t)(c0 + cc) * B];
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
        for (int cc = 0; cc < H; ++cc) el.sd[cc] = 0.0;
        el.md = el.mo = el.kb = el.ahd = el.aho = el.lodu = el.lodl = el.lodup = el.lodlp = 0.0;
    };
    Elem eL, eR, eN;
    zero_elem(eL); zero_elem(eR); zero_elem(eN);
    const size_t xstep = (size_t)D * B;
    const double *xptr = p.sysvar + (size_t)r * B + b;
    double *lodptr = p.syslod + (size_t)r * 2 * B + b;
    double *spptr = p.spill + (size_t)r * (2 * D + 2) * B + b;
    double x_prev = 0.0, x_cur = 0.0, x_nxt = 0.0, x_pre = 0.0, sl0_next = 0.0, sl1_next = 0.0;
    if (act) {
        load_elem(gptr, eR);
        x_cur = xptr[0];
        x_nxt = xptr[xstep];
        sl0_next = lodptr[0];
        sl1_next = lodptr[B];
    }

    auto node = [&](auto HL, auto HR, const int n) {
        constexpr bool hasL = decltype(HL)::value, hasR = decltype(HR)::value;
        double *pub_cur = pub + (size_t)(n & 1) * D * PW;
        const double *pub_prev = pub + (size_t)((n + 1) & 1) * D * PW;
        const long long i_row = (long long)n * D + r;
        double sl0 = sl0_next, sl1 = sl1_next;
        if (act) {
            if (n + 1 < N - 1) load_elem(gptr + gstep, eN);
            if (n + 2 < N) x_pre = xptr[2 * xstep];
            if (n + 1 < N) { sl0_next = lodptr[2 * xstep]; sl1_next = lodptr[2 * xstep + B]; }
        }
        double rowL[H], rowD[H], rowU[H], rhs = 0.0;
        {
            // SYSLOD (smc:1165-1318): both lanes of a row compute it, lane h = 0 stores it
            if (shift) { sl1 = sl0; sl0 = 0.0; }
            if (hasL) { sl0 += eL.lodl; if (first) sl1 += eL.lodlp; }
            if (hasR) { sl0 += eR.lodu; if (first) sl1 += eR.lodup; }
            if (act && h == 0) { lodptr[0] = sl0; lodptr[B] = sl1; }

            const bool is_bc = (!hasL && bc_first) || (!hasR && bc_last);
            double mx = 0.0, known = 0.0;
            if (!is_bc) {
            const double masL = hasL ? osc2_div_flag(eL.mo, rdt, range_flag) : 0.0;
            const double masD = osc2_div_flag((hasL ? eL.md : 0.0) + (hasR ? eR.md : 0.0), rdt, range_flag);
            const double masU = hasR ? osc2_div_flag(eR.mo, rdt, range_flag) : 0.0;
#pragma unroll
            for (int cc = 0; cc < H; ++cc) {
                const int c = c0 + cc;
                const bool dg = (c == r);
                const double avL = dg ? eL.ahd : (c == aoffcol ? eL.aho : 0.0);
                const double avR = dg ? eR.ahd : (c == aoffcol ? eR.aho : 0.0);
                {
                    double sys = 0.0;
                    if (hasL) {
                        const double fds = (-avL + (dg ? -eL.kb : 0.0)) + eL.sd[cc] / 2.;
                        const double m = dg ? masL : 0.0;
                        sys = m + fds;
                    }
                    rowL[cc] = sys;
                    { const double a_ = fabs(sys); if (a_ > mx) mx = a_; }
                }
                {
                    double flx = 0.0, dif = 0.0, sor = 0.0;
                    if (hasL) { if (dg) dif += eL.kb; flx += avL; sor += eL.sd[cc]; }
                    if (hasR) { if (dg) dif += eR.kb; flx += -avR; sor += eR.sd[cc]; }
                    const double fds = (flx + dif) + sor;
                    const double m = dg ? masD : 0.0;
                    const double sys = m + fds;
                    rowD[cc] = sys;
                    { const double a_ = fabs(sys); if (a_ > mx) mx = a_; }
                }
                {
                    double sys = 0.0;
                    if (hasR) {
                        const double fds = (avR + (dg ? -eR.kb : 0.0)) + eR.sd[cc] / 2.;
                        const double m = dg ? masU : 0.0;
                        sys = m + fds;
                    }
                    rowU[cc] = sys;
                    { const double a_ = fabs(sys); if (a_ > mx) mx = a_; }
                }
            }
            // Known (smc:1366-1443), Backward Euler closed form: only MASMAT/dt * x_old survives
            if (hasL) known = masL * x_prev;
            known = hasL ? known + masD * x_cur : masD * x_cur;
            if (hasR) known = known + masU * x_nxt;
            known += sl0;
            }
            // max |row| over both column halves (tsf:538-551); both lanes of a row take the same branch
            { const double o = osc2_shfl_xor(mx, D); if (o > mx) mx = o; }
            if (is_bc) {
#pragma unroll
                for (int cc = 0; cc < H; ++cc) { rowL[cc] = 0.0; rowD[cc] = (c0 + cc == r) ? 1.0 : 0.0; rowU[cc] = 0.0; }
                rhs = (!hasL) ? val_first : val_last;
            } else {
                const Osc2Rcp rmx = osc2_rcp_prepare(mx);
                osc2_rcp_check(rmx, range_flag);
#pragma unroll
                for (int cc = 0; cc < H; ++cc) {
                    rowL[cc] = osc2_div_flag(rowL[cc], rmx, range_flag);
                    rowD[cc] = osc2_div_flag(rowD[cc], rmx, range_flag);
                    rowU[cc] = osc2_div_flag(rowU[cc], rmx, range_flag);
                }
                rhs = osc2_div_flag(known, rmx, range_flag);
            }
            if (CAPTURE && act) {
#pragma unroll
                for (int cc = 0; cc < H; ++cc) {
                    const int c = c0 + cc;
                    const long long cols[3] = {(long long)(n - 1) * D + c, (long long)n * D + c, (long long)(n + 1) * D + c};
                    const double vals[3] = {rowL[cc], rowD[cc], rowU[cc]};
                    for (int q = 0; q < 3; ++q) {
                        if (cols[q] < 0 || cols[q] >= Total) continue;
                        const long long kb_ = (2 * D - 1) + (cols[q] - i_row);
                        p.cap_sysmat[((size_t)kb_ * Total + i_row) * B + b] = vals[q];
                    }
                }
                if (h == 0) p.cap_known[(size_t)i_row * B + b] = rhs;
            }
        }
        // pivots of node n-1 (gredub / gbacsb forward part, tsf:602-761)
        if (hasL) {
#pragma unroll
            for (int k = 0; k < D; ++k) {
                const double own = rowL[k % H];
                const double oth = osc2_shfl_xor(own, D);
                const double num = ((k / H) == h) ? own : oth;
                const double *pk = pub_prev + k * PW;
                const double q = osc2_div_flag(num, osc2_rcp_make(pk[k], pk[2 * D + 1]), range_flag);
#pragma unroll
                for (int cc = 0; cc < H; ++cc)
                    if (c0 + cc > k) rowL[cc] = rowL[cc] - pk[c0 + cc] * q;
#pragma unroll
                for (int cc = 0; cc < H; ++cc) rowD[cc] = rowD[cc] - pk[D + c0 + cc] * q;
                rhs = rhs - pk[2 * D] * q;
            }
        }
        // pivots inside node n
        Osc2Rcp my_rcp = osc2_rcp_make(1.0, 1.0);
#pragma unroll
        for (int k = 0; k < D; ++k) {
            const double own = rowD[k % H];
            const double oth = osc2_shfl_xor(own, D);
            if (r == k) {
                double *pk = pub_cur + k * PW;
#pragma unroll
                for (int cc = 0; cc < H; ++cc) { pk[c0 + cc] = rowD[cc]; pk[D + c0 + cc] = rowU[cc]; }
                if (h == 0) pk[2 * D] = rhs;
                if (h == k / H) {            // the lane that holds the pivot itself
                    my_rcp = osc2_rcp_prepare(own);
                    pk[2 * D + 1] = my_rcp.y;
                    if (fabs(own) <= OSC2_TINY_PIVOT) { if (bad_row < 0) bad_row = i_row; }
                    else osc2_rcp_check(my_rcp, range_flag);
                }
            }
            __syncwarp();
            if (r > k) {
                const double num = ((k / H) == h) ? own : oth;
                const double *pk = pub_cur + k * PW;
                const double q = osc2_div_flag(num, osc2_rcp_make(pk[k], pk[2 * D + 1]), range_flag);
#pragma unroll
                for (int cc = 0; cc < H; ++cc)
                    if (c0 + cc > k) rowD[cc] = rowD[cc] - pk[c0 + cc] * q;
#pragma unroll
                for (int cc = 0; cc < H; ++cc) rowU[cc] = rowU[cc] - pk[D + c0 + cc] * q;
                rhs = rhs - pk[2 * D] * q;
            }
        }
        if (act) {
#pragma unroll
            for (int cc = 0; cc < H; ++cc) {
                spptr[(size_t)(c0 + cc) * B] = rowD[cc];
                spptr[(size_t)(D + c0 + cc) * B] = rowU[cc];
            }
            if (h == 0) spptr[(size_t)(2 * D) * B] = rhs;
            if (h == r / H) spptr[(size_t)(2 * D + 1) * B] = my_rcp.y;
        }
        eL = eR;
        eR = eN;
        x_prev = x_cur; x_cur = x_nxt; x_nxt = x_pre;
        gptr += gstep;
        xptr += xstep;
        lodptr += 2 * xstep;
        spptr += (size_t)D * (2 * D + 2) * B;
        __syncwarp();
    };
    node(std::false_type(), std::true_type(), 0);
    for (int n = 1; n < N - 1; ++n) node(std::true_type(), std::true_type(), n);
    node(std::true_type(), std::false_type(), N - 1);
    __syncwarp();
    double *scratch = sm_dyn + (size_t)grp * SM::PER_COND;
    scratch[l16] = (double)bad_row;
    scratch[LPC + l16] = (double)range_flag;
    __syncwarp();
    if (act && l16 == 0) {
        long long worst = -1;
        bool range = false;
        for (int q = 0; q < LPC; ++q) {
            const long long v = (long long)scratch[q];
            if (v >= 0 && (worst < 0 || v < worst)) worst = v;
            range = range || (scratch[LPC + q] != 0.0);
        }
        p.status[b] = (worst >= 0) ? -(int)(worst + 1) : (range ? OSC2_STATUS_RANGE : 0);
    }
}
