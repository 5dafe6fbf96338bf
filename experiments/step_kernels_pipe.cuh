// K_F, software-pipelined over nodes (Backward Euler, D a multiple of 8: closed-form Known).
//
// osc2_forward_fast_kernel walks one node at a time: build the row of node n, eliminate it with
// the 8 published pivot rows of node n-1 ("L phase"), then run the 8 pivot steps inside node n ("D
// phase") -- 16 strictly serial rounds of  load pivot row -> exact division -> 17 products, each
// round mostly latency (tools/fp64_latency_probe.cu).  But round k of the L phase of node n+1 needs
// nothing but pivot row k of node n, which is final the moment round k of node n's D phase
// publishes it.  This kernel therefore builds the row of node n+1 FIRST and then runs both phases
// in the same 8 rounds: one shared-memory exchange and one pivot-row load per round serve two
// independent update streams (rows of node n below the pivot, rows of node n+1), so the serial
// rounds per node drop from 16 to 8 and each round carries twice the independent arithmetic.
// Every matrix entry still sees exactly the operations of the reference, in the reference's order.
//
// Registers: the U block of node n+1 and the element record that the next construction needs are
// parked in a per-lane shared-memory slot while the rounds run.
#pragma once

#include "step_kernels_fast.cuh"

template <int D>
struct PipeFwdSmem {
    static constexpr int PW = 2 * D + 2;                  // published pivot row: D' | U | rhs | 1/pivot
    static constexpr int PUB = D * PW;                    // the pivot rows of ONE node (row k is rewritten 8 rounds later)
    static constexpr int PK_FIELDS = 2 * D + 9;           // parked: U block of node n+1 + one element record
    static constexpr int PK = PK_FIELDS + ((6 - (PK_FIELDS % 4)) % 4);   // == 2 (mod 4): conflict-free 16-byte accesses
    static constexpr int PF_FIELDS = D + 12;              // asynchronous prefetch: element record, x, SYSLOD row
    static constexpr int PF = PF_FIELDS + ((6 - (PF_FIELDS % 4)) % 4);
    static constexpr int RAW = D * PK + D * PF + PUB;
    static constexpr int PER_COND = RAW + ((18 - (RAW % 16)) % 16);      // == 2 (mod 16): bank spread
};

template <int D, int LPC, bool FULL>
__global__ void osc2_forward_pipe_kernel(const osc2_topology *__restrict__ tp, StepDev p)
{
    static_assert(D % 8 == 0, "closed-form Known: D % 8 == 0");
    static_assert(!FULL || LPC == D, "FULL: every lane owns a row");
    OSC2_DYN_SMEM(sm_dyn);
    using SM = PipeFwdSmem<D>;
    constexpr int PW = SM::PW, RW = D + 9;
    const int M = p.M, N = p.N;
    const size_t B = (size_t)p.B;
    const unsigned Bu = (unsigned)p.B;
    auto fld = [Bu](auto *ptr, const int k) {
        return (decltype(ptr))((const char *)ptr + (unsigned long long)Bu * (unsigned)(8 * k));
    };
    const int cpb = blockDim.x / LPC;
    const int grp = threadIdx.x / LPC, r = threadIdx.x % LPC;
    const long long b_ll = (long long)blockIdx.x * cpb + grp;
    const bool act = FULL || ((b_ll < (long long)p.B) && (r < D));
    const size_t b = act ? (size_t)b_ll : 0;
    const int rr_ = (FULL || r < D) ? r : 0;
    const Osc2Rcp rdt = osc2_rcp_prepare(p.dt);
    const bool first = (p.num_step == 1);
    const bool shift = (p.num_step > 1);
    unsigned div_span = osc2_abs_hi(p.dt) - OSC2_HI_DIV_LO;     // divisor range: dt, then every pivot

    double *base = sm_dyn + (size_t)grp * SM::PER_COND;
    double *park = base + (size_t)rr_ * SM::PK;           // this lane's parking slot
    double *pre = base + D * SM::PK + (size_t)rr_ * SM::PF;      // this lane's prefetch slot
    double *pub = base + D * SM::PK + D * SM::PF;         // [D][PW]

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

    struct Elem {
        double sd[D];
        double sdr, sdo, kb, md, mo, ahd, aho, lodu, lodl;
    };
    const size_t gstep = (size_t)p.NS * B;                       // one element further
    const double *gptr = p.gm + (size_t)rr_ * RW * B + b;        // record of the element right of the node under construction
    const size_t off_r = (size_t)rr_ * B, off_o = (size_t)(aoffcol >= 0 ? aoffcol : 0) * B;
    const size_t xstep = (size_t)D * B;
    const double *xptr = p.sysvar + (size_t)rr_ * B + b;                 // x[node under construction][r]
    double *lodptr = p.syslod + (size_t)rr_ * 2 * B + b;                 // SYSLOD[node under construction][r][0]
    double *spptr = p.spill + (size_t)rr_ * (2 * D + 2) * B + b;         // spill[node under elimination][r][0]
    // What the construction of node m reads, copied global -> shared without passing through
    // registers (they are all taken during the rounds): element m, x[m+1], SYSLOD[m]
    enum { PF_SDR = D, PF_SDO, PF_KB, PF_MD, PF_MO, PF_AHD, PF_AHO, PF_LODU, PF_LODL, PF_X, PF_SL0, PF_SL1 };
    static_assert(PF_SL1 + 1 == SM::PF_FIELDS, "prefetch slot layout");
    auto prefetch = [&](const int m, const double *g, const double *xq, const double *lq) {
        if (m < N - 1) {
#pragma unroll
            for (int c = 0; c < D; ++c) osc2_async_copy8(pre + c, fld(g, c));
            osc2_async_copy8(pre + PF_SDR, g + off_r);
            osc2_async_copy8(pre + PF_SDO, g + off_o);
#pragma unroll
            for (int q = 0; q < 7; ++q) osc2_async_copy8(pre + PF_KB + q, fld(g, D + q));
            osc2_async_copy8(pre + PF_X, xq);
        }
        if (m < N) {
            osc2_async_copy8(pre + PF_SL0, lq);
            osc2_async_copy8(pre + PF_SL1, fld(lq, 1));
        }
    };
    auto zero_elem = [&](Elem &el) {
#pragma unroll
        for (int c = 0; c < D; ++c) el.sd[c] = 0.0;
        el.sdr = el.sdo = 0.0;
        el.md = el.mo = el.kb = el.ahd = el.aho = el.lodu = el.lodl = 0.0;
    };
    Elem eA, eB;                                                         // left / right element of the node under construction
    zero_elem(eA); zero_elem(eB);
    double xm = 0.0, x0 = 0.0, xp = 0.0;                                 // own old solution at nodes m-1, m, m+1
    double mas_carry = 0.0;                                              // MASMAT/dt of the element to the left

    // ---- row of node m, scaled (the arithmetic of osc2_forward_fast_kernel, closed-form Known) ----
    auto construct = [&](auto HL, auto HR, const int m, Elem &A, Elem &Bv, double (&rowL)[D], double (&rowD)[D],
                         double (&rowU)[D], double &rhs) {
        constexpr bool hasL = decltype(HL)::value, hasR = decltype(HR)::value;
        rhs = 0.0;
        if (act) {
            // prefetched by the previous stage (collected by its wait): element m, x[m+1], SYSLOD row m
            if (hasR) {
#pragma unroll
                for (int c = 0; c < D; ++c) Bv.sd[c] = pre[c];
                Bv.sdr = pre[PF_SDR]; Bv.sdo = pre[PF_SDO]; Bv.kb = pre[PF_KB]; Bv.md = pre[PF_MD];
                Bv.mo = pre[PF_MO]; Bv.ahd = pre[PF_AHD]; Bv.aho = pre[PF_AHO];
                Bv.lodu = pre[PF_LODU]; Bv.lodl = pre[PF_LODL];
                xp = pre[PF_X];
            }
            double sl0 = pre[PF_SL0], sl1 = pre[PF_SL1];
            if (hasR) prefetch(m + 1, gptr + gstep, xptr + 2 * xstep, lodptr + 2 * xstep);
            // SYSLOD (smc:1165-1318); the previous-step halves exist at the first step only
            if (shift) { sl1 = sl0; sl0 = 0.0; }
            if (hasL) { sl0 += A.lodl; if (first) sl1 += *fld(gptr - gstep, D + 8); }
            if (hasR) { sl0 += Bv.lodu; if (first) sl1 += *fld(gptr, D + 7); }
            lodptr[0] = sl0;
            *fld(lodptr, 1) = sl1;

            const double masL = hasL ? mas_carry : 0.0;
            const double masU = hasR ? osc2_div_p(Bv.mo, rdt) : 0.0;
            mas_carry = masU;
            const bool is_bc = (!hasL && bc_first) || (!hasR && bc_last);
            if (is_bc) {
#pragma unroll
                for (int c = 0; c < D; ++c) { rowL[c] = 0.0; rowD[c] = (c == r) ? 1.0 : 0.0; rowU[c] = 0.0; }
                rhs = (!hasL) ? val_first : val_last;
            } else {
                const double masD = osc2_div_p((hasL ? A.md : 0.0) + (hasR ? Bv.md : 0.0), rdt);
                double fLd = 0.0, fLo = 0.0, fUd = 0.0, fUo = 0.0, fDd, fDo;
                if (hasL) { fLd = (-A.ahd + -A.kb) + A.sdr / 2.; fLo = -A.aho + A.sdo / 2.; }
                if (hasR) { fUd = (Bv.ahd + -Bv.kb) + Bv.sdr / 2.; fUo = Bv.aho + Bv.sdo / 2.; }
                if (hasL && hasR) {
                    fDd = ((A.ahd + -Bv.ahd) + (A.kb + Bv.kb)) + (A.sdr + Bv.sdr);
                    fDo = (A.aho + -Bv.aho) + (A.sdo + Bv.sdo);
                } else if (hasL) {
                    fDd = (A.ahd + A.kb) + A.sdr;
                    fDo = A.aho + A.sdo;
                } else {
                    fDd = (-Bv.ahd + Bv.kb) + Bv.sdr;
                    fDo = -Bv.aho + Bv.sdo;
                }
                double mxL = 0.0, mxD = 0.0, mxU = 0.0;
                const double sLd = masL + fLd, sDd = masD + fDd, sUd = masU + fUd;
#pragma unroll
                for (int c = 0; c < D; ++c) {
                    const bool dg = (c == r), of = (c == aoffcol);
                    const double gD = (hasL && hasR) ? A.sd[c] + Bv.sd[c] : (hasL ? A.sd[c] : Bv.sd[c]);
                    rowL[c] = hasL ? (dg ? sLd : (of ? fLo : A.sd[c] / 2.)) : 0.0;
                    rowD[c] = dg ? sDd : (of ? fDo : gD);
                    rowU[c] = hasR ? (dg ? sUd : (of ? fUo : Bv.sd[c] / 2.)) : 0.0;
                    if (hasL) mxL = osc2_absmax(mxL, rowL[c]);
                    mxD = osc2_absmax(mxD, rowD[c]);
                    if (hasR) mxU = osc2_absmax(mxU, rowU[c]);
                }
                const double mx = osc2_absmax(osc2_absmax(mxL, mxD), mxU);
                // Known (smc:1366-1443), closed form: only MASMAT/dt * x_old survives
                double known = 0.0;
                if (hasL) known = masL * xm;
                known = hasL ? known + masD * x0 : masD * x0;
                if (hasR) known = known + masU * xp;
                known += sl0;
                // row scaling (tsf:538-551), one range check for the whole row (osc2_div.cuh)
                Osc2Rcp rmx = osc2_rcp_prepare(mx);
                unsigned kmin = osc2_num_key(known);
#pragma unroll
                for (int c = 0; c < D; ++c) {
                    if (hasL) kmin = osc2_umin(kmin, osc2_num_key(rowL[c]));
                    kmin = osc2_umin(kmin, osc2_num_key(rowD[c]));
                    if (hasR) kmin = osc2_umin(kmin, osc2_num_key(rowU[c]));
                }
                osc2_rcp_poison_if(rmx, (kmin < OSC2_KEY_MIN) || (osc2_abs_hi(known) >= OSC2_HI_2P900) ||
                                            (osc2_abs_hi(mx) - OSC2_HI_DIV_LO >= OSC2_HI_DIV_SPAN));
#pragma unroll
                for (int c = 0; c < D; ++c) {
                    if (hasL) rowL[c] = osc2_div_fast(rowL[c], rmx);
                    rowD[c] = osc2_div_fast(rowD[c], rmx);
                    if (hasR) rowU[c] = osc2_div_fast(rowU[c], rmx);
                }
                rhs = osc2_div_fast(known, rmx);
            }
            xm = x0; x0 = xp;       // rotate the old solution
        } else {
#pragma unroll
            for (int c = 0; c < D; ++c) { rowL[c] = 0.0; rowD[c] = 0.0; rowU[c] = 0.0; }
        }
        gptr += gstep;
        xptr += xstep;
        lodptr += 2 * xstep;
    };

    // ---- the 8 rounds of node n: its own pivots (rows r > k) and, with WN, the pivots of node n
    //      applied to the row of node n+1 (gredub / gbacsb forward part, tsf:602-761) -------------
    auto eliminate = [&](auto WN, const int n, double (&cD)[D], double (&cU)[D], double &crhs, double (&nL)[D],
                         double (&nD)[D], double &nrhs) {
        constexpr bool withNext = decltype(WN)::value;
        double *pub_cur = pub;
        const long long i_row = (long long)n * D + r;
        Osc2Rcp my_rcp = osc2_rcp_make(1.0, 1.0);
        Osc2Rcp nxt = osc2_rcp_prepare(cD[0]);          // every lane; lane k's is the one that gets published
#pragma unroll
        for (int k = 0; k < D; ++k) {
            if (act && r == k) {
                double *pk = pub_cur + k * PW;
#pragma unroll
                for (int c = k; c < D; ++c) pk[c] = cD[c];
#pragma unroll
                for (int c = 0; c < D; ++c) pk[D + c] = cU[c];
                pk[2 * D] = crhs;
                my_rcp = nxt;
                pk[2 * D + 1] = nxt.y;
            }
            __syncwarp();
            const double *pk = pub_cur + k * PW;
            const Osc2Rcp rp = osc2_rcp_make(pk[k], pk[2 * D + 1]);
            {
                // rows above the pivot (and idle lanes) take the multiplier 0: x - p * 0 leaves x unchanged
                const bool on = act && r > k;
                double q = osc2_div_p(cD[k], rp);
                q = on ? q : 0.0;
                if (k + 1 < D) {
                    cD[k + 1] = cD[k + 1] - pk[k + 1] * q;
                    nxt = osc2_rcp_prepare(cD[k + 1]);
                }
#pragma unroll
                for (int c = k + 2; c < D; ++c) cD[c] = cD[c] - pk[c] * q;
#pragma unroll
                for (int c = 0; c < D; ++c) cU[c] = cU[c] - pk[D + c] * q;
                crhs = crhs - pk[2 * D] * q;
            }
            if (withNext) {
                // every row of node n+1 lies below every row of node n
                double q = osc2_div_p(nL[k], rp);
                if (!FULL) q = act ? q : 0.0;
#pragma unroll
                for (int c = k + 1; c < D; ++c) nL[c] = nL[c] - pk[c] * q;
#pragma unroll
                for (int c = 0; c < D; ++c) nD[c] = nD[c] - pk[D + c] * q;
                nrhs = nrhs - pk[2 * D] * q;
            }
        }
        // the copies for the next construction were started before these rounds: collect them BEFORE the
        // stores below (the wait also drains younger stores from the load/store queue if it comes after them)
        osc2_async_wait_all();
        if (act) {
            // this lane's pivot: singular (tsf:676-682) or outside the exact range of the division
            const bool tiny = fabs(my_rcp.b) <= OSC2_TINY_PIVOT;
            if (tiny && bad_row < 0) bad_row = i_row;
            div_span = osc2_umax(div_span, tiny ? 0u : osc2_abs_hi(my_rcp.b) - OSC2_HI_DIV_LO);
#pragma unroll
            for (int c = 0; c < D; ++c) *fld(spptr, c) = cD[c];
#pragma unroll
            for (int c = 0; c < D; ++c) *fld(spptr, D + c) = cU[c];
            *fld(spptr, 2 * D) = crhs;
            *fld(spptr, 2 * D + 1) = my_rcp.y;
        }
        spptr += (size_t)D * (2 * D + 2) * B;
    };

    // parking slot: [0, D) U block of node n+1, [D, 2D + 9) element record
    auto park_put = [&](const double (&u)[D], const Elem &el) {
#pragma unroll
        for (int c = 0; c < D; ++c) park[c] = u[c];
#pragma unroll
        for (int c = 0; c < D; ++c) park[D + c] = el.sd[c];
        park[2 * D + 0] = el.sdr; park[2 * D + 1] = el.sdo; park[2 * D + 2] = el.kb; park[2 * D + 3] = el.md;
        park[2 * D + 4] = el.mo; park[2 * D + 5] = el.ahd; park[2 * D + 6] = el.aho; park[2 * D + 7] = el.lodu;
        park[2 * D + 8] = el.lodl;
    };
    auto park_get = [&](double (&u)[D], Elem &el) {
#pragma unroll
        for (int c = 0; c < D; ++c) u[c] = park[c];
#pragma unroll
        for (int c = 0; c < D; ++c) el.sd[c] = park[D + c];
        el.sdr = park[2 * D + 0]; el.sdo = park[2 * D + 1]; el.kb = park[2 * D + 2]; el.md = park[2 * D + 3];
        el.mo = park[2 * D + 4]; el.ahd = park[2 * D + 5]; el.aho = park[2 * D + 6]; el.lodu = park[2 * D + 7];
        el.lodl = park[2 * D + 8];
    };

    double cL[D], cD[D], cU[D], crhs = 0.0;       // node under elimination (cL: node 0 has no L block)
    double nL[D], nD[D], nU[D], nrhs = 0.0;       // node n+1
    if (act) {
        prefetch(0, gptr, xptr + xstep, lodptr);
        x0 = xptr[0];
    }
    osc2_async_wait_all();
    construct(std::false_type(), std::true_type(), 0, eA, eB, cL, cD, cU, crhs);
    eA = eB;                                                     // left element of node 1
    osc2_async_wait_all();                                       // (no rounds yet between this prefetch and its use)
    // one step of the pipeline: row of node n+1, then the rounds of node n
    auto stage = [&](auto HR, const int n) {
        construct(std::true_type(), HR, n + 1, eA, eB, nL, nD, nU, nrhs);
        if (act) park_put(nU, eB);
        eliminate(std::true_type(), n, cD, cU, crhs, nL, nD, nrhs);
#pragma unroll
        for (int c = 0; c < D; ++c) cD[c] = nD[c];
        crhs = nrhs;
        if (act) {
            park_get(cU, eA);                                    // U block of node n+1, element n+1 (left of node n+2)
        } else {
#pragma unroll
            for (int c = 0; c < D; ++c) cU[c] = 0.0;
        }
    };
    for (int n = 0; n < N - 2; ++n) stage(std::true_type(), n);
    stage(std::false_type(), N - 2);
    eliminate(std::false_type(), N - 1, cD, cU, crhs, nL, nD, nrhs);
    __syncwarp();
    double *scratch = base;
    if (r < D) { scratch[r] = (double)bad_row; scratch[D + r] = (div_span >= OSC2_HI_DIV_SPAN) ? 1.0 : 0.0; }
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
