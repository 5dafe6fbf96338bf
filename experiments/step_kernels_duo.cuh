// EXPERIMENT (round 2, not built): two warps per tile.  Bit-exact on the B200 (all 60 GPU tests with it as the default),
// but 42.5 ms against 28.9 ms of the one-warp pipelined kernel at 4096 x 10 001 nodes: the per-round critical path
// (publish -> barrier -> load -> quotient -> update -> reciprocal) is not shortened by moving the second update stream to
// another warp, and the named barriers add their wake-up latency nine times per node (ncu: barrier 0.47 and
// short_scoreboard 1.64 stall cycles per issued instruction at 128 registers; profiles/r2_kf_duo_ncu_summary.txt).
// Include after step_kernels_tile.cuh to build it.
#pragma once

// ---------------------------------------------------------------------------
// K_F, two warps per tile ("duo").  The two update streams of the node-pipelined sweep above are independent until
// the end of a node, so they get a warp each:
//   warp A (threads  0..31)  the pivots INSIDE node n: publishes pivot row k, updates its rows r > k, and sends the
//                            finished factor rows of the node to `spill` with one bulk store;
//   warp B (threads 32..63)  builds the row of node n+1 (element records from the TMA ring) and eliminates it against
//                            the pivot rows of node n as they are published; then hands the row (D block, U block,
//                            right-hand side) over to A, for which it becomes node n+1's starting row.
// A's rounds of node n+1 run while B builds and eliminates node n+2, so the serial instruction stream of a node is split
// over two warps (about 750 instead of 1500 warp instructions each), and a batch of 4096 conductors is 2048 warps.
// Synchronisation: named barriers with 64 participants, bar.arrive by the producer, bar.sync by the consumer
// (ids 1..D: pivot row k published; id D + 1: row handed over).  A starts node n+1 only after the hand-over, which B
// performs after it has read all pivot rows of node n: one pivot-row buffer and one hand-over buffer suffice.
// Same arithmetic, same bits as osc2_forward_fast_kernel / osc2_forward_tile_kernel.
// ---------------------------------------------------------------------------
#ifdef OSC2_EMU
#define OSC2_NBAR_SYNC(id) emu_named_barrier(id)->arrive_and_wait()
#define OSC2_NBAR_ARRIVE(id) (void)emu_named_barrier(id)->arrive()
#else
#define OSC2_NBAR_SYNC(id) asm volatile("bar.sync %0, 64;" ::"r"(id) : "memory")
#define OSC2_NBAR_ARRIVE(id) asm volatile("bar.arrive %0, 64;" ::"r"(id) : "memory")
#endif

template <int D, int LPC>
struct DuoFwdSmem {
    static constexpr int CPW = 32 / LPC;
    static constexpr int PW = 2 * D + 2;
    static constexpr int NF = D + 9;
    static constexpr int NST = 3;
    static constexpr int REC = NF * 32;
    static constexpr int STG = PW * 32;
    static constexpr int PUBC = D * PW + 2;
    static constexpr int HAND = (2 * D + 1) * 32;         // D block | U block | rhs of one row per lane
    static constexpr int DOUBLES = NST * REC + STG + CPW * PUBC + HAND + NST;
    static constexpr size_t BYTES = (size_t)DOUBLES * sizeof(double);
};

template <int D, int LPC, bool FULL>
__global__ void __launch_bounds__(64, 8) osc2_forward_duo_kernel(const osc2_topology *__restrict__ tp, StepDev p)
{
    static_assert(D % 8 == 0, "closed-form Known: D % 8 == 0");
    static_assert(!FULL || LPC == D, "FULL: every lane owns a row");
    OSC2_DYN_SMEM_ALIGNED(sm_dyn);
    using SM = DuoFwdSmem<D, LPC>;
    constexpr int CPW = SM::CPW, PW = SM::PW, NST = SM::NST;
    const int M = p.M, N = p.N;
    const size_t B = (size_t)p.B;
    const int lane = threadIdx.x & 31;
    const bool warpA = threadIdx.x < 32;
    const int r = lane / CPW, bb = lane % CPW;
    const size_t tile = blockIdx.x;
    const long long b_ll = (long long)tile * CPW + bb;
    const bool act = FULL || ((b_ll < (long long)p.B) && (r < D));
    const size_t b = act ? (size_t)b_ll : 0;
    const int rr_ = (FULL || r < D) ? r : 0;

    double *rec = sm_dyn;                                        // [NST][NF][32]       B reads
    double *stg = rec + NST * SM::REC;                           // [PW][32]            A writes, bulk store reads
    double *pub = stg + SM::STG + (size_t)bb * SM::PUBC;         // this conductor's [D][PW]: A writes, A and B read
    double *hand = stg + SM::STG + CPW * SM::PUBC + lane;        // [2D + 1][32]        B writes, A reads
    unsigned long long *bars = (unsigned long long *)(sm_dyn + NST * SM::REC + SM::STG + CPW * SM::PUBC + SM::HAND);
    const size_t sstep = (size_t)p.NT * SM::STG;

    if (threadIdx.x == 32) {
        for (int s = 0; s < NST; ++s) osc2_mbar_init(bars + s, 1);
        osc2_mbar_init_fence();
    }
    __syncthreads();

    if (warpA) {
        // =====================================================================================================
        // warp A: the pivots inside each node (gredub, tsf:602-698) and the factor rows
        // =====================================================================================================
        double *sptile = p.spill + tile * SM::STG;
        unsigned div_span = osc2_abs_hi(p.dt) - OSC2_HI_DIV_LO;     // divisor range: dt, then every pivot
        long long bad_row = -1;
        double cD[D], cU[D], crhs;
        for (int n = 0; n < N; ++n) {
            // the row of node n, eliminated against node n-1 (node 0: as built), comes from warp B
            OSC2_NBAR_SYNC(D + 1);
#pragma unroll
            for (int c = 0; c < D; ++c) cD[c] = hand[c * 32];
#pragma unroll
            for (int c = 0; c < D; ++c) cU[c] = hand[(D + c) * 32];
            crhs = hand[2 * D * 32];
            const long long i_row = (long long)n * D + r;
            Osc2Rcp my_rcp = osc2_rcp_make(1.0, 1.0);
            Osc2Rcp nxt = osc2_rcp_prepare(cD[0]);          // every lane; lane k's is the one that gets published
#pragma unroll
            for (int k = 0; k < D; ++k) {
                if (act && r == k) {
                    double *pk = pub + k * PW;
#pragma unroll
                    for (int c = k; c < D; ++c) pk[c] = cD[c];
#pragma unroll
                    for (int c = 0; c < D; ++c) pk[D + c] = cU[c];
                    pk[2 * D] = crhs;
                    my_rcp = nxt;
                    pk[2 * D + 1] = nxt.y;
                }
                __syncwarp();
                if (n + 1 < N) OSC2_NBAR_ARRIVE(1 + k);          // warp B may read pivot row k of node n
                const double *pk = pub + k * PW;
                const Osc2Rcp rp = osc2_rcp_make(pk[k], pk[2 * D + 1]);
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
            // the staging block is free once the bulk store of node n-1 has READ it
            if (lane == 0) osc2_bulk_store_wait_read();
            __syncwarp();
            if (act) {
                // this lane's pivot: singular (tsf:676-682) or outside the exact range of the division
                const bool tiny = fabs(my_rcp.b) <= OSC2_TINY_PIVOT;
                if (tiny && bad_row < 0) bad_row = i_row;
                div_span = osc2_umax(div_span, tiny ? 0u : osc2_abs_hi(my_rcp.b) - OSC2_HI_DIV_LO);
            }
            {
                double *sq = stg + lane;
#pragma unroll
                for (int c = 0; c < D; ++c) sq[c * 32] = cD[c];
#pragma unroll
                for (int c = 0; c < D; ++c) sq[(D + c) * 32] = cU[c];
                sq[2 * D * 32] = crhs;
                sq[(2 * D + 1) * 32] = my_rcp.y;
            }
            osc2_fence_async_smem();
            __syncwarp();
            if (lane == 0) osc2_bulk_store(sptile + (size_t)n * sstep, stg, (unsigned)(SM::STG * sizeof(double)));
        }
        if (lane == 0) osc2_bulk_store_wait_all();
        __syncwarp();
        double *scratch = stg;
        if (r < D) { scratch[lane] = (double)bad_row; scratch[32 + lane] = (div_span >= OSC2_HI_DIV_SPAN) ? 1.0 : 0.0; }
        __syncwarp();
        if (act && r == 0) {
            long long worst = -1;
            bool range = false;
            for (int q = 0; q < D; ++q) {
                const long long v = (long long)scratch[q * CPW + bb];
                if (v >= 0 && (worst < 0 || v < worst)) worst = v;
                range = range || (scratch[32 + q * CPW + bb] != 0.0);
            }
            // a singular pivot wins (the reference raises there); otherwise report the range problem; sticky
            if ((worst >= 0 || range) && p.status[b] == 0) p.status[b] = (worst >= 0) ? -(int)(worst + 1) : OSC2_STATUS_RANGE;
        }
        return;
    }

    // =========================================================================================================
    // warp B: rows (step() assembly, boundary rows, row scaling: smc:1015-1474, fc:117-565, tsf:538-551) and their
    // elimination against the previous node (gredub / gbacsb forward part)
    // =========================================================================================================
    const Osc2Rcp rdt = osc2_rcp_prepare(p.dt);
    const bool first = (p.num_step == 1);
    const bool shift = (p.num_step > 1);
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
    const double *gtile = p.gm + tile * SM::REC;
    const size_t gstep = (size_t)p.NT * SM::REC;
    auto issue = [&](const int e) {                              // lane 0: start the copy of element e (stage e % NST is free)
        if (lane == 0 && e < N - 1) {
            const int s = e % NST;
            osc2_mbar_expect_tx(bars + s, (unsigned)(SM::REC * sizeof(double)));
            osc2_bulk_load(rec + (size_t)s * SM::REC, gtile + (size_t)e * gstep, (unsigned)(SM::REC * sizeof(double)), bars + s);
        }
    };
    auto arrived = [&](const int e) {
        if (e < N - 1) osc2_mbar_wait(bars + e % NST, (unsigned)((e / NST) & 1));
    };
    issue(0);
    issue(1);
    const int f_own = lane;
    const int f_sdr = rr_ * 32 + lane, f_sdo = (aoffcol >= 0 ? aoffcol : 0) * 32 + lane;
    const size_t xstep = (size_t)D * B;
    const double *xptr = p.sysvar + (size_t)rr_ * B + b;
    double *lodptr = p.syslod + (size_t)rr_ * 2 * B + b;
    double xm = 0.0, x0 = 0.0, xp = 0.0;
    double x_pre = 0.0, sl0_pre = 0.0, sl1_pre = 0.0;
    double mas_carry = 0.0;

    auto construct = [&](auto HL, auto HR, const int m, double (&rowL)[D], double (&rowD)[D], double (&rowU)[D], double &rhs) {
        constexpr bool hasL = decltype(HL)::value, hasR = decltype(HR)::value;
        const double *A = rec + (size_t)((m + NST - 1) % NST) * SM::REC + f_own;   // element m-1 (left)
        const double *Bv = rec + (size_t)(m % NST) * SM::REC + f_own;              // element m (right)
        const double *Ar = rec + (size_t)((m + NST - 1) % NST) * SM::REC, *Br = rec + (size_t)(m % NST) * SM::REC;
        rhs = 0.0;
        if (act) {
            if (hasR) xp = x_pre;
            double sl0 = sl0_pre, sl1 = sl1_pre;
            if (m + 2 < N) x_pre = xptr[2 * xstep];
            if (m + 1 < N) { sl0_pre = lodptr[2 * xstep]; sl1_pre = lodptr[2 * xstep + B]; }
            if (shift) { sl1 = sl0; sl0 = 0.0; }
            if (hasL) { sl0 += A[(D + 6) * 32]; if (first) sl1 += A[(D + 8) * 32]; }
            if (hasR) { sl0 += Bv[(D + 5) * 32]; if (first) sl1 += Bv[(D + 7) * 32]; }
            lodptr[0] = sl0;
            lodptr[B] = sl1;
            const double masL = hasL ? mas_carry : 0.0;
            const double masU = hasR ? osc2_div_p(Bv[(D + 2) * 32], rdt) : 0.0;
            mas_carry = masU;
            const bool is_bc = (!hasL && bc_first) || (!hasR && bc_last);
            if (is_bc) {
#pragma unroll
                for (int c = 0; c < D; ++c) { rowL[c] = 0.0; rowD[c] = (c == r) ? 1.0 : 0.0; rowU[c] = 0.0; }
                rhs = (!hasL) ? val_first : val_last;
            } else {
                const double Amd = hasL ? A[(D + 1) * 32] : 0.0, Bmd = hasR ? Bv[(D + 1) * 32] : 0.0;
                const double masD = osc2_div_p(Amd + Bmd, rdt);
                double Akb = 0.0, Aahd = 0.0, Aaho = 0.0, Asdr = 0.0, Asdo = 0.0;
                double Bkb = 0.0, Bahd = 0.0, Baho = 0.0, Bsdr = 0.0, Bsdo = 0.0;
                if (hasL) { Akb = A[D * 32]; Aahd = A[(D + 3) * 32]; Aaho = A[(D + 4) * 32]; Asdr = Ar[f_sdr]; Asdo = Ar[f_sdo]; }
                if (hasR) { Bkb = Bv[D * 32]; Bahd = Bv[(D + 3) * 32]; Baho = Bv[(D + 4) * 32]; Bsdr = Br[f_sdr]; Bsdo = Br[f_sdo]; }
                double fLd = 0.0, fLo = 0.0, fUd = 0.0, fUo = 0.0, fDd, fDo;
                if (hasL) { fLd = (-Aahd + -Akb) + Asdr / 2.; fLo = -Aaho + Asdo / 2.; }
                if (hasR) { fUd = (Bahd + -Bkb) + Bsdr / 2.; fUo = Baho + Bsdo / 2.; }
                if (hasL && hasR) {
                    fDd = ((Aahd + -Bahd) + (Akb + Bkb)) + (Asdr + Bsdr);
                    fDo = (Aaho + -Baho) + (Asdo + Bsdo);
                } else if (hasL) {
                    fDd = (Aahd + Akb) + Asdr;
                    fDo = Aaho + Asdo;
                } else {
                    fDd = (-Bahd + Bkb) + Bsdr;
                    fDo = -Baho + Bsdo;
                }
                double mxL = 0.0, mxD = 0.0, mxU = 0.0;
                const double sLd = masL + fLd, sDd = masD + fDd, sUd = masU + fUd;
#pragma unroll
                for (int c = 0; c < D; ++c) {
                    const bool dg = (c == r), of = (c == aoffcol);
                    const double As = hasL ? A[c * 32] : 0.0, Bs = hasR ? Bv[c * 32] : 0.0;
                    const double gD = (hasL && hasR) ? As + Bs : (hasL ? As : Bs);
                    rowL[c] = hasL ? (dg ? sLd : (of ? fLo : As / 2.)) : 0.0;
                    rowD[c] = dg ? sDd : (of ? fDo : gD);
                    rowU[c] = hasR ? (dg ? sUd : (of ? fUo : Bs / 2.)) : 0.0;
                    if (hasL) mxL = osc2_absmax(mxL, rowL[c]);
                    mxD = osc2_absmax(mxD, rowD[c]);
                    if (hasR) mxU = osc2_absmax(mxU, rowU[c]);
                }
                const double mx = osc2_absmax(osc2_absmax(mxL, mxD), mxU);
                double known = 0.0;
                if (hasL) known = masL * xm;
                known = hasL ? known + masD * x0 : masD * x0;
                if (hasR) known = known + masU * xp;
                known += sl0;
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
            xm = x0; x0 = xp;
        } else {
#pragma unroll
            for (int c = 0; c < D; ++c) { rowL[c] = 0.0; rowD[c] = 0.0; rowU[c] = 0.0; }
        }
        xptr += xstep;
        lodptr += 2 * xstep;
    };
    // hand the row over to warp A
    auto hand_over = [&](const double (&rowD)[D], const double (&rowU)[D], const double rhs) {
#pragma unroll
        for (int c = 0; c < D; ++c) hand[c * 32] = rowD[c];
#pragma unroll
        for (int c = 0; c < D; ++c) hand[(D + c) * 32] = rowU[c];
        hand[2 * D * 32] = rhs;
        OSC2_NBAR_ARRIVE(D + 1);
    };

    double nL[D], nD[D], nU[D], nrhs = 0.0;
    if (act) {
        x0 = xptr[0];
        x_pre = xptr[xstep];
        sl0_pre = lodptr[0];
        sl1_pre = lodptr[B];
    }
    arrived(0);
    construct(std::false_type(), std::true_type(), 0, nL, nD, nU, nrhs);
    hand_over(nD, nU, nrhs);
    auto stage = [&](auto HR, const int m) {                    // node m >= 1: build, eliminate against node m-1, hand over
        // element m+1 goes where element m-2 was: its last readers (the construction of node m-1, this warp) are done
        __syncwarp();
        issue(m + 1);
        arrived(m);
        construct(std::true_type(), HR, m, nL, nD, nU, nrhs);
#pragma unroll
        for (int k = 0; k < D; ++k) {
            OSC2_NBAR_SYNC(1 + k);                               // pivot row k of node m-1 is published
            const double *pk = pub + k * PW;
            const Osc2Rcp rp = osc2_rcp_make(pk[k], pk[2 * D + 1]);
            double q = osc2_div_p(nL[k], rp);
            if (!FULL) q = act ? q : 0.0;
#pragma unroll
            for (int c = k + 1; c < D; ++c) nL[c] = nL[c] - pk[c] * q;
#pragma unroll
            for (int c = 0; c < D; ++c) nD[c] = nD[c] - pk[D + c] * q;
            nrhs = nrhs - pk[2 * D] * q;
        }
        hand_over(nD, nU, nrhs);
    };
    for (int m = 1; m < N - 1; ++m) stage(std::true_type(), m);
    stage(std::false_type(), N - 1);
}
