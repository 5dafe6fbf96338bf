// K_F split across two warps per group of conductors (warp specialisation).
//
// The register-resident forward sweep (step_kernels_fast.cuh) is issue/latency bound: one warp
// runs ~2200 instructions per node for its 32/LPC conductors and only ~1.7 warps per scheduler
// exist for a batch of 4096 (profiles/r1_ncu_full_summary.txt).  Here the same arithmetic is cut
// into two instruction streams that run concurrently:
//   warp 0 (assembler):  row records + x_old + SYSLOD -> scaled block row [L D U | rhs]
//                        (SYSMAT/Known smc:1320-1474, boundary rows fc:255-565, scaling tsf:538-551)
//   warp 1 (eliminator): gredub/gbacsb forward part (tsf:602-761) on the handed-over rows, spill.
// Rows travel through a two-stage shared-memory ring guarded by named barriers (bar.sync /
// bar.arrive with 64 participants; ids 0/1 = stage full, 2/3 = stage empty), so each warp only
// ever waits for the other one.
// Every floating-point operation and its operand order is that of the single-warp kernel.
#pragma once

#include "step_kernels_fast.cuh"

#ifdef OSC2_EMU
#define OSC2_BAR_SYNC(id, n) emu_named_barrier(id)->arrive_and_wait()
#define OSC2_BAR_ARRIVE(id, n) (void)emu_named_barrier(id)->arrive()
#define OSC2_FENCE_BLOCK() std::atomic_thread_fence(std::memory_order_seq_cst)
#else
// barrier ids are immediates so that ptxas reserves exactly the four named barriers in use
// (a register id makes it reserve all 16, which caps the SM at two CTAs)
#define OSC2_BAR_SYNC(id, n) asm volatile("bar.sync %0, %1;" ::"n"(id), "n"(n) : "memory")
#define OSC2_BAR_ARRIVE(id, n) asm volatile("bar.arrive %0, %1;" ::"n"(id), "n"(n) : "memory")
#define OSC2_FENCE_BLOCK() __threadfence_block()
#endif

template <int D>
struct SplitFwdSmem {
    static constexpr int RWS = (3 * D + 5) & ~1;          // handed-over row: L | D | U | Known | max|row| | its reciprocal | range flag; 16-byte rows
    static constexpr int PW = 2 * D + 2;                   // published pivot row: D' | U | rhs | 1/pivot
    static constexpr int STAGE = D * RWS;                  // one node of one conductor
    static constexpr int PUB = 2 * D * PW;
    static constexpr int XS = 3 * D;
    static constexpr int STRIP = 5 * D + 1;
    // per conductor: [2 stages][D][RWS] | pub | xs | strips (strips only when Known needs them)
    __host__ __device__ static constexpr int per_cond(bool closed_known)
    {
        const int raw = 2 * STAGE + PUB + XS + (closed_known ? 0 : D * STRIP);
        return raw + ((18 - (raw % 16)) % 16);             // == 2 (mod 16): bank spread between conductors
    }
};

template <int D, int LPC, bool BE, bool CAPTURE>
__global__ void __maxnreg__(128) osc2_forward_split_kernel(const osc2_topology *__restrict__ tp, StepDev p)
{
    OSC2_DYN_SMEM(sm_dyn);
    using SM = SplitFwdSmem<D>;
    constexpr int PW = SM::PW, STRIP = SM::STRIP, RW = D + 9, RWS = SM::RWS;
    constexpr bool CLOSED_KNOWN = BE && (D % 8 == 0);
    constexpr int PER_COND = SM::per_cond(CLOSED_KNOWN);
    constexpr int cpb = 32 / LPC;
    const int M = p.M, N = p.N;
    const size_t B = (size_t)p.B;
    const int role = threadIdx.x >> 5;                    // 0: assembler, 1: eliminator
    const int lane = threadIdx.x & 31;
    const int grp = lane / LPC, r = lane % LPC;
    const long long b_ll = (long long)blockIdx.x * cpb + grp;
    const bool act = (b_ll < (long long)p.B) && (r < D);
    const size_t b = act ? (size_t)b_ll : 0;
    const int rr_ = (r < D) ? r : 0;
    const long long Total = (long long)D * N;

    double *base = sm_dyn + (size_t)grp * PER_COND;
    double *stage = base;                                  // [2][D][RWS]
    double *pub = base + 2 * SM::STAGE;                    // [2][D][PW]
    double *xs = pub + SM::PUB;                            // [3][D]
    double *flags = sm_dyn + (size_t)cpb * PER_COND;       // [2 roles][32 lanes][2]: bad row, range flag

    long long bad_row = -1;
    int range_flag = 0;

    if (role == 0) {
        // ------------------------------------------------------------------ assembler
        constexpr int Half = 2 * D;
        const double theta = tp->theta, omt = 1.0 - tp->theta;
        const Osc2Rcp rdt = osc2_rcp_prepare(p.dt);
        const bool first = (p.num_step == 1);
        const bool shift = (p.num_step > 1);
        osc2_rcp_check(rdt, range_flag);
        double *strip = xs + SM::XS + (size_t)rr_ * STRIP;   // only dereferenced when !CLOSED_KNOWN

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

        struct Elem {
            double sd[D];
            double kb, md, mo, ahd, aho, lodu, lodl, lodup, lodlp;
        };
        const size_t gstep = (size_t)p.NS * B;
        const double *gptr = p.gm + (size_t)rr_ * RW * B + b;
        auto load_elem = [&](const double *g, Elem &el) {
#pragma unroll
            for (int c = 0; c < D; ++c) el.sd[c] = g[(size_t)c * B];
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
            for (int c = 0; c < D; ++c) el.sd[c] = 0.0;
            el.md = el.mo = el.kb = el.ahd = el.aho = el.lodu = el.lodl = el.lodup = el.lodlp = 0.0;
        };
        Elem eL, eR;       // eL is refilled with element n+1 as soon as node n no longer needs element n-1
        zero_elem(eL); zero_elem(eR);
        const size_t xstep = (size_t)D * B;
        const double *xptr = p.sysvar + (size_t)rr_ * B + b;
        double *lodptr = p.syslod + (size_t)rr_ * 2 * B + b;
        // old solution of this lane's equation at nodes n-1, n, n+1 (+ prefetch of n+2)
        double x_prev = 0.0, x_cur = 0.0, x_nxt = 0.0, x_pre = 0.0, sl0_next = 0.0, sl1_next = 0.0;
        if (act) {
            load_elem(gptr, eR);
            x_cur = xptr[0];
            x_nxt = xptr[xstep];
            sl0_next = lodptr[0];
            sl1_next = lodptr[B];
            if (!CLOSED_KNOWN) { xs[0 * D + r] = x_cur; xs[1 * D + r] = x_nxt; }
        }
        if (!CLOSED_KNOWN && r < D) for (int q = 0; q < STRIP; ++q) strip[q] = 0.0;
        double *xa = xs + 2 * D, *xb = xs, *xc = xs + D;     // ring slots of nodes n-1, n, n+1

        auto node = [&](auto HL, auto HR, const int n) {
            constexpr bool hasL = decltype(HL)::value, hasR = decltype(HR)::value;
            const int s = n & 1;
            const long long i_row = (long long)n * D + r;
            double sl0 = sl0_next, sl1 = sl1_next;
            if (act) {
                if (n + 2 < N) x_pre = xptr[2 * xstep];
                if (n + 1 < N) { sl0_next = lodptr[2 * xstep]; sl1_next = lodptr[2 * xstep + B]; }
            }
            if (!CLOSED_KNOWN) __syncwarp();
            // the stage of node n-2 must have been drained by the eliminator before it is refilled
            if (s) OSC2_BAR_SYNC(3, 64); else OSC2_BAR_SYNC(2, 64);

            if (act) {
                double *dst = stage + ((size_t)s * D + r) * RWS;
                if (shift) { sl1 = sl0; sl0 = 0.0; }
                if (hasL) { sl0 += eL.lodl; if (first) sl1 += eL.lodlp; }
                if (hasR) { sl0 += eR.lodu; if (first) sl1 += eR.lodup; }
                lodptr[0] = sl0;
                lodptr[B] = sl1;

                const bool is_bc = (!hasL && bc_first) || (!hasR && bc_last);
                if (is_bc) {
#pragma unroll
                    for (int c = 0; c < D; ++c) { dst[c] = 0.0; dst[D + c] = (c == r) ? 1.0 : 0.0; dst[2 * D + c] = 0.0; }
                    dst[3 * D] = (!hasL) ? val_first : val_last;
                    dst[3 * D + 1] = 1.0;
                    dst[3 * D + 2] = 1.0;
                    if (n + 1 < N - 1) load_elem(gptr + gstep, eL);
                } else {
                    double mx = 0.0;
                    const double masL = hasL ? osc2_div_flag(eL.mo, rdt, range_flag) : 0.0;
                    const double masD = osc2_div_flag((hasL ? eL.md : 0.0) + (hasR ? eR.md : 0.0), rdt, range_flag);
                    const double masU = hasR ? osc2_div_flag(eR.mo, rdt, range_flag) : 0.0;
#pragma unroll
                    for (int c = 0; c < D; ++c) {
                        const bool dg = (c == r);
                        const double avL = dg ? eL.ahd : (c == aoffcol ? eL.aho : 0.0);
                        const double avR = dg ? eR.ahd : (c == aoffcol ? eR.aho : 0.0);
                        {
                            double sys = 0.0;
                            if (hasL) {
                                const double fds = (-avL + (dg ? -eL.kb : 0.0)) + eL.sd[c] / 2.;
                                const double m = dg ? masL : 0.0;
                                sys = BE ? m + fds : m + theta * fds;
                                if (!CLOSED_KNOWN) strip[D + c] = (m - omt * fds) * xa[c];
                            } else if (!CLOSED_KNOWN) strip[D + c] = 0.0;
                            dst[c] = sys;
                            { const double a_ = fabs(sys); if (a_ > mx) mx = a_; }
                        }
                        {
                            double flx = 0.0, dif = 0.0, sor = 0.0;
                            if (hasL) { if (dg) dif += eL.kb; flx += avL; sor += eL.sd[c]; }
                            if (hasR) { if (dg) dif += eR.kb; flx += -avR; sor += eR.sd[c]; }
                            const double fds = (flx + dif) + sor;
                            const double m = dg ? masD : 0.0;
                            const double sys = BE ? m + fds : m + theta * fds;
                            if (!CLOSED_KNOWN) strip[2 * D + c] = (m - omt * fds) * xb[c];
                            dst[D + c] = sys;
                            { const double a_ = fabs(sys); if (a_ > mx) mx = a_; }
                        }
                        {
                            double sys = 0.0;
                            if (hasR) {
                                const double fds = (avR + (dg ? -eR.kb : 0.0)) + eR.sd[c] / 2.;
                                const double m = dg ? masU : 0.0;
                                sys = BE ? m + fds : m + theta * fds;
                                if (!CLOSED_KNOWN) strip[3 * D + c] = (m - omt * fds) * xc[c];
                            } else if (!CLOSED_KNOWN) strip[3 * D + c] = 0.0;
                            dst[2 * D + c] = sys;
                            { const double a_ = fabs(sys); if (a_ > mx) mx = a_; }
                        }
                    }
                    double known;
                    if (CLOSED_KNOWN) {
                        known = 0.0;
                        if (hasL) known = masL * x_prev;
                        known = hasL ? known + masD * x_cur : masD * x_cur;
                        if (hasR) known = known + masU * x_nxt;
                    } else {
                        long long lo, hi;
                        if (i_row <= Half - 1) { lo = 0; hi = Half + i_row; }
                        else if (i_row >= Total - (Half - 1)) { lo = i_row - (Half - 1); hi = Total; }
                        else { lo = i_row - (Half - 1); hi = i_row + Half; }
                        const int Lfull = (int)(hi - lo);
                        const double *w = strip + D + (int)(lo - (long long)(n - 1) * D);
                        if (Lfull < 8) {
                            known = 0.0;
                            for (int q = 0; q < Lfull; ++q) known += w[q];
                        } else {
                            double a0 = w[0], a1 = w[1], a2 = w[2], a3 = w[3], a4 = w[4], a5 = w[5], a6 = w[6], a7 = w[7];
                            int q = 8;
                            for (; q < Lfull - (Lfull % 8); q += 8) {
                                a0 += w[q]; a1 += w[q + 1]; a2 += w[q + 2]; a3 += w[q + 3];
                                a4 += w[q + 4]; a5 += w[q + 5]; a6 += w[q + 6]; a7 += w[q + 7];
                            }
                            known = ((a0 + a1) + (a2 + a3)) + ((a4 + a5) + (a6 + a7));
                            for (; q < Lfull; ++q) known += w[q];
                        }
                    }
                    known += BE ? sl0 : (theta * sl0 + omt * sl1);
                    // element n-1 is dead: fetch element n+1 into its registers (used one node later)
                    if (n + 1 < N - 1) load_elem(gptr + gstep, eL);
                    // row scaling (tsf:538-551) is done by the eliminator with this reciprocal
                    const Osc2Rcp rmx = osc2_rcp_prepare(mx);
                    dst[3 * D] = known;
                    dst[3 * D + 1] = mx;
                    dst[3 * D + 2] = rmx.y;
                }
                // flags ride along: bit 0 = the assembler's cumulative range guard, bit 1 = boundary row (not scaled)
                dst[3 * D + 3] = (double)((range_flag ? 1 : 0) | (is_bc ? 2 : 0));
            }
            OSC2_FENCE_BLOCK();
            if (s) OSC2_BAR_ARRIVE(1, 64); else OSC2_BAR_ARRIVE(0, 64);
            if (!CLOSED_KNOWN) __syncwarp();      // every lane is done reading the old-solution ring of this node
            if (act) {
                { const Elem t = eL; eL = eR; eR = t; }     // eL <- element n, eR <- element n+1
                x_prev = x_cur; x_cur = x_nxt; x_nxt = x_pre;
                if (!CLOSED_KNOWN && n + 2 < N) xa[r] = x_pre;      // slot of node n-1 now holds node n+2
            }
            { double *t = xa; xa = xb; xb = xc; xc = t; }
            gptr += gstep;
            xptr += xstep;
            lodptr += 2 * xstep;
        };
        node(std::false_type(), std::true_type(), 0);
        for (int n = 1; n < N - 1; ++n) node(std::true_type(), std::true_type(), n);
        node(std::true_type(), std::false_type(), N - 1);
    } else {
        // ------------------------------------------------------------------ eliminator
        double *spptr = p.spill + (size_t)rr_ * (2 * D + 2) * B + b;
        OSC2_BAR_ARRIVE(2, 64);          // both stages start empty
        OSC2_BAR_ARRIVE(3, 64);
        for (int n = 0; n < N; ++n) {
            const int s = n & 1;
            const bool hasL = n > 0;
            double *pub_cur = pub + (size_t)(n & 1) * D * PW;
            const double *pub_prev = pub + (size_t)((n + 1) & 1) * D * PW;
            const long long i_row = (long long)n * D + r;
            double rowL[D], rowD[D], rowU[D], rhs = 0.0;
            if (s) OSC2_BAR_SYNC(1, 64); else OSC2_BAR_SYNC(0, 64);
            double kn = 0.0, mxv = 1.0, mxy = 1.0;
            int rowflags = 2;
            if (act) {
                const double *src = stage + ((size_t)s * D + r) * RWS;
#pragma unroll
                for (int c = 0; c < D; ++c) { rowL[c] = src[c]; rowD[c] = src[D + c]; rowU[c] = src[2 * D + c]; }
                kn = src[3 * D];
                mxv = src[3 * D + 1];
                mxy = src[3 * D + 2];
                rowflags = (int)src[3 * D + 3];
                if (rowflags & 1) range_flag = 1;
            } else {
#pragma unroll
                for (int c = 0; c < D; ++c) { rowL[c] = 0.0; rowD[c] = 0.0; rowU[c] = 0.0; }
            }
            if (n + 2 < N) {
                OSC2_FENCE_BLOCK();
                if (s) OSC2_BAR_ARRIVE(3, 64); else OSC2_BAR_ARRIVE(2, 64);
            }
            if (act) {
                // row scaling (tsf:538-551); boundary rows stay as they are
                if (!(rowflags & 2)) {
                    const Osc2Rcp rmx = osc2_rcp_make(mxv, mxy);
                    osc2_rcp_check(rmx, range_flag);
#pragma unroll
                    for (int c = 0; c < D; ++c) {
                        rowL[c] = osc2_div_flag(rowL[c], rmx, range_flag);
                        rowD[c] = osc2_div_flag(rowD[c], rmx, range_flag);
                        rowU[c] = osc2_div_flag(rowU[c], rmx, range_flag);
                    }
                    rhs = osc2_div_flag(kn, rmx, range_flag);
                } else {
                    rhs = kn;
                }
                if (CAPTURE) {
#pragma unroll
                    for (int c = 0; c < 3 * D; ++c) {
                        const long long col = (long long)(n - 1) * D + c;
                        if (col < 0 || col >= Total) continue;
                        const long long kb_ = (2 * D - 1) + (col - i_row);
                        const double v = (c < D) ? rowL[c % D] : (c < 2 * D ? rowD[c % D] : rowU[c % D]);
                        p.cap_sysmat[((size_t)kb_ * Total + i_row) * B + b] = v;
                    }
                    p.cap_known[(size_t)i_row * B + b] = rhs;
                }
            }
            // pivots of node n-1 (their rows were published by this warp one iteration ago)
            if (act && hasL) {
#pragma unroll
                for (int k = 0; k < D; ++k) {
                    const double *pk = pub_prev + k * PW;
                    const double q = osc2_div_flag(rowL[k], osc2_rcp_make(pk[k], pk[2 * D + 1]), range_flag);
#pragma unroll
                    for (int c = k + 1; c < D; ++c) rowL[c] = rowL[c] - pk[c] * q;
#pragma unroll
                    for (int c = 0; c < D; ++c) rowD[c] = rowD[c] - pk[D + c] * q;
                    rhs = rhs - pk[2 * D] * q;
                }
            }
            Osc2Rcp my_rcp = osc2_rcp_make(1.0, 1.0);
#pragma unroll
            for (int k = 0; k < D; ++k) {
                if (act && r == k) {
                    double *pk = pub_cur + k * PW;
#pragma unroll
                    for (int c = k; c < D; ++c) pk[c] = rowD[c];
#pragma unroll
                    for (int c = 0; c < D; ++c) pk[D + c] = rowU[c];
                    pk[2 * D] = rhs;
                    my_rcp = osc2_rcp_prepare(rowD[k]);
                    pk[2 * D + 1] = my_rcp.y;
                    if (fabs(rowD[k]) <= OSC2_TINY_PIVOT) { if (bad_row < 0) bad_row = i_row; }
                    else osc2_rcp_check(my_rcp, range_flag);
                }
                __syncwarp();
                if (act && r > k) {
                    const double *pk = pub_cur + k * PW;
                    const double q = osc2_div_flag(rowD[k], osc2_rcp_make(pk[k], pk[2 * D + 1]), range_flag);
#pragma unroll
                    for (int c = k + 1; c < D; ++c) rowD[c] = rowD[c] - pk[c] * q;
#pragma unroll
                    for (int c = 0; c < D; ++c) rowU[c] = rowU[c] - pk[D + c] * q;
                    rhs = rhs - pk[2 * D] * q;
                }
            }
            if (act) {
#pragma unroll
                for (int c = 0; c < D; ++c) spptr[(size_t)c * B] = rowD[c];
#pragma unroll
                for (int c = 0; c < D; ++c) spptr[(size_t)(D + c) * B] = rowU[c];
                spptr[(size_t)(2 * D) * B] = rhs;
                spptr[(size_t)(2 * D + 1) * B] = my_rcp.y;
            }
            spptr += (size_t)D * (2 * D + 2) * B;
            __syncwarp();      // pub_prev of the next node is this node's pub_cur: all lanes done reading the older one
        }
    }
    // status: first singular row wins (the reference raises there), else the range guard.  Only the
    // eliminator holds state here (the assembler's guard arrived with the rows), so a warp-level
    // exchange suffices -- no CTA barrier: the named barriers may still hold the assembler's arrivals.
    if (role == 1) {
        flags[(size_t)lane * 2 + 0] = (double)bad_row;
        flags[(size_t)lane * 2 + 1] = (double)range_flag;
        __syncwarp();
        if (act && r == 0) {
            long long worst = -1;
            bool range = false;
            for (int q = 0; q < LPC; ++q) {
                const double *f = flags + ((size_t)grp * LPC + q) * 2;
                const long long v = (long long)f[0];
                if (v >= 0 && (worst < 0 || v < worst)) worst = v;
                range = range || (f[1] != 0.0);
            }
            p.status[b] = (worst >= 0) ? -(int)(worst + 1) : (range ? OSC2_STATUS_RANGE : 0);
        }
    }
}
