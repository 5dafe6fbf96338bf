// K_F / K_B for NODOFS = 8, Backward Euler: tiled records moved by TMA bulk copies, rounds pipelined over nodes.
//
// What is different from osc2_forward_fast_kernel (same arithmetic, same bits; file/line references there):
//
//  * DATA MOVEMENT.  `gm` and `spill` are tiled by warp (osc2_dev.h): everything the warp needs for one node is
//    one contiguous 4352-byte (element record) / 4608-byte (factor rows) block.  Lane 0 moves them with
//    cp.async.bulk (TMA, SASS UBLKCP): global -> shared two elements ahead of their use, completion on an mbarrier;
//    shared -> global for the factor rows.  No per-lane LDG/STG of records, no 64-bit address arithmetic and no
//    prefetch registers in the elimination chain; every DRAM burst is used whole (the round-1 kernel read 32-byte
//    sectors of 64-byte bursts from two SMs: 1.32x the algorithmic traffic).
//  * ELEMENT RECORDS STAY IN SHARED MEMORY and are read where they are used (the round-1 kernel kept two records,
//    68 registers, live across the whole node).
//  * ROUNDS.  Round k of the elimination of node n+1 against node n needs only pivot row k of node n, which is final
//    the moment round k of node n's own elimination publishes it.  The row of node n+1 is therefore built FIRST and
//    both eliminations run in the same 8 rounds: 8 serial publish -> load -> divide exchanges per node instead of 16,
//    one pivot-row load serving two independent update streams.  (Round 1 had this as an experiment that spilled at
//    255 registers; with the records in shared memory it fits.)
//
// Lane = row * CPW + conductor of the tile, one warp per CTA, CTA = tile.
#pragma once

#include "step_kernels_fast.cuh"

// ---- TMA bulk copy + mbarrier (PTX ISA 8.x, sm_90+); synchronous stand-ins under the CPU emulator -------------
#ifdef OSC2_EMU
inline void osc2_mbar_init(unsigned long long *, unsigned) {}
inline void osc2_mbar_init_fence() {}
inline void osc2_mbar_expect_tx(unsigned long long *, unsigned) {}
inline void osc2_bulk_load(void *dst, const void *src, unsigned bytes, unsigned long long *) { std::memcpy(dst, src, bytes); }
inline void osc2_mbar_wait(unsigned long long *, unsigned) { __syncwarp(); }     // every lane calls it: lane 0's copy is done
inline void osc2_bulk_store(void *dst, const void *src, unsigned bytes) { std::memcpy(dst, src, bytes); }
inline void osc2_bulk_store_wait_read() {}
inline void osc2_bulk_store_wait_all() {}
inline void osc2_fence_async_smem() {}
// CTA-scope mbarrier between warps (fused kernel): 16 bytes of shared memory per barrier
struct EmuMbar { std::atomic<unsigned> pending, phase; unsigned count, pad; };
inline void osc2_cbar_init(void *bar, unsigned count) { auto *m = new (bar) EmuMbar; m->pending.store(count); m->phase.store(0); m->count = count; }
inline void osc2_cbar_arrive(void *bar)
{
    auto *m = (EmuMbar *)bar;
    if (m->pending.fetch_sub(1, std::memory_order_acq_rel) == 1) { m->pending.store(m->count); m->phase.fetch_add(1, std::memory_order_release); }
}
inline void osc2_cbar_wait(void *bar, unsigned parity)
{
    auto *m = (EmuMbar *)bar;
    while ((m->phase.load(std::memory_order_acquire) & 1u) == parity) std::this_thread::yield();
}
#else
__device__ __forceinline__ unsigned osc2_smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void osc2_mbar_init(unsigned long long *bar, unsigned count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(osc2_smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void osc2_mbar_init_fence()
{
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void osc2_mbar_expect_tx(unsigned long long *bar, unsigned bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(osc2_smem_u32(bar)), "r"(bytes) : "memory");
}
// global -> shared, `bytes` a multiple of 16, both addresses 16-byte aligned; completes on `bar`
__device__ __forceinline__ void osc2_bulk_load(void *dst, const void *src, unsigned bytes, unsigned long long *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(osc2_smem_u32(dst)), "l"(src), "r"(bytes), "r"(osc2_smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void osc2_mbar_wait(unsigned long long *bar, unsigned parity)
{
    unsigned ok;
    const unsigned a = osc2_smem_u32(bar);
    do {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(a), "r"(parity) : "memory");
    } while (!ok);
}
// shared -> global
__device__ __forceinline__ void osc2_bulk_store(void *dst, const void *src, unsigned bytes)
{
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(osc2_smem_u32(src)), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
__device__ __forceinline__ void osc2_bulk_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void osc2_bulk_store_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
// CTA-scope mbarrier between warps (fused kernel): arrive has release, the wait acquire semantics
__device__ __forceinline__ void osc2_cbar_init(void *bar, unsigned count) { osc2_mbar_init((unsigned long long *)bar, count); }
__device__ __forceinline__ void osc2_cbar_arrive(void *bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(osc2_smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void osc2_cbar_wait(void *bar, unsigned parity) { osc2_mbar_wait((unsigned long long *)bar, parity); }
// generic-proxy writes to shared memory -> visible to the async proxy (the bulk store engine)
__device__ __forceinline__ void osc2_fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
#endif

#ifndef OSC2_DYN_SMEM_ALIGNED
#define OSC2_DYN_SMEM_ALIGNED(name) extern __shared__ __align__(128) double name##_tile[]; double *name = name##_tile
#endif

template <int D, int LPC>
struct TileFwdSmem {
    static constexpr int CPW = 32 / LPC;
    static constexpr int PW = 2 * D + 2;                  // published / spilled row: D' | U | rhs | 1/pivot
    static constexpr int NF = D + 9;                      // fields of an element record
    static constexpr int NST = 3;                         // ring: elements m-1 and m in use, m+1 in flight
    static constexpr int REC = NF * 32;                   // doubles of one (element, tile) record
    static constexpr int STG = PW * 32;                   // doubles of one (node, tile) block of factor rows
    static constexpr int PUBC = D * PW + 2;               // pivot rows of one conductor; + 2: the conductors of a tile sit
                                                          // 4 banks apart, their lanes read the same offsets at once
    static constexpr int DOUBLES = NST * REC + STG + CPW * PUBC + NST;   // + the mbarriers (8 bytes each)
    static constexpr size_t BYTES = (size_t)DOUBLES * sizeof(double);
};

// Where one warp of the forward sweep finds its shared memory.  FUSED (step_kernels_fused.cuh): the element records
// are not copied in from `gm` but written into the ring by a producer warp of the same CTA; `full` / `empty` are the
// mbarriers of the ring stages (16 bytes apart), shared by the consumer warps one producer serves.
struct TileFwdCtx {
    double *rec;                 // [NST][NF][32] ring of element records
    double *stg;                 // [PW][32] factor rows of one node on their way out
    double *pub;                 // [CPW][PUBC] published pivot rows
    unsigned long long *bars;    // !FUSED: the NST mbarriers of the bulk loads
    double *full, *empty;        // FUSED
    size_t tile;
    bool tile_valid;             // a CTA that covers several tiles may own tiles past the batch
    // split roles (osc2_forward_split_kernel): scaled rows handed from the assembler warp to the eliminator warp
    double *hand;                // [2][3 D + 1][32]
    double *hfull, *hempty;      // mbarriers of the two hand-over slots, 16 bytes apart
};
enum { OSC2_ROLE_BOTH = 0, OSC2_ROLE_ASM = 1, OSC2_ROLE_ELIM = 2 };

template <int D, int LPC, bool FULL, bool FUSED, int NST, int ROLE = OSC2_ROLE_BOTH>
__device__ __forceinline__ void osc2_forward_tile_body(const osc2_topology *__restrict__ tp, const StepDev &p, const TileFwdCtx &cx)
{
    static_assert(D % 8 == 0, "closed-form Known: D % 8 == 0");
    static_assert(!FULL || LPC == D, "FULL: every lane owns a row");
    using SM = TileFwdSmem<D, LPC>;
    constexpr int CPW = SM::CPW, PW = SM::PW;
    const int M = p.M, N = p.N;
    const size_t B = (size_t)p.B;
    const int lane = threadIdx.x & 31;
    const int r = lane / CPW, bb = lane % CPW;
    const size_t tile = cx.tile;
    const long long b_ll = (long long)tile * CPW + bb;
    const bool act = FULL || ((b_ll < (long long)p.B) && (r < D) && cx.tile_valid);
    const size_t b = act ? (size_t)b_ll : 0;
    const int rr_ = (FULL || r < D) ? r : 0;
    const Osc2Rcp rdt = osc2_rcp_prepare(p.dt);
    const bool first = (p.num_step == 1);
    const bool shift = (p.num_step > 1);
    unsigned div_span = osc2_abs_hi(p.dt) - OSC2_HI_DIV_LO;     // divisor range: dt, then every pivot

    double *rec = cx.rec;                                        // [NST][NF][32]
    double *stg = cx.stg;                                        // [PW][32]
    double *pub = cx.pub + (size_t)bb * SM::PUBC;                // this conductor's [D][PW]
    unsigned long long *bars = cx.bars;

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

    // ---- the ring of element records --------------------------------------------------------------------------
    const double *gtile = p.gm + tile * SM::REC;                 // record of element 0 of this tile
    const size_t gstep = (size_t)p.NT * SM::REC;
    double *sptile = p.spill + tile * SM::STG;                   // factor rows of node 0 of this tile
    const size_t sstep = (size_t)p.NT * SM::STG;
    auto issue = [&](const int e) {                              // lane 0: start the copy of element e (stage e % NST is free)
        if (!FUSED && lane == 0 && e < N - 1) {
            const int s = e % NST;
            osc2_mbar_expect_tx(bars + s, (unsigned)(SM::REC * sizeof(double)));
            osc2_bulk_load(rec + (size_t)s * SM::REC, gtile + (size_t)e * gstep, (unsigned)(SM::REC * sizeof(double)), bars + s);
        }
    };
    auto arrived = [&](const int e) {                            // every lane: element e is in its stage
        if (e < N - 1) {
            if (FUSED) osc2_cbar_wait(cx.full + 2 * (e % NST), (unsigned)((e / NST) & 1));
            else osc2_mbar_wait(bars + e % NST, (unsigned)((e / NST) & 1));
        }
    };
    auto release = [&](const int e) {                            // every lane: this lane has read element e for the last time
        if (FUSED && e >= 0 && e < N - 1) osc2_cbar_arrive(cx.empty + 2 * (e % NST));
    };
    if (!FUSED && ROLE != OSC2_ROLE_ELIM) {
        if (ROLE == OSC2_ROLE_BOTH) {          // split roles: the kernel initialises every barrier before the roles part
            if (lane == 0) {
                for (int s = 0; s < NST; ++s) osc2_mbar_init(bars + s, 1);
                osc2_mbar_init_fence();
            }
            __syncwarp();
        }
        issue(0);
        issue(1);
    }

    // per-lane field offsets inside a record
    const int f_own = lane;                                      // field f of this lane: rec[f * 32 + lane]
    const int f_sdr = rr_ * 32 + lane, f_sdo = (aoffcol >= 0 ? aoffcol : 0) * 32 + lane;

    const size_t xstep = (size_t)D * B;
    const double *xptr = p.sysvar + (size_t)rr_ * B + b;                 // x[node under construction][r]
    double *lodptr = p.syslod + (size_t)rr_ * 2 * B + b;                 // SYSLOD[node under construction][r][0..1]
    double xm = 0.0, x0 = 0.0, xp = 0.0;                                 // own old solution at nodes m-1, m, m+1
    double x_pre = 0.0, sl0_pre = 0.0, sl1_pre = 0.0;                    // register prefetch, one node ahead
    double mas_carry = 0.0;                                              // MASMAT/dt of the element to the left

    // ---- row of node m, scaled (the arithmetic of osc2_forward_fast_kernel, closed-form Known) ----------------
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
            // SYSLOD (smc:1165-1318); the previous-step halves exist at the first step only
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
        xptr += xstep;
        lodptr += 2 * xstep;
    };

    // ---- the 8 rounds of node n: its own pivots (rows r > k) and, with WN, the pivots of node n applied to the row
    //      of node n+1 (gredub / gbacsb forward part, tsf:602-761); then the factor rows of node n leave through the
    //      staging block with one bulk store ----------------------------------------------------------------------
    auto eliminate = [&](auto WN, const int n, double (&cD)[D], double (&cU)[D], double &crhs, double (&nL)[D],
                         double (&nD)[D], double &nrhs) {
        constexpr bool withNext = decltype(WN)::value;
        const long long i_row = (long long)n * D + r;
        Osc2Rcp my_rcp = osc2_rcp_make(1.0, 1.0);
        Osc2Rcp nxt = osc2_rcp_prepare(cD[0]);          // every lane; lane k's is the one that gets published
        unsigned qkey = 0xffffffffu;                    // smallest numerator key of this node's 16 multipliers
        // (publishing row k + 1 inside round k, ahead of the updates of the next node's row, was tried: the longer live
        // ranges cost 255 registers and spills -- 36.9 instead of 28.9 ms)
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
            const double *pk = pub + k * PW;
            const Osc2Rcp rp = osc2_rcp_make(pk[k], pk[2 * D + 1]);
            if (k + 1 < D) {          // no row lies below the last pivot: its round updates the next node's row only
                // rows above the pivot (and idle lanes) take the multiplier 0: x - p * 0 leaves x unchanged
                const bool on = act && r > k;
                double q = osc2_div_k(cD[k], rp, qkey);
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
                double q = osc2_div_k(nL[k], rp, qkey);
                if (!FULL) q = act ? q : 0.0;
#pragma unroll
                for (int c = k + 1; c < D; ++c) nL[c] = nL[c] - pk[c] * q;
#pragma unroll
                for (int c = 0; c < D; ++c) nD[c] = nD[c] - pk[D + c] * q;
                nrhs = nrhs - pk[2 * D] * q;
            }
        }
        // the staging block is free once the bulk store of node n-1 has READ it
        if (ROLE != OSC2_ROLE_ELIM) {
            if (lane == 0) osc2_bulk_store_wait_read();
            __syncwarp();
        }
        if (act) {
            // this lane's pivot: singular (tsf:676-682) or outside the exact range of the division
            const bool tiny = fabs(my_rcp.b) <= OSC2_TINY_PIVOT;
            if (tiny && bad_row < 0) bad_row = i_row;
            div_span = osc2_umax(div_span, tiny ? 0u : osc2_abs_hi(my_rcp.b) - OSC2_HI_DIV_LO);
            // a multiplier's numerator in (0, 2^-900): its quotient is not verified exact -> the conductor reports RANGE
            if (qkey < OSC2_KEY_MIN) div_span = 0xffffffffu;
        }
        if (ROLE == OSC2_ROLE_ELIM) {
            // split roles: no staging block (its shared memory went to the hand-over slots); a warp-wide store of one
            // field is 256 contiguous bytes of the tiled `spill`
            if (cx.tile_valid) {
                double *sq = sptile + (size_t)n * sstep + lane;
#pragma unroll
                for (int c = 0; c < D; ++c) sq[c * 32] = cD[c];
#pragma unroll
                for (int c = 0; c < D; ++c) sq[(D + c) * 32] = cU[c];
                sq[2 * D * 32] = crhs;
                sq[(2 * D + 1) * 32] = my_rcp.y;
            }
        } else {
            double *sq = stg + lane;
#pragma unroll
            for (int c = 0; c < D; ++c) sq[c * 32] = cD[c];
#pragma unroll
            for (int c = 0; c < D; ++c) sq[(D + c) * 32] = cU[c];
            sq[2 * D * 32] = crhs;
            sq[(2 * D + 1) * 32] = my_rcp.y;
            osc2_fence_async_smem();
            __syncwarp();
            if (lane == 0 && cx.tile_valid) osc2_bulk_store(sptile + (size_t)n * sstep, stg, (unsigned)(SM::STG * sizeof(double)));
        }
    };

    double cL[D], cD[D], cU[D], crhs = 0.0;       // node under elimination (cL: node 0 has no L block)
    double nL[D], nD[D], nU[D], nrhs = 0.0;       // node n+1
    constexpr int HREC = (3 * D + 1) * 32;        // doubles of one hand-over slot: L | D | U | rhs, lane innermost
    if (ROLE != OSC2_ROLE_ELIM && act) {
        x0 = xptr[0];
        x_pre = xptr[xstep];
        sl0_pre = lodptr[0];
        sl1_pre = lodptr[B];
    }
    if constexpr (ROLE == OSC2_ROLE_BOTH) {
        arrived(0);
        construct(std::false_type(), std::true_type(), 0, cL, cD, cU, crhs);
        release(-1);
        // one step of the pipeline: row of node n+1, then the rounds of node n
        auto stage = [&](auto HR, const int n) {
            // element n+2 goes where element n-1 was: its last readers (the construction of node n) are past the
            // __syncwarp()s of the rounds that followed
            issue(n + 2);
            arrived(n + 1);
            construct(std::true_type(), HR, n + 1, nL, nD, nU, nrhs);
            release(n);                                  // element n was the left neighbour of node n + 1
            eliminate(std::true_type(), n, cD, cU, crhs, nL, nD, nrhs);
#pragma unroll
            for (int c = 0; c < D; ++c) { cD[c] = nD[c]; cU[c] = nU[c]; }
            crhs = nrhs;
        };
        for (int n = 0; n < N - 2; ++n) stage(std::true_type(), n);
        stage(std::false_type(), N - 2);
        eliminate(std::false_type(), N - 1, cD, cU, crhs, nL, nD, nrhs);
        if (lane == 0) osc2_bulk_store_wait_all();
        __syncwarp();
    } else if constexpr (ROLE == OSC2_ROLE_ASM) {
        // ---- assembler warp: the scaled row of node m into hand-over slot m % 2, up to two nodes ahead of the eliminator.
        //      Every lane writes and (in the eliminator warp) reads its own row only ----------------------------------
        auto hand_over = [&](const int m, const double (&rowL)[D], const double (&rowD)[D], const double (&rowU)[D], const double rhs) {
            const int s = m & 1;
            if (m >= 2) osc2_cbar_wait(cx.hempty + 2 * s, (unsigned)(((m >> 1) - 1) & 1));
            double *h = cx.hand + (size_t)s * HREC + lane;
#pragma unroll
            for (int c = 0; c < D; ++c) { h[c * 32] = rowL[c]; h[(D + c) * 32] = rowD[c]; h[(2 * D + c) * 32] = rowU[c]; }
            h[3 * D * 32] = rhs;
            osc2_cbar_arrive(cx.hfull + 2 * s);
        };
        arrived(0);
        construct(std::false_type(), std::true_type(), 0, cL, cD, cU, crhs);
        hand_over(0, cL, cD, cU, crhs);
        for (int m = 1; m < N - 1; ++m) {
            __syncwarp();                                // every lane is past its reads of element m - 2
            issue(m + 1);
            arrived(m);
            construct(std::true_type(), std::true_type(), m, nL, nD, nU, nrhs);
            hand_over(m, nL, nD, nU, nrhs);
        }
        construct(std::true_type(), std::false_type(), N - 1, nL, nD, nU, nrhs);
        hand_over(N - 1, nL, nD, nU, nrhs);
        return;
    } else {
        // ---- eliminator warp ----------------------------------------------------------------------------------------
        auto take = [&](const int m) { osc2_cbar_wait(cx.hfull + 2 * (m & 1), (unsigned)((m >> 1) & 1)); };
        {
            take(0);
            const double *h = cx.hand + lane;
#pragma unroll
            for (int c = 0; c < D; ++c) { cD[c] = h[(D + c) * 32]; cU[c] = h[(2 * D + c) * 32]; }
            crhs = h[3 * D * 32];
            osc2_cbar_arrive(cx.hempty);
        }
        for (int n = 0; n < N - 1; ++n) {
            const int s = (n + 1) & 1;
            take(n + 1);
            const double *h = cx.hand + (size_t)s * HREC + lane;
#pragma unroll
            for (int c = 0; c < D; ++c) { nL[c] = h[c * 32]; nD[c] = h[(D + c) * 32]; }
            nrhs = h[3 * D * 32];
            eliminate(std::true_type(), n, cD, cU, crhs, nL, nD, nrhs);
#pragma unroll
            for (int c = 0; c < D; ++c) { cD[c] = nD[c]; cU[c] = h[(2 * D + c) * 32]; }     // the U block is untouched by node n's pivots
            crhs = nrhs;
            osc2_cbar_arrive(cx.hempty + 2 * s);
        }
        eliminate(std::false_type(), N - 1, cD, cU, crhs, nL, nD, nrhs);
        __syncwarp();
    }
    double *scratch = (ROLE == OSC2_ROLE_ELIM) ? cx.hand : rec;     // every element / row has been consumed
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
}

template <int D, int LPC, bool FULL>
__global__ void __launch_bounds__(32) osc2_forward_tile_kernel(const osc2_topology *__restrict__ tp, StepDev p)
{
    OSC2_DYN_SMEM_ALIGNED(sm_dyn);
    using SM = TileFwdSmem<D, LPC>;
    TileFwdCtx cx;
    cx.rec = sm_dyn;
    cx.stg = cx.rec + SM::NST * SM::REC;
    cx.pub = cx.stg + SM::STG;
    cx.bars = (unsigned long long *)(sm_dyn + SM::NST * SM::REC + SM::STG + SM::CPW * SM::PUBC);
    cx.full = cx.empty = cx.hand = cx.hfull = cx.hempty = nullptr;
    cx.tile = blockIdx.x;
    cx.tile_valid = true;
    osc2_forward_tile_body<D, LPC, FULL, false, SM::NST>(tp, p, cx);
}

// ---------------------------------------------------------------------------
// K_B on the tiled factor rows: back substitution (gbacsb, tsf:772-802), node N-1 down to 0, the arithmetic of
// osc2_backward_fast_kernel.  The (node, tile) blocks of `spill` arrive by bulk copy three nodes ahead into a
// four-stage ring (the round-1 kernel held three rows, 108 registers, in flight per lane).
// ---------------------------------------------------------------------------
template <int D, int LPC>
struct TileBwdSmem {
    static constexpr int W = 2 * D + 2;
    static constexpr int NST = 4;
    static constexpr int STG = W * 32;
    static constexpr size_t BYTES = (size_t)(NST * STG + NST) * sizeof(double);
};

template <int D, int LPC>
__global__ void __launch_bounds__(32) osc2_backward_tile_kernel(StepDev p)
{
    OSC2_DYN_SMEM_ALIGNED(sm_dyn);
    using SM = TileBwdSmem<D, LPC>;
    constexpr int CPW = 32 / LPC, W = SM::W, NST = SM::NST;
    const int N = p.N;
    const size_t B = (size_t)p.B;
    const int lane = threadIdx.x & 31;
    const int r = lane / CPW, bb = lane % CPW;
    const size_t tile = blockIdx.x;
    const long long b_ll = (long long)tile * CPW + bb;
    const bool act = (b_ll < (long long)p.B) && (r < D);
    const size_t b = act ? (size_t)b_ll : 0;
    const int lane0 = bb;                                 // row rr of this conductor sits at lane0 + rr * CPW
    double *ring = sm_dyn;
    unsigned long long *bars = (unsigned long long *)(sm_dyn + NST * SM::STG);
    const double *sptile = p.spill + tile * SM::STG;
    const size_t sstep = (size_t)p.NT * SM::STG;
    // i counts the nodes from the last one: node n = N - 1 - i lives in stage i % NST
    auto issue = [&](const int i) {
        if (lane == 0 && i < N) {
            const int s = i % NST;
            osc2_mbar_expect_tx(bars + s, (unsigned)(SM::STG * sizeof(double)));
            osc2_bulk_load(ring + (size_t)s * SM::STG, sptile + (size_t)(N - 1 - i) * sstep, (unsigned)(SM::STG * sizeof(double)), bars + s);
        }
    };
    if (lane == 0) {
        for (int s = 0; s < NST; ++s) osc2_mbar_init(bars + s, 1);
        osc2_mbar_init_fence();
    }
    __syncwarp();
    for (int i = 0; i < NST - 1; ++i) issue(i);

    int range_flag = 0;
    double f[W], xc[D], xn[D];
#pragma unroll
    for (int c = 0; c < D; ++c) { xc[c] = 0.0; xn[c] = 0.0; }
    for (int i = 0; i < N; ++i) {
        const int n = N - 1 - i;
        __syncwarp();                 // every lane has taken its row of node n+1 out of the stage that is refilled now
        issue(i + NST - 1);
        osc2_mbar_wait(bars + i % NST, (unsigned)((i / NST) & 1));
        {
            const double *sq = ring + (size_t)(i % NST) * SM::STG + lane;
#pragma unroll
            for (int c = 0; c < W; ++c) f[c] = sq[c * 32];
        }
        if (!act) {                   // idle lanes divide by 1 (no dynamic index: f[] must stay in registers)
#pragma unroll
            for (int c = 0; c < W; ++c) f[c] = (c == (r < D ? r : 0) || c == 2 * D + 1) ? 1.0 : 0.0;
        }
        // products with the already known node n+1 are off the dependency chain
        double pu[D];
        const bool hasU = n < N - 1;
#pragma unroll
        for (int c = 0; c < D; ++c) pu[c] = hasU ? f[D + c] * xn[c] : 0.0;
        double mine = 0.0;
#pragma unroll
        for (int rr = D - 1; rr >= 0; --rr) {
            double q = f[2 * D];
#pragma unroll
            for (int c = rr + 1; c < D; ++c) q = q - f[c] * xc[c];
            if (hasU) {
#pragma unroll
                for (int c = 0; c < D; ++c) q = q - pu[c];
            }
            int flag = 0;
            const double x = osc2_div_flag(q, osc2_rcp_make(f[rr], f[2 * D + 1]), flag);
            if (r == rr) { mine = x; range_flag |= flag; }
            xc[rr] = osc2_bcast(x, lane0 + rr * CPW);
        }
        if (act) p.sysvar_new[((size_t)n * D + r) * B + b] = mine;
#pragma unroll
        for (int c = 0; c < D; ++c) xn[c] = xc[c];
    }
    if (act && range_flag && p.status[b] == 0) p.status[b] = OSC2_STATUS_RANGE;
}
