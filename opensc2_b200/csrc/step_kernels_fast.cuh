// Register-resident fast path of the batched step() for small systems
// (NODOFS D known at compile time, D <= LPC lanes per conductor).
//
// Same arithmetic as step_kernels.cuh (and therefore as the reference, file
// and line references there); what changes is where the numbers live:
//   * K_G2 stores, per element and per matrix ROW, the element-matrix values
//     themselves (S*dz/3, K/dz, A/2, dz/3*M, dz/6*M, load halves), so the
//     sweep does no per-entry divisions and every lane reads one contiguous
//     record with compile-time offsets;
//   * K_F2 keeps the lane's block row [L D U | rhs] in registers, prefetches
//     the next element's record while the current node is eliminated, and
//     publishes finished pivot rows (with the pivot's reciprocal) through
//     shared memory;
//   * K_B2 keeps the lane's factor row in registers, prefetches the next
//     node's row, and exchanges the solved unknowns through shared memory.
#pragma once

#include <type_traits>

#include "step_kernels.cuh"
#include "osc2_div.cuh"

// Row record r of one element in `gm` (fast path): D + 9 fields.  The records are tiled by warp
// (osc2_dev.h: StepDev::cpw): field f of row r of conductor b lives at double
//   ((e * NT + b / cpw) * (D + 9) + f) * 32 + r * cpw + b % cpw
// the methods below return the offset inside the (element, tile) block, without the b % cpw part.
struct GmSlots2 {
    int d, M, S, cpw;
    __host__ __device__ int width() const { return d + 9; }
    __host__ __device__ int at(int r, int f) const { return f * 32 + r * cpw; }
    __host__ __device__ int sd(int r, int c) const { return at(r, c); }             // S[r][c]*dz/3
    __host__ __device__ int kb(int r) const { return at(r, d); }                    // K[r]/dz
    __host__ __device__ int md(int r) const { return at(r, d + 1); }                // dz/3*M[r]
    __host__ __device__ int mo(int r) const { return at(r, d + 2); }                // dz/6*M[r]
    __host__ __device__ int ahd(int r) const { return at(r, d + 3); }               // A[r][r]/2
    __host__ __device__ int aho(int r) const { return at(r, d + 4); }               // A[r][offcol(r)]/2
    __host__ __device__ int lod(int r, int which) const { return at(r, d + 5 + which); }
    __host__ __device__ int block() const { return (d + 9) * 32; }                  // doubles of one (element, tile)
};
// doubles of one (node, tile) block of `spill`
__host__ __device__ inline int osc2_spill_block(int d) { return (2 * d + 2) * 32; }
enum { LOD_UP = 0, LOD_LO = 1, LOD_UP_PREV = 2, LOD_LO_PREV = 3 };

// read-only input of this launch through the non-coherent path: such loads may be hoisted over the stores of the
// kernel and issued in batches (with plain loads every one of K_G's ~50 input loads was waited for on its own,
// ~660 cycles each: the compiler cannot prove that the record stores do not alias the inputs)
#ifdef OSC2_EMU
#define OSC2_LDG(ptr) (*(ptr))
#else
#define OSC2_LDG(ptr) __ldg(ptr)
#endif

// request the line that holds *ptr from DRAM into L2 (no register, no completion to wait for)
__device__ __forceinline__ void osc2_prefetch_l2(const void *ptr)
{
#ifndef OSC2_EMU
    asm volatile("prefetch.global.L2 [%0];" ::"l"(ptr));
#else
    (void)ptr;
#endif
}

// ---------------------------------------------------------------------------
// K_G2: same statement order as osc2_gauss_kernel for SMAT, then converts the
// Gauss-point values into the element-matrix values (smc:1015-1216).
// ---------------------------------------------------------------------------
// body for element e of conductor b; sm_dyn: s_slots * blockDim.x doubles of shared memory
// `g`: the record of (element e, conductor b) without the lane offset of the row -- in `gm` (K_G) or in a consumer's
// shared-memory ring (INREC: the producer warps of osc2_forward_fused_kernel); Sb, Ss: per-thread SMAT staging
// column and its stride (unused with INREC, where SMAT is accumulated in the record itself)
template <bool SMEM_S, bool INREC>
__device__ __forceinline__ void osc2_gauss2_body(const osc2_topology *__restrict__ tp, const StepDev &p, double *g, double *Sb,
                                                 const size_t Ss, const int e, const size_t b, const unsigned short *sinv_in = nullptr)
{
    const size_t B = (size_t)p.B;
    const int M = p.M, Sn = p.S, d = p.d, E = p.E;
    const GmSlots2 sl{d, M, Sn, p.cpw};
    static_assert(SMEM_S, "the fast path always stages SMAT in shared memory");
    const unsigned char *__restrict__ smap = p.smap;
    // `gm` was zeroed when the handle was created and nothing but this kernel writes it: fields that are +0 for
    // every element of this topology (SMAT entries no statement can touch, the load halves of the fluid rows,
    // the advection entries of the solid rows) are not written again -- 440 of the 1088 bytes of a CASE_1 record
    const bool skipz = INREC || p.gm_prezeroed != 0;
#define SM_(r, c) (*(INREC ? &g[(c) * 32 + (r) * sl.cpw] : &Sb[(size_t)smap[(r) * d + (c)] * Ss]))
#define FG(q, j) OSC2_LDG(p.fl_gauss + ((((size_t)(q) * M + (j)) * E) + e) * B + b)
#define GOUT(slot) g[slot]
    // The kernel is bound by load latency (ncu: 67 % of the stall samples wait for global loads that sit inside
    // loops with run-time trip counts, 6 warps per scheduler): every input of this (element, conductor) is
    // requested from L2 up front, the loads below then hit L2
    {
        const size_t EB = (size_t)E * B, eb = (size_t)e * B + b;
        // `n` lines, `stride` doubles apart (running pointers: a 64-bit multiply per address costs three instructions)
        auto run = [](const double *q0, int n, size_t stride) { for (int i = 0; i < n; ++i, q0 += stride) osc2_prefetch_l2(q0); };
        osc2_prefetch_l2(p.dz + eb);
        run(p.fl_gauss + eb, 9 * M, EB);
        run(p.peri_ff + eb, 2 * p.nff, EB); run(p.htc_ff + eb, 2 * p.nff, EB);
        run(p.peri_fs + eb, p.nfs, EB); run(p.htc_fs + eb, p.nfs, EB);
        run(p.peri_ss + eb, p.nss, EB); run(p.htc_ss + eb, p.nss, EB);
        run(p.peri_es + eb, p.nes, EB); run(p.htc_es + eb, p.nes, 3 * EB);
        run(p.sol_gauss + eb, 3 * Sn, EB);
        run(p.sol_q + (size_t)e * 2 * B + b, 2 * Sn, 2 * EB); run(p.sol_q + (size_t)e * 2 * B + B + b, 2 * Sn, 2 * EB);
        run(p.qsource + (size_t)e * Sn * B + b, 2 * Sn, B);
    }
    if (INREC) { for (int slot = 1; slot < p.s_slots; ++slot) g[p.sinv[slot]] = 0.0; }
    else for (int i = 0; i < p.s_slots; ++i) Sb[(size_t)i * Ss] = 0.0;
    const double dz = OSC2_LDG(p.dz + ((size_t)e * B + b));
    const double alpha = 0.0;
    const double cdiag = dz * (1. / 3. + alpha), coff = dz * (1. / 6. - alpha);

    for (int j = 0; j < M; ++j) {
        const int iv = j, ip = M + j, it = 2 * M + j;
        const double v = FG(FL_V, j), rho = FG(FL_RHO, j), c0 = FG(FL_C0, j);
        const double grun = FG(FL_GRUN, j);
        const double kd = dz * 1.0 * fabs(v) / 2.0;
        const double kb = kd / dz;
        const double ah_off[3] = {(1. / rho) / 2., (rho * (c0 * c0)) / 2., (grun * FG(FL_T, j)) / 2.};
        for (int q = 0; q < 3; ++q) {
            const int row = q * M + j;              // velocity, pressure, temperature equation
            GOUT(sl.ahd(row)) = v / 2.;
            GOUT(sl.aho(row)) = ah_off[q];
            GOUT(sl.kb(row)) = kb;
            GOUT(sl.md(row)) = cdiag;              // MMAT = 1 on the fluid equations
            GOUT(sl.mo(row)) = coff;
            if (!skipz) for (int w = 0; w < 4; ++w) GOUT(sl.lod(row, w)) = 0.0;
        }
        const double svv = 2.0 * FG(FL_F, j) * fabs(v) / tp->ch_dh[j];
        SM_(iv, iv) = svv;
        SM_(ip, iv) = -svv * grun * rho * v;
        SM_(it, iv) = -svv / FG(FL_CV, j) * v;
    }
    for (int k = 0; k < p.nff; ++k) {
        const int ca = tp->ff[k][0], cb = tp->ff[k][1];
        const double pa = FG(FL_P, ca), pb = FG(FL_P, cb);
        double delta_p = fabs(pa - pb);
        if (delta_p < tp->delta_p_min) delta_p = tp->delta_p_min;
        const int cu = (pb < pa) ? ca : cb;
        const double vel = FG(FL_V, cu);
        const double peri_o = OSC2_LDG(p.peri_ff + ((((size_t)0 * p.nff + k) * E + e) * B + b));
        const double peri_c = OSC2_LDG(p.peri_ff + ((((size_t)1 * p.nff + k) * E + e) * B + b));
        const double K1 = peri_o * sqrt(2. * FG(FL_RHO, cu) / (tp->k_loc * delta_p));
        const double K2 = K1 * tp->lambda_v * vel;
        const double vl = vel * tp->lambda_v;
        const double K3 = K1 * (FG(FL_H, cu) + .5 * (vl * vl));
        p.k123[(((size_t)0 * p.nff + k) * E + e) * B + b] = K1;
        p.k123[(((size_t)1 * p.nff + k) * E + e) * B + b] = K2;
        p.k123[(((size_t)2 * p.nff + k) * E + e) * B + b] = K3;
        const double coef_htc = peri_o * OSC2_LDG(p.htc_ff + ((((size_t)0 * p.nff + k) * E + e) * B + b))
                              + peri_c * OSC2_LDG(p.htc_ff + ((((size_t)1 * p.nff + k) * E + e) * B + b));
        for (int side = 0; side < 2; ++side) {
            const int c1 = side ? cb : ca, c2 = side ? ca : cb;
            const int v1 = c1, p1 = M + c1, t1 = 2 * M + c1, p2 = M + c2, t2 = 2 * M + c2;
            const double v = FG(FL_V, c1), rho = FG(FL_RHO, c1), A = tp->ch_area[c1];
            const double phi = FG(FL_GRUN, c1), h = FG(FL_H, c1), cv = FG(FL_CV, c1);
            const double c0 = FG(FL_C0, c1), T = FG(FL_T, c1);
            const double s_vj_pj = (K1 * v - K2) / (A * rho);
            SM_(v1, p1) -= s_vj_pj;
            SM_(v1, p2) = s_vj_pj;
            const double cga = phi / A;
            const double s_pj_pj = cga * (K3 - v * K2 - (h - (v * v) / 2. - (c0 * c0) / phi) * K1);
            SM_(p1, p1) += s_pj_pj;
            SM_(p1, p2) = -s_pj_pj;
            const double s_pj_tj = cga * coef_htc;
            SM_(p1, t1) += s_pj_tj;
            SM_(p1, t2) = -s_pj_tj;
            const double crca = 1. / (rho * cv * A);
            const double s_tj_pj = crca * (K3 - v * K2 - (h - (v * v) / 2. - phi * cv * T) * K1);
            SM_(t1, p1) += s_tj_pj;
            SM_(t1, p2) = -s_tj_pj;
            const double s_tj_tj = crca * coef_htc;
            SM_(t1, t1) += s_tj_tj;
            SM_(t1, t2) = -s_tj_tj;
        }
    }
    for (int k = 0; k < p.nfs; ++k) {
        const int j = tp->fs[k][0], l = tp->fs[k][1];
        const int ip = M + j, it = 2 * M + j, is = 3 * M + l;
        const double A = tp->ch_area[j];
        const double cga = FG(FL_GRUN, j) / A;
        const double coef_htc = OSC2_LDG(p.peri_fs + (((size_t)k * E + e) * B + b)) * OSC2_LDG(p.htc_fs + (((size_t)k * E + e) * B + b));
        const double s_pj_tj = cga * coef_htc;
        SM_(ip, it) += s_pj_tj;
        SM_(ip, is) = -s_pj_tj;
        const double crca = 1. / (FG(FL_RHO, j) * FG(FL_CV, j) * A);
        const double s_tj_tj = crca * coef_htc;
        SM_(it, it) += s_tj_tj;
        SM_(it, is) = -s_tj_tj;
        SM_(is, is) += coef_htc;
        SM_(is, it) = -coef_htc;
    }
    const bool first = (p.num_step == 1);
    const double dz_6 = dz / 6.;
    for (int k = 0; k < p.nss; ++k) {
        const int a = 3 * M + tp->ss[k][0], c = 3 * M + tp->ss[k][1];
        const double coef = OSC2_LDG(p.peri_ss + (((size_t)k * E + e) * B + b)) * OSC2_LDG(p.htc_ss + (((size_t)k * E + e) * B + b));
        SM_(a, a) += coef;
        SM_(a, c) = -coef;
        SM_(c, c) += coef;
        SM_(c, a) = -coef;
    }
    // per-solid env heat (added to SVEC after build_svec, smc:950-1013); only one
    // environment interface per solid exists in valid decks, but keep the order
    for (int l = 0; l < Sn; ++l) {
        const int is = 3 * M + l;
        const double A = tp->sol_area[l], cs = tp->sol_costeta[l];
        const double rho = OSC2_LDG(p.sol_gauss + ((((size_t)0 * Sn + l) * E + e) * B + b));
        const double cp = OSC2_LDG(p.sol_gauss + ((((size_t)1 * Sn + l) * E + e) * B + b));
        const double kk = OSC2_LDG(p.sol_gauss + ((((size_t)2 * Sn + l) * E + e) * B + b));
        const double mm = A * rho * cp / cs;
        GOUT(sl.md(is)) = cdiag * mm;
        GOUT(sl.mo(is)) = coff * mm;
        GOUT(sl.kb(is)) = (A * kk / cs) / dz;
        if (!skipz) { GOUT(sl.ahd(is)) = 0.0; GOUT(sl.aho(is)) = 0.0; }
        const size_t q1 = ((((size_t)0 * Sn + l) * E + e) * 2) * B + b;
        const size_t q2 = ((((size_t)1 * Sn + l) * E + e) * 2) * B + b;
        const double qs0 = OSC2_LDG(p.qsource + (((size_t)e * Sn + l) * B + b));
        const double qs1 = OSC2_LDG(p.qsource + (((size_t)(e + 1) * Sn + l) * B + b));
        double sv0 = OSC2_LDG(p.sol_q + (q1)) - qs0, sv1 = OSC2_LDG(p.sol_q + (q2)) - qs1;
        double sp0 = first ? OSC2_LDG(p.sol_q + (q1 + B)) - qs0 : 0.0, sp1 = first ? OSC2_LDG(p.sol_q + (q2 + B)) - qs1 : 0.0;
        for (int k = 0; k < p.nes; ++k) {
            if (tp->es[k] != l) continue;
            const double h0 = OSC2_LDG(p.htc_es + ((((size_t)k * 3 + 0) * E + e) * B + b));
            const double peri = OSC2_LDG(p.peri_es + (((size_t)k * E + e) * B + b));
            const bool rect = tp->es_choice2[k] && tp->is_rectangular;
            double coef_htc;
            if (rect) {
                const double h1 = OSC2_LDG(p.htc_es + ((((size_t)k * 3 + 1) * E + e) * B + b));
                const double h2 = OSC2_LDG(p.htc_es + ((((size_t)k * 3 + 2) * E + e) * B + b));
                coef_htc = 2. * tp->height * h0 + tp->width * (h1 + h2);
            } else {
                coef_htc = h0 * peri;
            }
            SM_(is, is) += coef_htc;
            if (!rect) {
                const double env_heat = (peri * h0) * p.t_env;
                sv0 += env_heat; sv1 += env_heat;
                if (first) { sp0 += env_heat; sp1 += env_heat; }
            }
        }
        GOUT(sl.lod(is, LOD_UP)) = (2. * sv0 + sv1) * dz_6;
        GOUT(sl.lod(is, LOD_LO)) = (sv0 + 2. * sv1) * dz_6;
        GOUT(sl.lod(is, LOD_UP_PREV)) = (2. * sp0 + sp1) * dz_6;
        GOUT(sl.lod(is, LOD_LO_PREV)) = (sp0 + 2. * sp1) * dz_6;
    }
    // S * dz / 3 (smc:1126-1163): 3 is a shared divisor, so its reciprocal is refined once (bit-identical
    // quotients, osc2_div.cuh); entries the topology cannot touch are +0 and need no arithmetic at all
    const Osc2Rcp r3 = osc2_rcp_prepare(3.0);
    if (skipz) {
        // slot -> offset of S[r][c] in the record, for the entries the topology can touch (osc2_build_smap, inverted
        // on the host: a division by the run-time d here cost 27 instructions per entry)
        const unsigned short *__restrict__ sinv = sinv_in ? sinv_in : p.sinv;
        if (INREC) { for (int slot = 1; slot < p.s_slots; ++slot) { double *q = g + sinv[slot]; *q = osc2_div(*q * dz, r3); } }
        else for (int slot = 1; slot < p.s_slots; ++slot) GOUT(sinv[slot]) = osc2_div(Sb[(size_t)slot * Ss] * dz, r3);
    } else {
        for (int r = 0; r < d; ++r)
            for (int c = 0; c < d; ++c) {
                const int slot = smap[r * d + c];
                GOUT(sl.sd(r, c)) = slot ? osc2_div(Sb[(size_t)slot * Ss] * dz, r3) : 0.0;
            }
    }
#undef SM_
#undef FG
#undef GOUT
}

template <bool SMEM_S>
__global__ void osc2_gauss2_kernel(const osc2_topology *__restrict__ tp, StepDev p)
{
    OSC2_DYN_SMEM(sm_dyn);
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    // slot -> record offset table next to the staging columns (its loads from global memory, one per touched entry in
    // the final pass and each waited for, were 17 % of the kernel's stall samples)
    unsigned short *sinv_s = (unsigned short *)(sm_dyn + (size_t)p.s_slots * blockDim.x);
    if (p.gm_prezeroed) {
        for (int i = threadIdx.x; i < p.s_slots; i += blockDim.x) sinv_s[i] = p.sinv[i];
        __syncthreads();
    }
    if (idx >= (long long)p.E * p.B) return;
    const int e = (int)(idx / p.B);
    const size_t b = (size_t)(idx % p.B);
    double *g = p.gm + ((size_t)e * p.NT + b / p.cpw) * ((size_t)(p.d + 9) * 32) + b % p.cpw;
    osc2_gauss2_body<SMEM_S, false>(tp, p, g, sm_dyn + threadIdx.x, blockDim.x, e, b, p.gm_prezeroed ? sinv_s : nullptr);
}

// max(m, |v|) as one compare and one select (a NaN never wins, as in `if (|v| > m) m = |v|`)
__device__ __forceinline__ double osc2_absmax(double m, double v)
{
    const double a = fabs(v);
    return (a > m) ? a : m;
}

// shared-memory doubles per conductor of the fast forward sweep
template <int D, bool CLOSED>       // CLOSED: closed-form Known, no old-solution ring and no product strips
struct FastFwdSmem {
    static constexpr int PW = 2 * D + 2;                 // published row: D' | U | rhs | 1/pivot
    static constexpr int PUB = 2 * D * PW;               // two nodes (double buffer)
    static constexpr int XS = 3 * D;                     // old solution ring
    static constexpr int STRIP = 5 * D + 1;              // per-lane product strip with zero pads
    // per-lane landing slot of the asynchronous prefetch (element record, next old solution value,
    // SYSLOD row); stride == 2 (mod 4) doubles keeps the 16-byte reads of eight lanes conflict-free
    static constexpr int PF_FIELDS = D + 12;
    static constexpr int PF = PF_FIELDS + ((6 - (PF_FIELDS % 4)) % 4);
    static constexpr int PRE = D * PF;
    static constexpr int RAW = PRE + PUB + (CLOSED ? 0 : XS + D * STRIP);
    static constexpr int PER_COND = RAW + ((18 - (RAW % 16)) % 16);   // == 2 (mod 16): bank spread
};

// 8-byte global -> shared copy that occupies no register while in flight (LDGSTS)
__device__ __forceinline__ void osc2_async_copy8(double *dst_shared, const double *src_global)
{
#ifdef OSC2_EMU
    *dst_shared = *src_global;
#else
    const unsigned d = (unsigned)__cvta_generic_to_shared(dst_shared);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(d), "l"(src_global) : "memory");
#endif
}

__device__ __forceinline__ void osc2_async_wait_all()
{
#ifndef OSC2_EMU
    asm volatile("cp.async.wait_all;" ::: "memory");
#endif
}

// ---------------------------------------------------------------------------
// K_F2.  BE: theta == 1 (Backward Euler; the Known coefficients off the block
// diagonals vanish and, for D % 8 == 0, numpy's pairwise sum of the three
// surviving products is (pL + pD) + pU for every row).  CAPTURE: also write
// the pre-solve system in the reference band layout.
//
// Lane = matrix row.  A row of the three blocks differs from "S*dz/3 halves" in
// two columns only -- its own (mass, advection, diffusion) and, for fluid rows,
// the partner column of the AMAT entry -- so those two entries are formed as
// scalars and every other column costs one multiply or add and two selects.
// The element records alternate between two register sets (nodes are processed
// in pairs), the next element is fetched straight into the set that died.
// ---------------------------------------------------------------------------
template <int D, int LPC, bool BE, bool CAPTURE, bool FULL = false>
__global__ void osc2_forward_fast_kernel(const osc2_topology *__restrict__ tp, StepDev p)
{
    static_assert(!FULL || LPC == D, "FULL: every lane owns a row");
    OSC2_DYN_SMEM(sm_dyn);
    constexpr bool CLOSED_KNOWN = BE && (D % 8 == 0);
    using SM = FastFwdSmem<D, CLOSED_KNOWN>;
    constexpr int PW = SM::PW, STRIP = SM::STRIP, RW = D + 9;
    const int M = p.M, N = p.N;
    const size_t B = (size_t)p.B;
    // field k of a batch-innermost record: ONE multiply-add on the 32-bit batch size, so that the
    // compiler has no 64-bit k * B to hoist out of the node loop (dozens of registers otherwise)
    const unsigned Bu = (unsigned)p.B;
    auto fld = [Bu](auto *ptr, const int k) {
        return (decltype(ptr))((const char *)ptr + (unsigned long long)Bu * (unsigned)(8 * k));
    };
    // lane = row * CPW + conductor of the tile (the layout of the tiled records, osc2_dev.h)
    constexpr int CPW = 32 / LPC;
    const int cpb = (blockDim.x / 32) * CPW;
    const int lane_ = threadIdx.x & 31;
    const int grp = (threadIdx.x >> 5) * CPW + lane_ % CPW, r = lane_ / CPW;
    const long long b_ll = (long long)blockIdx.x * cpb + grp;
    const bool act = FULL || ((b_ll < (long long)p.B) && (r < D));
    const size_t b = act ? (size_t)b_ll : 0;
    const int rr_ = (FULL || r < D) ? r : 0;
    const size_t tile_ = (size_t)blockIdx.x * (blockDim.x / 32) + (threadIdx.x >> 5);    // < NT by the launch geometry
    auto tfld = [](auto *ptr, const int k) { return ptr + 32 * k; };                     // field k of a tiled record
    const long long Total = (long long)D * N;
    constexpr int Half = 2 * D;
    const double theta = tp->theta, omt = 1.0 - tp->theta;
    const Osc2Rcp rdt = osc2_rcp_prepare(p.dt);
    const bool first = (p.num_step == 1);
    const bool shift = (p.num_step > 1);
    // divisor range: dt here, the pivots at the end of every node (max of hi(|b|) - hi(2^-100))
    unsigned div_span = osc2_abs_hi(p.dt) - OSC2_HI_DIV_LO;

    double *base = sm_dyn + (size_t)grp * SM::PER_COND;
    double *pre = base + (size_t)rr_ * SM::PF;           // this lane's prefetch slot
    double *pub = base + SM::PRE;                        // [2][D][PW]
    double *xs = pub + SM::PUB;                          // [3][D]   (general Known only)
    double *strip = xs + SM::XS + (size_t)rr_ * STRIP;   // [5D+1], products live at [D, 4D)

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

    // This lane's row record of one element, plus the two entries of its S row that meet the
    // special columns (fetched a second time with a lane-dependent offset).  The record of the
    // next element, the next old-solution value and the next SYSLOD row travel global -> shared
    // asynchronously while the current node is eliminated: a register-held prefetch would be
    // spilled at 255 registers, and a spilled load target stalls the warp for the DRAM latency.
    struct Elem {
        double sd[D];
        double sdr, sdo, kb, md, mo, ahd, aho, lodu, lodl;
    };
    enum { PF_SDR = D, PF_SDO, PF_KB, PF_MD, PF_MO, PF_AHD, PF_AHO, PF_LODU, PF_LODL, PF_X, PF_SL0, PF_SL1 };
    static_assert(PF_SL1 + 1 == SM::PF_FIELDS, "prefetch slot layout");
    const size_t gstep = (size_t)p.NT * RW * 32;                 // one element further
    const double *gptr = p.gm + tile_ * (RW * 32) + rr_ * CPW + lane_ % CPW;    // row record of element 0
    const size_t off_r = (size_t)rr_ * 32, off_o = (size_t)(aoffcol >= 0 ? aoffcol : 0) * 32;
    const size_t xstep = (size_t)D * B;
    const double *xptr = p.sysvar + (size_t)rr_ * B + b;                 // x[node][r]
    double *lodptr = p.syslod + (size_t)rr_ * 2 * B + b;                 // SYSLOD[node][r][0]
    double *spptr = p.spill + tile_ * (PW * 32) + rr_ * CPW + lane_ % CPW;      // spill[node][tile][0][lane]
    // start the copies node m needs: element m, x[m+1], SYSLOD[m] (g, xq, lq point at them)
    auto prefetch = [&](const int m, const double *g, const double *xq, const double *lq) {
        if (m < N - 1) {
#pragma unroll
            for (int c = 0; c < D; ++c) osc2_async_copy8(pre + c, tfld(g, c));
            osc2_async_copy8(pre + PF_SDR, g + off_r);
            osc2_async_copy8(pre + PF_SDO, g + off_o);
#pragma unroll
            for (int q = 0; q < 7; ++q) osc2_async_copy8(pre + PF_KB + q, tfld(g, D + q));
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

    Elem eA, eB;
    zero_elem(eA); zero_elem(eB);
    double xm = 0.0, x0 = 0.0, xp = 0.0;                                 // own old solution at nodes n-1, n, n+1
    double mas_carry = 0.0;                                              // MASMAT/dt of the element to the left
    double x_pre = 0.0, sl0_pre = 0.0, sl1_pre = 0.0;
    auto load_elem = [&](const double *g, Elem &el) {
#pragma unroll
        for (int c = 0; c < D; ++c) el.sd[c] = *tfld(g, c);
        el.sdr = g[off_r];
        el.sdo = g[off_o];
        el.kb = *tfld(g, D + 0); el.md = *tfld(g, D + 1); el.mo = *tfld(g, D + 2); el.ahd = *tfld(g, D + 3);
        el.aho = *tfld(g, D + 4); el.lodu = *tfld(g, D + 5); el.lodl = *tfld(g, D + 6);
    };
    if (act) {
        load_elem(gptr, eB);
        x0 = xptr[0];
        x_pre = xptr[xstep];
        sl0_pre = lodptr[0];
        sl1_pre = *fld(lodptr, 1);
        if (!CLOSED_KNOWN) xs[0 * D + r] = x0;
    }
    if (!CLOSED_KNOWN) { if (r < D) for (int q = 0; q < STRIP; ++q) strip[q] = 0.0; }   // zero pads, once
    double *xa = xs, *xb = xs + D, *xc = xs + 2 * D;             // nodes n-1, n, n+1 (rotating)
    {   // at n = 0: node -1 does not exist; ring holds node 0 in slot 0, node 1 arrives in slot 1
        double *t = xa; xa = xc; xc = xb; xb = t;                 // xb = slot0 (node 0), xc = slot1 (node 1)
    }

    // One node of the sweep: A is the element to its left, B the one to its right.  The first
    // and the last node are compiled separately (no left / right element, boundary rows
    // possible), so the interior nodes have neither of those branches.
    auto node = [&](auto HL, auto HR, const int n, Elem &A, Elem &Bv) {
        constexpr bool hasL = decltype(HL)::value, hasR = decltype(HR)::value;
        double *pub_cur = pub + (size_t)(n & 1) * D * PW;
        const double *pub_prev = pub + (size_t)((n + 1) & 1) * D * PW;
        const long long i_row = (long long)n * D + r;
        double sl0 = 0.0, sl1 = 0.0;
        // ---- prefetched by the previous node: element n (in Bv), x[n+1], SYSLOD row n ----------
        if (act) {
            if (hasR) { xp = x_pre; if (!CLOSED_KNOWN) xc[r] = xp; }
            sl0 = sl0_pre; sl1 = sl1_pre;
            if (n + 2 < N) x_pre = xptr[2 * xstep];
            if (n + 1 < N) { sl0_pre = lodptr[2 * xstep]; sl1_pre = *fld(lodptr + 2 * xstep, 1); }
        }
        if (!CLOSED_KNOWN) __syncwarp();   // old-solution ring of this node visible

        double rowL[D], rowD[D], rowU[D], rhs = 0.0;
        if (act) {
            // SYSLOD (smc:1165-1318); fluid rows carry zero loads
            if (shift) { sl1 = sl0; sl0 = 0.0; }
            // (the previous-step halves exist at the first step only: read in place, not prefetched)
            if (hasL) { sl0 += A.lodl; if (first) sl1 += *tfld(gptr - gstep, D + 8); }
            if (hasR) { sl0 += Bv.lodu; if (first) sl1 += *tfld(gptr, D + 7); }
            lodptr[0] = sl0;
            *fld(lodptr, 1) = sl1;

            // MASMAT/dt is non-zero on the diagonal entry of each block only; the U-block
            // value of this node is the L-block value of the next one
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
                // FDS = FLX + DIF + SOR (smc:1015-1163) of the two special columns, same operation
                // order as the reference; adding the zeros of the other columns changes no bit
                //   L block: lower-left of element n-1;  U block: upper-right of element n
                //   D block: lower-right of element n-1 + upper-left of element n
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
                if (CLOSED_KNOWN) {
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
                } else {
#pragma unroll
                    for (int c = 0; c < D; ++c) {
                        const bool dg = (c == r), of = (c == aoffcol);
                        const double gD = (hasL && hasR) ? A.sd[c] + Bv.sd[c] : (hasL ? A.sd[c] : Bv.sd[c]);
                        {
                            double sys = 0.0;
                            if (hasL) {
                                const double fds = dg ? fLd : (of ? fLo : A.sd[c] / 2.);
                                const double m = dg ? masL : 0.0;
                                sys = BE ? m + fds : m + theta * fds;
                                strip[D + c] = (m - omt * fds) * xa[c];
                            } else strip[D + c] = 0.0;
                            rowL[c] = sys;
                        }
                        {
                            const double fds = dg ? fDd : (of ? fDo : gD);
                            const double m = dg ? masD : 0.0;
                            rowD[c] = BE ? m + fds : m + theta * fds;
                            strip[2 * D + c] = (m - omt * fds) * xb[c];
                        }
                        {
                            double sys = 0.0;
                            if (hasR) {
                                const double fds = dg ? fUd : (of ? fUo : Bv.sd[c] / 2.);
                                const double m = dg ? masU : 0.0;
                                sys = BE ? m + fds : m + theta * fds;
                                strip[3 * D + c] = (m - omt * fds) * xc[c];
                            } else strip[3 * D + c] = 0.0;
                            rowU[c] = sys;
                        }
                        if (hasL) mxL = osc2_absmax(mxL, rowL[c]);
                        mxD = osc2_absmax(mxD, rowD[c]);
                        if (hasR) mxU = osc2_absmax(mxU, rowU[c]);
                    }
                }
                const double mx = osc2_absmax(osc2_absmax(mxL, mxD), mxU);
                // Known (smc:1366-1443)
                double known;
                if (CLOSED_KNOWN) {
                    // only MASMAT/dt * x_old survives; the three products sit 8k band
                    // positions apart, i.e. in one numpy accumulator (or its tail)
                    known = 0.0;
                    if (hasL) known = masL * xm;
                    known = hasL ? known + masD * x0 : masD * x0;
                    if (hasR) known = known + masU * xp;
                } else {
                    long long lo, hi;
                    if (i_row <= Half - 1) { lo = 0; hi = Half + i_row; }
                    else if (i_row >= Total - (Half - 1)) { lo = i_row - (Half - 1); hi = Total; }
                    else { lo = i_row - (Half - 1); hi = i_row + Half; }
                    const int Lfull = (int)(hi - lo);
                    // band position jj <-> strip index D + (lo + jj - (n-1)*D)
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
                // row scaling (tsf:538-551)
                // one range check for the whole row (runs beside the Newton steps of 1/mx): a numerator
                // or the divisor outside the exact range poisons the reciprocal, hence every quotient
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

            // pivots of node n-1 (gredub / gbacsb forward part)
            if (hasL) {
#pragma unroll
                for (int k = 0; k < D; ++k) {
                    const double *pk = pub_prev + k * PW;
                    const double q = osc2_div_p(rowL[k], osc2_rcp_make(pk[k], pk[2 * D + 1]));
#pragma unroll
                    for (int c = k + 1; c < D; ++c) rowL[c] = rowL[c] - pk[c] * q;
#pragma unroll
                    for (int c = 0; c < D; ++c) rowD[c] = rowD[c] - pk[D + c] * q;
                    rhs = rhs - pk[2 * D] * q;
                }
            }
            // element n-1 and the L block are dead: fetch element n+1 into the left element's registers
            if (hasR && n + 1 < N - 1) load_elem(gptr + gstep, A);
        } else {
#pragma unroll
            for (int c = 0; c < D; ++c) { rowL[c] = 0.0; rowD[c] = 0.0; rowU[c] = 0.0; }
        }
        // pivots inside node n.  The reciprocal of pivot k+1 is started by EVERY lane on its own
        // rowD[k+1] as soon as pivot step k has updated that entry (only lane k+1's is meaningful):
        // outside the divergent publish block its seed + five dependent DFMA interleave with the other
        // sixteen updates of step k instead of holding the whole warp at the head of step k+1
        Osc2Rcp my_rcp = osc2_rcp_make(1.0, 1.0);
        Osc2Rcp nxt = osc2_rcp_prepare(rowD[0]);
#pragma unroll
        for (int k = 0; k < D; ++k) {
            if (act && r == k) {
                double *pk = pub_cur + k * PW;
#pragma unroll
                for (int c = k; c < D; ++c) pk[c] = rowD[c];
#pragma unroll
                for (int c = 0; c < D; ++c) pk[D + c] = rowU[c];
                pk[2 * D] = rhs;
                my_rcp = nxt;
                pk[2 * D + 1] = nxt.y;
            }
            __syncwarp();
            {
                // rows above the pivot (and idle lanes) take the multiplier 0: x - p * 0 leaves x unchanged
                // bit for bit, and ONE select on the multiplier replaces a select per updated entry
                const double *pk = pub_cur + k * PW;
                const bool on = act && r > k;
                double q = osc2_div_p(rowD[k], osc2_rcp_make(pk[k], pk[2 * D + 1]));
                q = on ? q : 0.0;
                if (k + 1 < D) {
                    rowD[k + 1] = rowD[k + 1] - pk[k + 1] * q;
                    nxt = osc2_rcp_prepare(rowD[k + 1]);
                }
#pragma unroll
                for (int c = k + 2; c < D; ++c) rowD[c] = rowD[c] - pk[c] * q;
#pragma unroll
                for (int c = 0; c < D; ++c) rowU[c] = rowU[c] - pk[D + c] * q;
                rhs = rhs - pk[2 * D] * q;
            }
        }
        if (act) {
            // this lane's pivot: singular (tsf:676-682) or outside the exact range of the division
            const bool tiny = fabs(my_rcp.b) <= OSC2_TINY_PIVOT;
            if (tiny && bad_row < 0) bad_row = i_row;
            div_span = osc2_umax(div_span, tiny ? 0u : osc2_abs_hi(my_rcp.b) - OSC2_HI_DIV_LO);
#pragma unroll
            for (int c = 0; c < D; ++c) *tfld(spptr, c) = rowD[c];
#pragma unroll
            for (int c = 0; c < D; ++c) *tfld(spptr, D + c) = rowU[c];
            *tfld(spptr, 2 * D) = rhs;
            *tfld(spptr, 2 * D + 1) = my_rcp.y;
            xm = x0; x0 = xp;       // rotate the old solution
            if (hasR) { const Elem t = A; A = Bv; Bv = t; }   // left <- right, right <- the prefetched element
        }
        { double *t = xa; xa = xb; xb = xc; xc = t; }
        gptr += gstep;
        xptr += xstep;
        lodptr += 2 * xstep;
        spptr += (size_t)p.NT * (PW * 32);
    };
    // (one body for all interior nodes: a two-node body that alternates the register sets saves the
    //  copy of the element record but doubles the loop to 50 KB of code, and instruction-cache
    //  misses then cost more than the copy)
    node(std::false_type(), std::true_type(), 0, eA, eB);
    for (int n = 1; n < N - 1; ++n) node(std::true_type(), std::true_type(), n, eA, eB);
    node(std::true_type(), std::false_type(), N - 1, eA, eB);
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
        // sticky: the first failing step is kept (osc2_upload_state starts over)
        if ((worst >= 0 || range) && p.status[b] == 0) p.status[b] = (worst >= 0) ? -(int)(worst + 1) : OSC2_STATUS_RANGE;
    }
}

// ---------------------------------------------------------------------------
// K_B2: back substitution (gbacsb, tsf:772-802), node N-1 down to 0.  Every lane keeps the
// solved unknowns of the current and of the next node in registers; a freshly solved x travels
// to the other lanes of its conductor by warp shuffle, so the serial chain of a node
// (the subtractions of one row must follow the reference's column order) carries no
// shared-memory round trip.  All lanes run the same arithmetic on their own factor row; only
// lane rr's quotient of turn rr is kept.
// ---------------------------------------------------------------------------
#ifdef OSC2_EMU
inline double osc2_bcast(double v, int src_lane) { return emu_shfl(v, src_lane); }
#else
__device__ __forceinline__ double osc2_bcast(double v, int src_lane) { return __shfl_sync(0xffffffffu, v, src_lane); }
#endif

template <int D, int LPC>
__global__ void osc2_backward_fast_kernel(StepDev p)
{
    const int N = p.N;
    const size_t B = (size_t)p.B;
    constexpr int CPW = 32 / LPC;
    const int cpb = (blockDim.x / 32) * CPW;
    const int lane_ = threadIdx.x & 31;
    const int grp = (threadIdx.x >> 5) * CPW + lane_ % CPW, r = lane_ / CPW;    // lane = row * CPW + conductor of the tile
    const long long b_ll = (long long)blockIdx.x * cpb + grp;
    const bool act = (b_ll < (long long)p.B) && (r < D);
    const size_t b = act ? (size_t)b_ll : 0;
    const int lane0 = lane_ % CPW;                       // lane of row 0 of this conductor; row rr sits at lane0 + rr * CPW

    constexpr int W = 2 * D + 2;      // D' | U | rhs | 1/pivot
    const size_t tile_ = (size_t)blockIdx.x * (blockDim.x / 32) + (threadIdx.x >> 5);
    const size_t sstep = (size_t)p.NT * (W * 32);        // one node further in the tiled spill
    const double *spbase = p.spill + tile_ * (W * 32) + lane_;
    int range_flag = 0;
    // factor rows of the current node and of the two nodes below it: a node takes ~2500 cycles, about
    // one DRAM round trip under load, so a one-node prefetch still left 0.6 of them exposed (ncu)
    double f[W], fn[W], fnn[W], xc[D], xn[D];
#pragma unroll
    for (int c = 0; c < W; ++c) { f[c] = 0.0; fn[c] = 0.0; fnn[c] = 0.0; }
#pragma unroll
    for (int c = 0; c < D; ++c) { xc[c] = 0.0; xn[c] = 0.0; }
    if (act) {
        const double *sp = spbase + (size_t)(N - 1) * sstep;
#pragma unroll
        for (int c = 0; c < W; ++c) fn[c] = sp[32 * c];
        if (N > 1) {
            const double *sq = spbase + (size_t)(N - 2) * sstep;
#pragma unroll
            for (int c = 0; c < W; ++c) fnn[c] = sq[32 * c];
        }
    } else {
        fn[r < D ? r : 0] = 1.0;      // idle lanes divide by 1
        fn[2 * D + 1] = 1.0;
#pragma unroll
        for (int c = 0; c < W; ++c) fnn[c] = fn[c];
    }
    for (int n = N - 1; n >= 0; --n) {
#pragma unroll
        for (int c = 0; c < W; ++c) { f[c] = fn[c]; fn[c] = fnn[c]; }
        if (act && n > 1) {
            const double *sp = spbase + (size_t)(n - 2) * sstep;
#pragma unroll
            for (int c = 0; c < W; ++c) fnn[c] = sp[32 * c];
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
