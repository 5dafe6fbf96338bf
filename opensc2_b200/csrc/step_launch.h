// Launch geometry of one batched step(), shared by the CUDA library and by
// the CPU thread emulator used in the kernel-logic tests (tests/cuda_emu).
//
// `Launcher` provides
//   template <class K, class... A>
//   void launch(int kernel_id, K kernel, unsigned grid, unsigned block, size_t smem, A... args);
//   size_t max_smem() const;     // opt-in dynamic shared memory per CTA
//   int sm_count() const;
#pragma once

#include "step_kernels.cuh"
#include "step_kernels_fast.cuh"
#include "step_kernels_tile.cuh"
#ifdef OSC2_WITH_FUSED_GF          // experiments/step_kernels_fused.cuh: K_G inside K_F (bit-exact, 3x slower: DESIGN.md 5)
#include "../../experiments/step_kernels_fused.cuh"
#endif
#ifdef OSC2_WITH_SPLIT_ROLES       // experiments/step_kernels_split_roles.cuh: assembler + eliminator warps per tile
#include "../../experiments/step_kernels_split_roles.cuh"
#endif
#include "step_kernels_d64.cuh"

#include <cstdlib>
#include <vector>

// numpy pairwise summation (loops_utils.h.src, PW_BLOCKSIZE = 128) of n values:
// leaves (start, length) in order and the postfix program of its additions.
inline void osc2_build_norm_plan(long long start, long long n, std::vector<int> &leaves, std::vector<int> &prog)
{
    if (n <= 128) {
        prog.push_back((int)(leaves.size() / 2));
        leaves.push_back((int)start);
        leaves.push_back((int)n);
        return;
    }
    long long n2 = n / 2;
    n2 -= n2 % 8;
    osc2_build_norm_plan(start, n2, leaves, prog);
    osc2_build_norm_plan(start + n2, n - n2, leaves, prog);
    prog.push_back(-1);
}

// Sparsity pattern of the source Jacobian SMAT for a topology: exactly the (row, col) pairs the
// statements of K_G can write (smc:262-311, 440-555, 596-668, 786-886).  Slot 0 is the shared zero.
inline int osc2_build_smap(const osc2_topology &t, std::vector<unsigned char> &map)
{
    const int M = t.n_ch, d = 3 * M + t.n_sol;
    std::vector<int> slot((size_t)d * d, 0);
    int next = 1;
    auto add = [&](int r, int c) { if (!slot[(size_t)r * d + c]) slot[(size_t)r * d + c] = next++; };
    for (int j = 0; j < M; ++j) { add(j, j); add(M + j, j); add(2 * M + j, j); }
    for (int k = 0; k < t.nff; ++k)
        for (int side = 0; side < 2; ++side) {
            const int c1 = t.ff[k][side], c2 = t.ff[k][1 - side];
            const int v1 = c1, p1 = M + c1, t1 = 2 * M + c1, p2 = M + c2, t2 = 2 * M + c2;
            add(v1, p1); add(v1, p2); add(p1, p1); add(p1, p2); add(p1, t1); add(p1, t2);
            add(t1, p1); add(t1, p2); add(t1, t1); add(t1, t2);
        }
    for (int k = 0; k < t.nfs; ++k) {
        const int j = t.fs[k][0], ip = M + j, it = 2 * M + j, is = 3 * M + t.fs[k][1];
        add(ip, it); add(ip, is); add(it, it); add(it, is); add(is, is); add(is, it);
    }
    for (int k = 0; k < t.nss; ++k) {
        const int a = 3 * M + t.ss[k][0], c = 3 * M + t.ss[k][1];
        add(a, a); add(a, c); add(c, c); add(c, a);
    }
    for (int k = 0; k < t.nes; ++k) add(3 * M + t.es[k], 3 * M + t.es[k]);
    if (next > 255) return -1;                 // does not fit a byte: the caller keeps the dense staging
    map.resize((size_t)d * d);
    for (size_t i = 0; i < map.size(); ++i) map[i] = (unsigned char)slot[i];
    return next;
}

struct Osc2LaunchPlan {
    int lpc;            // lanes per conductor (8, 16, 32, 64)
    int cpb_fwd;        // conductors per CTA, forward sweep
    int cpb_bwd;        // conductors per CTA, backward sweep
    size_t smem_fwd, smem_bwd;
    int gauss_block;
    size_t smem_gauss;  // 0: SMAT staged in the output record
};

// NODOFS values with a register-resident instantiation (CASE_1: 8, CASE_3: 13).
inline bool osc2_has_fast_path(int d) { return d == 4 || d == 8 || d == 13; }

// number of doubles of one element record in `gm` per conductor (generic path; the fast path tiles, below)
inline int osc2_gm_slots(int d, int M, int S, bool fast)
{
    if (fast) return d * (d + 9);
    const GmSlots s{d, M, S};
    return s.count();
}

// fast path: conductors per warp of the sweeps, tiles of the batch, sizes of the tiled buffers (osc2_dev.h)
inline int osc2_fast_cpw(int d) { return 32 / (d <= 8 ? 8 : 16); }
inline long long osc2_fast_tiles(int d, long long B) { const int c = osc2_fast_cpw(d); return (B + c - 1) / c; }
inline size_t osc2_gm_doubles(int d, int M, int S, bool fast, size_t E, size_t B)
{
    if (fast) return E * (size_t)osc2_fast_tiles(d, (long long)B) * (size_t)(d + 9) * 32;
    return E * (size_t)osc2_gm_slots(d, M, S, false) * B;
}
inline size_t osc2_spill_doubles(int d, bool fast, size_t N, size_t B)
{
    if (fast) return N * (size_t)osc2_fast_tiles(d, (long long)B) * (size_t)(2 * d + 2) * 32;
    return N * (size_t)d * (2 * d + 2) * B;
}

inline int osc2_pick_lpc(int d)
{
    if (d <= 8) return 8;
    if (d <= 16) return 16;
    if (d <= 32) return 32;
    return 64;
}

inline Osc2LaunchPlan osc2_make_plan(int d, int batch, size_t max_smem, int sm_count)
{
    Osc2LaunchPlan pl;
    pl.lpc = osc2_pick_lpc(d);
    const size_t per_f = (size_t)osc2_fwd_smem_per_conductor(d) * sizeof(double);
    const size_t per_b = (size_t)osc2_bwd_smem_per_conductor(d) * sizeof(double);
    // one warp per CTA for small systems keeps the grid wide (a batch of a few
    // thousand conductors only fills a few warps per SM); 64-lane groups are
    // one conductor per CTA because they synchronise with __syncthreads.
    int cpb = (pl.lpc <= 32) ? 32 / pl.lpc : 1;
    while (cpb > 1 && per_f * cpb > max_smem) cpb /= 2;
    pl.cpb_fwd = cpb;
    pl.cpb_bwd = cpb;
    pl.smem_fwd = per_f * pl.cpb_fwd;
    pl.smem_bwd = per_b * pl.cpb_bwd;
    // Gauss kernel: stage SMAT in shared memory when a useful block fits.
    pl.gauss_block = 128;
    const size_t per_thread = (size_t)d * d * sizeof(double);
    while (pl.gauss_block > 32 && per_thread * pl.gauss_block > 96 * 1024) pl.gauss_block /= 2;
    pl.smem_gauss = (per_thread * pl.gauss_block <= 96 * 1024) ? per_thread * pl.gauss_block : 0;
    if (pl.smem_gauss == 0) pl.gauss_block = 128;
    (void)batch; (void)sm_count;
    return pl;
}

// experiment (-DOSC2_WITH_FUSED_GF and OSC2_FUSED=1): K_G inside the forward sweep, experiments/step_kernels_fused.cuh
template <class Launcher>
bool osc2_use_fused(Launcher &L, const StepDev &p, double theta)
{
#ifdef OSC2_WITH_FUSED_GF
    static const bool on = [] { const char *e = getenv("OSC2_FUSED"); return e && e[0] == '1'; }();
    return on && p.d == 8 && theta == 1.0 && p.cap_sysmat == nullptr && p.sinv != nullptr && FusedFwdSmem<8, 8>::BYTES <= L.max_smem();
#else
    (void)L; (void)p; (void)theta;
    return false;
#endif
}

template <int D, int LPC, class Launcher>
void osc2_run_fast_sweeps(Launcher &L, const osc2_topology *tp_dev, const StepDev &p, double theta, bool fused = false)
{
    // one warp per CTA: a batch of a few thousand conductors then still spreads
    // over all SMs (4096 conductors x 8 lanes = 1024 warps on 148 SMs)
    constexpr int cpb = 32 / LPC;
    const bool be = (theta == 1.0), cap = (p.cap_sysmat != nullptr);
    const size_t smem_f = (size_t)((be && D % 8 == 0) ? FastFwdSmem<D, true>::PER_COND : FastFwdSmem<D, false>::PER_COND) *
                          cpb * sizeof(double);
    const size_t smem_b = (size_t)(2 * D + 2) * cpb * sizeof(double);
    if constexpr (D % 8 == 0 && LPC == D) {
        // Backward Euler, no capture: tiled records by TMA bulk copies + node-pipelined rounds (step_kernels_tile.cuh);
        // OSC2_FWD_TILE=0 keeps the round-1 kernel (same bits) for A/B timing
        static const bool tile_on = [] { const char *e = getenv("OSC2_FWD_TILE"); return !(e && e[0] == '0'); }();
        if (be && !cap && tile_on) {
            const unsigned nt = (unsigned)((p.B + cpb - 1) / cpb);
#ifdef OSC2_WITH_SPLIT_ROLES
            // experiment: assembler / eliminator warps (experiments/step_kernels_split_roles.cuh), OSC2_FWD_SPLIT=1
            static const bool split_on = [] { const char *e = getenv("OSC2_FWD_SPLIT"); return e && e[0] == '1'; }();
            if (split_on && SplitFwdSmem<D, LPC>::BYTES <= L.max_smem()) {
                using SS = SplitFwdSmem<D, LPC>;
                const unsigned nc = (nt + SS::TILES - 1) / SS::TILES;
                if (p.B % cpb != 0) L.launch(OSC2_K_FORWARD, osc2_forward_split_kernel<D, LPC, false>, nc, 512u, SS::BYTES, tp_dev, p);
                else L.launch(OSC2_K_FORWARD, osc2_forward_split_kernel<D, LPC, true>, nc, 512u, SS::BYTES, tp_dev, p);
                L.launch(OSC2_K_BACKWARD, osc2_backward_tile_kernel<D, LPC>, nt, 32u, TileBwdSmem<D, LPC>::BYTES, p);
                return;
            }
#endif
#ifdef OSC2_WITH_FUSED_GF
            if constexpr (D == 8) {
                if (fused) {
                    using FS = FusedFwdSmem<D, LPC>;
                    const unsigned nc = (nt + FS::NCW - 1) / FS::NCW;
                    if (p.B % (cpb * FS::NCW) == 0) L.launch(OSC2_K_FORWARD, osc2_forward_fused_kernel<D, LPC, true>, nc, 32u * (FS::NCW + FS::NPW), FS::BYTES, tp_dev, p);
                    else L.launch(OSC2_K_FORWARD, osc2_forward_fused_kernel<D, LPC, false>, nc, 32u * (FS::NCW + FS::NPW), FS::BYTES, tp_dev, p);
                    L.launch(OSC2_K_BACKWARD, osc2_backward_tile_kernel<D, LPC>, nt, 32u, TileBwdSmem<D, LPC>::BYTES, p);
                    return;
                }
            }
#endif
            (void)fused;
            if (p.B % cpb == 0) L.launch(OSC2_K_FORWARD, osc2_forward_tile_kernel<D, LPC, true>, nt, 32u, TileFwdSmem<D, LPC>::BYTES, tp_dev, p);
            else L.launch(OSC2_K_FORWARD, osc2_forward_tile_kernel<D, LPC, false>, nt, 32u, TileFwdSmem<D, LPC>::BYTES, tp_dev, p);
            L.launch(OSC2_K_BACKWARD, osc2_backward_tile_kernel<D, LPC>, nt, 32u, TileBwdSmem<D, LPC>::BYTES, p);
            return;
        }
    }
    constexpr int W = 1;     // one warp per CTA
    const unsigned blk = 32u * W, gridW = (unsigned)((p.B + cpb * W - 1) / (cpb * W));
    bool full = false;       // every lane of every warp owns a row: the kernel drops its lane predicates
    if constexpr (LPC == D) full = be && !cap && (p.B % (cpb * W) == 0);
    if (full) {
        if constexpr (LPC == D)
            L.launch(OSC2_K_FORWARD, osc2_forward_fast_kernel<D, LPC, true, false, true>, gridW, blk, smem_f * W, tp_dev, p);
    }
    else if (be && !cap) L.launch(OSC2_K_FORWARD, osc2_forward_fast_kernel<D, LPC, true, false>, gridW, blk, smem_f * W, tp_dev, p);
    else if (be) L.launch(OSC2_K_FORWARD, osc2_forward_fast_kernel<D, LPC, true, true>, gridW, blk, smem_f * W, tp_dev, p);
    else if (!cap) L.launch(OSC2_K_FORWARD, osc2_forward_fast_kernel<D, LPC, false, false>, gridW, blk, smem_f * W, tp_dev, p);
    else L.launch(OSC2_K_FORWARD, osc2_forward_fast_kernel<D, LPC, false, true>, gridW, blk, smem_f * W, tp_dev, p);
    L.launch(OSC2_K_BACKWARD, osc2_backward_fast_kernel<D, LPC>, gridW, blk, smem_b * W, p);
}

// `records_ready`: the element row records were already produced by the fused K_PG kernel
template <class Launcher>
void osc2_run_step(Launcher &L, const osc2_topology *tp_dev, const StepDev &p, bool fast, double theta,
                   bool records_ready = false)
{
    const Osc2LaunchPlan pl = osc2_make_plan(p.d, p.B, L.max_smem(), L.sm_count());
    if (fast) {
        const long long work = (long long)p.E * p.B;
        const unsigned grid = (unsigned)((work + pl.gauss_block - 1) / pl.gauss_block);
        const bool fused = !records_ready && osc2_use_fused(L, p, theta);
        if (!records_ready && !fused)
            L.launch(OSC2_K_GAUSS, osc2_gauss2_kernel<true>, grid, pl.gauss_block,
                     (size_t)p.s_slots * sizeof(double) * pl.gauss_block + (((size_t)p.s_slots * 2 + 15) / 16) * 16, tp_dev, p);
        switch (p.d) {
        case 4: osc2_run_fast_sweeps<4, 8>(L, tp_dev, p, theta); break;
        case 8: osc2_run_fast_sweeps<8, 8>(L, tp_dev, p, theta, fused); break;
        default: osc2_run_fast_sweeps<13, 16>(L, tp_dev, p, theta); break;
        }
        {
            const long long nl = (long long)p.d * p.B * p.n_leaves, nw = (long long)p.d * p.B;
            L.launch(OSC2_K_NORMS, osc2_norm_leaves_kernel, (unsigned)((nl + 127) / 128), 128u, (size_t)0, p);
            L.launch(OSC2_K_NORMS, osc2_norm_combine_kernel, (unsigned)((nw + 127) / 128), 128u, (size_t)0, p);
        }
        return;
    }
    // K_G
    {
        const long long work = (long long)p.E * p.B;
        const unsigned grid = (unsigned)((work + pl.gauss_block - 1) / pl.gauss_block);
        if (pl.smem_gauss) L.launch(OSC2_K_GAUSS, osc2_gauss_kernel<true>, grid, pl.gauss_block, pl.smem_gauss, tp_dev, p);
        else L.launch(OSC2_K_GAUSS, osc2_gauss_kernel<false>, grid, pl.gauss_block, (size_t)0, tp_dev, p);
    }
    // K_F, K_B
    // NODOFS = 64 (CASE_2): register-tiled sweeps, one 256-thread CTA (forward) / one warp (backward)
    // per conductor (step_kernels_d64.cuh); OSC2_D64=0 keeps the generic shared-memory sweeps
    static const bool d64_on = [] { const char *e = getenv("OSC2_D64"); return !(e && e[0] == '0'); }();
    if (p.d == 64 && d64_on && osc2_d64_fwd_smem_doubles() * sizeof(double) <= L.max_smem()) {
        const size_t smem64 = osc2_d64_fwd_smem_doubles() * sizeof(double);
        // OSC2_D64_SYNC=flags (experiment): per-row shared-memory flags instead of one __syncthreads per
        // in-node pivot round -- 76.3 vs 77.1 ms: the owner warp's own chain is the limit, not the waiting
        static const bool flags = [] { const char *e = getenv("OSC2_D64_SYNC"); return e && e[0] == 'f'; }();
        if (p.cap_sysmat && flags) L.launch(OSC2_K_FORWARD, osc2_forward64_kernel<true, true>, (unsigned)p.B, 256u, smem64, tp_dev, p);
        else if (p.cap_sysmat) L.launch(OSC2_K_FORWARD, osc2_forward64_kernel<true, false>, (unsigned)p.B, 256u, smem64, tp_dev, p);
        else if (flags) L.launch(OSC2_K_FORWARD, osc2_forward64_kernel<false, true>, (unsigned)p.B, 256u, smem64, tp_dev, p);
        else L.launch(OSC2_K_FORWARD, osc2_forward64_kernel<false, false>, (unsigned)p.B, 256u, smem64, tp_dev, p);
        L.launch(OSC2_K_BACKWARD, osc2_backward64_kernel, (unsigned)p.B, 32u, (size_t)(2 * 128 * sizeof(double)), p);
        const long long nl = (long long)p.d * p.B * p.n_leaves, nw = (long long)p.d * p.B;
        L.launch(OSC2_K_NORMS, osc2_norm_leaves_kernel, (unsigned)((nl + 127) / 128), 128u, (size_t)0, p);
        L.launch(OSC2_K_NORMS, osc2_norm_combine_kernel, (unsigned)((nw + 127) / 128), 128u, (size_t)0, p);
        return;
    }
    const unsigned gridf = (unsigned)((p.B + pl.cpb_fwd - 1) / pl.cpb_fwd);
    const unsigned blockf = (unsigned)(pl.lpc * pl.cpb_fwd);
    const unsigned gridb = (unsigned)((p.B + pl.cpb_bwd - 1) / pl.cpb_bwd);
    const unsigned blockb = (unsigned)(pl.lpc * pl.cpb_bwd);
    switch (pl.lpc) {
    case 8:
        L.launch(OSC2_K_FORWARD, osc2_forward_kernel<8>, gridf, blockf, pl.smem_fwd, tp_dev, p);
        L.launch(OSC2_K_BACKWARD, osc2_backward_kernel<8>, gridb, blockb, pl.smem_bwd, p);
        break;
    case 16:
        L.launch(OSC2_K_FORWARD, osc2_forward_kernel<16>, gridf, blockf, pl.smem_fwd, tp_dev, p);
        L.launch(OSC2_K_BACKWARD, osc2_backward_kernel<16>, gridb, blockb, pl.smem_bwd, p);
        break;
    case 32:
        L.launch(OSC2_K_FORWARD, osc2_forward_kernel<32>, gridf, blockf, pl.smem_fwd, tp_dev, p);
        L.launch(OSC2_K_BACKWARD, osc2_backward_kernel<32>, gridb, blockb, pl.smem_bwd, p);
        break;
    default:
        L.launch(OSC2_K_FORWARD, osc2_forward_kernel<64>, gridf, blockf, pl.smem_fwd, tp_dev, p);
        L.launch(OSC2_K_BACKWARD, osc2_backward_kernel<64>, gridb, blockb, pl.smem_bwd, p);
        break;
    }
    // K_N
    {
        const long long nl = (long long)p.d * p.B * p.n_leaves, nw = (long long)p.d * p.B;
        L.launch(OSC2_K_NORMS, osc2_norm_leaves_kernel, (unsigned)((nl + 127) / 128), 128u, (size_t)0, p);
        L.launch(OSC2_K_NORMS, osc2_norm_combine_kernel, (unsigned)((nw + 127) / 128), 128u, (size_t)0, p);
    }
}
