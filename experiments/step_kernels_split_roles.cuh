// EXPERIMENT (not in the product build; -DOSC2_WITH_SPLIT_ROLES compiles it in, OSC2_FWD_SPLIT=1 selects it): the forward
// sweep of a tile as two warps.  Bit-exact (emulator + the -m gpu step parity suite).  Measured at 4096 x 10 001 nodes:
// 28.1 ms against 28.9 ms for the one-warp tile kernel when it was written; after the instruction trims of the same
// session (range keys of the multipliers folded into one check per node, the empty last own-row round dropped) the
// one-warp kernel runs 27.9 ms and this one 28.9 ms -- the scheduler is issue bound either way (FP64 instructions take
// two issue cycles: 2 x (1480 + 700) slots of the 5700 cycles per node), so splitting the stream over two warps moves
// no work off it and adds the hand-over.  The role switches of osc2_forward_tile_body (step_kernels_tile.cuh) stay.
#pragma once

#include "../opensc2_b200/csrc/step_kernels_tile.cuh"

// ---------------------------------------------------------------------------
// K_F with split roles.  The forward sweep of a tile is two instruction streams with a one-way dependence: building
// and scaling the row of node m (tables of selects, the range keys, 75 quotients: ~550 instructions, no dependence
// on the elimination) and the eight pivot rounds (~930 instructions, the serial chain).  In
// osc2_forward_tile_kernel one warp runs both, and with 200 registers per thread only two such warps fit a
// scheduler -- which then issues on half of its cycles.  Here an ASSEMBLER warp builds the rows one or two nodes
// ahead and hands them over through two shared-memory slots, an ELIMINATOR warp runs the pivot rounds: four warps
// per scheduler instead of two for the same 4096 conductors.  One CTA per SM: 7 tiles = 7 eliminators (warps 0..6)
// + 7 assemblers (warps 8..14), 147 CTAs for 4096 conductors; setmaxnreg moves registers from the assembler
// warpgroups (120) to the eliminator warpgroups (136): 8 * 136 + 8 * 120 = 16 * 128 (152 / 104 and 144 / 112 spill in
// the assembler: 34.6 and 31.8 ms; 128 / 128: 29.2 ms; 136 / 120: 28.0 ms).  Shared memory per tile: record
// ring 3 x 4352 B (TMA, assembler) + hand-over 2 x 6400 B + pivot rows 4672 B; the factor rows leave by plain
// coalesced stores.  Same arithmetic, same bits.
// ---------------------------------------------------------------------------
template <int R> __device__ __forceinline__ void osc2_setmaxnreg_inc()
{
#ifndef OSC2_EMU
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(R));
#endif
}
template <int R> __device__ __forceinline__ void osc2_setmaxnreg_dec()
{
#ifndef OSC2_EMU
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(R));
#endif
}

template <int D, int LPC>
struct SplitFwdSmem {
    using T = TileFwdSmem<D, LPC>;
    static constexpr int TILES = 7;                        // per CTA: 7 x 30.5 KB of the 227 KB
    static constexpr int NST = 3;                          // record ring
    static constexpr int HREC = (3 * D + 1) * 32;          // one hand-over slot
    static constexpr int PER_TILE = NST * T::REC + 2 * HREC + T::CPW * T::PUBC;
    static constexpr int BAR = 16;                         // doubles per tile: ring barriers (8 bytes each) in 0..3, hand-over
                                                           // full x 2 at 4 and 6, empty x 2 at 8 and 10 (16 bytes each)
    static constexpr int OFF_BAR = TILES * PER_TILE;
    static constexpr int DOUBLES = OFF_BAR + TILES * BAR;
    static constexpr size_t BYTES = (size_t)DOUBLES * sizeof(double);
    static_assert(PER_TILE % 2 == 0 && T::REC % 2 == 0, "16-byte alignment of the bulk-copy destinations");
};

template <int D, int LPC, bool FULL, int RE = 136>       // RE: registers of an eliminator thread; an assembler thread gets 256 - RE
__global__ void __launch_bounds__(512, 1) osc2_forward_split_kernel(const osc2_topology *__restrict__ tp, StepDev p)
{
    OSC2_DYN_SMEM_ALIGNED(sm_dyn);
    using SM = SplitFwdSmem<D, LPC>;
    using T = TileFwdSmem<D, LPC>;
    const int warp = threadIdx.x >> 5;
    if (threadIdx.x < SM::TILES) {
        double *bar = sm_dyn + SM::OFF_BAR + (size_t)threadIdx.x * SM::BAR;
        for (int s = 0; s < SM::NST; ++s) osc2_mbar_init((unsigned long long *)bar + s, 1);
        for (int s = 0; s < 4; ++s) osc2_cbar_init(bar + 4 + 2 * s, 32);
        osc2_mbar_init_fence();
    }
    __syncthreads();
    const int role = (warp < 8) ? OSC2_ROLE_ELIM : OSC2_ROLE_ASM;
    const int t = warp & 7;
    if (RE > 128) { if (role == OSC2_ROLE_ELIM) osc2_setmaxnreg_inc<RE>(); else osc2_setmaxnreg_dec<256 - RE>(); }
    if (t >= SM::TILES || (size_t)blockIdx.x * SM::TILES + t >= (size_t)p.NT) return;     // the barriers are per tile
    double *base = sm_dyn + (size_t)t * SM::PER_TILE;
    double *bar = sm_dyn + SM::OFF_BAR + (size_t)t * SM::BAR;
    TileFwdCtx cx;
    cx.rec = base;
    cx.hand = base + SM::NST * T::REC;
    cx.pub = cx.hand + 2 * SM::HREC;
    cx.stg = nullptr;
    cx.bars = (unsigned long long *)bar;
    cx.full = cx.empty = nullptr;
    cx.hfull = bar + 4;
    cx.hempty = bar + 8;
    cx.tile = (size_t)blockIdx.x * SM::TILES + t;
    cx.tile_valid = true;
    if (role == OSC2_ROLE_ELIM) osc2_forward_tile_body<D, LPC, FULL, false, SM::NST, OSC2_ROLE_ELIM>(tp, p, cx);
    else osc2_forward_tile_body<D, LPC, FULL, false, SM::NST, OSC2_ROLE_ASM>(tp, p, cx);
}

