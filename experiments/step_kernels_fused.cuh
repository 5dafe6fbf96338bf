// K_GF: assembly (K_G) and forward elimination (K_F) in ONE kernel, NODOFS = 8, Backward Euler.
//
// The element records K_G writes and K_F reads back (1088 B per node-timestep each way) exist only because assembly
// and elimination were separate launches.  Here a CTA is 8 consumer warps -- each the forward sweep of one tile of
// 4 conductors, the code of osc2_forward_tile_kernel (step_kernels_tile.cuh) -- and a warpgroup of producer warps
// that runs the K_G statements for the next elements and writes the records straight into the consumers'
// shared-memory rings: `gm` is not touched, K_G's launch disappears, and the producers' global loads sit in the
// latency shadow of the elimination chain (K_F issues on 50 % of its cycles).
//
//   warps 0..7   consumers, 208 registers each (setmaxnreg.inc)            tile = 8 * blockIdx.x + warp
//   warps 8..11  producers,  88 registers each (setmaxnreg.dec); warps 8 and 9 work, each serves 4 consumers:
//                lane = (element parity i, conductor c of 16): one pass builds elements e, e + 1 for 16 conductors
//   ring         4 element records per consumer warp: m - 1 and m in use, two being built
//   mbarriers    full[quad][stage]  (16 producer lanes arrive, the 4 consumer warps wait)
//                empty[quad][stage] (the 128 consumer lanes arrive after their last read, the producer lanes wait)
//
// 8 * 208 + 4 * 88 = 2016 = 12 * 168 registers per thread slot: one CTA per SM, 128 CTAs for 4096 conductors.
#pragma once

#include "../opensc2_b200/csrc/step_kernels_tile.cuh"

#ifdef OSC2_EMU
inline void osc2_setmaxnreg_inc_208() {}
inline void osc2_setmaxnreg_dec_88() {}
#else
__device__ __forceinline__ void osc2_setmaxnreg_inc_208() { asm volatile("setmaxnreg.inc.sync.aligned.u32 208;"); }
__device__ __forceinline__ void osc2_setmaxnreg_dec_88() { asm volatile("setmaxnreg.dec.sync.aligned.u32 88;"); }
#endif

template <int D, int LPC>
struct FusedFwdSmem {
    using T = TileFwdSmem<D, LPC>;
    static constexpr int NCW = 8;                          // consumer warps per CTA
    static constexpr int NPW = 4;                          // producer warps (one warpgroup), two of them work
    static constexpr int QUAD = 4;                         // consumer warps per working producer
    static constexpr int NST = 4;                          // ring stages
    // + 4: the rings of the four consumers one producer serves start 4 doubles (8 banks) apart, so the 16 lanes of a
    // producer half-warp (4 tiles x 4 conductors) hit 16 different doubles
    static constexpr int RING = NST * T::REC + 4;
    static constexpr int OFF_STG = NCW * RING;
    static constexpr int OFF_PUB = OFF_STG + NCW * T::STG;
    static constexpr int OFF_BAR = OFF_PUB + NCW * T::CPW * T::PUBC;
    static constexpr int NBAR = 2 * (NCW / QUAD) * NST;    // full + empty, per quad and stage; 16 bytes each
    static constexpr int DOUBLES = OFF_BAR + 2 * NBAR;
    static constexpr size_t BYTES = (size_t)DOUBLES * sizeof(double);
    static_assert(RING % 2 == 0 && OFF_STG % 2 == 0 && T::STG % 2 == 0, "16-byte alignment of the bulk-store source");
};

// Producer: the K_G statements (osc2_gauss2_body, INREC) for elements e0 + i of 16 conductors, records into the rings
template <int D, int LPC>
__device__ __forceinline__ void osc2_fused_producer(const osc2_topology *__restrict__ tp, const StepDev &p, double *sm_dyn, const int quad)
{
    using SM = FusedFwdSmem<D, LPC>;
    using T = TileFwdSmem<D, LPC>;
    constexpr int NST = SM::NST, CPW = T::CPW;
    const int lane = threadIdx.x & 31;
    const int i = lane >> 4, c = lane & 15;                 // element parity, conductor of the quad
    const int w = quad * SM::QUAD + c / CPW;                // consumer warp that owns this conductor
    const long long b_ll = ((long long)blockIdx.x * SM::NCW + w) * CPW + c % CPW;
    const bool valid = b_ll < (long long)p.B;
    const size_t b = valid ? (size_t)b_ll : 0;
    double *ring = sm_dyn + (size_t)w * SM::RING + c % CPW;
    double *full = sm_dyn + SM::OFF_BAR + 2 * (quad * NST);
    double *empty = sm_dyn + SM::OFF_BAR + 2 * ((SM::NCW / SM::QUAD) * NST + quad * NST);
    const int E = p.E;
    for (int e0 = 0; e0 < E; e0 += 2) {
        const int e = e0 + i;
        if (e < E) {
            const int s = e % NST;
            if (e >= NST) osc2_cbar_wait(empty + 2 * s, (unsigned)(((e / NST) - 1) & 1));
            if (valid) osc2_gauss2_body<true, true>(tp, p, ring + (size_t)s * T::REC, nullptr, 0, e, b);
            osc2_cbar_arrive(full + 2 * s);
        }
        __syncwarp();
    }
}

template <int D, int LPC, bool FULL>
__global__ void __launch_bounds__(384, 1) osc2_forward_fused_kernel(const osc2_topology *__restrict__ tp, StepDev p)
{
    OSC2_DYN_SMEM_ALIGNED(sm_dyn);
    using SM = FusedFwdSmem<D, LPC>;
    using T = TileFwdSmem<D, LPC>;
    const int warp = threadIdx.x >> 5;
    // the rings start as zeros: the producers leave the fields that are +0 for every element of the topology alone
    for (int k = threadIdx.x; k < SM::OFF_STG; k += blockDim.x) sm_dyn[k] = 0.0;
    if (threadIdx.x == 0) {
        for (int q = 0; q < SM::NCW / SM::QUAD; ++q)
            for (int s = 0; s < SM::NST; ++s) {
                osc2_cbar_init(sm_dyn + SM::OFF_BAR + 2 * (q * SM::NST + s), 16);                                     // full
                osc2_cbar_init(sm_dyn + SM::OFF_BAR + 2 * ((SM::NCW / SM::QUAD) * SM::NST + q * SM::NST + s), 32 * SM::QUAD);   // empty
            }
        osc2_mbar_init_fence();
    }
    __syncthreads();
    if (warp >= SM::NCW) {
        osc2_setmaxnreg_dec_88();
        if (warp < SM::NCW + SM::NCW / SM::QUAD) osc2_fused_producer<D, LPC>(tp, p, sm_dyn, warp - SM::NCW);
        return;
    }
    osc2_setmaxnreg_inc_208();
    TileFwdCtx cx;
    cx.rec = sm_dyn + (size_t)warp * SM::RING;
    cx.stg = sm_dyn + SM::OFF_STG + (size_t)warp * T::STG;
    cx.pub = sm_dyn + SM::OFF_PUB + (size_t)warp * T::CPW * T::PUBC;
    cx.bars = nullptr;
    const int quad = warp / SM::QUAD;
    cx.full = sm_dyn + SM::OFF_BAR + 2 * (quad * SM::NST);
    cx.empty = sm_dyn + SM::OFF_BAR + 2 * ((SM::NCW / SM::QUAD) * SM::NST + quad * SM::NST);
    cx.tile = (size_t)blockIdx.x * SM::NCW + warp;
    cx.tile_valid = cx.tile < (size_t)p.NT;
    osc2_forward_tile_body<D, LPC, FULL, true, SM::NST>(tp, p, cx);
}
