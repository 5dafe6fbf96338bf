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

struct Osc2LaunchPlan {
    int lpc;            // lanes per conductor (8, 16, 32, 64)
    int cpb_fwd;        // conductors per CTA, forward sweep
    int cpb_bwd;        // conductors per CTA, backward sweep
    size_t smem_fwd, smem_bwd;
    int gauss_block;
    size_t smem_gauss;  // 0: SMAT staged in the output record
};

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

template <class Launcher>
void osc2_run_step(Launcher &L, const osc2_topology *tp_dev, const StepDev &p)
{
    const Osc2LaunchPlan pl = osc2_make_plan(p.d, p.B, L.max_smem(), L.sm_count());
    // K_G
    {
        const long long work = (long long)p.E * p.B;
        const unsigned grid = (unsigned)((work + pl.gauss_block - 1) / pl.gauss_block);
        if (pl.smem_gauss) L.launch(OSC2_K_GAUSS, osc2_gauss_kernel<true>, grid, pl.gauss_block, pl.smem_gauss, tp_dev, p);
        else L.launch(OSC2_K_GAUSS, osc2_gauss_kernel<false>, grid, pl.gauss_block, (size_t)0, tp_dev, p);
    }
    // K_F, K_B
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
        const long long work = (long long)p.d * p.B;
        L.launch(OSC2_K_NORMS, osc2_norms_kernel, (unsigned)((work + 127) / 128), 128u, (size_t)0, p);
    }
}
