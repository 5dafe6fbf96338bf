// Device-side parameter blocks shared by the kernels and the C-ABI layer.
//
// Device layout is batch-innermost: an array the reference keeps as
// x[...] per conductor is x[...][b] here, so a warp touching 32 consecutive
// conductors issues one 256-byte FP64 transaction per value.
#pragma once

#include "../../include/opensc2_b200.h"

#ifndef OSC2_EMU
#include <cuda_runtime.h>
#endif

// status[b] value of a conductor whose numbers left the range in which the fast
// kernels' division is verified exact (osc2_div.cuh); never produced by the generic path
#define OSC2_STATUS_RANGE (-2000000000)
// osc2_eval_tcs: Jop is not below Jc(3.5 K) (the reference's bisect raises ValueError)
#define OSC2_STATUS_TCS (-2000000001)

struct StepDev {
    int B, N, E, M, S, d, NS;
    int nff, nfs, nss, nes;
    double dt, t_env;
    int num_step;
    // inputs ([...][B])
    const double *dz;         // [E]
    const double *fl_gauss;   // [9][M][E]
    const double *fl_node;    // [4][M]
    const double *fl_bc;      // [6][M]
    const double *sol_gauss;  // [3][S][E]
    const double *sol_q;      // [2][S][E][2]
    const double *peri_ff, *htc_ff;   // [2][nff][E]
    const double *peri_fs, *htc_fs;   // [nfs][E]
    const double *peri_ss, *htc_ss;   // [nss][E]
    const double *peri_es;            // [nes][E]
    const double *htc_es;             // [nes][3][E]
    const double *qsource;            // [N][S]
    const double *peri_ss_nodal;      // [nss][N]  nodal contact perimeters (radiative interfaces)
    // state
    const double *sysvar;     // [N][d]   old solution
    double *sysvar_new;       // [N][d]   new solution
    double *syslod;           // [N][d][2]
    // work
    double *gm;               // [E][NS]  Gauss-point matrices (see gm_slots)
    double *k123;             // [3][nff][E]
    double *spill;            // [N][d][2d+1]  U factors + forward-substituted rhs
    // outputs
    double *norm_solution, *norm_change, *eqteig;   // [d]
    int *status;              // [B] 0 or -(row+1)
    double *cap_sysmat;       // optional [Full][Total]
    double *cap_known;        // optional [Total]
    // norms: numpy pairwise-sum tree of N values (osc2_build_norm_plan)
    const int *norm_leaf;     // [n_leaves][2]  start node, length
    const int *norm_prog;     // [n_prog]       >= 0: push leaf, -1: add the two on top
    double *norm_part;        // [n_leaves][3][d][B]
    int n_leaves, n_prog;
    // K_G: SMAT is staged per thread in shared memory; only the entries the topology can touch get a
    // slot (osc2_build_smap), every other (row, col) maps to one shared all-zero slot
    const unsigned char *smap; // [d*d] -> slot
    int s_slots;              // number of slots incl. the zero slot
    // fast path (step_kernels_fast.cuh / step_kernels_tile.cuh): `gm` and `spill` are TILED by warp -- a tile is the
    // cpw = 32 / LPC conductors one warp of the sweeps owns, NT = ceil(B / cpw) tiles:
    //   gm    [E][NT][d + 9 fields][32 lanes]      lane = row * cpw + (b % cpw)
    //   spill [N][NT][2d + 2 fields][32 lanes]
    // so that everything a warp needs for one node is ONE contiguous block (4352 / 4608 bytes at d = 8) that a
    // single bulk copy (cp.async.bulk, TMA) moves
    int cpw, NT;
    // K_G (fast path): slot -> offset of S[row][column] inside the tiled record (column * 32 + row * cpw), the inverse
    // of `smap`; gm_prezeroed: `gm` was zeroed at creation and K_G leaves the fields that are +0 for every element of
    // this topology alone
    const unsigned short *sinv;
    int gm_prezeroed;
};

// Slot map of one element's Gauss-point record in `gm`:
//   S      d*d   source Jacobian SMAT, row major
//   Msol   S     MMAT diagonal of the solid equations (fluid rows are 1)
//   K      d     KMAT diagonal
//   A      4*M   per channel: v, 1/rho, rho*c0^2, Gruneisen*T (AMAT non-zeros)
//   SV     2*S   SVEC present  [l][0|1]
//   SVP    2*S   SVEC previous [l][0|1]  (first step only)
struct GmSlots {
    int d, M, S;
    __host__ __device__ int s(int r, int c) const { return r * d + c; }
    __host__ __device__ int msol(int l) const { return d * d + l; }
    __host__ __device__ int k(int r) const { return d * d + S + r; }
    __host__ __device__ int a(int j, int q) const { return d * d + S + d + 4 * j + q; }
    __host__ __device__ int sv(int l, int t) const { return d * d + S + d + 4 * M + 2 * l + t; }
    __host__ __device__ int svp(int l, int t) const { return d * d + S + d + 4 * M + 2 * S + 2 * l + t; }
    __host__ __device__ int count() const { return d * d + S + d + 4 * M + 4 * S; }
};
