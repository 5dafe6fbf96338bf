/* opensc2_b200 -- C ABI of the batched B200 (sm_100a) OPENSC2 time-step path.
 *
 * The reference has no FFI: its hot path is the Python function
 *   step(conductor, environment, qsource, num_step)
 *   (/root/reference/source_code/utility_functions/transient_solution_functions.py:186-599)
 * called once per conductor per time step from Simulation.conductor_solution
 *   (/root/reference/source_code/simulation.py:407-412).
 * This header is what a ctypes binding of that call site binds instead
 * (INTEGRATION.md shows the stub).  One handle = one GPU = one batch of B
 * independent conductors sharing a topology (same deck, swept scalars).
 *
 * Conventions
 *  - every function returns 0 on success or a negative osc2_status; the text
 *    of the last failure (per host thread) is osc2_last_error();
 *  - host arrays are C-contiguous FP64/int32 with the batch axis leading and
 *    are only read or written during the call (the caller owns them);
 *  - device memory lives behind the opaque handle, laid out batch-innermost;
 *  - calls on one handle are ordered on the handle's CUDA stream and are not
 *    re-entrant; use one handle per GPU per host thread.
 */
#ifndef OPENSC2_B200_H
#define OPENSC2_B200_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define OSC2_MAX_CHANNELS 16
#define OSC2_MAX_SOLIDS 64
#define OSC2_MAX_INTERFACES 160

typedef enum {
    OSC2_OK = 0,
    OSC2_ERR_INVALID = -1,     /* bad argument / inconsistent sizes (reference: ValueError) */
    OSC2_ERR_CUDA = -2,        /* CUDA runtime failure, see osc2_last_error() */
    OSC2_ERR_NO_DEVICE = -3,   /* no usable CUDA device: there is no CPU fallback */
    OSC2_ERR_SINGULAR = -4,    /* at least one conductor hit abs(pivot) <= 1e-20 (gredub/gbacsb, tsf:676,763) */
    OSC2_ERR_STATE = -5,       /* call order violated (e.g. step before upload) */
    OSC2_ERR_RANGE = -6        /* a non-zero value below 2^-900 (or a divisor outside 2^+-100) met the fast
                                  kernels' exactness guard; rerun with OSC2_FORCE_GENERIC=1 */
} osc2_status;

/* Uniform-across-the-batch description of the conductor; mirrors what step()
 * reads from Conductor: dict_N_equation / equation_index (conductor.py:897-946,
 * 1070-1093), inventory, interface lists (conductor.py:2570-2597), channel and
 * solid inputs, Delta_p_min/k_loc/lambda_v (conductor.py:89-94), theta_method. */
typedef struct {
    int n_ch, n_sol;
    int nff, nfs, nss, nes;
    double ch_area[OSC2_MAX_CHANNELS];      /* channel.inputs["CROSSECTION"] */
    double ch_dh[OSC2_MAX_CHANNELS];        /* channel.inputs["HYDIAMETER"] */
    int ch_forward[OSC2_MAX_CHANNELS];      /* channel.flow_dir[0] == "forward" */
    int ch_intial[OSC2_MAX_CHANNELS];       /* abs(coolant.operations["INTIAL"]) in {1,2,3} */
    double sol_area[OSC2_MAX_SOLIDS];       /* solid.inputs["CROSSECTION"] */
    double sol_costeta[OSC2_MAX_SOLIDS];    /* solid.inputs["COSTETA"] */
    int ff[OSC2_MAX_INTERFACES][2];         /* fluid-fluid: channel, channel */
    int fs[OSC2_MAX_INTERFACES][2];         /* fluid-solid: channel, solid */
    int ss[OSC2_MAX_INTERFACES][2];         /* solid-solid: solid, solid */
    int es[OSC2_MAX_INTERFACES];            /* environment-solid: solid */
    int es_choice2[OSC2_MAX_INTERFACES];    /* HTC_choice.at[env, jacket] == 2 */
    int is_rectangular;
    double height, width;
    double delta_p_min, k_loc, lambda_v;
    double theta;                           /* conductor.theta_method */
    int method;                             /* 0 = "BE", 1 = "CN" (AM4: not supported) */
} osc2_topology;

/* What step() reads from the conductor's Gauss/nodal dictionaries, batch leading.
 * M = n_ch, S = n_sol, N = n_nodes, E = N-1, d = 3M+S, Total = d*N. */
typedef struct {
    const double *dz;         /* [B][E]            grid_features["delta_z"] */
    const double *fl_gauss;   /* [B][9][M][E]      v,p,T,rho,c0,Gruneisen,h,cv,f_total (coolant.dict_Gauss_pt, channel.dict_friction_factor[False]["total"]) */
    const double *fl_node;    /* [B][4][M]         nodal v[0], v[-1], rho[0], rho[-1] */
    const double *fl_bc;      /* [B][6][M]         PREINL,PREOUT,TEMINL,TEMOUT,MDTIN,MDTOUT */
    const double *sol_gauss;  /* [B][3][S][E]      rho, cp, k */
    const double *sol_q;      /* [B][2][S][E][2]   Q1,Q2 x [present, previous] */
    const double *peri_ff;    /* [B][2][nff][E]    dict_interf_peri["ch_ch"]["Open"|"Close"]["Gauss"] */
    const double *htc_ff;     /* [B][2][nff][E]    dict_Gauss_pt["HTC"]["ch_ch"]["Open"|"Close"] */
    const double *peri_fs;    /* [B][nfs][E] */
    const double *htc_fs;     /* [B][nfs][E] */
    const double *peri_ss;    /* [B][nss][E] */
    const double *htc_ss;     /* [B][nss][E]       HTC["sol_sol"]["cond"] */
    const double *peri_es;    /* [B][nes][E] */
    const double *htc_es;     /* [B][nes][3][E]    conv (or side), bottom, top */
    const double *qsource;    /* [B][N][S]         step() argument */
    double t_env;             /* environment.inputs["Temperature"] */
} osc2_step_inputs;

typedef struct osc2_context *osc2_handle;

/* Kernel ids for osc2_get_kernel_ms. */
enum { OSC2_K_TRANSPOSE = 0, OSC2_K_GAUSS = 1, OSC2_K_FORWARD = 2, OSC2_K_BACKWARD = 3,
       OSC2_K_NORMS = 4, OSC2_K_PROPS = 5, OSC2_K_COUNT = 8 };

const char *osc2_last_error(void);
int osc2_version(void);
int osc2_device_count(void);

/* Allocate device state for `batch` conductors of `n_nodes` nodes (>= 4: the
 * reference's build_known_therm_vector indexes out of range below that,
 * step_matrix_construction.py:1401-1443) on CUDA device `device`. */
int osc2_create(const osc2_topology *topo, int batch, int n_nodes, int device, osc2_handle *out);
int osc2_destroy(osc2_handle h);

/* dict_Step["SYSVAR"][:,0] ([B][Total]) and dict_Step["SYSLOD"] ([B][Total][2]). */
int osc2_upload_state(osc2_handle h, const double *sysvar, const double *syslod);
int osc2_upload_step_inputs(osc2_handle h, const osc2_step_inputs *in);

/* One reference step() for every conductor of the batch, on device-resident
 * inputs: Gauss matrices -> block rows + BC + row scaling -> no-pivot banded
 * LU (gredub) -> substitution (gbacsb) -> norms / EQTEIG.  Asynchronous. */
int osc2_step(osc2_handle h, double dt, int num_step);

/* Wait for the stream; returns OSC2_ERR_SINGULAR if any conductor is singular. */
int osc2_synchronize(osc2_handle h);

/* Any pointer may be NULL.  status[b] = 0 or -(row+1) of the first tiny pivot. */
int osc2_download_state(osc2_handle h, double *sysvar, double *syslod, int *status);
/* norm_solution/norm_change/eqteig: [B][d]; k123: [B][3][nff][E]. */
int osc2_download_diagnostics(osc2_handle h, double *norm_solution, double *norm_change,
                              double *eqteig, double *k123);
/* Pre-solve system of the LAST step (debug/parity): sysmat [B][Full][Total] in the
 * reference's band layout, known [B][Total]; needs osc2_set_capture(h,1) before. */
int osc2_set_capture(osc2_handle h, int on);
int osc2_download_system(osc2_handle h, double *sysmat, double *known);

/* Timing: CUDA events on the handle's stream. */
int osc2_set_profiling(osc2_handle h, int on);               /* per-kernel events */
int osc2_get_kernel_ms(osc2_handle h, double ms[OSC2_K_COUNT], long long launches[OSC2_K_COUNT]);
int osc2_reset_kernel_ms(osc2_handle h);
int osc2_timer_start(osc2_handle h);
int osc2_timer_stop(osc2_handle h, double *elapsed_ms);      /* synchronises */

size_t osc2_device_bytes(osc2_handle h);

/* Device self test: the shared-divisor division used by the fast kernels
 * (csrc/osc2_div.cuh) against the compiler's a / b on n pseudo-random pairs. */
int osc2_selftest_division(int device, long long n, unsigned long long seed, unsigned long long *mismatches);

#ifdef __cplusplus
}
#endif
#endif
