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
    OSC2_ERR_RANGE = -6,       /* a non-zero value below 2^-900 (or a divisor outside 2^+-100) met the fast
                                  kernels' exactness guard; rerun with OSC2_FORCE_GENERIC=1 */
    OSC2_ERR_TCS = -7          /* osc2_eval_tcs: Jop >= Jc(3.5 K) somewhere (scipy.optimize.bisect: "f(a) and f(b)
                                  must have different signs", niobium3_tin.py:613) */
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
    const double *peri_ss_nodal; /* [B][nss][N]    dict_interf_peri["sol_sol"]["nodal"]: only read for radiative
                                                   interfaces (|HTC_choice| == 3), may be NULL otherwise */
} osc2_step_inputs;

/* ---- Operating conditions at the Gauss points (what the reference evaluates every
 * time step BEFORE step(): Conductor.operating_conditions_em / _th, conductor.py:4591-4744,
 * eval_transp_coeff conductor.py:4822-5406, build_heat_source conductor.py:4423-4581).
 * osc2_model is uniform across the batch; osc2_sweep holds the swept per-conductor scalars. */
#define OSC2_MAX_MATERIALS 8
#define OSC2_MAX_FLUIDS 4
#define OSC2_N_FLUID_PROPS 10
/* property order of a fluid table = the reference's alias table (simulation.py:89-100) */
enum { OSC2_FP_BETA = 0,   /* isobaric_expansion_coefficient */
       OSC2_FP_KAPPA,      /* isothermal_compressibility */
       OSC2_FP_PRANDTL, OSC2_FP_RHO, OSC2_FP_MU, OSC2_FP_H, OSC2_FP_CP, OSC2_FP_CV,
       OSC2_FP_SOUND, OSC2_FP_K };
/* materials (the fits of properties_of_materials/) */
enum { OSC2_MAT_CU_NIST = 0, OSC2_MAT_NB3SN = 1, OSC2_MAT_SS = 2, OSC2_MAT_GE = 3, OSC2_MAT_AL = 4,
       OSC2_MAT_RE123 = 5,
       OSC2_MAT_AG = 6,        /* silver.py */
       OSC2_MAT_HC276 = 7,     /* hastelloy_c276.py */
       OSC2_MAT_SN60PB40 = 8,  /* solder_sn60_pb40.py */
       OSC2_MAT_NBTI = 9,      /* niobium_titanium.py: superconductor of a StrandMixedComponent (Tc from
                                  critical_temperature_nbti; sol_tc0m, sol_bc20m as for Nb3Sn) */
       OSC2_MAT_MGB2 = 10,     /* magnesium_diboride.py: conductivity (T, B) at its fixed RRR = 100, cp(T) */
       OSC2_MAT_COUNT };
/* solid kinds: StrandMixedComponent, StrandStabilizerComponent, JacketComponent, StackComponent (HTS tapes:
 * layers homogenised by thickness, stack_component.py:398-476; weights = material_thickness,
 * weight_sum = tape_thickness) */
enum { OSC2_SOLID_STRAND_MIXED = 0, OSC2_SOLID_STRAND_STAB = 1, OSC2_SOLID_JACKET = 2, OSC2_SOLID_STACK = 3 };

typedef struct {
    /* channels (channel.py:43-169) */
    int ch_fluid[OSC2_MAX_CHANNELS];            /* which fluid table */
    int ch_ifriction[OSC2_MAX_CHANNELS];        /* IFRICTION: -99, 100-115, 118-122, 124, 200-207, 209-211 (every law of channel.py that
                                                   runs; 116 / 117 / 123 fail and 208 raises in the reference) */
    double ch_friction_mult[OSC2_MAX_CHANNELS]; /* FRICTION_MULTIPLIER */
    int ch_htc_corr[OSC2_MAX_CHANNELS];         /* Flag_htc_steady_corr: 1, 119, 120, 121, 211 or 212 */
    /* solids */
    int sol_kind[OSC2_MAX_SOLIDS];
    int sol_nmat[OSC2_MAX_SOLIDS];
    int sol_mat[OSC2_MAX_SOLIDS][OSC2_MAX_MATERIALS];
    double sol_weight[OSC2_MAX_SOLIDS][OSC2_MAX_MATERIALS]; /* homogenization_cefficients | cross_sections */
    double sol_weight_sum[OSC2_MAX_SOLIDS];     /* homogenization_cefficients_sum | inputs["CROSSECTION"] */
    double sol_rrr[OSC2_MAX_SOLIDS];
    double sol_tc0m[OSC2_MAX_SOLIDS], sol_bc20m[OSC2_MAX_SOLIDS];
    int sol_heated[OSC2_MAX_SOLIDS];            /* IQFUN == 1 (square wave, solid_component.py:415-576) */
    /* interfaces, in the order of the topology lists.  choice = HTC_choice; htc = contact_HTC;
     * mlt = HTC_multiplier (conductor.py:4867-5160) */
    int fs_choice[OSC2_MAX_INTERFACES];         /* 2 or -2 */
    double fs_htc[OSC2_MAX_INTERFACES], fs_mlt[OSC2_MAX_INTERFACES];
    int ff_choice[OSC2_MAX_INTERFACES];         /* 2 or -2 */
    double ff_htc[OSC2_MAX_INTERFACES], ff_mlt[OSC2_MAX_INTERFACES], ff_thick[OSC2_MAX_INTERFACES];
    int ss_choice[OSC2_MAX_INTERFACES];         /* 1, -1 (conduction) or 3, -3 (radiation between jackets) */
    double ss_htc[OSC2_MAX_INTERFACES], ss_mlt[OSC2_MAX_INTERFACES];
    double ss_thick_rc[OSC2_MAX_INTERFACES], ss_thick_cr[OSC2_MAX_INTERFACES], ss_rcontact[OSC2_MAX_INTERFACES];
    double es_htc[OSC2_MAX_INTERFACES];         /* constant convective HTC (choices -2, -4), Phi_conv and multiplier applied */
    double ch_roughness[OSC2_MAX_CHANNELS];     /* Channel.inputs["Roughness"] (IFRICTION 122) */
    /* radiative exchange between two jackets, HTC_choice 3 (conductor.py:5161-5219, 5428-5443): the
     * temperature-independent denominator (1-e_i)/e_i + 1/F + (1-e_j)/e_j * (P_out,i / P_in,j) with i the
     * inner jacket (inf when an emissivity or the view factor is 0), and the multiplier the reference
     * applies in its second branch only (1 otherwise) */
    double ss_rad_den[OSC2_MAX_INTERFACES];
    double ss_rad_mlt[OSC2_MAX_INTERFACES];
    double ch_void_fraction[OSC2_MAX_CHANNELS]; /* Channel.inputs["VOID_FRACTION"] (bundle laws 204, 205, 206, 209) */
    double sol_c0[OSC2_MAX_SOLIDS];             /* strand inputs["c0"]: scaling constant of Jc (A T / m^2), physical definition
                                                   (strand_component.py:266-276); only read by osc2_eval_tcs */
    double ch_lam_coef[OSC2_MAX_CHANNELS];      /* IFRICTION 119 (rectangular duct, channel.py:664-686): numerator of its laminar law,
                                                   24 (1 - 1.3553 a + 1.9467 a^2 - 1.7012 a^3 + 0.9564 a^4 - 0.2537 a^5), a = SIDE2 / SIDE1 */
} osc2_model;

/* Swept per-conductor scalars, batch leading. */
typedef struct {
    const double *b_field;   /* [B][S][2]  BISS, BOSS (IBIFUN == 0: linspace, solid_component.py:383-388) */
    const double *eps;       /* [B][S]     strain (IEPS 0/1, strand_component.py:353-399) */
    const double *q0;        /* [B][S]     Q0, W/m */
    const double *q_window;  /* [B][S][4]  first node, last node (heat0_where), TQBEG, TQEND */
    const double *tcs;       /* [B][S][E]  T_cur_sharing_min at the Gauss points (evaluated once at
                                           step 0, conductor.py:4615-4619); NULL when no solid holds a
                                           superconductor, or when the library computes it: osc2_eval_tcs */
    const double *jop;       /* [B][S]     operating current density abs(op_current) / cross_section["sc"], A/m^2
                                           (strand_component.py:266-276), the input of osc2_eval_tcs; may be NULL
                                           (then tcs is the caller's) */
} osc2_sweep;

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

/* Device-side operating conditions.  osc2_set_model + osc2_set_fluid_table + osc2_upload_sweep
 * once; then osc2_eval_operating_conditions(h, t) computes, from the device-resident state, every
 * per-step array osc2_step consumes except the geometry (dz, contact perimeters, qsource) and the
 * boundary values fl_bc, which stay as uploaded by osc2_upload_step_inputs (NULL members of
 * osc2_step_inputs are then allowed for the computed arrays).
 * Fluid table: regular grid, props[OSC2_N_FLUID_PROPS][n_p][n_t], bilinear in (p, T). */
int osc2_set_model(osc2_handle h, const osc2_model *model);
int osc2_set_fluid_table(osc2_handle h, int fluid, int n_p, int n_t, double p_min, double p_step,
                         double t_min, double t_step, const double *props);
int osc2_upload_sweep(osc2_handle h, const osc2_sweep *sweep);
/* Current sharing temperature on the device (StrandComponent.eval_tcs, strand_component.py:266-349; Nb3Sn:
 * current_sharing_temperature_nb3sn, properties_of_materials/niobium3_tin.py:493-623 -- the bisection of
 * scipy.optimize.bisect(xtol = 1e-5, rtol = 4 eps) on Jc(T, B, eps) - Jop between 3.5 K and Tc(B, eps), iterate for
 * iterate).  Fills T_cur_sharing_min at the Gauss points (what the Nb3Sn specific heat reads every step) and at the
 * nodes (what save_properties prints) from sweep.b_field / eps / jop and model.sol_c0 / sol_tc0m / sol_bc20m for every
 * StrandMixedComponent whose superconductor is Nb3Sn.  Called once, at step 0, like the reference does with
 * TCS_EVALUATION False (conductor.py:4615-4619).  A conductor whose Jop is not below Jc(3.5 K) gets status
 * OSC2_STATUS_TCS (the reference's bisect raises ValueError there). */
int osc2_eval_tcs(osc2_handle h);
/* tcs_gauss [B][S][E], tcs_nodal [B][S][N]; either may be NULL. */
int osc2_download_tcs(osc2_handle h, double *tcs_gauss, double *tcs_nodal);
/* Temperature margin min over the Gauss points of (T_cur_sharing_min - T) per conductor and solid, [B][S], of the
 * state the LAST osc2_eval_operating_conditions / osc2_advance looked at (+inf for solids without a superconductor
 * or before any evaluation): the quantity a temperature-margin sweep is run for. */
int osc2_download_margin(osc2_handle h, double *margin);
/* Nodal coolant pass on the CURRENT state (after a step: the new solution) -- what the reference evaluates for its
 * outputs with FluidComponent._eval_properties_nodal_gauss(nodal = True) and
 * _compute_density_and_mass_flow_rates (coolant.py:134-263) and Channel.eval_friction_factor(nodal = True)
 * (channel.py:1119-1138).  osc2_download_nodal_properties: chan [OSC2_N_NODAL = 14][B][M][N] = total_density,
 * isobaric_expansion_coefficient, isothermal_compressibility, Prandtl, total_dynamic_viscosity, total_enthalpy,
 * total_isobaric_specific_heat, total_isochoric_specific_heat, total_speed_of_sound, total_thermal_conductivity,
 * Reynolds, Gruneisen, mass_flow_rate, friction_factor (the property columns of save_properties' channel files,
 * utility_functions/output.py:8-92).  The device buffer (14 M N B doubles) is allocated by the first call. */
#define OSC2_N_NODAL_PROPS 14
int osc2_eval_nodal_properties(osc2_handle h);
int osc2_download_nodal_properties(osc2_handle h, double *chan);
/* `time` = conductor.cond_time[-1] of the step about to be taken. */
int osc2_eval_operating_conditions(osc2_handle h, double time);
/* One full time step: osc2_eval_operating_conditions(time) + osc2_step(dt, num_step). */
int osc2_advance(osc2_handle h, double time, double dt, int num_step);
/* Device -> host copy of the computed step inputs (parity tests); any member may be NULL. */
int osc2_download_step_inputs(osc2_handle h, double *fl_gauss, double *fl_node, double *sol_gauss,
                              double *sol_q, double *htc_ff, double *htc_fs, double *htc_ss, double *htc_es);

/* One reference step() for every conductor of the batch, on device-resident
 * inputs: Gauss matrices -> block rows + BC + row scaling -> no-pivot banded
 * LU (gredub) -> substitution (gbacsb) -> norms / EQTEIG.  Asynchronous. */
int osc2_step(osc2_handle h, double dt, int num_step);

/* Wait for the stream; returns OSC2_ERR_SINGULAR if any conductor hit a singular pivot in ANY step since the
 * state was uploaded (status is sticky: the first failing step is kept, like the reference's ValueError at that
 * step; the state after it is meaningless).  osc2_upload_state starts it over. */
int osc2_synchronize(osc2_handle h);

/* Any pointer may be NULL.  status[b] = 0 or -(row+1) of the first tiny pivot of the first failing step. */
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
