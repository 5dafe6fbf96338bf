/* CPU ORACLE (test infrastructure, not product code) -- plain C restatement of
 * the OPENSC2 reference `step()` for one conductor.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline/reference
 * legs may link or call this.  The shipped CUDA path never does.
 *
 * Follows /root/reference/source_code/utility_functions/
 *   transient_solution_functions.py:186-599 (step), :602-698 (gredub),
 *   :701-808 (gbacsb), :810-900 (norms, eigenvalues)
 *   step_matrix_construction.py:62-1474 (builders, assembly, SYSMAT, Known)
 * and /root/reference/source_code/fluid_component.py:117-565 (boundary rows).
 *
 * Parity pin: tests/golden/step_*.npz were produced by the unmodified
 * reference step() in the build container (tests/golden/make_golden.py); the
 * oracle reproduces them bit for bit (tests/test_oracle_golden.py).
 */
#ifndef OSC2_STEP_ORACLE_H
#define OSC2_STEP_ORACLE_H

#ifdef __cplusplus
extern "C" {
#endif

typedef struct {
    int n_ch, n_sol;
    int nff, nfs, nss, nes;
    const double *ch_area;    /* [M] */
    const double *ch_dh;      /* [M] */
    const int *ch_forward;    /* [M] 1 = forward */
    const int *ch_intial;     /* [M] abs(INTIAL) in {1,2,3} */
    const double *sol_area;   /* [S] */
    const double *sol_costeta;/* [S] */
    const int *ff;            /* [nff][2] channel indices */
    const int *fs;            /* [nfs][2] channel, solid */
    const int *ss;            /* [nss][2] solid indices */
    const int *es;            /* [nes] solid index */
    const int *es_choice2;    /* [nes] HTC_choice == 2 */
    int is_rectangular;
    double height, width;
    double delta_p_min, k_loc, lambda_v;
    double theta;
    int method;               /* 0 = BE, 1 = CN */
} osc2o_topology;

typedef struct {
    int n_nodes;
    const double *dz;         /* [E] */
    const double *fl_gauss;   /* [9][M][E] v,p,T,rho,c0,Gruneisen,h,cv,f */
    const double *fl_node;    /* [4][M] v[0], v[-1], rho[0], rho[-1] */
    const double *fl_bc;      /* [6][M] PREINL,PREOUT,TEMINL,TEMOUT,MDTIN,MDTOUT */
    const double *sol_gauss;  /* [3][S][E] rho,cp,k */
    const double *sol_q;      /* [2][S][E][2] Q1,Q2 x [present, previous] */
    const double *peri_ff;    /* [2][nff][E] open, close */
    const double *htc_ff;     /* [2][nff][E] */
    const double *peri_fs;    /* [nfs][E] */
    const double *htc_fs;     /* [nfs][E] */
    const double *peri_ss;    /* [nss][E] */
    const double *htc_ss;     /* [nss][E] */
    const double *peri_es;    /* [nes][E] */
    const double *htc_es;     /* [nes][3][E] conv|side, bottom, top */
    const double *qsource;    /* [N][S] */
    double t_env, dt;
    int num_step;
    /* in/out */
    double *sysvar;           /* [Total] */
    double *syslod;           /* [Total][2] */
    /* out */
    double *k123;             /* [3][nff][E] or NULL */
    double *norm_solution;    /* [d] or NULL */
    double *norm_change;      /* [d] or NULL */
    double *eqteig;           /* [d] or NULL */
    double *sysmat;           /* [Full][Total] reference band layout, pre-solve; or NULL */
    double *known;            /* [Total] pre-solve; or NULL */
    int status;               /* 0 ok; -(row+1) singular pivot at `row` */
} osc2o_step_io;

/* One reference step() for one conductor.  Returns io->status. */
int osc2o_step(const osc2o_topology *topo, osc2o_step_io *io);

/* Same over `batch` independent conductors; arrays carry a leading batch axis
 * and the io pointers address slot 0.  n_threads <= 0: all online cores (pthreads). */
int osc2o_step_batch(const osc2o_topology *topo, const osc2o_step_io *io0, int batch,
                     int n_threads, int *status);

/* numpy's pairwise summation (np.sum on a strided 1-D double array). */
double osc2o_pairwise_sum(const double *a, long n, long stride);

#ifdef __cplusplus
}
#endif
#endif
