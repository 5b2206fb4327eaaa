/* scs_oracle.h -- TEST INFRASTRUCTURE ONLY.
 *
 * Plain-C, single-threaded CPU restatement of the SCS hot path (reference
 * cvxgrp/scs 3.2.11): CSC SpMV, Jacobi-PCG KKT solve, cone projections
 * (zero / LP / box / SOC / PSD), Anderson acceleration and the ADMM loop with
 * equilibration, residuals, stopping and adaptive scale.  Every function cites
 * the reference file:line it restates (see scs_oracle.c).
 *
 * PARITY PINNED: checked against the UNMODIFIED reference built from
 * /root/reference (oracle/_ref, `make -C oracle ref`) by
 * oracle/make_golden.py, whose outputs are committed under tests/golden/ and
 * re-checked by tests/test_oracle_cpu.py.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
 * reference legs may load this. The product (scs_b200/libscs_b200.so) never
 * links or calls it.
 */
#ifndef SCS_ORACLE_H
#define SCS_ORACLE_H
#ifdef __cplusplus
extern "C" {
#endif

typedef struct {
  double *x;
  int *i;
  int *p;
  int m, n;
} OrcMatrix; /* CSC, same layout as ScsMatrix */

typedef struct {
  int z, l;
  double *bu, *bl;
  int bsize;
  int *q;
  int qsize;
  int *s;
  int ssize;
} OrcCone;

typedef struct {
  int normalize;
  double scale;
  int adaptive_scale;
  double rho_x;
  int max_iters;
  double eps_abs, eps_rel, eps_infeas, alpha;
  int acceleration_lookback, acceleration_interval, acceleration_type_1;
  double acceleration_regularization, acceleration_relaxation;
} OrcSettings;

typedef struct {
  int iter, status_val, scale_updates;
  double pobj, dobj, res_pri, res_dual, gap, scale;
  long long cg_iters, lin_sys_solves;
  int accepted_accel_steps, rejected_accel_steps;
  double solve_time_ms, setup_time_ms;
} OrcInfo;

void orc_default_settings(OrcSettings *s);

/* SpMV (reference linsys/scs_matrix.c:161-203) */
void orc_accum_by_atrans(const OrcMatrix *A, const double *x, double *y);
void orc_accum_by_a(const OrcMatrix *A, const double *x, double *y);

/* KKT solve (reference linsys/cpu/indirect/private.c) */
typedef struct OrcLinSys OrcLinSys;
OrcLinSys *orc_linsys_init(const OrcMatrix *A, const double *diag_r);
void orc_linsys_free(OrcLinSys *w);
int orc_linsys_solve(OrcLinSys *w, double *b, const double *s, double tol);
void orc_linsys_update_diag_r(OrcLinSys *w, const double *diag_r);
int orc_linsys_last_cg_its(const OrcLinSys *w);

/* cone projection (reference src/cones.c) -- in place; r_y may be NULL */
typedef struct OrcConeWork OrcConeWork;
OrcConeWork *orc_cone_init(const OrcCone *k, int m);
void orc_cone_free(OrcConeWork *c);
int orc_proj_dual_cone(OrcConeWork *c, double *x, const double *r_y);

/* Anderson acceleration (reference src/aa.c) */
typedef struct OrcAa OrcAa;
OrcAa *orc_aa_init(int dim, int mem, int min_len, int type1, double regularization,
                   double relaxation, double safeguard_factor, double max_weight_norm,
                   int ir_max_steps);
double orc_aa_apply(OrcAa *a, double *f, const double *x);
int orc_aa_safeguard(OrcAa *a, double *f_new, double *x_new);
void orc_aa_reset(OrcAa *a);
void orc_aa_free(OrcAa *a);

/* whole solve (reference src/scs.c); x (n), y, s (m) are outputs */
int orc_solve(const OrcMatrix *A, const double *b, const double *c, const OrcCone *k,
              const OrcSettings *stgs, double *x, double *y, double *s, OrcInfo *info);

#ifdef __cplusplus
}
#endif
#endif
