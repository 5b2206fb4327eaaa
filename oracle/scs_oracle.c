/* scs_oracle.c -- TEST INFRASTRUCTURE ONLY (see scs_oracle.h).
 *
 * Plain C, strictly sequential CPU restatement of the SCS hot path for the
 * in-scope cones (zero, LP, box, SOC, PSD) and P = 0.  Citations are into
 * /root/reference (cvxgrp/scs 3.2.11).  Dense kernels the reference delegates
 * to BLAS/LAPACK (ddot, dnrm2, dsyevr, dsyrk, dgeqp3, dormqr, dgesv, dtrsv) are
 * restated with textbook algorithms (sequential sums, cyclic Jacobi
 * eigen-solver, Householder QR with column pivoting, LU with partial pivoting);
 * LAPACK itself is an unvendored dependency of the reference (any BLAS; the
 * oracle/_ref build links OpenBLAS 0.3.15).
 */
#include "scs_oracle.h"
#include <float.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#define MAXV(a, b) (((a) > (b)) ? (a) : (b))
#define MINV(a, b) (((a) < (b)) ? (a) : (b))
#define DIV_EPS (1e-18)
#define SAFEDIV(X, Y) ((Y) < DIV_EPS ? ((X) / DIV_EPS) : (X) / (Y))

static double now_ms(void) {
  struct timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  return 1e3 * (double)ts.tv_sec + 1e-6 * (double)ts.tv_nsec;
}

/* include/glbopts.h:35-50 */
void orc_default_settings(OrcSettings *s) {
  s->normalize = 1; s->scale = 0.1; s->adaptive_scale = 1; s->rho_x = 1e-6;
  s->max_iters = 100000; s->eps_abs = 1e-4; s->eps_rel = 1e-4; s->eps_infeas = 1e-7;
  s->alpha = 1.5; s->acceleration_lookback = 10; s->acceleration_interval = 10;
  s->acceleration_type_1 = 1; s->acceleration_regularization = 1e-8;
  s->acceleration_relaxation = 1.0;
}

/* ---- src/linalg.c:36-102 (plain-C branch) ---- */
static double dot(const double *x, const double *y, int n) {
  double s = 0.0; int i;
  for (i = 0; i < n; ++i) s += x[i] * y[i];
  return s;
}
static double norm_inf(const double *a, int n) {
  double mx = 0.0; int i;
  for (i = 0; i < n; ++i) { double t = fabs(a[i]); if (t > mx) mx = t; }
  return mx;
}
static double norm_2(const double *a, int n) { return sqrt(dot(a, a, n)); }

/* ---- linsys/scs_matrix.c:161-186 ---- */
void orc_accum_by_atrans(const OrcMatrix *A, const double *x, double *y) {
  int j, p;
  for (j = 0; j < A->n; j++) {
    double yj = y[j];
    for (p = A->p[j]; p < A->p[j + 1]; p++) yj += A->x[p] * x[A->i[p]];
    y[j] = yj;
  }
}
/* ---- linsys/scs_matrix.c:188-203 ---- */
void orc_accum_by_a(const OrcMatrix *A, const double *x, double *y) {
  int j, p;
  for (j = 0; j < A->n; j++)
    for (p = A->p[j]; p < A->p[j + 1]; p++) y[A->i[p]] += A->x[p] * x[j];
}

/* ======================= KKT solve: linsys/cpu/indirect/private.c ======================= */
struct OrcLinSys {
  int n, m;
  const OrcMatrix *A;
  OrcMatrix At;
  const double *diag_r;
  double *p, *r, *Gp, *z, *M, *tmp;
  int last_its;
  long long tot_its;
};

/* private.c:7-46 */
static void transpose(const OrcMatrix *A, OrcMatrix *C) {
  int m = A->m, n = A->n, nnz = A->p[n], i, j, k;
  int *z = (int *)calloc((size_t)m + 1, sizeof(int));
  C->m = n; C->n = m;
  C->p = (int *)calloc((size_t)m + 1, sizeof(int));
  C->i = (int *)malloc(((size_t)nnz + 1) * sizeof(int));
  C->x = (double *)malloc(((size_t)nnz + 1) * sizeof(double));
  for (k = 0; k < nnz; k++) z[A->i[k]]++;
  for (i = 0; i < m; i++) C->p[i + 1] = C->p[i] + z[i];
  for (i = 0; i < m; i++) z[i] = C->p[i];
  for (j = 0; j < n; j++)
    for (k = A->p[j]; k < A->p[j + 1]; k++) {
      int q = z[A->i[k]]++;
      C->i[q] = j;
      C->x[q] = A->x[k];
    }
  free(z);
}
/* private.c:50-82 */
static void set_preconditioner(OrcLinSys *w) {
  int i, k;
  for (i = 0; i < w->n; ++i) {
    double acc = w->diag_r[i];
    for (k = w->A->p[i]; k < w->A->p[i + 1]; ++k)
      acc += w->A->x[k] * w->A->x[k] / w->diag_r[w->n + w->A->i[k]];
    w->M[i] = 1. / acc;
  }
}
OrcLinSys *orc_linsys_init(const OrcMatrix *A, const double *diag_r) {
  OrcLinSys *w = (OrcLinSys *)calloc(1, sizeof(OrcLinSys));
  w->n = A->n; w->m = A->m; w->A = A; w->diag_r = diag_r;
  transpose(A, &w->At);
  w->p = (double *)calloc((size_t)A->n, 8); w->r = (double *)calloc((size_t)A->n, 8);
  w->Gp = (double *)calloc((size_t)A->n, 8); w->z = (double *)calloc((size_t)A->n, 8);
  w->M = (double *)calloc((size_t)A->n, 8); w->tmp = (double *)calloc((size_t)A->m, 8);
  set_preconditioner(w);
  return w;
}
void orc_linsys_free(OrcLinSys *w) {
  if (!w) return;
  free(w->At.p); free(w->At.i); free(w->At.x);
  free(w->p); free(w->r); free(w->Gp); free(w->z); free(w->M); free(w->tmp); free(w);
}
void orc_linsys_update_diag_r(OrcLinSys *w, const double *diag_r) { /* private.c:327-331 */
  w->diag_r = diag_r;
  set_preconditioner(w);
}
int orc_linsys_last_cg_its(const OrcLinSys *w) { return w->last_its; }

/* private.c:106-119: y = (R_x + A' R_y^-1 A) x */
static void mat_vec(OrcLinSys *w, const double *x, double *y) {
  int i;
  memset(w->tmp, 0, (size_t)w->m * 8);
  memset(y, 0, (size_t)w->n * 8);
  orc_accum_by_atrans(&w->At, x, w->tmp);            /* tmp = A x */
  for (i = 0; i < w->m; ++i) w->tmp[i] /= w->diag_r[w->n + i];
  orc_accum_by_atrans(w->A, w->tmp, y);              /* y += A' tmp */
  for (i = 0; i < w->n; ++i) y[i] += w->diag_r[i] * x[i];
}
/* private.c:133-217 */
static int pcg(OrcLinSys *w, const double *s, double *b, int max_its, double tol) {
  int i, k, n = w->n;
  double ztr, ztr_prev, alpha;
  double *p = w->p, *Gp = w->Gp, *r = w->r, *z = w->z, *M = w->M;
  if (!s) {
    memcpy(r, b, (size_t)n * 8);
    memset(b, 0, (size_t)n * 8);
  } else {
    mat_vec(w, s, r);
    for (k = 0; k < n; ++k) r[k] += -1. * b[k];
    for (k = 0; k < n; ++k) r[k] *= -1.;
    memcpy(b, s, (size_t)n * 8);
  }
  if (norm_inf(r, n) < MAXV(tol, 1e-12)) return 0;
  for (k = 0; k < n; ++k) z[k] = r[k] * M[k];
  ztr = dot(z, r, n);
  memcpy(p, z, (size_t)n * 8);
  for (i = 0; i < max_its; ++i) {
    double norm_r = 0.0, beta;
    mat_vec(w, p, Gp);
    alpha = ztr / dot(p, Gp, n);
    for (k = 0; k < n; ++k) b[k] += alpha * p[k];
    for (k = 0; k < n; ++k) r[k] += -alpha * Gp[k];
    ztr_prev = ztr;
    ztr = 0.0;
    for (k = 0; k < n; ++k) {
      double rk = r[k], zk = rk * M[k], ark = fabs(rk);
      z[k] = zk;
      ztr += zk * rk;
      if (ark > norm_r) norm_r = ark;
    }
    if (norm_r < tol) return i + 1;
    if (ztr_prev == 0.) break;
    beta = ztr / ztr_prev;
    for (k = 0; k < n; ++k) p[k] = z[k] + beta * p[k];
  }
  return i;
}
/* private.c:284-324 */
int orc_linsys_solve(OrcLinSys *w, double *b, const double *s, double tol) {
  int i, its, n = w->n, m = w->m;
  if (norm_inf(b, n + m) <= 1e-12) {
    memset(b, 0, ((size_t)n + m) * 8);
    w->last_its = 0;
    return 0;
  }
  memcpy(w->tmp, b + n, (size_t)m * 8);
  for (i = 0; i < m; ++i) w->tmp[i] /= w->diag_r[n + i];
  orc_accum_by_atrans(w->A, w->tmp, b);
  its = pcg(w, s, b, 10 * n, tol);
  for (i = 0; i < m; ++i) b[n + i] *= -1.;
  orc_accum_by_atrans(&w->At, b, b + n);
  for (i = 0; i < m; ++i) b[n + i] /= w->diag_r[n + i];
  w->last_its = its;
  w->tot_its += its;
  return 0;
}

/* ======================= cones: src/cones.c ======================= */
struct OrcConeWork {
  OrcCone k; /* shallow copy; bl/bu owned copies */
  int m;
  double *s;
  double box_t;
  double *Xs, *Z, *e; /* PSD scratch (max order) */
  int nmax;
};

OrcConeWork *orc_cone_init(const OrcCone *k, int m) { /* cones.c:1498-1540 */
  int i, nmax = 1;
  OrcConeWork *c = (OrcConeWork *)calloc(1, sizeof(OrcConeWork));
  c->k = *k;
  c->m = m;
  c->s = (double *)calloc((size_t)(m > 0 ? m : 1), 8);
  c->box_t = 1.0; /* cones.c:1562 */
  for (i = 0; i < k->ssize; ++i) nmax = MAXV(nmax, k->s[i]);
  c->nmax = nmax;
  c->Xs = (double *)calloc((size_t)nmax * nmax, 8);
  c->Z = (double *)calloc((size_t)nmax * nmax, 8);
  c->e = (double *)calloc((size_t)nmax, 8);
  return c;
}
void orc_cone_free(OrcConeWork *c) {
  if (!c) return;
  free(c->s); free(c->Xs); free(c->Z); free(c->e); free(c);
}

/* cones.c:1250-1279 */
static void proj_soc(double *x, int q) {
  double v1, s, alpha;
  int i;
  if (q <= 0) return;
  if (q == 1) { x[0] = MAXV(x[0], 0.); return; }
  v1 = x[0];
  s = norm_2(x + 1, q - 1);
  alpha = (s + v1) / 2.0;
  if (s <= v1) return;
  if (s <= -v1) { memset(x, 0, (size_t)q * 8); return; }
  x[0] = alpha;
  for (i = 1; i < q; ++i) x[i] *= alpha / s;
}

/* cones.c:1182-1245 */
static double proj_box_cone(double *tx, const double *bl, const double *bu, int bsize, double t_wm,
                            const double *r_box) {
  double *x = tx + 1, gt, ht, t = t_wm, t_prev, rho_t = 1.0;
  int iter, j;
  if (bsize == 1) { tx[0] = MAXV(tx[0], 0.0); return tx[0]; }
  if (r_box) rho_t = 1.0 / r_box[0];
  for (iter = 0; iter < 25; iter++) {
    t_prev = t;
    gt = rho_t * (t - tx[0]);
    ht = rho_t;
    for (j = 0; j < bsize - 1; j++) {
      const double r = r_box ? 1.0 / r_box[1 + j] : 1.0;
      if (x[j] > t * bu[j]) {
        gt += r * (t * bu[j] - x[j]) * bu[j];
        ht += r * bu[j] * bu[j];
      } else if (x[j] < t * bl[j]) {
        gt += r * (t * bl[j] - x[j]) * bl[j];
        ht += r * bl[j] * bl[j];
      }
    }
    t = MAXV(t - gt / MAXV(ht, 1e-8), 0.0);
    if (fabs(gt / MAXV(ht, 1e-6)) < 1e-12 * MAXV(t, 1.) || fabs(t - t_prev) < 1e-11 * MAXV(t, 1.))
      break;
  }
  for (j = 0; j < bsize - 1; j++) {
    if (x[j] > t * bu[j]) x[j] = t * bu[j];
    else if (x[j] < t * bl[j]) x[j] = t * bl[j];
  }
  tx[0] = t;
  return t;
}

/* symmetric eigen-decomposition A = Z diag(e) Z' by cyclic Jacobi (stands in for LAPACK dsyevr,
 * cones.c:1028-1031). A is n x n column-major, destroyed; Z gets the eigenvectors (columns). */
static void jacobi_eig(double *A, int n, double *Z, double *e) {
  int i, j, k, sweep;
  for (i = 0; i < n * n; ++i) Z[i] = 0.0;
  for (i = 0; i < n; ++i) Z[i + i * n] = 1.0;
  for (sweep = 0; sweep < 100; ++sweep) {
    double off = 0.0, diag = 0.0;
    for (j = 0; j < n; ++j)
      for (i = 0; i < n; ++i) {
        if (i != j) off += A[i + j * n] * A[i + j * n];
        else diag += A[i + j * n] * A[i + j * n];
      }
    if (off <= 1e-32 * (diag + off) || off == 0.0) break;
    for (j = 0; j < n - 1; ++j)
      for (k = j + 1; k < n; ++k) {
        const double apq = A[j + k * n];
        double theta, t, c, s2;
        if (fabs(apq) < 1e-300) continue;
        theta = (A[k + k * n] - A[j + j * n]) / (2.0 * apq);
        t = (theta >= 0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
        c = 1.0 / sqrt(t * t + 1.0);
        s2 = t * c;
        for (i = 0; i < n; ++i) { /* columns j,k */
          const double aij = A[i + j * n], aik = A[i + k * n];
          A[i + j * n] = c * aij - s2 * aik;
          A[i + k * n] = s2 * aij + c * aik;
        }
        for (i = 0; i < n; ++i) { /* rows j,k */
          const double aji = A[j + i * n], aki = A[k + i * n];
          A[j + i * n] = c * aji - s2 * aki;
          A[k + i * n] = s2 * aji + c * aki;
        }
        for (i = 0; i < n; ++i) {
          const double zij = Z[i + j * n], zik = Z[i + k * n];
          Z[i + j * n] = c * zij - s2 * zik;
          Z[i + k * n] = s2 * zij + c * zik;
        }
      }
  }
  for (i = 0; i < n; ++i) e[i] = A[i + i * n];
}

/* cones.c:999-1067 */
static void proj_psd(double *X, int n, OrcConeWork *c) {
  int i, j, k;
  const double sqrt2 = sqrt(2.0), sqrt2_inv = 1.0 / sqrt(2.0);
  double *Xs = c->Xs, *Z = c->Z, *e = c->e;
  if (n == 0) return;
  if (n == 1) { X[0] = MAXV(X[0], 0.); return; }
  /* unpack lower triangle (column-major packed) to full symmetric, diagonal * sqrt2 */
  for (j = 0; j < n; ++j) {
    const double *col = X + (j * n - ((j - 1) * j) / 2);
    for (i = j; i < n; ++i) {
      Xs[i + j * n] = col[i - j];
      Xs[j + i * n] = col[i - j];
    }
    Xs[j + j * n] *= sqrt2;
  }
  jacobi_eig(Xs, n, Z, e);
  /* Xs = sum_{e_k > 0} e_k z_k z_k'   (dsyrk on sqrt(e) * Z, cones.c:1036-1053) */
  for (j = 0; j < n; ++j)
    for (i = j; i < n; ++i) {
      double acc = 0.0;
      for (k = 0; k < n; ++k)
        if (e[k] > 0) acc += (sqrt(e[k]) * Z[i + k * n]) * (sqrt(e[k]) * Z[j + k * n]);
      Xs[i + j * n] = acc;
    }
  for (j = 0; j < n; ++j) {
    double *col = X + (j * n - ((j - 1) * j) / 2);
    Xs[j + j * n] *= sqrt2_inv;
    for (i = j; i < n; ++i) col[i - j] = Xs[i + j * n];
  }
}

/* cones.c:1340-1494 (zero, LP, box, SOC, PSD branches) */
static void proj_cone(double *x, OrcConeWork *c, const double *r_y) {
  const OrcCone *k = &c->k;
  int i, count = 0;
  if (k->z) { memset(x, 0, (size_t)k->z * 8); count += k->z; }
  if (k->l) {
    for (i = count; i < count + k->l; ++i) x[i] = MAXV(x[i], 0.0);
    count += k->l;
  }
  if (k->bsize) {
    c->box_t = proj_box_cone(x + count, k->bl, k->bu, k->bsize, c->box_t, r_y ? r_y + count : NULL);
    count += k->bsize;
  }
  for (i = 0; i < k->qsize; ++i) { proj_soc(x + count, k->q[i]); count += k->q[i]; }
  for (i = 0; i < k->ssize; ++i) {
    proj_psd(x + count, k->s[i], c);
    count += (k->s[i] * (k->s[i] + 1)) / 2;
  }
}
/* cones.c:1552-1596 */
int orc_proj_dual_cone(OrcConeWork *c, double *x, const double *r_y) {
  int i;
  memcpy(c->s, x, (size_t)c->m * 8);
  if (r_y) for (i = 0; i < c->m; ++i) x[i] *= -r_y[i];
  else for (i = 0; i < c->m; ++i) x[i] = -x[i];
  proj_cone(x, c, r_y);
  if (r_y) for (i = 0; i < c->m; ++i) x[i] = x[i] / r_y[i] + c->s[i];
  else for (i = 0; i < c->m; ++i) x[i] += c->s[i];
  return 0;
}

/* ======================= Anderson acceleration: src/aa.c ======================= */
struct OrcAa {
  int type1, mem, min_len, dim, iter, success, ir_max_steps;
  double relaxation, regularization, safeguard_factor, max_weight_norm, norm_g;
  double *x, *f, *g, *g_prev, *Y, *S, *D, *nrm_s, *nrm_y, *A_aug, *B_aug, *c_aug, *work;
};
OrcAa *orc_aa_init(int dim, int mem, int min_len, int type1, double regularization,
                   double relaxation, double safeguard_factor, double max_weight_norm,
                   int ir_max_steps) { /* aa.c:657-820 */
  OrcAa *a = (OrcAa *)calloc(1, sizeof(OrcAa));
  int memc = MINV(mem, dim);
  size_t aug = (size_t)dim + memc;
  a->type1 = type1; a->dim = dim; a->mem = memc; a->min_len = memc > 0 ? MINV(min_len, memc) : 0;
  a->regularization = regularization; a->relaxation = relaxation;
  a->safeguard_factor = safeguard_factor; a->max_weight_norm = max_weight_norm;
  a->ir_max_steps = ir_max_steps;
  if (memc <= 0) return a;
  a->x = (double *)calloc((size_t)dim, 8); a->f = (double *)calloc((size_t)dim, 8);
  a->g = (double *)calloc((size_t)dim, 8); a->g_prev = (double *)calloc((size_t)dim, 8);
  a->Y = (double *)calloc((size_t)dim * memc, 8); a->S = (double *)calloc((size_t)dim * memc, 8);
  a->D = (double *)calloc((size_t)dim * memc, 8);
  a->nrm_s = (double *)calloc((size_t)memc, 8); a->nrm_y = (double *)calloc((size_t)memc, 8);
  a->A_aug = (double *)calloc(aug * memc, 8); a->B_aug = (double *)calloc(aug * memc, 8);
  a->c_aug = (double *)calloc(aug, 8); a->work = (double *)calloc((size_t)MAXV(dim, memc), 8);
  return a;
}
void orc_aa_free(OrcAa *a) {
  if (!a) return;
  free(a->x); free(a->f); free(a->g); free(a->g_prev); free(a->Y); free(a->S); free(a->D);
  free(a->nrm_s); free(a->nrm_y); free(a->A_aug); free(a->B_aug); free(a->c_aug); free(a->work);
  free(a);
}
void orc_aa_reset(OrcAa *a) { /* aa.c:934-964 */
  a->iter = 0; a->success = 0; a->norm_g = 0;
  if (a->nrm_s) memset(a->nrm_s, 0, (size_t)a->mem * 8);
  if (a->nrm_y) memset(a->nrm_y, 0, (size_t)a->mem * 8);
}
static double frob(const double *nrm, int mem) { /* aa.c:257-270 */
  double m = 0, ss = 0; int i;
  for (i = 0; i < mem; ++i) if (nrm[i] > m) m = nrm[i];
  if (m == 0) return 0;
  for (i = 0; i < mem; ++i) { double t = nrm[i] / m; ss += t * t; }
  return m * sqrt(ss);
}
/* Householder QR with column pivoting of the first `len` columns of H (rows x len), applying
 * the reflectors also to `nextra` extra column blocks (stand-in for dgeqp3 + dormqr). */
static void pivoted_qr(double *H, int rows, int len, int *jpvt, double *E1, int n1, double *E2,
                       int n2) {
  int i, j, c;
  for (j = 0; j < len; ++j) jpvt[j] = j;
  for (j = 0; j < len && j < rows; ++j) {
    int best = j; double bestn = -1.0, sg = 0.0, alpha, nrm, beta, tau, scale;
    double *cj;
    for (c = j; c < len; ++c) {
      double s = 0.0;
      for (i = j; i < rows; ++i) s += H[(size_t)c * rows + i] * H[(size_t)c * rows + i];
      if (s > bestn) { bestn = s; best = c; }
    }
    if (best != j) {
      for (i = 0; i < rows; ++i) {
        double t = H[(size_t)j * rows + i];
        H[(size_t)j * rows + i] = H[(size_t)best * rows + i];
        H[(size_t)best * rows + i] = t;
      }
      i = jpvt[j]; jpvt[j] = jpvt[best]; jpvt[best] = i;
    }
    cj = H + (size_t)j * rows;
    for (i = j + 1; i < rows; ++i) sg += cj[i] * cj[i];
    if (sg == 0.0) continue;
    alpha = cj[j];
    nrm = sqrt(alpha * alpha + sg);
    beta = alpha >= 0.0 ? -nrm : nrm;
    tau = (beta - alpha) / beta;
    scale = 1.0 / (alpha - beta);
    for (i = j + 1; i < rows; ++i) cj[i] *= scale;
#define APPLY(MAT, NC)                                                           \
    for (c = 0; c < (NC); ++c) {                                                 \
      double *cc = (MAT) + (size_t)c * rows, w = cc[j];                          \
      for (i = j + 1; i < rows; ++i) w += cj[i] * cc[i];                         \
      w *= tau;                                                                  \
      cc[j] -= w;                                                                \
      for (i = j + 1; i < rows; ++i) cc[i] -= w * cj[i];                         \
    }
    APPLY(H + (size_t)(j + 1) * rows, len - j - 1)
    if (E1) { APPLY(E1, n1) }
    if (E2) { APPLY(E2, n2) }
#undef APPLY
    cj[j] = beta;
    for (i = j + 1; i < rows; ++i) cj[i] = 0.0;
  }
}
static int lu_factor(double *A, int n, int ld, int *ipiv) {
  int i, k, c;
  for (k = 0; k < n; ++k) {
    int p = k; double mx = fabs(A[(size_t)k * ld + k]), piv;
    for (i = k + 1; i < n; ++i) if (fabs(A[(size_t)k * ld + i]) > mx) { mx = fabs(A[(size_t)k * ld + i]); p = i; }
    ipiv[k] = p;
    if (mx == 0.0) return k + 1;
    if (p != k) for (c = 0; c < n; ++c) { double t = A[(size_t)c * ld + k]; A[(size_t)c * ld + k] = A[(size_t)c * ld + p]; A[(size_t)c * ld + p] = t; }
    piv = A[(size_t)k * ld + k];
    for (i = k + 1; i < n; ++i) A[(size_t)k * ld + i] /= piv;
    for (c = k + 1; c < n; ++c) { const double akc = A[(size_t)c * ld + k]; for (i = k + 1; i < n; ++i) A[(size_t)c * ld + i] -= A[(size_t)k * ld + i] * akc; }
  }
  return 0;
}
static void lu_solve(const double *LU, int n, int ld, const int *ipiv, double *b) {
  int i, k;
  for (k = 0; k < n; ++k) {
    if (ipiv[k] != k) { double t = b[k]; b[k] = b[ipiv[k]]; b[ipiv[k]] = t; }
    for (i = k + 1; i < n; ++i) b[i] -= LU[(size_t)k * ld + i] * b[k];
  }
  for (k = n - 1; k >= 0; --k) {
    b[k] /= LU[(size_t)k * ld + k];
    for (i = 0; i < k; ++i) b[i] -= LU[(size_t)k * ld + i] * b[k];
  }
}
/* aa.c:422-652 */
static double aa_solve(OrcAa *a, double *f, int len) {
  const int dim = a->dim, mem = a->mem, rows = dim + mem;
  const double *A_src = a->type1 ? a->S : a->Y;
  double r, sqrt_r, aa_norm;
  int i, k, c, rank = 0, info = 0;
  int jpvt[64], ipiv[64];
  double W[64 * 64], W0[64 * 64], gamma_red[64], c_top[64], ir[64], gamma[64];
  if (a->regularization > 0) {
    const double ny = frob(a->nrm_y, mem), na = a->type1 ? frob(a->nrm_s, mem) : ny;
    r = a->regularization * na * ny;
  } else if (a->regularization < 0) r = -a->regularization;
  else r = 0.0;
  sqrt_r = r > 0 ? sqrt(r) : 0.0;
  for (i = 0; i < len; ++i) { /* build_augmented, aa.c:297-307 */
    double *col = a->A_aug + (size_t)i * rows;
    memcpy(col, A_src + (size_t)i * dim, (size_t)dim * 8);
    memset(col + dim, 0, (size_t)mem * 8);
    col[dim + i] = sqrt_r;
  }
  memcpy(a->c_aug, a->g, (size_t)dim * 8);
  memset(a->c_aug + dim, 0, (size_t)mem * 8);
  if (a->type1)
    for (i = 0; i < len; ++i) { /* unpermuted [Y; sqrt(r) I]; permutation applied when reading */
      double *col = a->B_aug + (size_t)i * rows;
      memcpy(col, a->Y + (size_t)i * dim, (size_t)dim * 8);
      memset(col + dim, 0, (size_t)mem * 8);
      col[dim + i] = sqrt_r;
    }
  pivoted_qr(a->A_aug, rows, len, jpvt, a->c_aug, 1, a->type1 ? a->B_aug : NULL, len);
  {
    const double r11 = fabs(a->A_aug[0]);
    if (r11 > 0) {
      const double tol = r11 * (double)len * DBL_EPSILON;
      for (rank = 0; rank < len; ++rank) if (fabs(a->A_aug[(size_t)rank * rows + rank]) < tol) break;
    }
    if (rank == 0) info = 1;
  }
  if (info == 0) {
    for (i = 0; i < rank; ++i) c_top[i] = a->c_aug[i];
    if (a->type1) {
      double prev = 0.0;
      for (i = 0; i < rank; ++i)
        for (k = 0; k < rank; ++k) W[i * 64 + k] = W0[i * 64 + k] = a->B_aug[(size_t)jpvt[i] * rows + k];
      memcpy(gamma_red, c_top, (size_t)rank * 8);
      info = lu_factor(W, rank, 64, ipiv);
      if (info == 0) {
        lu_solve(W, rank, 64, ipiv, gamma_red);
        for (k = 0; k < a->ir_max_steps; ++k) {
          double dn = 0.0;
          for (i = 0; i < rank; ++i) {
            double s = c_top[i];
            for (c = 0; c < rank; ++c) s -= W0[c * 64 + i] * gamma_red[c];
            ir[i] = s;
          }
          lu_solve(W, rank, 64, ipiv, ir);
          for (i = 0; i < rank; ++i) dn += ir[i] * ir[i];
          dn = sqrt(dn);
          for (i = 0; i < rank; ++i) gamma_red[i] += ir[i];
          if (k > 0 && dn >= 0.5 * prev) break;
          prev = dn;
        }
      }
    } else {
      double prev = 0.0;
      memcpy(gamma_red, c_top, (size_t)rank * 8);
      for (k = rank - 1; k >= 0; --k) {
        gamma_red[k] /= a->A_aug[(size_t)k * rows + k];
        for (i = 0; i < k; ++i) gamma_red[i] -= a->A_aug[(size_t)k * rows + i] * gamma_red[k];
      }
      for (c = 0; c < a->ir_max_steps; ++c) {
        double dn = 0.0;
        for (i = 0; i < rank; ++i) {
          double s = 0.0;
          for (k = i; k < rank; ++k) s += a->A_aug[(size_t)k * rows + i] * gamma_red[k];
          ir[i] = c_top[i] - s;
        }
        for (k = rank - 1; k >= 0; --k) {
          ir[k] /= a->A_aug[(size_t)k * rows + k];
          for (i = 0; i < k; ++i) ir[i] -= a->A_aug[(size_t)k * rows + i] * ir[k];
        }
        for (i = 0; i < rank; ++i) dn += ir[i] * ir[i];
        dn = sqrt(dn);
        for (i = 0; i < rank; ++i) gamma_red[i] += ir[i];
        if (c > 0 && dn >= 0.5 * prev) break;
        prev = dn;
      }
    }
    if (info == 0) {
      memset(gamma, 0, (size_t)len * 8);
      for (i = 0; i < rank; ++i) gamma[jpvt[i]] = gamma_red[i];
    }
  }
  aa_norm = -1.0;
  if (info == 0) { aa_norm = 0.0; for (i = 0; i < len; ++i) aa_norm += gamma[i] * gamma[i]; aa_norm = sqrt(aa_norm); }
  if (info != 0 || !isfinite(aa_norm) || aa_norm >= a->max_weight_norm) {
    a->success = 0;
    orc_aa_reset(a);
    if (!isfinite(aa_norm)) aa_norm = -1.0;
    return aa_norm < 0 ? aa_norm : -aa_norm;
  }
  for (c = 0; c < len; ++c) for (i = 0; i < dim; ++i) f[i] -= a->D[(size_t)c * dim + i] * gamma[c];
  if (a->relaxation != 1.0) { /* aa.c:393-408; x_work = x of this call = a->x after the advance */
    for (i = 0; i < dim; ++i) a->work[i] = a->x[i];
    for (c = 0; c < len; ++c) for (i = 0; i < dim; ++i) a->work[i] -= a->S[(size_t)c * dim + i] * gamma[c];
    for (i = 0; i < dim; ++i) f[i] = a->relaxation * f[i] + (1. - a->relaxation) * a->work[i];
  }
  a->success = 1;
  return aa_norm;
}
double orc_aa_apply(OrcAa *a, double *f, const double *x) { /* aa.c:822-854, 310-390 */
  double aa_norm = 0;
  int i, len = MINV(a->iter, a->mem), dim = a->dim;
  a->success = 0;
  if (a->mem <= 0) return 0;
  if (a->iter == 0) {
    for (i = 0; i < dim; ++i) { a->x[i] = x[i]; a->f[i] = f[i]; a->g_prev[i] = x[i] - f[i]; }
    a->iter++;
    return 0;
  }
  {
    const int idx = (a->iter - 1) % a->mem;
    double *sc = a->S + (size_t)idx * dim, *dc = a->D + (size_t)idx * dim, *yc = a->Y + (size_t)idx * dim;
    for (i = 0; i < dim; ++i) {
      sc[i] = x[i] - a->x[i];
      dc[i] = f[i] - a->f[i];
      a->g[i] = x[i] - f[i];
      yc[i] = a->g[i] - a->g_prev[i];
    }
    a->nrm_s[idx] = norm_2(sc, dim);
    a->nrm_y[idx] = norm_2(yc, dim);
    memcpy(a->x, x, (size_t)dim * 8);
    memcpy(a->f, f, (size_t)dim * 8);
    memcpy(a->g_prev, a->g, (size_t)dim * 8);
    a->norm_g = norm_2(a->g, dim);
  }
  if (a->iter >= a->min_len) aa_norm = aa_solve(a, f, len);
  a->iter++;
  return aa_norm;
}
int orc_aa_safeguard(OrcAa *a, double *f_new, double *x_new) { /* aa.c:856-901 */
  int i; double nd = 0.0;
  if (a->mem <= 0 || !a->success) return 0;
  a->success = 0;
  for (i = 0; i < a->dim; ++i) { double d = x_new[i] - f_new[i]; nd += d * d; }
  nd = sqrt(nd);
  if (nd > a->safeguard_factor * a->norm_g) {
    memcpy(f_new, a->f, (size_t)a->dim * 8);
    memcpy(x_new, a->x, (size_t)a->dim * 8);
    orc_aa_reset(a);
    return -1;
  }
  return 0;
}

/* ======================= equilibration: linsys/scs_matrix.c:229-496, src/normalize.c ======================= */
static double lim(double x) { x = x < 1e-4 ? 1.0 : x; return x > 1e4 ? 1e4 : x; }
static void enforce(const int *bnd, int nb, double *v, int mean) {
  int i, j, count = bnd[0];
  for (i = 1; i < nb; ++i) {
    double w = 0.0;
    if (!mean) { for (j = count; j < count + bnd[i]; ++j) w = MAXV(w, fabs(v[j])); }
    else { for (j = count; j < count + bnd[i]; ++j) w += v[j]; w = bnd[i] > 0 ? w / bnd[i] : 0.0; }
    for (j = count; j < count + bnd[i]; ++j) v[j] = w;
    count += bnd[i];
  }
}
static void equilibrate(OrcMatrix *A, const int *bnd, int nb, double *D, double *E) {
  int pass, i, j, m = A->m, n = A->n;
  double *Dt = (double *)calloc((size_t)m, 8), *Et = (double *)calloc((size_t)n, 8);
  for (i = 0; i < m; ++i) D[i] = 1.0;
  for (i = 0; i < n; ++i) E[i] = 1.0;
  for (pass = 0; pass < 26; ++pass) {
    const int l2 = pass == 25;
    for (i = 0; i < m; ++i) Dt[i] = 0.0;
    for (i = 0; i < n; ++i)
      for (j = A->p[i]; j < A->p[i + 1]; ++j) {
        if (l2) Dt[A->i[j]] += A->x[j] * A->x[j];
        else Dt[A->i[j]] = MAXV(Dt[A->i[j]], fabs(A->x[j]));
      }
    if (l2) for (i = 0; i < m; ++i) Dt[i] = sqrt(Dt[i]);
    enforce(bnd, nb, Dt, l2);
    for (i = 0; i < m; ++i) { Dt[i] = sqrt(lim(Dt[i])); Dt[i] = SAFEDIV(1.0, Dt[i]); }
    for (i = 0; i < n; ++i) {
      double e = 0.0;
      for (j = A->p[i]; j < A->p[i + 1]; ++j) {
        if (l2) e += A->x[j] * A->x[j];
        else e = MAXV(e, fabs(A->x[j]));
      }
      if (l2) e = sqrt(e);
      Et[i] = sqrt(lim(e));
      Et[i] = SAFEDIV(1.0, Et[i]);
    }
    for (i = 0; i < n; ++i) {
      const double ei = Et[i];
      for (j = A->p[i]; j < A->p[i + 1]; ++j) A->x[j] *= Dt[A->i[j]] * ei;
    }
    for (i = 0; i < m; ++i) D[i] *= Dt[i];
    for (i = 0; i < n; ++i) E[i] *= Et[i];
  }
  free(Dt); free(Et);
}

/* ======================= ADMM driver: src/scs.c ======================= */
static double root_plus_coeffs(double a, double b, double c) { /* scs.c:689-708 */
  double rad, sq, q;
  if (!isfinite(a) || !isfinite(b) || !isfinite(c) || a <= 0.) return NAN;
  rad = b * b - 4 * a * c;
  if (!isfinite(rad)) return NAN;
  if (rad < 0.) return -b / (2 * a);
  sq = sqrt(rad);
  if (b <= 0.) return (-b + sq) / (2 * a);
  q = -0.5 * (b + sq);
  return q != 0. ? c / q : 0.;
}

/* populate_residual_struct + unnormalize_residuals + compute_residuals, scs.c:463-607 */
#define RESIDUALS()                                                                                \
  do {                                                                                             \
    const double pd_ = sigma * sigma, inv_ = 1.0 / sigma;                                          \
    tau = fabs(u[l - 1]); kap = fabs(rsk[l - 1]);                                                  \
    memset(ax, 0, (size_t)m * 8); orc_accum_by_a(&A, u, ax);                                       \
    memset(aty, 0, (size_t)n * 8); orc_accum_by_atrans(&A, u + n, aty);                            \
    nm_axsb_n = 0; nm_axsb_o = 0; nm_ax_o = 0; nm_s_o = 0; nm_axs_o = 0;                           \
    for (i = 0; i < m; ++i) {                                                                      \
      const double d_ = D ? D[i] : 1.0, f_ = inv_ / d_, axs_ = ax[i] + rsk[n + i],                 \
                   rr_ = axs_ - tau * b[i];                                                        \
      nm_axsb_n = MAXV(nm_axsb_n, fabs(rr_)); nm_axsb_o = MAXV(nm_axsb_o, fabs(rr_ * f_));         \
      nm_ax_o = MAXV(nm_ax_o, fabs(ax[i] * f_)); nm_axs_o = MAXV(nm_axs_o, fabs(axs_ * f_));       \
      nm_s_o = MAXV(nm_s_o, fabs(rsk[n + i] / (d_ * sigma)));                                      \
    }                                                                                              \
    nm_pxatyc_n = 0; nm_atyc_o = 0; nm_aty_o = 0;                                                  \
    for (i = 0; i < n; ++i) {                                                                      \
      const double e_ = E ? E[i] : 1.0, f_ = inv_ / e_, rr_ = aty[i] + tau * c[i];                 \
      nm_pxatyc_n = MAXV(nm_pxatyc_n, fabs(rr_)); nm_atyc_o = MAXV(nm_atyc_o, fabs(rr_ * f_));     \
      nm_aty_o = MAXV(nm_aty_o, fabs(aty[i] * f_));                                                \
    }                                                                                              \
    bty_tau = dot(u + n, b, m) / pd_; ctx_tau = dot(u, c, n) / pd_;                                \
    bty_ = SAFEDIV(bty_tau * pd_, tau) / pd_; ctx_ = SAFEDIV(ctx_tau * pd_, tau) / pd_;            \
    gap = fabs(ctx_ + bty_); pobj = ctx_; dobj = -bty_;                                            \
    res_pri = SAFEDIV(nm_axsb_o, tau); res_dual = SAFEDIV(nm_atyc_o, tau);                         \
    res_unbdd_a = NAN; res_infeas = NAN;                                                           \
    if (ctx_tau < -1e-9 / pd_) res_unbdd_a = SAFEDIV(nm_axs_o, -ctx_tau);                          \
    if (bty_tau < -1e-9 / pd_) res_infeas = SAFEDIV(nm_aty_o, -bty_tau);                           \
  } while (0)

int orc_solve(const OrcMatrix *A_in, const double *b_in, const double *c_in, const OrcCone *k_in,
              const OrcSettings *stgs, double *xo, double *yo, double *so, OrcInfo *info) {
  const int n = A_in->n, m = A_in->m, l = n + m + 1;
  const int nnz = A_in->p[n];
  int i, it, status = 0, nb, *bnd;
  double t0 = now_ms(), t1;
  OrcMatrix A;
  OrcCone k = *k_in;
  double *b = (double *)malloc((size_t)m * 8), *c = (double *)malloc((size_t)n * 8);
  double *D = NULL, *E = NULL, sigma = 1.0, scale = stgs->scale;
  double *u = (double *)calloc((size_t)l, 8), *u_t = (double *)calloc((size_t)l, 8), *v = (double *)calloc((size_t)l, 8);
  double *v_prev = (double *)calloc((size_t)l, 8), *rsk = (double *)calloc((size_t)l, 8), *g = (double *)calloc((size_t)l, 8);
  double *R = (double *)calloc((size_t)l, 8), *ws = (double *)calloc((size_t)n, 8);
  double *ax = (double *)calloc((size_t)m, 8), *aty = (double *)calloc((size_t)n, 8);
  double nm_axsb_n = 0.0, nm_pxatyc_n = 0.0; /* normalised residual norms, stale like the reference */
  double nm_b = norm_inf(b_in, m), nm_c = norm_inf(c_in, n);
  double sum_log = 0.0, aa_norm = 0.0;
  int n_log = 0, last_update = 0, last_res_iter = -1;
  /* residual scalars (un-normalised) */
  double tau = 0, kap = 0, bty_tau = 0, ctx_tau = 0, res_pri = NAN, res_dual = NAN, gap = NAN, pobj = NAN, dobj = NAN;
  double bty_ = 0, ctx_ = 0;
  double nm_ax_o = 0, nm_s_o = 0, nm_aty_o = 0, nm_axsb_o = 0, nm_atyc_o = 0, nm_axs_o = 0, res_infeas = NAN, res_unbdd_a = NAN;
  OrcLinSys *ls;
  OrcConeWork *cw;
  OrcAa *aa = NULL;
  memset(info, 0, sizeof(*info));
  A.m = m; A.n = n;
  A.p = (int *)malloc(((size_t)n + 1) * 4); A.i = (int *)malloc(((size_t)nnz + 1) * 4); A.x = (double *)malloc(((size_t)nnz + 1) * 8);
  memcpy(A.p, A_in->p, ((size_t)n + 1) * 4); memcpy(A.i, A_in->i, (size_t)nnz * 4); memcpy(A.x, A_in->x, (size_t)nnz * 8);
  memcpy(b, b_in, (size_t)m * 8); memcpy(c, c_in, (size_t)n * 8);
  if (k.bsize > 1) {
    k.bl = (double *)malloc((size_t)(k.bsize - 1) * 8); k.bu = (double *)malloc((size_t)(k.bsize - 1) * 8);
    memcpy(k.bl, k_in->bl, (size_t)(k.bsize - 1) * 8); memcpy(k.bu, k_in->bu, (size_t)(k.bsize - 1) * 8);
  }
  /* cone boundaries, cones.c:386-424 */
  nb = k.qsize + k.ssize + 1;
  bnd = (int *)calloc((size_t)nb, 4);
  bnd[0] = k.z + k.l + k.bsize;
  for (i = 0; i < k.qsize; ++i) bnd[1 + i] = k.q[i];
  for (i = 0; i < k.ssize; ++i) bnd[1 + k.qsize + i] = (k.s[i] * (k.s[i] + 1)) / 2;
  if (stgs->normalize) {
    D = (double *)malloc((size_t)m * 8); E = (double *)malloc((size_t)n * 8);
    equilibrate(&A, bnd, nb, D, E);
    if (k.bsize > 1) { /* cones.c:1160-1177 */
      const double *Db = D + k.z + k.l;
      for (i = 0; i < k.bsize - 1; ++i) {
        const double fct = Db[i + 1] / Db[0];
        if (k.bu[i] >= 1e15) k.bu[i] = INFINITY; else k.bu[i] *= fct;
        if (k.bl[i] <= -1e15) k.bl[i] = -INFINITY; else k.bl[i] *= fct;
      }
    }
    for (i = 0; i < n; ++i) c[i] *= E[i];   /* normalize.c:33-61 */
    for (i = 0; i < m; ++i) b[i] *= D[i];
    sigma = MAXV(norm_inf(c, n), norm_inf(b, m));
    sigma = sigma < 1e-4 ? 1.0 : sigma;
    sigma = sigma > 1e4 ? 1e4 : sigma;
    sigma = SAFEDIV(1.0, sigma);
    for (i = 0; i < n; ++i) c[i] *= sigma;
    for (i = 0; i < m; ++i) b[i] *= sigma;
  }
#define SET_DIAG_R()                                                           \
  do {                                                                         \
    for (i = 0; i < n; ++i) R[i] = stgs->rho_x;                                \
    for (i = 0; i < k.z; ++i) R[n + i] = 1.0 / (1000. * scale);                \
    for (i = k.z; i < m; ++i) R[n + i] = 1.0 / scale;                          \
    R[n + m] = 10.;                                                            \
  } while (0)
  SET_DIAG_R();
  ls = orc_linsys_init(&A, R);
  cw = orc_cone_init(&k, m);
  if (stgs->acceleration_lookback)
    aa = orc_aa_init(l, stgs->acceleration_lookback, stgs->acceleration_lookback, stgs->acceleration_type_1,
                     stgs->acceleration_regularization, stgs->acceleration_relaxation, 1.0, 1e10, 5);
  info->setup_time_ms = now_ms() - t0;
  t1 = now_ms();
  /* cold start + g (scs.c:681-685, 1118-1128) */
  v[l - 1] = 1.0;
#define WORK_CACHE()                                                           \
  do {                                                                         \
    memcpy(g, c, (size_t)n * 8);                                               \
    for (i = 0; i < m; ++i) g[n + i] = -b[i];                                  \
    orc_linsys_solve(ls, g, NULL, 1e-12);                                      \
    info->cg_iters += orc_linsys_last_cg_its(ls); info->lin_sys_solves++;     \
  } while (0)
  WORK_CACHE();

  for (it = 0; it < stgs->max_iters; ++it) {
    double tol, nm_ws, tau_t;
    if (aa && it > 0 && it % stgs->acceleration_interval == 0) aa_norm = orc_aa_apply(aa, v, v_prev);
    if (it >= 1) { /* normalize_v, scs.c:813-821 */
      const double nv = norm_2(v, l);
      if (nv != 0.) { const double f = sqrt((double)l) * 1. / nv; for (i = 0; i < l; ++i) v[i] *= f; }
    }
    if (aa) memcpy(v_prev, v, (size_t)l * 8);
    /* project_lin_sys, scs.c:733-771 */
    for (i = 0; i < n; ++i) u_t[i] = v[i] * R[i];
    for (i = n; i < l - 1; ++i) u_t[i] = -v[i] * R[i];
    u_t[l - 1] = v[l - 1];
    memcpy(ws, u, (size_t)n * 8);
    for (i = 0; i < n; ++i) ws[i] += u[l - 1] * g[i];
    tol = MINV(nm_axsb_n, nm_pxatyc_n);
    nm_ws = norm_inf(ws, n) / pow((double)it + 1, 1.5);
    tol = 0.2 * MINV(tol, nm_ws);
    tol = MAXV(1e-12, tol);
    orc_linsys_solve(ls, u_t, ws, tol);
    info->cg_iters += orc_linsys_last_cg_its(ls); info->lin_sys_solves++;
    if (it < 1) tau_t = 1.;
    else { /* root_plus, scs.c:710-730 */
      double gg = 0, mug = 0, pg = 0, pp = 0, pmu = 0;
      for (i = 0; i < n + m; ++i) {
        const double ri = R[i], gi = g[i], pi = u_t[i], mui = v[i];
        gg += gi * gi * ri; mug += mui * gi * ri; pg += pi * gi * ri; pp += pi * pi * ri; pmu += pi * mui * ri;
      }
      tau_t = root_plus_coeffs(R[n + m] + gg, mug - 2 * pg - v[l - 1] * R[n + m], pp - pmu);
    }
    u_t[l - 1] = tau_t;
    for (i = 0; i < l - 1; ++i) u_t[i] += -tau_t * g[i];
    /* project_cones, scs.c:796-810 */
    for (i = 0; i < l; ++i) u[i] = 2 * u_t[i] - v[i];
    orc_proj_dual_cone(cw, u + n, R + n);
    u[l - 1] = it < 1 ? 1.0 : MAXV(u[l - 1], 0.);
    for (i = 0; i < l; ++i) rsk[i] = (v[i] + u[i] - 2 * u_t[i]) * R[i]; /* scs.c:781-786 */

    if (it % 25 == 0) { /* has_converged, scs.c:611-649 */
      double grl, prl, drl;
      RESIDUALS();
      last_res_iter = it;
      if (tau > 0.) {
        grl = MAXV(fabs(ctx_), fabs(bty_));
        prl = MAXV(MAXV(nm_b * tau, nm_s_o), nm_ax_o) / tau;
        drl = MAXV(nm_c * tau, nm_aty_o) / tau;
        if (isless(res_pri, stgs->eps_abs + stgs->eps_rel * prl) && isless(res_dual, stgs->eps_abs + stgs->eps_rel * drl) &&
            isless(gap, stgs->eps_abs + stgs->eps_rel * grl)) { status = 1; break; }
      }
      if (isless(res_unbdd_a, stgs->eps_infeas)) { status = -1; break; } /* res_unbdd_p = 0 when P = 0 */
      if (isless(res_infeas, stgs->eps_infeas)) { status = -2; break; }
    }
    if (stgs->adaptive_scale && it == last_res_iter) { /* update_scale, scs.c:1164-1241 */
      double rp, rd, factor, ns;
      rp = SAFEDIV(nm_axsb_o, MAXV(MAXV(nm_ax_o, nm_s_o), nm_b * tau));
      rd = SAFEDIV(nm_atyc_o, MAXV(nm_aty_o, nm_c * tau));
      rp = MAXV(rp, DIV_EPS); rd = MAXV(rd, DIV_EPS);
      sum_log += log(rp) - log(rd); n_log++;
      factor = sqrt(exp(sum_log / (double)n_log));
      if (it - last_update >= 100) {
        ns = MINV(MAXV(scale * factor, 1e-6), 1e6);
        if (ns != scale && (factor > sqrt(10.) || factor < 1. / sqrt(10.))) {
          info->scale_updates++; sum_log = 0; n_log = 0; last_update = it; scale = ns;
          SET_DIAG_R();
          orc_linsys_update_diag_r(ls, R);
          WORK_CACHE();
          if (aa) orc_aa_reset(aa);
          for (i = 0; i < l; i++) v[i] = rsk[i] / R[i] + 2 * u_t[i] - u[i];
        }
      }
    }
    for (i = 0; i < l; ++i) v[i] += stgs->alpha * (u[i] - u_t[i]); /* scs.c:788-793 */
    if (aa && it % stgs->acceleration_interval == 0 && aa_norm > 0) {
      if (orc_aa_safeguard(aa, v, v_prev) < 0) info->rejected_accel_steps++;
      else info->accepted_accel_steps++;
    }
  }
  /* finalize (solved / inaccurate only), scs.c:916-966: residuals are recomputed at `it` */
  if (last_res_iter != it) RESIDUALS();
  for (i = 0; i < n; ++i) xo[i] = (E ? u[i] * (E[i] / sigma) : u[i]);
  for (i = 0; i < m; ++i) {
    yo[i] = D ? u[n + i] * (D[i] / sigma) : u[n + i];
    so[i] = D ? rsk[n + i] / (D[i] * sigma) : rsk[n + i];
  }
  if (status == 1 || status == 0) {
    const double f = SAFEDIV(1.0, tau);
    for (i = 0; i < n; ++i) xo[i] *= f;
    for (i = 0; i < m; ++i) { yo[i] *= f; so[i] *= f; }
  }
  info->iter = it; info->status_val = status == 0 ? 2 : status; info->pobj = pobj; info->dobj = dobj;
  info->res_pri = res_pri; info->res_dual = res_dual; info->gap = gap; info->scale = scale;
  info->solve_time_ms = now_ms() - t1;
  (void)kap; (void)res_infeas;
  orc_linsys_free(ls); orc_cone_free(cw); orc_aa_free(aa);
  free(A.p); free(A.i); free(A.x); free(b); free(c); free(D); free(E); free(u); free(u_t); free(v);
  free(v_prev); free(rsk); free(g); free(R); free(ws); free(ax); free(aty); free(bnd);
  if (k.bsize > 1) { free(k.bl); free(k.bu); }
  return info->status_val;
}
