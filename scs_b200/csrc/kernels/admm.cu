// admm.cu -- fused l-vector kernels of the SCS ADMM iteration (l = n + m + 1),
// all iterates resident in HBM, every scalar (||v||, CG tolerance, tau, ...)
// produced and consumed on the device.
//
// Replaces, from reference src/scs.c:
//   normalize_v            :813-821      -> k_sumsq + k_prep_linsys
//   project_lin_sys        :733-771      -> k_prep_linsys (+ cg.cu) + k_root_plus + k_cone_pre
//   root_plus(_from_coeffs):689-730      -> k_root_plus (last block solves the quadratic)
//   project_cones          :796-810      -> k_cone_pre (+ cones.cu for box/SOC/PSD rows)
//   proj_dual_cone wrapper  src/cones.c:1552-1596 (Moreau pre/post scaling; zero & LP rows inline)
//   compute_rsk            :781-786      -> k_rsk_dual
//   update_dual_vars       :788-793      -> k_rsk_dual / k_dual_update
//   update_scale remap     :1236-1238    -> k_remap_v
//   populate_residual_struct :535-607 and unnormalize_residuals :487-531 -> k_resid_rows / k_resid_cols
#include "../common.cuh"
#include "../dev_api.h"
#include "../admm_api.h"
#include <math.h>

#define VT B200_RED_THREADS

static inline int vgrid(long long n) {
  long long g = (n + (long long)VT * 4 - 1) / ((long long)VT * 4);
  long long cap = 4LL * b200_num_sms();
  if (g > cap) g = cap;
  if (g < 1) g = 1;
  return (int)g;
}
#define GS_LOOP(i, n)                                                                   \
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < (n);        \
       i += (long long)gridDim.x * blockDim.x)

// ------------------------------------------------------------- sum of squares
__global__ void __launch_bounds__(VT)
k_sumsq(long long len, const double *__restrict__ v, double *out, double *partials,
        unsigned int *counter) {
  __shared__ double s_red[64];
  double acc[1] = {0.0};
  GS_LOOP(i, len) acc[0] = fma(v[i], v[i], acc[0]);
  block_sum<1>(acc, s_red);
  if (grid_finish<1>(acc, partials, counter, 0u, s_red))
    if (threadIdx.x == 0) *out = acc[0];
}

// ------------------------------------------------------------- before the KKT solve
// v *= sqrt(l)/||v|| (iter >= FEASIBLE_ITERS); v_prev = v; u_t = [R_x v_x; -R_y v_y; v_tau];
// ws = u_x + tau g_x; tol = max(1e-12, 0.2 min(min(res norms), ||ws||_inf / (iter+1)^1.5)).
__global__ void __launch_bounds__(VT)
k_prep_linsys(int n, int m, int do_normalize, int store_prev, double *__restrict__ v,
              double *__restrict__ v_prev, double *__restrict__ u_t, const double *__restrict__ u,
              const double *__restrict__ g, const double *__restrict__ R, double *__restrict__ ws,
              double *sc, double pw, double *partials, unsigned int *counter) {
  __shared__ double s_red[64];
  const long long l = (long long)n + m + 1;
  double scale = 1.0;
  if (do_normalize) {
    const double nrm = sqrt(sc[SC_VNORM2]);
    if (nrm != 0.0) scale = sqrt((double)l) * 1.0 / nrm;
  }
  const double tau = u[l - 1];
  double mx[1] = {0.0};
  GS_LOOP(i, l) {
    double vi = v[i];
    if (do_normalize) {
      vi *= scale;
      v[i] = vi;
    }
    if (store_prev) v_prev[i] = vi;
    double ut;
    if (i < n) {
      ut = vi * R[i];
      const double w = fma(tau, g[i], u[i]);
      ws[i] = w;
      mx[0] = fmax(mx[0], fabs(w));
    } else if (i < l - 1) {
      ut = -vi * R[i];
    } else {
      ut = vi;
    }
    u_t[i] = ut;
  }
  block_max<1>(mx, s_red);
  if (grid_finish<1>(mx, partials, counter, 1u, s_red)) {
    if (threadIdx.x == 0) {
      sc[SC_NM_WS] = mx[0];
      double tol = fmin(sc[SC_NM_AXSB], sc[SC_NM_PXATYC]);
      const double nm_ws = mx[0] / pw;
      tol = 0.2 * fmin(tol, nm_ws);
      tol = fmax(1e-12, tol);
      sc[SC_TOL] = tol;
    }
  }
}

// ------------------------------------------------------------- root_plus
__device__ double root_plus_from_coeffs(double a, double b, double c) {
  if (!isfinite(a) || !isfinite(b) || !isfinite(c) || a <= 0.0) return NAN;
  const double rad = b * b - 4 * a * c;
  if (!isfinite(rad)) return NAN;
  if (rad < 0.0) return -b / (2 * a);
  const double sq = sqrt(rad);
  if (b <= 0.0) return (-b + sq) / (2 * a);
  const double q = -0.5 * (b + sq);
  return q != 0.0 ? c / q : 0.0;
}

// five R-weighted dots over n+m, then u_t[l-1] = root (or 1 while iter < FEASIBLE_ITERS)
__global__ void __launch_bounds__(VT)
k_root_plus(int n, int m, int feasible_iter, double *__restrict__ u_t, const double *__restrict__ v,
            const double *__restrict__ g, const double *__restrict__ R, double *sc, double *partials,
            unsigned int *counter) {
  __shared__ double s_red[5 * 32];
  const long long nm = (long long)n + m;
  double a5[5] = {0.0, 0.0, 0.0, 0.0, 0.0};  // gg, mug, pg, pp, pmu
  if (!feasible_iter) {
    GS_LOOP(i, nm) {
      const double ri = R[i], gi = g[i], pi = u_t[i], mui = v[i];
      a5[0] += gi * gi * ri;
      a5[1] += mui * gi * ri;
      a5[2] += pi * gi * ri;
      a5[3] += pi * pi * ri;
      a5[4] += pi * mui * ri;
    }
  }
  block_sum<5>(a5, s_red);
  if (grid_finish<5>(a5, partials, counter, 0u, s_red)) {
    if (threadIdx.x == 0) {
      double root = 1.0;
      if (!feasible_iter) {
        const double tau_scale = R[nm];
        const double eta = v[nm];
        const double a = tau_scale + a5[0];
        const double b = a5[1] - 2 * a5[2] - eta * tau_scale;
        const double c = a5[3] - a5[4];
        root = root_plus_from_coeffs(a, b, c);
      }
      u_t[nm] = root;
      sc[SC_TAU_T] = root;
    }
  }
}

// ------------------------------------------------------------- cone pre-pass
// u_t[0:l-1] -= tau_t g ; u = 2 u_t - v ; Moreau wrapper on the y rows:
//   zero rows: dual cone is free     -> u_y unchanged
//   LP rows:  u_y = max(-r u, 0)/r + u
//   other rows (box/SOC/PSD): save s = u_y in cs[], write x = -r u for cones.cu
// u_tau = 1 (iter < FEASIBLE_ITERS) or max(u_tau, 0).
__global__ void __launch_bounds__(VT)
k_cone_pre(int n, int m, int nz, int nl, int feasible_iter, double *__restrict__ u_t,
           double *__restrict__ u, const double *__restrict__ v, const double *__restrict__ g,
           const double *__restrict__ R, double *__restrict__ cs) {
  const long long l = (long long)n + m + 1;
  const double ntau = -u_t[l - 1];
  GS_LOOP(i, l) {
    double ut = u_t[i];
    if (i < l - 1) {
      ut = fma(ntau, g[i], ut);
      u_t[i] = ut;
    }
    double ui = 2 * ut - v[i];
    if (i >= n && i < l - 1) {
      const long long k = i - n;
      if (k < nz) {
        // free
      } else if (k < (long long)nz + nl) {
        const double r = R[i];
        double x = ui * -r;
        x = fmax(x, 0.0);
        ui = x / r + ui;
      } else {
        cs[k] = ui;
        ui = ui * -R[i];
      }
    } else if (i == l - 1) {
      ui = feasible_iter ? 1.0 : fmax(ui, 0.0);
    }
    u[i] = ui;
  }
}

// ------------------------------------------------------------- rsk and dual update
__global__ void __launch_bounds__(VT)
k_rsk_dual(long long l, int do_dual, double alpha, double *__restrict__ rsk, double *__restrict__ v,
           const double *__restrict__ u, const double *__restrict__ u_t,
           const double *__restrict__ R) {
  GS_LOOP(i, l) {
    const double vi = v[i], ui = u[i], ut = u_t[i];
    rsk[i] = (vi + ui - 2 * ut) * R[i];
    if (do_dual) v[i] = vi + alpha * (ui - ut);
  }
}
__global__ void __launch_bounds__(VT)
k_dual_update(long long l, double alpha, double *__restrict__ v, const double *__restrict__ u,
              const double *__restrict__ u_t) {
  GS_LOOP(i, l) v[i] += alpha * (u[i] - u_t[i]);
}
// v = rsk / R+ + 2 u_t - u   (scs.c:1236-1238)
__global__ void __launch_bounds__(VT)
k_remap_v(long long l, double *__restrict__ v, const double *__restrict__ rsk,
          const double *__restrict__ u, const double *__restrict__ u_t,
          const double *__restrict__ R) {
  GS_LOOP(i, l) v[i] = rsk[i] / R[i] + 2 * u_t[i] - u[i];
}

// diag_r = [rho_x]*n ++ [1/(1000 scale)]*z ++ [1/scale]*(m-z) ++ [TAU_FACTOR]
__global__ void k_set_diag_r(int n, int m, int nz, double rho_x, double scale, double *__restrict__ R) {
  const long long l = (long long)n + m + 1;
  GS_LOOP(i, l) {
    double r;
    if (i < n) r = rho_x;
    else if (i < (long long)n + nz) r = 1.0 / (1000. * scale);
    else if (i < l - 1) r = 1.0 / scale;
    else r = 10.0;
    R[i] = r;
  }
}

// g = [c; -b]  (scs.c:1118-1126)
__global__ void k_build_h(int n, int m, const double *__restrict__ c, const double *__restrict__ b,
                          double *__restrict__ g) {
  GS_LOOP(i, (long long)n + m) g[i] = (i < n) ? c[i] : -b[i - n];
}

// v = [x; y + s / R_y; 1] with NaN -> 0   (warm_start_vars, scs.c:660-679; x,y,s already normalised)
__global__ void k_warm_start(int n, int m, const double *__restrict__ x, const double *__restrict__ y,
                             const double *__restrict__ s, const double *__restrict__ R,
                             double *__restrict__ v) {
  const long long l = (long long)n + m + 1;
  GS_LOOP(i, l) {
    double r;
    if (i < n) {
      r = x[i];
    } else if (i < l - 1) {
      r = y[i - n] + s[i - n] / R[i];
    } else {
      r = 1.0;
    }
    v[i] = (r != r) ? 0.0 : r;
  }
}
__global__ void k_cold_start(long long l, double *__restrict__ v) {
  GS_LOOP(i, l) v[i] = (i == l - 1) ? 1.0 : 0.0;
}

// ------------------------------------------------------------- residuals
// rows (length m): ax given (A x). s = rsk_y, y = u_y. Produces
//   max: |ax+s-tau b| (normalised), and un-normalised |.| of ax_s_btau, ax, ax_s, s ; sum: y'b
// f_i = inv_ds / D_i  (un-normalisation of primal quantities); s_orig = s / (D dual_scale).
__global__ void __launch_bounds__(VT)
k_resid_rows(int m, const double *__restrict__ ax, const double *__restrict__ s,
             const double *__restrict__ y, const double *__restrict__ b,
             const double *__restrict__ D, double inv_ds, double dual_scale, const double *tau_p,
             double *sc, double *partials, unsigned int *counter) {
  __shared__ double s_red[6 * 32];
  const double tau = fabs(*tau_p);
  double a[6] = {0, 0, 0, 0, 0, 0};  // 0..4 max, 5 sum
  GS_LOOP(i, m) {
    const double axi = ax[i], si = s[i];
    const double ax_s = axi + si;
    const double r = ax_s - tau * b[i];
    const double d = D ? D[i] : 1.0;
    const double f = inv_ds / d;
    a[0] = fmax(a[0], fabs(r));
    a[1] = fmax(a[1], fabs(r * f));
    a[2] = fmax(a[2], fabs(axi * f));
    a[3] = fmax(a[3], fabs(ax_s * f));
    a[4] = fmax(a[4], fabs(si / (d * dual_scale)));
    a[5] = fma(y[i], b[i], a[5]);
  }
  double mx[5] = {a[0], a[1], a[2], a[3], a[4]};
  double sm[1] = {a[5]};
  block_max<5>(mx, s_red);
  block_sum<1>(sm, s_red + 5 * 32);
  double all[6] = {mx[0], mx[1], mx[2], mx[3], mx[4], sm[0]};
  if (grid_finish<6>(all, partials, counter, 0x1fu, s_red)) {
    if (threadIdx.x == 0) {
      sc[SC_NM_AXSB] = all[0];
      sc[SC_O_AXSB] = all[1];
      sc[SC_O_AX] = all[2];
      sc[SC_O_AXS] = all[3];
      sc[SC_O_S] = all[4];
      sc[SC_BTY_TAU] = all[5];
      sc[SC_TAU] = tau;
    }
  }
}
// cols (length n): aty given (A'y), px given or NULL.
__global__ void __launch_bounds__(VT)
k_resid_cols(int n, const double *__restrict__ aty, const double *__restrict__ px,
             const double *__restrict__ x, const double *__restrict__ c,
             const double *__restrict__ E, double inv_ps, const double *tau_p, const double *kap_p,
             double *sc, double *partials, unsigned int *counter) {
  __shared__ double s_red[6 * 32];
  const double tau = fabs(*tau_p);
  double a[6] = {0, 0, 0, 0, 0, 0};  // 0..3 max, 4..5 sum
  GS_LOOP(i, n) {
    const double pxi = px ? px[i] : 0.0;
    const double atyi = aty[i];
    const double r = pxi + atyi + tau * c[i];
    const double e = E ? E[i] : 1.0;
    const double f = inv_ps / e;
    a[0] = fmax(a[0], fabs(r));
    a[1] = fmax(a[1], fabs(r * f));
    a[2] = fmax(a[2], fabs(pxi * f));
    a[3] = fmax(a[3], fabs(atyi * f));
    a[4] = fma(x[i], c[i], a[4]);
    a[5] = fma(pxi, x[i], a[5]);
  }
  double mx[4] = {a[0], a[1], a[2], a[3]};
  double sm[2] = {a[4], a[5]};
  block_max<4>(mx, s_red);
  block_sum<2>(sm, s_red + 4 * 32);
  double all[6] = {mx[0], mx[1], mx[2], mx[3], sm[0], sm[1]};
  if (grid_finish<6>(all, partials, counter, 0xfu, s_red)) {
    if (threadIdx.x == 0) {
      sc[SC_NM_PXATYC] = all[0];
      sc[SC_O_PXATYC] = all[1];
      sc[SC_O_PX] = all[2];
      sc[SC_O_ATY] = all[3];
      sc[SC_CTX_TAU] = all[4];
      sc[SC_XPX_TAU] = all[5];
      sc[SC_KAP] = fabs(*kap_p);
    }
  }
}

// ------------------------------------------------------------- finalize helpers
// out = in .* (E / dual_scale)  | in .* (D / primal_scale) | in ./ (D dual_scale)   (un_normalize_sol)
__global__ void k_unnormalize_sol(int n, int m, const double *__restrict__ u,
                                  const double *__restrict__ rsk, const double *__restrict__ D,
                                  const double *__restrict__ E, double primal_scale,
                                  double dual_scale, double *__restrict__ xo, double *__restrict__ yo,
                                  double *__restrict__ so) {
  GS_LOOP(i, (long long)n + m) {
    if (i < n) {
      xo[i] = E ? u[i] * (E[i] / dual_scale) : u[i];
    } else {
      const long long k = i - n;
      yo[k] = D ? u[i] * (D[k] / primal_scale) : u[i];
      so[k] = D ? rsk[i] / (D[k] * dual_scale) : rsk[i];
    }
  }
}
// normalise a warm start (normalize_sol, normalize.c:64-75)
__global__ void k_normalize_sol(int n, int m, const double *__restrict__ D,
                                const double *__restrict__ E, double primal_scale,
                                double dual_scale, double *__restrict__ x, double *__restrict__ y,
                                double *__restrict__ s) {
  GS_LOOP(i, (long long)n + m) {
    if (i < n) {
      x[i] /= (E[i] / dual_scale);
    } else {
      const long long k = i - n;
      y[k] /= (D[k] / primal_scale);
      s[k] *= (D[k] * dual_scale);
    }
  }
}
// three generic reductions used by finalize: |a|_inf, |b|_inf, a'b
__global__ void __launch_bounds__(VT)
k_norms_dot(long long len, const double *__restrict__ a, const double *__restrict__ b, double *out3,
            double *partials, unsigned int *counter) {
  __shared__ double s_red[3 * 32];
  double mx[2] = {0, 0}, sm[1] = {0};
  GS_LOOP(i, len) {
    mx[0] = fmax(mx[0], fabs(a[i]));
    mx[1] = fmax(mx[1], fabs(b[i]));
    sm[0] = fma(a[i], b[i], sm[0]);
  }
  block_max<2>(mx, s_red);
  block_sum<1>(sm, s_red + 64);
  double all[3] = {mx[0], mx[1], sm[0]};
  if (grid_finish<3>(all, partials, counter, 0x3u, s_red))
    if (threadIdx.x == 0) { out3[0] = all[0]; out3[1] = all[1]; out3[2] = all[2]; }
}
__global__ void k_scale(long long len, double *__restrict__ a, double f) { GS_LOOP(i, len) a[i] *= f; }
__global__ void k_fill(long long len, double *__restrict__ a, double f) { GS_LOOP(i, len) a[i] = f; }
// b *= D*sigma or c *= E*sigma  (normalize_b_c applied with a known sigma)
__global__ void k_scale_by(long long len, double *__restrict__ a, const double *__restrict__ d, double f) {
  GS_LOOP(i, len) a[i] = a[i] * d[i] * f;
}

// ------------------------------------------------------------- launchers
#define ST ((cudaStream_t)b200_stream())
#define DONE(nk) do { b200_count_launch(nk); CUDA_OK(cudaGetLastError()); return 0; } while (0)

extern "C" int b200_admm_sumsq(long long len, const double *d_v, double *d_out, double *d_part,
                               unsigned int *d_cnt) {
  k_sumsq<<<vgrid(len), VT, 0, ST>>>(len, d_v, d_out, d_part, d_cnt);
  DONE(1);
}
extern "C" int b200_admm_prep_linsys(const B200Admm *w, int iter, int store_prev, double pw) {
  const long long l = (long long)w->n + w->m + 1;
  const int do_norm = iter >= 1;
  if (do_norm) {
    k_sumsq<<<vgrid(l), VT, 0, ST>>>(l, w->d_v, w->d_sc + SC_VNORM2, w->d_part, w->d_cnt);
    b200_count_launch(1);
  }
  k_prep_linsys<<<vgrid(l), VT, 0, ST>>>(w->n, w->m, do_norm, store_prev, w->d_v, w->d_v_prev,
                                          w->d_u_t, w->d_u, w->d_g, w->d_R, w->d_ws, w->d_sc, pw,
                                          w->d_part, w->d_cnt);
  DONE(1);
}
extern "C" int b200_admm_root_plus(const B200Admm *w, int iter) {
  k_root_plus<<<vgrid((long long)w->n + w->m), VT, 0, ST>>>(w->n, w->m, iter < 1, w->d_u_t, w->d_v,
                                                            w->d_g, w->d_R, w->d_sc, w->d_part,
                                                            w->d_cnt);
  DONE(1);
}
extern "C" int b200_admm_cone_pre(const B200Admm *w, int iter, int nz, int nl, double *d_cs) {
  const long long l = (long long)w->n + w->m + 1;
  k_cone_pre<<<vgrid(l), VT, 0, ST>>>(w->n, w->m, nz, nl, iter < 1, w->d_u_t, w->d_u, w->d_v,
                                       w->d_g, w->d_R, d_cs);
  DONE(1);
}
extern "C" int b200_admm_rsk_dual(const B200Admm *w, int do_dual, double alpha) {
  const long long l = (long long)w->n + w->m + 1;
  k_rsk_dual<<<vgrid(l), VT, 0, ST>>>(l, do_dual, alpha, w->d_rsk, w->d_v, w->d_u, w->d_u_t, w->d_R);
  DONE(1);
}
extern "C" int b200_admm_dual_update(const B200Admm *w, double alpha) {
  const long long l = (long long)w->n + w->m + 1;
  k_dual_update<<<vgrid(l), VT, 0, ST>>>(l, alpha, w->d_v, w->d_u, w->d_u_t);
  DONE(1);
}
extern "C" int b200_admm_remap_v(const B200Admm *w) {
  const long long l = (long long)w->n + w->m + 1;
  k_remap_v<<<vgrid(l), VT, 0, ST>>>(l, w->d_v, w->d_rsk, w->d_u, w->d_u_t, w->d_R);
  DONE(1);
}
extern "C" int b200_admm_set_diag_r(const B200Admm *w, int nz, double rho_x, double scale) {
  const long long l = (long long)w->n + w->m + 1;
  k_set_diag_r<<<vgrid(l), 256, 0, ST>>>(w->n, w->m, nz, rho_x, scale, w->d_R);
  DONE(1);
}
extern "C" int b200_admm_build_h(const B200Admm *w, const double *d_c, const double *d_b) {
  k_build_h<<<vgrid((long long)w->n + w->m), 256, 0, ST>>>(w->n, w->m, d_c, d_b, w->d_g);
  DONE(1);
}
extern "C" int b200_admm_warm_start(const B200Admm *w, const double *d_x, const double *d_y,
                                    const double *d_s) {
  const long long l = (long long)w->n + w->m + 1;
  k_warm_start<<<vgrid(l), 256, 0, ST>>>(w->n, w->m, d_x, d_y, d_s, w->d_R, w->d_v);
  DONE(1);
}
extern "C" int b200_admm_cold_start(const B200Admm *w) {
  const long long l = (long long)w->n + w->m + 1;
  k_cold_start<<<vgrid(l), 256, 0, ST>>>(l, w->d_v);
  DONE(1);
}
extern "C" int b200_admm_resid_rows(const B200Admm *w, const double *d_ax, const double *d_b,
                                    const double *d_D, double inv_ds, double dual_scale) {
  const long long l = (long long)w->n + w->m + 1;
  k_resid_rows<<<vgrid(w->m), VT, 0, ST>>>(w->m, d_ax, w->d_rsk + w->n, w->d_u + w->n, d_b, d_D,
                                           inv_ds, dual_scale, w->d_u + (l - 1), w->d_sc, w->d_part,
                                           w->d_cnt);
  DONE(1);
}
extern "C" int b200_admm_resid_cols(const B200Admm *w, const double *d_aty, const double *d_px,
                                    const double *d_c, const double *d_E, double inv_ps) {
  const long long l = (long long)w->n + w->m + 1;
  k_resid_cols<<<vgrid(w->n), VT, 0, ST>>>(w->n, d_aty, d_px, w->d_u, d_c, d_E, inv_ps,
                                           w->d_u + (l - 1), w->d_rsk + (l - 1), w->d_sc, w->d_part,
                                           w->d_cnt);
  DONE(1);
}
extern "C" int b200_admm_unnormalize_sol(const B200Admm *w, const double *d_D, const double *d_E,
                                         double primal_scale, double dual_scale, double *d_x,
                                         double *d_y, double *d_s) {
  k_unnormalize_sol<<<vgrid((long long)w->n + w->m), 256, 0, ST>>>(
      w->n, w->m, w->d_u, w->d_rsk, d_D, d_E, primal_scale, dual_scale, d_x, d_y, d_s);
  DONE(1);
}
extern "C" int b200_admm_normalize_sol(int n, int m, const double *d_D, const double *d_E,
                                       double primal_scale, double dual_scale, double *d_x,
                                       double *d_y, double *d_s) {
  k_normalize_sol<<<vgrid((long long)n + m), 256, 0, ST>>>(n, m, d_D, d_E, primal_scale, dual_scale,
                                                          d_x, d_y, d_s);
  DONE(1);
}
extern "C" int b200_vec_norms_dot(long long len, const double *d_a, const double *d_b, double *d_out3,
                                  double *d_part, unsigned int *d_cnt) {
  k_norms_dot<<<vgrid(len), VT, 0, ST>>>(len, d_a, d_b, d_out3, d_part, d_cnt);
  DONE(1);
}
extern "C" int b200_vec_scale(long long len, double *d_a, double f) {
  k_scale<<<vgrid(len), 256, 0, ST>>>(len, d_a, f);
  DONE(1);
}
extern "C" int b200_vec_fill(long long len, double *d_a, double f) {
  k_fill<<<vgrid(len), 256, 0, ST>>>(len, d_a, f);
  DONE(1);
}
extern "C" int b200_vec_scale_by(long long len, double *d_a, const double *d_d, double f) {
  k_scale_by<<<vgrid(len), 256, 0, ST>>>(len, d_a, d_d, f);
  DONE(1);
}
