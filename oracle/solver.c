/*
 * solver.c -- oracle (TEST INFRASTRUCTURE, see sgw_oracle.h) for
 *   algo/linear_solver/src/bicgstab.f90       (Frommer multishift BiCGStab(l))
 *   algo/linear_solver/src/linear_solver.f90  (SGW Krylov-subspace solver)
 *   algo/linear_solver/src/select_solver.f90  (priority / fallback chain)
 *   data/algebra/src/gram_schmidt.f90, norm.f90
 *   data/parallel/src/parallel.f90:80-138     (parallel_task)
 * written in the reference's execution order with unfused BLAS-1 loops where the reference
 * calls ZDOTU/ZDOTC/ZAXPY/ZSCAL/ZCOPY.  Line references are to the files above.
 */
#include "sgw_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

/* ---------------------------------------------------------------- BLAS-1 (reference BLAS semantics) */
static zcplx zdotu(int n, const zcplx *x, const zcplx *y) {
  zcplx s = 0.0;
  for (int i = 0; i < n; ++i) s += x[i] * y[i];
  return s;
}
static zcplx zdotc(int n, const zcplx *x, const zcplx *y) {
  zcplx s = 0.0;
  for (int i = 0; i < n; ++i) s += conj(x[i]) * y[i];
  return s;
}
static void zaxpy(int n, zcplx a, const zcplx *x, zcplx *y) {
  for (int i = 0; i < n; ++i) y[i] += a * x[i];
}
static void zscal(int n, zcplx a, zcplx *x) {
  for (int i = 0; i < n; ++i) x[i] = a * x[i];
}
static void zcopy(int n, const zcplx *x, zcplx *y) { memcpy(y, x, (size_t)n * sizeof(zcplx)); }

/* reference-BLAS DNRM2 / LAPACK ZLASSQ: scaled sum of squares */
double orc_dnrm2(int n, const double *x) {
  if (n < 1) return 0.0;
  double scale = 0.0, ssq = 1.0;
  for (int i = 0; i < n; ++i) {
    if (x[i] != 0.0) {
      double a = fabs(x[i]);
      if (scale < a) {
        ssq = 1.0 + ssq * (scale / a) * (scale / a);
        scale = a;
      } else {
        ssq += (a / scale) * (a / scale);
      }
    }
  }
  return scale * sqrt(ssq);
}

/* norm.f90:73-107: ZLANGE('F', 1, n, v) -> ZLASSQ over (re, im) pairs */
double orc_norm(const zcplx *v, int n) { return orc_dnrm2(2 * n, (const double *)v); }

static int has_nan(const zcplx *x, long n) { /* util/src/debug.f90:57-92 test_nan */
  const double *d = (const double *)x;
  for (long i = 0; i < 2 * n; ++i)
    if (d[i] != d[i]) return 1;
  return 0;
}

/* ---------------------------------------------------------------- gram_schmidt.f90:35-135 */
void orc_gram_schmidt(int first, int n, int nb, zcplx *basis, zcplx *vector) {
  /* first is 1-based like the Fortran argument */
  for (int ib = first - 1; ib < nb; ++ib) {             /* :94 */
    for (int jb = 0; jb < first - 1; ++jb) {            /* :96 */
      zcplx nrm = zdotc(n, basis + (size_t)jb * n, basis + (size_t)ib * n);          /* :98 */
      zaxpy(n, -nrm, basis + (size_t)jb * n, basis + (size_t)ib * n);                /* :99 */
      if (vector) zaxpy(n, -nrm, vector + (size_t)jb * n, vector + (size_t)ib * n);  /* :102 */
    }
  }
  for (int ib = first - 1; ib < nb; ++ib) {             /* :111 */
    zcplx nrm = 1.0 / orc_norm(basis + (size_t)ib * n, n);                           /* :114 */
    zscal(n, nrm, basis + (size_t)ib * n);
    if (vector) zscal(n, nrm, vector + (size_t)ib * n);
    for (int jb = ib + 1; jb < nb; ++jb) {              /* :121 */
      nrm = zdotc(n, basis + (size_t)ib * n, basis + (size_t)jb * n);                /* :124 */
      zaxpy(n, -nrm, basis + (size_t)ib * n, basis + (size_t)jb * n);
      if (vector) zaxpy(n, -nrm, vector + (size_t)ib * n, vector + (size_t)jb * n);
    }
  }
}

/* ---------------------------------------------------------------- bicgstab.f90 */
typedef struct {
  zcplx *uu, *xx;     /* uu: n x (L+1) */
  zcplx sigma, inv_phi_old, inv_phi, inv_phi_new, inv_theta, alpha, beta;
  zcplx *mu, *gamma, *gamma_p, *gamma_pp;
} shift_sys;

typedef struct {
  zcplx *uu, *rr, *xx, *tilde_r0;
  zcplx sigma, rho, rho_old, alpha, alpha_old, beta, omega;
  zcplx *gamma, *gamma_p, *gamma_pp;
} seed_sys;

/* horner_scheme, bicgstab.f90:941-999 (arrays 1-based in the reference: g[j-1]) */
static void horner_scheme(int lmax, const zcplx *seed_gamma, zcplx sigma, zcplx *sg, zcplx *psi) {
  sg[lmax - 1] = -seed_gamma[lmax - 1];                                   /* :969 */
  for (int jj = lmax - 1; jj >= 1; --jj) sg[jj - 1] = -sigma * sg[jj] - seed_gamma[jj - 1]; /* :975 */
  *psi = -sigma * sg[0] + 1.0;                                            /* :980 */
  for (int ii = 1; ii <= lmax - 1; ++ii)
    for (int jj = lmax - 1; jj >= ii; --jj) sg[jj - 1] = -sigma * sg[jj] + sg[jj - 1];      /* :986 */
  for (int jj = 1; jj <= lmax; ++jj) sg[jj - 1] = -sg[jj - 1] / *psi;     /* :995 */
}

int orc_bicgstab(int lmax, double threshold, int max_iter, orc_op_fn AA, void *ctx, int n,
                 const zcplx *bb, int nshift_tot, const zcplx *sigma, zcplx *xx, orc_stats *st) {
  const int L = lmax;
  const int ns = nshift_tot - 1;
  const size_t N = (size_t)n;
  int ierr = 0;
  long n_op = 0;

  /* init_seed :304-350 */
  seed_sys sd;
  sd.uu = calloc(N * (L + 1), sizeof(zcplx));
  sd.rr = calloc(N * (L + 1), sizeof(zcplx));
  sd.xx = calloc(N, sizeof(zcplx));
  sd.tilde_r0 = malloc(N * sizeof(zcplx));
  sd.gamma = calloc(L, sizeof(zcplx));
  sd.gamma_p = calloc(L, sizeof(zcplx));
  sd.gamma_pp = calloc(L, sizeof(zcplx));
  zcopy(n, bb, sd.rr);
  zcopy(n, bb, sd.tilde_r0);
  sd.sigma = sigma[0];
  sd.rho_old = 1.0;
  sd.alpha_old = 1.0;
  sd.alpha = 0.0;
  sd.omega = 1.0;
  sd.rho = 0.0;
  sd.beta = 0.0;

  /* init_shift :370-490 */
  const int mu_size = L * (L + 1) / 2;
  double *binomial = malloc(sizeof(double) * (mu_size > 0 ? mu_size : 1));
  zcplx *sigma_pow = malloc(sizeof(zcplx) * (L > 0 ? L : 1));
  for (int jj = 0; jj <= L - 1; ++jj) {
    int offset = jj * (jj + 1) / 2 + 1;
    for (int ii = 0; ii <= jj; ++ii) {
      int ij = offset + ii;
      if (ii == 0) binomial[ij - 1] = 1.0;
      else binomial[ij - 1] = (binomial[ij - 2] * (jj - ii + 1)) / ii;   /* :428 */
    }
  }
  shift_sys *sh = ns > 0 ? calloc(ns, sizeof(shift_sys)) : NULL;
  for (int is = 0; is < ns; ++is) {
    sh[is].uu = calloc(N * (L + 1), sizeof(zcplx));
    sh[is].xx = calloc(N, sizeof(zcplx));
    sh[is].gamma = calloc(L, sizeof(zcplx));
    sh[is].gamma_p = calloc(L, sizeof(zcplx));
    sh[is].gamma_pp = calloc(L, sizeof(zcplx));
    sh[is].mu = calloc(mu_size, sizeof(zcplx));
    sh[is].inv_phi_old = 1.0;
    sh[is].inv_phi = 1.0;
    sh[is].inv_theta = 1.0;
    sh[is].sigma = sigma[is + 1] - sigma[0];                             /* :458 */
    sigma_pow[0] = 1.0;
    for (int ii = 1; ii <= L - 1; ++ii) sigma_pow[ii] = sh[is].sigma * sigma_pow[ii - 1];
    for (int jj = 0; jj <= L - 1; ++jj) {
      int offset = jj * (jj + 1) / 2 + 1;
      for (int ii = 0; ii <= jj; ++ii) {
        int ij = offset + ii;
        sh[is].mu[ij - 1] = binomial[ij - 1] * sigma_pow[jj - ii];      /* :480 */
      }
    }
  }
  free(binomial);
  free(sigma_pow);

  zcplx *nu = calloc(L + 1, sizeof(zcplx));
  zcplx *tau = calloc((L + 1) * (L + 2) / 2 + 2, sizeof(zcplx));
#define UU(s, i) ((s).uu + (size_t)(i) * N)
#define RR(i) (sd.rr + (size_t)(i) * N)

  int iter;
  for (iter = 1; iter <= max_iter; ++iter) {                             /* :229 */
    /* ------------------------------------------------ bicg_part :520-702 */
    sd.rho_old = -sd.omega * sd.rho_old;                                 /* :587 */
    for (int jj = 0; jj <= L - 1; ++jj) {                                /* :590 */
      sd.rho = zdotu(n, RR(jj), sd.tilde_r0);                            /* :595 */
      sd.beta = sd.alpha * sd.rho / sd.rho_old;                          /* :597 */
      sd.rho_old = sd.rho;                                               /* :599 */
      for (int ii = 0; ii <= jj; ++ii) {                                 /* :602 */
        zscal(n, -sd.beta, UU(sd, ii));                                  /* :605 */
        zaxpy(n, 1.0, RR(ii), UU(sd, ii));                               /* :606 */
      }
      AA(ctx, sd.sigma, UU(sd, jj), UU(sd, jj + 1), n);                  /* :611 */
      ++n_op;
      sd.alpha = sd.rho / zdotu(n, UU(sd, jj + 1), sd.tilde_r0);         /* :614 */
      for (int is = 0; is < ns; ++is) {                                  /* :621 */
        shift_sys *a = &sh[is];
        a->inv_phi_new = a->inv_phi / (1.0 + sd.alpha * a->sigma +
                                       sd.alpha * sd.beta / sd.alpha_old *
                                           (a->inv_phi / a->inv_phi_old - 1.0)); /* :629-631 */
        zcplx ratio = a->inv_phi / a->inv_phi_old;
        a->beta = ratio * ratio * sd.beta;                               /* :633 */
        a->alpha = (a->inv_phi_new / a->inv_phi) * sd.alpha;             /* :635 */
        zcplx factor = a->inv_theta * a->inv_phi;                        /* :638 */
        for (int ii = 0; ii <= jj; ++ii) {                               /* :641 */
          zscal(n, -a->beta, UU(*a, ii));                                /* :644 */
          zaxpy(n, factor, RR(ii), UU(*a, ii));                          /* :645 */
        }
        zaxpy(n, a->alpha, UU(*a, 0), a->xx);                            /* :650 */
        a->inv_phi_old = a->inv_phi;                                     /* :655 */
        a->inv_phi = a->inv_phi_new;                                     /* :657 */
        zcopy(n, RR(jj), UU(*a, jj + 1));                                /* :661 */
        zscal(n, factor, UU(*a, jj + 1));                                /* :662 */
      }
      sd.alpha_old = sd.alpha;                                           /* :670 */
      for (int ii = 0; ii <= jj; ++ii) zaxpy(n, -sd.alpha, UU(sd, ii + 1), RR(ii)); /* :676 */
      AA(ctx, sd.sigma, RR(jj), RR(jj + 1), n);                          /* :682 */
      ++n_op;
      zaxpy(n, sd.alpha, UU(sd, 0), sd.xx);                              /* :685 */
      for (int is = 0; is < ns; ++is) {                                  /* :688 */
        shift_sys *a = &sh[is];
        zcplx factor = -a->inv_theta * a->inv_phi;                       /* :693 */
        zaxpy(n, factor, RR(jj), UU(*a, jj + 1));                        /* :694 */
        zscal(n, 1.0 / a->alpha, UU(*a, jj + 1));                        /* :695 */
        zaxpy(n, -a->sigma, UU(*a, jj), UU(*a, jj + 1));                 /* :696 */
      }
    }
    if (orc_norm(RR(0), n) < threshold) break;                           /* :237 */

    /* ------------------------------------------------ mr_part :708-931 */
    for (int jj = 1; jj <= L; ++jj) {                                    /* :780 */
      int offset = jj * (jj + 1) / 2 + 1;
      for (int ii = 1; ii <= jj - 1; ++ii) {                             /* :785 */
        int ij = offset + ii;
        tau[ij - 1] = zdotu(n, RR(jj), RR(ii)) / nu[ii - 1];             /* :790 */
        zaxpy(n, -tau[ij - 1], RR(ii), RR(jj));                          /* :792 */
      }
      nu[jj - 1] = zdotu(n, RR(jj), RR(jj));                             /* :797 */
      sd.gamma_p[jj - 1] = zdotu(n, RR(0), RR(jj)) / nu[jj - 1];         /* :799 */
    }
    sd.gamma[L - 1] = sd.gamma_p[L - 1];                                 /* :804 */
    sd.omega = sd.gamma[L - 1];                                          /* :806 */
    for (int jj = L - 1; jj >= 1; --jj) {                                /* :809 */
      sd.gamma[jj - 1] = sd.gamma_p[jj - 1];
      for (int ii = jj + 1; ii <= L; ++ii) {
        int ij = ii * (ii + 1) / 2 + jj + 1;
        sd.gamma[jj - 1] -= tau[ij - 1] * sd.gamma[ii - 1];              /* :815 */
      }
    }
    for (int jj = 1; jj <= L - 1; ++jj) {                                /* :821 */
      sd.gamma_pp[jj - 1] = sd.gamma[jj];
      for (int ii = jj + 1; ii <= L - 1; ++ii) {
        int ij = ii * (ii + 1) / 2 + jj + 1;
        sd.gamma_pp[jj - 1] += tau[ij - 1] * sd.gamma[ii];               /* :827 */
      }
    }
    zaxpy(n, sd.gamma[0], RR(0), sd.xx);                                 /* :833 */
    zaxpy(n, -sd.gamma[L - 1], UU(sd, L), UU(sd, 0));                    /* :835 */
    for (int jj = 1; jj <= L - 1; ++jj) {                                /* :838 */
      zaxpy(n, -sd.gamma[jj - 1], UU(sd, jj), UU(sd, 0));                /* :841 */
      zaxpy(n, sd.gamma_pp[jj - 1], RR(jj), sd.xx);                      /* :844 */
    }
    for (int is = 0; is < ns; ++is) {                                    /* :850 */
      shift_sys *a = &sh[is];
      zcplx psi;
      horner_scheme(L, sd.gamma, a->sigma, a->gamma, &psi);              /* :855 */
      zcplx inv_xi = a->inv_theta * a->inv_phi;                          /* :858 */
      a->inv_theta = a->inv_theta / psi;                                 /* :860 */
      for (int jj = 1; jj <= L; ++jj) {                                  /* :863 */
        a->gamma_p[jj - 1] = 0.0;
        for (int ii = jj; ii <= L; ++ii) {
          int ij = (ii - 1) * ii / 2 + jj;
          a->gamma_p[jj - 1] += a->mu[ij - 1] * a->gamma[ii - 1];        /* :869 */
        }
      }
      for (int jj = 1; jj <= L - 1; ++jj) {                              /* :875 */
        a->gamma_pp[jj - 1] = a->gamma_p[jj];
        for (int ii = jj + 1; ii <= L - 1; ++ii) {
          int ij = ii * (ii + 1) / 2 + jj + 1;
          a->gamma_pp[jj - 1] += tau[ij - 1] * a->gamma_p[ii];           /* :881 */
        }
      }
      zcplx factor = a->gamma_p[0] * inv_xi;                             /* :888 */
      zaxpy(n, factor, RR(0), a->xx);                                    /* :889 */
      zaxpy(n, -sd.gamma[L - 1], UU(*a, L), UU(*a, 0));                  /* :893 */
      for (int jj = 1; jj <= L - 1; ++jj) {                              /* :897 */
        zaxpy(n, -sd.gamma[jj - 1], UU(*a, jj), UU(*a, 0));              /* :900 */
        factor = a->gamma_pp[jj - 1] * inv_xi;                           /* :904 */
        zaxpy(n, factor, RR(jj), a->xx);                                 /* :905 */
      }
      factor = 1.0 / psi;                                                /* :910 */
      zscal(n, factor, UU(*a, 0));                                       /* :911 */
    }
    for (int jj = 1; jj <= L; ++jj) zaxpy(n, -sd.gamma_p[jj - 1], RR(jj), RR(0)); /* :919-925 */

    if (orc_norm(RR(0), n) < threshold) break;                           /* :245 */
  }
  if (iter > max_iter) ierr = 1;                                         /* :249-253 */

  zcopy(n, sd.xx, xx);                                                   /* :258 */
  for (int is = 0; is < ns; ++is) zcopy(n, sh[is].xx, xx + (size_t)(is + 1) * N); /* :260 */
  if (has_nan(xx, (long)N * nshift_tot)) ierr = 2;                       /* :264-267 */

  if (st) {
    st->n_op += n_op;
    st->n_outer = iter > max_iter ? max_iter : iter;
  }
  free(nu);
  free(tau);
  for (int is = 0; is < ns; ++is) {
    free(sh[is].uu); free(sh[is].xx); free(sh[is].gamma); free(sh[is].gamma_p);
    free(sh[is].gamma_pp); free(sh[is].mu);
  }
  free(sh);
  free(sd.uu); free(sd.rr); free(sd.xx); free(sd.tilde_r0);
  free(sd.gamma); free(sd.gamma_p); free(sd.gamma_pp);
#undef UU
#undef RR
  return ierr;
}

/* ---------------------------------------------------------------- linear_solver.f90:82-503 */
int orc_linear_solver(double threshold, int max_iter, orc_op_fn AA, void *ctx, int n,
                      const zcplx *bb, int nshift, const zcplx *sigma, zcplx *xx, orc_stats *st) {
  const size_t N = (size_t)n;
  int ierr = 0;
  long n_op = 0;
  int n_iter_total = 0;
  /* linear_solver_threshold :200-215 */
  const double abs_threshold = threshold * orc_norm(bb, n);
  /* recover_subspace (SCRATCH) :241-244 : empty subspace, sigma_old = 0 */
  zcplx sub_sigma = 0.0;
  int nb = 0, cap = 16;
  zcplx *vv = malloc(N * cap * sizeof(zcplx));
  zcplx *ww = malloc(N * cap * sizeof(zcplx));
  zcplx *residual = malloc(N * sizeof(zcplx));
  zcplx *new_vector = malloc(N * sizeof(zcplx));
  int conv = 0;

  for (int ishift = 0; ishift < nshift; ++ishift) {                      /* :153 */
    /* orthogonal_subspace :272-300 */
    zcplx diff_sigma = sigma[ishift] - sub_sigma;                        /* :289 */
    zaxpy((int)(N * nb), diff_sigma, vv, ww);                            /* :292 */
    orc_gram_schmidt(1, n, nb, ww, vv);                                  /* :295 */
    sub_sigma = sigma[ishift];                                           /* :298 */

    int iter;
    conv = 0;
    for (iter = 1; iter <= max_iter; ++iter) {                           /* :159 */
      /* residual :312-383 */
      zcopy(n, bb, residual);                                            /* :361 */
      for (int ib = 0; ib < nb; ++ib) {                                  /* :364 */
        zcplx overlap = zdotc(n, ww + (size_t)ib * N, bb);               /* :367 */
        for (int i = 0; i < n; ++i) residual[i] -= overlap * ww[(size_t)ib * N + i]; /* :368 */
      }
      double nrm = orc_dnrm2(2 * n, (const double *)residual);           /* :373 */
      conv = nrm < abs_threshold;                                        /* :374 */
      if (!conv && n == nb) ierr = 3; else ierr = 0;                     /* :376-381 */
      if (conv || ierr != 0) break;                                      /* :165 */
      AA(ctx, sigma[ishift], residual, new_vector, n);                   /* :168 */
      ++n_op;
      ++n_iter_total;
      /* expand_subspace :391-440 (reallocation replaced by capacity doubling; same arithmetic) */
      if (nb + 1 > cap) {
        cap *= 2;
        vv = realloc(vv, N * cap * sizeof(zcplx));
        ww = realloc(ww, N * cap * sizeof(zcplx));
      }
      zcopy(n, new_vector, ww + (size_t)nb * N);                         /* :426 */
      zcopy(n, residual, vv + (size_t)nb * N);                           /* :432 */
      ++nb;
      orc_gram_schmidt(nb, n, nb, ww, vv);                               /* :438 */
    }
    if (!conv) {                                                         /* :176-180 (overwrites ierr=3) */
      ierr = 1;
      break;
    }
    /* obtain_result :453-503 */
    zcplx *x = xx + (size_t)ishift * N;
    memset(x, 0, N * sizeof(zcplx));                                     /* :486 */
    for (int ib = 0; ib < nb; ++ib) {
      zcplx overlap = zdotc(n, ww + (size_t)ib * N, bb);                 /* :496 */
      zaxpy(n, overlap, vv + (size_t)ib * N, x);                         /* :499 */
    }
    if (has_nan(x, n)) {                                                 /* :186-190 */
      ierr = 2;
      break;
    }
  }
  if (st) {
    st->n_op += n_op;
    st->n_outer = n_iter_total;
  }
  free(vv); free(ww); free(residual); free(new_vector);
  return ierr;
}

/* ---------------------------------------------------------------- select_solver.f90:67-161 */
int orc_select_solver(const orc_solver_cfg *cfg, orc_op_fn AA, void *ctx, int n, const zcplx *bb,
                      int nshift, const zcplx *sigma, zcplx *xx, orc_stats *st) {
  int ierr = 1;                                                          /* :121 */
  for (int is = 0; is < cfg->npriority; ++is) {                          /* :122 */
    switch (cfg->priority[is]) {
      case 1:                                                            /* :134-138 */
        ierr = orc_bicgstab(cfg->bicg_lmax, cfg->threshold, cfg->max_iter, AA, ctx, n, bb, nshift, sigma, xx, st);
        break;
      case 2:                                                            /* :140-146: whole sigma each time */
        for (int ishift = 0; ishift < nshift; ++ishift)
          ierr = orc_bicgstab(cfg->bicg_lmax, cfg->threshold, cfg->max_iter, AA, ctx, n, bb, nshift, sigma, xx, st);
        break;
      case 3:                                                            /* :148-152 */
        ierr = orc_linear_solver(cfg->threshold, cfg->max_iter, AA, ctx, n, bb, nshift, sigma, xx, st);
        break;
      default:
        break;
    }
    if (st) st->solver_used = cfg->priority[is];
    if (ierr == 0) break;                                                /* :157 */
  }
  return ierr;
}

/* linear_solver.pf:106 : Ax = MATMUL(AA, xx) + sigma * xx */
void orc_dense_apply(void *ctx, zcplx sigma, const zcplx *x, zcplx *ax, int n) {
  const orc_dense_op *op = (const orc_dense_op *)ctx;
  for (int i = 0; i < n; ++i) ax[i] = sigma * x[i];
  for (int j = 0; j < n; ++j) {
    const zcplx xj = x[j];
    const zcplx *col = op->A + (size_t)j * n;
    for (int i = 0; i < n; ++i) ax[i] += col[i] * xj;
  }
}

/* ---------------------------------------------------------------- parallel.f90:80-138 */
void orc_parallel_task(int nproc, int rank, int ntotal, int *first, int *last, int *num_task) {
  int nmin = ntotal / nproc;                 /* :121 */
  int nrem = ntotal % nproc;                 /* :124 */
  int last_proc = nproc - nrem;              /* :127 */
  for (int p = 0; p < nproc; ++p) num_task[p] = p < last_proc ? nmin : nmin + 1;
  int sum = 0;
  for (int p = 0; p <= rank; ++p) sum += num_task[p];  /* my_rank = rank+1 ; :133 */
  *last = sum;
  *first = sum - num_task[rank] + 1;         /* :134 */
}
