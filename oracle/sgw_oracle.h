/*
 * sgw_oracle.h -- CPU restatement ("oracle") of the SternheimerGW Sternheimer hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in the product path (sternheimergw_b200/, include/,
 * the CUDA library) may include, link or call this.  Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs use it, as the checker / CPU baseline.
 *
 * The reference (Fortran 2003 + Quantum ESPRESSO 6.3 @ 7357cdb, main/src/version.f90.in:100)
 * cannot be compiled in this environment (no Fortran compiler, no QE, no MPI), so this is a
 * plain-C restatement that follows the reference file by file, in the reference's execution
 * order (one band, one vector at a time, unfused BLAS-1 exactly where the reference calls Z*).
 *
 * Parity pinning:
 *   - the solver half (solver.c) is pinned by the reference's own golden vector
 *     algo/linear_solver/test/lin_prob.xml.bz2 and the assertions of
 *     algo/linear_solver/test/linear_solver.pf:216,248,338 (tests/test_oracle_fixture.py);
 *   - the plane-wave half (pw.c: h_psi, dvqpsi_us, incdrhoscf, dv_of_drho, coulomb, ...) restates
 *     QE routines whose source is NOT under /root/reference: **parity unpinned** by reference
 *     tests; it is anchored instead by independent numpy formulas (dense H, numpy.fft,
 *     sum-over-states chi0) in tests/test_oracle_pw.py.
 *
 * All arrays are Fortran column-major; complex = interleaved double[2] (C99 double _Complex);
 * index arrays are 1-based int32 exactly as the Fortran caller holds them.
 */
#ifndef SGW_ORACLE_H
#define SGW_ORACLE_H

#include <complex.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef double _Complex zcplx;

/* The reference's operator callback AA(sigma, x, Ax)  (select_solver.f90:77-87). */
typedef void (*orc_op_fn)(void *ctx, zcplx sigma, const zcplx *x, zcplx *ax, int n);

/* select_solver_type (select_solver.f90:48-62) */
typedef struct {
  int npriority;
  int priority[4];   /* 1 = bicgstab multishift, 2 = bicgstab no multishift, 3 = SGW subspace solver */
  int max_iter;      /* default 10000 */
  double threshold;  /* default 1e-4 */
  int bicg_lmax;     /* default 4 */
} orc_solver_cfg;

/* statistics (not in the reference; used for iteration-count parity and the CPU baseline) */
typedef struct {
  long n_op;          /* operator applications */
  int n_outer;        /* outer iterations of the last bicgstab call / total subspace iterations */
  int solver_used;    /* priority entry that produced the result */
} orc_stats;

/* ---- data/algebra ---- */
double orc_norm(const zcplx *v, int n);                                   /* norm.f90:73-107 (ZLANGE 'F') */
double orc_dnrm2(int n, const double *x);                                 /* reference BLAS DNRM2 */
void orc_gram_schmidt(int first, int n, int nb, zcplx *basis, zcplx *vector); /* gram_schmidt.f90:35-135 */

/* ---- algo/linear_solver ---- */
int orc_bicgstab(int lmax, double threshold, int max_iter, orc_op_fn AA, void *ctx, int n,
                 const zcplx *bb, int nshift, const zcplx *sigma, zcplx *xx, orc_stats *st);
int orc_linear_solver(double threshold, int max_iter, orc_op_fn AA, void *ctx, int n,
                      const zcplx *bb, int nshift, const zcplx *sigma, zcplx *xx, orc_stats *st);
int orc_select_solver(const orc_solver_cfg *cfg, orc_op_fn AA, void *ctx, int n, const zcplx *bb,
                      int nshift, const zcplx *sigma, zcplx *xx, orc_stats *st);

/* dense fake backend of linear_solver.pf:106  (Ax = MATMUL(A,x) + sigma x) */
typedef struct {
  int n;
  const zcplx *A; /* n x n column-major */
} orc_dense_op;
void orc_dense_apply(void *ctx, zcplx sigma, const zcplx *x, zcplx *ax, int n);

/* ---- data/parallel ---- */
void orc_parallel_task(int nproc, int rank, int ntotal, int *first, int *last, int *num_task /*[nproc]*/);

/* ---- plane-wave half (pw.c) ---- */

/* FFT grid + local potential: QE dffts + vrs (gwq_setup.f90:92) */
typedef struct {
  int nr1, nr2, nr3;
  const double *vrs; /* nnr, real-space local potential (Ry), column-major (nr1,nr2,nr3) */
} orc_grid;

/* Operator data of one k-point, i.e. what init_us_2 + g2_kin + the globals evq/alpha_pv/nbnd_occ hold
 * (solve_linter.f90:315-316, green.f90:88-91). */
typedef struct {
  int npw, npwx;
  const int32_t *nl_igk; /* npw, 1-based linear FFT index dffts%nl(igk_k(ig,ik)) */
  const double *g2kin;   /* npw, |k+G|^2 tpiba2 (Ry) */
  int nkb;
  const zcplx *vkb;      /* npwx x nkb */
  const double *dion;    /* nkb x nkb (deeq expanded to the beta index, real) */
  int nbnd_occ;
  const zcplx *evq;      /* npwx x nbnd_occ */
  double alpha_pv;
} orc_kpoint;

typedef struct {
  const orc_grid *grid;
  const orc_kpoint *kp;
  double alpha_pv; /* alpha_pv used by the callback (0 for green_operator) */
  zcplx *work;     /* nnr scratch */
  zcplx *becp;     /* max(nkb, nbnd) scratch */
} orc_pw_op;

void orc_fft3d(zcplx *f, int nr1, int nr2, int nr3, int sign); /* sign=+1: invfft (unscaled), -1: fwfft (scaled 1/nnr) */
void orc_h_psi(const orc_grid *g, const orc_kpoint *kp, const zcplx *psi, zcplx *hpsi, zcplx *work, zcplx *becp);
void orc_linear_op(const orc_grid *g, const orc_kpoint *kp, zcplx omega, double alpha_pv, const zcplx *psi,
                   zcplx *apsi, zcplx *work, zcplx *becp); /* linear_op.f90:46-144 */
void orc_pw_apply(void *ctx, zcplx sigma, const zcplx *x, zcplx *ax, int n); /* coulomb_operator / green_operator */

/* one (k, k+q) pair of the W step (solve_linter.f90:288-316) */
typedef struct {
  orc_kpoint kq;          /* operator at k+q (evq inside) */
  int npw_k;              /* plane waves at k */
  const int32_t *nl_igk_k;/* npw_k, 1-based */
  int nbnd;               /* bands in evc (dvqpsi_us transforms all of them) */
  const zcplx *evc;       /* npwx x nbnd at k */
  const double *et;       /* nbnd eigenvalues at k (Ry) */
  double wk;              /* k-point weight (QE: sum_k wk = 2) */
} orc_kpair;

typedef struct {
  orc_grid grid;
  int nks;
  const orc_kpair *kp;
  double omega_cell;      /* cell volume (bohr^3) */
  double tpiba2;
  double xq[3];           /* q in units of 2pi/alat (cartesian) */
  int ngm;                /* G vectors of the density sphere */
  const double *g;        /* 3 x ngm (cartesian, 2pi/alat) */
  const int32_t *nl;      /* ngm, 1-based dffts%nl */
} orc_system;

/* metals: what [QE] orthogonalize (lgauss) and solve_linter.f90:373 read beyond the insulator case, per (k, k+q) pair */
typedef struct {
  int nbnd;                 /* all bands of evq (>= nbnd_occ(ikq)) */
  const zcplx *evq_all;     /* npwx x nbnd at k+q */
  const double *et_q;       /* nbnd eigenvalues at k+q (Ry) */
  int nocc_k;               /* nbnd_occ(ikk): bands of the solver loop */
  const double *wg_over_wk; /* nocc_k: wg(ibnd, ikk) / wk(ikk) */
} orc_metal_pair;
/* klist/ener globals; lgauss = 0 restores the insulator path.  `pairs` (nks entries) must outlive the calls that use it. */
void orc_set_smearing(int lgauss, double ef, double degauss, int ngauss, int nks, const orc_metal_pair *pairs);
double orc_wgauss(double x, int n);
double orc_w0gauss(double x, int n);
double orc_metal_weight(double e_i, double e_j, int j_in_projector, double alpha_pv, double ef, double degauss, int ngauss);

/* solve_linter.f90:55-624, direct branch (num_iter = 1).  drhoscf: nnr x nfreq */
int orc_solve_linter(const orc_system *sys, const orc_solver_cfg *cfg, const zcplx *dvbarein, int nfreq,
                     const zcplx *freq, zcplx *drhoscf, orc_stats *st, int nthreads);
/* solve_linter.f90:55-624 with num_iter > 1: the self-consistent branch (:376-460) + mix_potential_c (mix_pot_c.f90:25).
 * alpha_mix: num_iter entries (control_gw alpha_mix), tr2_gw, nmix_gw as in the reference; drhoscf = dvscfin on exit.
 * Returns 0, a solver ierr, 10 (not converged within num_iter, solve_linter.f90:588-591) or 11 (Broyden matrix singular). */
int orc_solve_linter_iter(const orc_system *sys, const orc_solver_cfg *cfg, int num_iter, const double *alpha_mix,
                          double tr2_gw, int nmix_gw, const zcplx *dvbarein, int nfreq, const zcplx *freq,
                          zcplx *drhoscf, orc_stats *st, int nthreads, int *iter_done);
/* bench-only: restrict the band loops of orc_solve_linter to lo <= ibnd < hi (bounded CPU-baseline samples) */
void orc_set_band_window(int lo, int hi);
/* coulomb.f90:29-176.  scrcoul: ngc x nfs x ntask.  ig_unique (1-based G indices), igstart 1-based. */
int orc_coulomb(const orc_system *sys, const orc_solver_cfg *cfg, int igstart, int ngc, int ntask,
                const int32_t *ig_unique, int nfs, const zcplx *fiu, zcplx *scrcoul, orc_stats *st,
                int nthreads);
/* coulomb_q0G0.f90:31-158 */
int orc_coulomb_q0G0(const orc_system *sys, const orc_solver_cfg *cfg, int nfs, const zcplx *fiu, zcplx *eps_m,
                     orc_stats *st);
/* unfold_w.f90:84 (identity-symmetry case) */
void orc_unfold_w(int ngc, int nfs, int ngmunique, const int32_t *ig_unique, const zcplx *in, zcplx *out);
/* invert_epsilon.f90:23-90 */
int orc_invert_epsilon(int ngc, int nfs, zcplx *scrcoul_g, int lgamma);
/* green.f90:105-226 : green_part(num_g, nfreq) per G', scattered through map.  green: ngc x ngp x nfreq */
int orc_green_function(const orc_grid *g, const orc_kpoint *kp, const orc_solver_cfg *cfg, int ngc,
                       const int32_t *map /*ngc, 1-based index into k sphere or 0*/, int ngp,
                       const int32_t *fft_map /*ngp, 1-based into map*/, int nfreq, const zcplx *omega,
                       zcplx *green, orc_stats *st, int nthreads);

#ifdef __cplusplus
}
#endif
#endif
