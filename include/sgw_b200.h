/*
 * sgw_b200.h -- C ABI of the B200-native (sm_100a) Sternheimer hot path.
 *
 * Drop-in boundary for QEF/SternheimerGW v0.15: the Fortran driver (phys/coul, phys/green, phys/corr)
 * stays the host and calls these entry points through ISO_C_BINDING in place of
 * algo/linear_solver (multishift BiCGStab(l) / SGW subspace solver) and the H.psi it applies.
 * The reference's plugin point is the procedure-argument callback AA(sigma, x, Ax)
 * (algo/linear_solver/src/select_solver.f90:77-87); because H.psi lives on the GPU too, the library takes
 * the operator's DATA at the places where the reference installs it into QE module globals
 * (phys/coul/src/solve_linter.f90:315-316, phys/green/src/green.f90:88-91, algo/setup/src/gwq_setup.f90:92).
 *
 * Conventions:
 *  - all arrays are HOST pointers owned by the caller (the library copies in/out and owns device memory);
 *  - Fortran column-major, COMPLEX(dp) = interleaved double[2], index arrays 1-based int32;
 *  - plane-wave vectors have leading dimension npwx and are zero padded beyond npw;
 *  - calls are blocking and non-re-entrant per context (the reference keeps its operator in QE globals);
 *  - solver error codes follow the reference: 0 converged, 1 max_iter reached (bicgstab.f90:249-253,
 *    linear_solver.f90:176-180), 2 NaN in the solution (bicgstab.f90:264-267, linear_solver.f90:186-190);
 *    API/CUDA failures are NEGATIVE return values (SGW_E_*), never confused with solver codes;
 *  - there is NO CPU fallback: without a CUDA device sgw_create fails with SGW_E_CUDA.
 */
#ifndef SGW_B200_H
#define SGW_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct sgw_ctx sgw_ctx;
typedef struct { double re, im; } sgw_cplx; /* COMPLEX(dp) */

#define SGW_OK 0
#define SGW_E_ARG (-1)     /* invalid argument */
#define SGW_E_CUDA (-2)    /* CUDA runtime / driver error (see sgw_last_error) */
#define SGW_E_STATE (-3)   /* required set_* call missing */
#define SGW_E_UNSUPPORTED (-4)

/* select_solver_type (algo/linear_solver/src/select_solver.f90:48-62) */
typedef struct {
  int32_t npriority;
  int32_t priority[4]; /* 1 bicgstab multishift, 2 bicgstab without multishift, 3 SGW subspace solver */
  int32_t max_iter;    /* 10000 */
  double threshold;    /* 1e-4 */
  int32_t bicg_lmax;   /* 4 */
} sgw_solver_cfg;

/* per-call statistics (device counters; labels follow data/timing/src/timing.f90) */
typedef struct {
  int64_t n_linear_op;     /* H.psi applications, summed over right-hand sides ("linear operator") */
  int64_t n_kernel_launch; /* kernels of THIS library launched during the call */
  int32_t n_outer_max;     /* max outer BiCGStab iterations over the batch */
  int32_t n_fallback;      /* right-hand sides that fell through to the next solver in the priority list */
  double ms_solver;        /* device time of the solver part ("coul solver" / "green") */
  double ms_linear_op;     /* device time inside H.psi (only when profiling is enabled) */
  double ms_total;         /* device time of the whole call (CUDA events on the library's stream) */
} sgw_stats;

/* ---- context (once per MPI rank / GPU, after mp_startup, main/src/gw.f90:83) ---- */
int sgw_create(int device, sgw_ctx **ctx);
int sgw_destroy(sgw_ctx *ctx);
/* Run the library's kernels on a caller-owned CUDA stream (a cudaStream_t passed as void*) instead of the context's own
 * non-blocking stream; NULL restores the context's stream.  The caller keeps ownership; calls stay blocking. */
int sgw_set_stream(sgw_ctx *ctx, void *cuda_stream);
const char *sgw_last_error(const sgw_ctx *ctx);
int sgw_get_stats(const sgw_ctx *ctx, sgw_stats *out);        /* stats of the last solver-level call */
int sgw_set_profiling(sgw_ctx *ctx, int on);                  /* time H.psi separately (adds syncs) */
/* return the solver workspace and the parked table buffers to the driver (the library keeps its scratch buffers between
 * calls and sizes its batches from the free memory; a host that needs device memory for something else calls this) */
int sgw_release_workspace(sgw_ctx *ctx);
int sgw_device_synchronize(sgw_ctx *ctx);
/* The reference writes solver warnings to stdout ("WARNING: BiCGstab algorithm did not converge in N iterations."
 * bicgstab.f90:250, "First choice of solver did not converge, try a different one" select_solver.f90:126, "WARNING:
 * SternheimerGW linear solver did not converge" linear_solver.f90:177).  The library never prints: it hands the same lines
 * (once per batch, with the number of right-hand sides concerned) to this callback, if one is set. */
typedef void (*sgw_message_fn)(const char *message, void *user);
int sgw_set_message_callback(sgw_ctx *ctx, sgw_message_fn fn, void *user);
/* per-kernel-class device time of the last solver-level call (needs sgw_set_profiling(ctx, 1)): class i took ms[i]
 * milliseconds in regions[i] timed regions (one region = one kernel launch for fft_plane, gemm_project, gemm_expand,
 * shift_fused and rho_plane).  Labels via sgw_profile_class_name; the reference's clocks `linear operator` /
 * `coul solver` (data/timing/src/timing.f90:64,104) are ms_linear_op / ms_solver of sgw_stats. */
int sgw_get_profile(const sgw_ctx *ctx, int max_classes, double *ms, int64_t *regions, int *nclasses);
const char *sgw_profile_class_name(int cls);

/* ---- L0: FFT grid and local potential (result of set_vrs, gwq_setup.f90:92; QE dffts) ---- */
/* nr1, nr2, nr3: FFT dimensions (any 2^a 3^b 5^c <= 960, optionally times one 7 and/or 11); nr1x >= nr1 ...: physical
 * dimensions of the caller's real-space arrays (dffts%nr1x; equal to nr except on padded builds).  With padding, vrs, dvbarein /
 * drhoscf and every 1-based FFT index (nl, nl_igk) are in the padded layout, exactly as the Fortran host holds them. */
int sgw_set_grid(sgw_ctx *ctx, int nr1, int nr2, int nr3, int nr1x, int nr2x, int nr3x);
int sgw_set_vloc(sgw_ctx *ctx, const double *vrs /* nnr */);

/* ---- L1: operator data of one k-point = init_us_2 + g2_kin + globals evq/alpha_pv/nbnd_occ
 *      (solve_linter.f90:315-316, green.f90:88-91).  slot: caller-chosen id >= 0. ---- */
int sgw_set_kpoint(sgw_ctx *ctx, int slot, int npw, int npwx, const int32_t *nl_igk /* npw, 1-based */,
                   const double *g2kin /* npw */, int nkb, const sgw_cplx *vkb /* npwx x nkb */,
                   const double *dion /* nkb x nkb */, int nbnd_occ, const sgw_cplx *evq /* npwx x nbnd_occ */,
                   double alpha_pv);
/* dense fake backend of the reference's unit test (linear_solver.pf:106: Ax = MATMUL(A,x) + sigma x) */
int sgw_set_dense_operator(sgw_ctx *ctx, int slot, int n, const sgw_cplx *A /* n x n */, int lda);

/* linear_op (algo/linear_solver/src/linear_op.f90:46): A_psi = (H + omega S + alpha_pv P_v) psi, batched */
int sgw_linear_op(sgw_ctx *ctx, int slot, int nvec, const sgw_cplx *omega /* nvec */, double alpha_pv,
                  const sgw_cplx *psi, int ldpsi, sgw_cplx *apsi, int ldapsi);

/* select_solver (select_solver.f90:67), batched over right-hand sides.
 * b: ldb x nrhs; sigma: nshift x nrhs (sigma(1,:) = seed system); x element (ig, ishift, irhs) at
 * x[ig + stride_shift*ishift + stride_rhs*irhs] so that x may alias dpsi(npwx, nbnd, num_omega)
 * (solve_linter.f90:368-369: stride_rhs = npwx, stride_shift = npwx*nbnd).  use_alpha_pv = 0 gives
 * green_operator (green.f90:283), 1 coulomb_operator (solve_linter.f90:708).  ierr: nrhs entries. */
int sgw_solve_multishift(sgw_ctx *ctx, int slot, const sgw_solver_cfg *cfg, int use_alpha_pv, int nrhs,
                         const sgw_cplx *b, int ldb, int nshift, const sgw_cplx *sigma, sgw_cplx *x,
                         int64_t stride_shift, int64_t stride_rhs, int32_t *ierr);

/* ---- L2/L3: screened Coulomb (phys/coul) ---- */
/* cell + density-sphere data used by dv_of_drho / coulomb: gvect g, dffts%nl, qpoint xq, cell omega, tpiba2 */
int sgw_set_system(sgw_ctx *ctx, double omega_cell, double tpiba2, int ngm, const double *g /* 3 x ngm */,
                   const int32_t *nl /* ngm, 1-based */);
int sgw_set_q(sgw_ctx *ctx, const double *xq /* 3 */);
/* number of (k, k+q) pairs of this pool (qpoint nksq) and the data of one pair: slot_kq must have been
 * set with sgw_set_kpoint; evc(npwx, nbnd) at k with its nl map, eigenvalues et(nbnd), weight wk(ikk) */
int sgw_set_nksq(sgw_ctx *ctx, int nksq);
int sgw_set_kpair(sgw_ctx *ctx, int ik, int slot_kq, int npw_k, const int32_t *nl_igk_k, int nbnd,
                  const sgw_cplx *evc, const double *et, double wk);

/* Metals.  [QE] klist (lgauss, degauss, ngauss) and ener (ef) as read by LR_Modules/orthogonalize.f90, which solve_linter.f90:337
 * calls: with lgauss != 0 the right-hand sides are built with the smeared projector of S. de Gironcoli, PRB 51, 6773 (1995) and
 * the solutions are scaled by wg(ibnd, ikk) / wk(ikk) (solve_linter.f90:373).  ngauss: -99 Fermi-Dirac, -1 Marzari-Vanderbilt,
 * 0 Gaussian, n > 0 Methfessel-Paxton.  lgauss = 0 (default) is the insulator path.  Direct and self-consistent branch;
 * sgw_coulomb_q0G0 (the insulator treatment of the head) refuses lgauss. */
int sgw_set_smearing(sgw_ctx *ctx, int lgauss, double ef, double degauss, int ngauss);
/* per (k, k+q) pair, after sgw_set_kpair: ALL nbnd bands of evq (npwx x nbnd, the first nbnd_occ(ikq) of them are the ones
 * given to sgw_set_kpoint) with et(:, ikq), the number of bands of the solver loop nbnd_occ(ikk) <= nbnd of sgw_set_kpair, and
 * wg(ibnd, ikk) / wk(ikk) for those bands */
int sgw_set_kpair_metal(sgw_ctx *ctx, int ik, int nbnd, const sgw_cplx *evq_all, const double *et_q, int nbnd_occ_k,
                        const double *wg_over_wk);

/* control_gw globals of the self-consistent branch (main/src/gw_input.yml: num_iter_coul -> niter_gw, alpha_mix,
 * tr2_gw, num_mix_coul -> nmix_gw <= 8 = maxter of mix_pot_c.f90:76): alpha_mix has niter_gw entries. */
int sgw_set_mixing(sgw_ctx *ctx, int niter_gw, const double *alpha_mix, double tr2_gw, int nmix_gw);
/* control_gw solve_direct (input solve_coul = 'direct' | 'iter', gwq_readin.f90): 1 (default) -> sgw_coulomb returns
 * eps = delta - v chi0 columns (coulomb.f90:153-157); 0 -> every perturbation runs the self-consistent solve_linter with
 * num_iter = niter_gw (needs sgw_set_mixing, niter_gw > 1, coulomb.f90:106-110) and scrcoul holds dV_scf = (eps^-1 - 1) columns */
int sgw_set_solve_direct(sgw_ctx *ctx, int solve_direct);
/* iterations the last self-consistent sgw_solve_linter / sgw_coulomb took ("iter #" line, solve_linter.f90:604) */
int sgw_get_scf_iterations(const sgw_ctx *ctx);
/* solve_linter (phys/coul/src/solve_linter.f90:55): dvbarein(nnr) real-space perturbation, freq(nfreq).
 *  num_iter = 1: direct branch, drhoscf(nnr, nfreq) = -dV_H (:598);
 *  num_iter > 1: self-consistent branch (:376-460 per-frequency solves with dV_scf psi added, :564-582 complex Broyden
 *  mixing mix_potential_c, needs sgw_set_mixing), drhoscf = dvscfin (:610).
 * ierr_out: solver code (0/1/2/3), or 10 when self-consistency was not reached within num_iter (:588-591). */
int sgw_solve_linter(sgw_ctx *ctx, const sgw_solver_cfg *cfg, int num_iter, const sgw_cplx *dvbarein, int nfreq,
                     const sgw_cplx *freq, sgw_cplx *drhoscf, int32_t *ierr_out);
/* coulomb (phys/coul/src/coulomb.f90:29): perturbations igstart..igstart+ntask-1 of ig_unique, batched over
 * perturbations x k x bands.  scrcoul(ngc, nfs, ntask). */
int sgw_coulomb(sgw_ctx *ctx, const sgw_solver_cfg *cfg, int igstart, int ngc, int ntask,
                const int32_t *ig_unique, int nfs, const sgw_cplx *fiu, sgw_cplx *scrcoul, int32_t *ierr_out);
/* Grid on which the last sgw_coulomb accumulated Delta-rho ([QE] incdrhoscf, solve_linter.f90:489-497).  coulomb.f90:143-157
 * keeps only the first ngc G vectors of Delta-rho, so the library uses the smallest FFT box that yields those components
 * without aliasing (n >= M_k + M_k+q + M_out + 1 per axis; identical to the full-box result up to rounding).
 * Returns 1 and the box in dims when a reduced box was used, 0 when the full dffts box was used. */
int sgw_get_rho_grid(const sgw_ctx *ctx, int *dims /* 3 */);
/* coulomb_q0G0 (phys/coul/src/coulomb_q0G0.f90:31): head element at the (shifted) q currently set */
int sgw_coulomb_q0G0(sgw_ctx *ctx, const sgw_solver_cfg *cfg, int nfs, const sgw_cplx *fiu, sgw_cplx *eps_m,
                     int32_t *ierr_out);
/* unfold_w (algo/symmetry/src/unfold_w.f90:84, identity-symmetry case) and invert_epsilon
 * (phys/coul/src/invert_epsilon.f90:23): scrcoul_g(ngc, ngc, nfs) in place */
int sgw_unfold_w(sgw_ctx *ctx, int ngc, int nfs, int ngmunique, const int32_t *ig_unique,
                 const sgw_cplx *scrcoul_in /* ngc x nfs x ngmunique */, sgw_cplx *scrcoul_out /* ngc x ngc x nfs */);
/* unfold_w with use_symm (algo/symmetry/src/unfold_w.f90:23-131, the default of main/src/gw_input.yml:138): the rows of the
 * symmetry-unique G (ig_unique, from stern_symm.f90) are filled as above, every other row ig is the row of sym_friend(ig) rotated by
 * the operation sym_ig(ig): out(ig, gmapsym(igp, invs(R)), iw) = out(sym_friend(ig), igp, iw) eigv(sym_friend(ig), R) CONJG(eigv(igp, R)).
 * sym_ig, sym_friend: num_g_corr entries (1-based, entries of unique G ignored); gmapsym, eigv: num_g_corr x nsym column-major as
 * gmap_sym.f90 returns them; invs: nsym (1-based).  scrcoul_out is fully overwritten.  Call with the identity only (nsymq = 1) is
 * sgw_unfold_w. */
int sgw_unfold_w_symm(sgw_ctx *ctx, int num_g_corr, int nfs, int ngmunique, const int32_t *ig_unique, int nsym,
                      const int32_t *sym_ig, const int32_t *sym_friend, const int32_t *gmapsym, const sgw_cplx *eigv,
                      const int32_t *invs, const sgw_cplx *scrcoul_in, sgw_cplx *scrcoul_out);

int sgw_invert_epsilon(sgw_ctx *ctx, int ngc, int nfs, sgw_cplx *scrcoul_g, int lgamma);

/* ---- L2': Green's function (phys/green/src/green.f90:105) ----
 * green(ngc, ngp, nfreq): for every igp, b = -e_{map(fft_map(igp))}, shifts -omega; rows scattered through map
 * with the reference's strict mask (map > 0 .AND. map < num_g, green.f90:212). */
int sgw_green_function(sgw_ctx *ctx, int slot, const sgw_solver_cfg *cfg, int ngc, const int32_t *map,
                       int ngp, const int32_t *fft_map, int nfreq, const sgw_cplx *omega, sgw_cplx *green,
                       int32_t *ierr_out);

/* ---- analytic continuation of W (algo/analytic/src/analytic.f90) ----
 * freqbins_type (algo/grid/src/freqbins.f90:42-105): the members the continuation and the G W convolution read. */
typedef struct {
  int32_t imag_sigma;       /* convolution along the imaginary (1) or real (0) axis */
  int32_t freq_symm_coul;   /* 0 no_symmetry, 1 even_symmetry, 2 square_symmetry (freqbins.f90:31-35) */
  int32_t num_solver;       /* SIZE(freq%solver): FREQUENCIES card */
  const sgw_cplx *solver;
  int32_t num_coul;         /* integration mesh of the convolution and its weights */
  const sgw_cplx *coul;
  const double *weight;
  int32_t num_sigma;        /* frequencies of the self-energy */
  const sgw_cplx *sigma;
} sgw_freqbins;
#define SGW_GODBY_NEEDS 1   /* analytic.f90:39-60 model_coul values */
#define SGW_PADE_APPROX 2
#define SGW_PADE_ROBUST 3   /* 'pade robust' (pade_robust.f90): solver frequencies on a circle, coefficients [deg_num, deg_den, num, den] */
#define SGW_AAA_APPROX 4    /* 'aaa' (vendor/analytic/src/aaa.f90): greedy AAA fit, coefficients [position | value | weight] */
#define SGW_AAA_POLE 5      /* 'aaa pole': AAA fit, then poles and residues; coefficients [pole | residue] */
/* freq%num_freq() = size of the symmetrised mesh (freqbins_symm, freqbins.f90:243-305); < 0 on error
 * (more than one frequency below 1e-14 with even symmetry, as the reference's errore). */
int sgw_freqbins_num_freq(const sgw_freqbins *freq);
/* coulpade (phys/coul/src/coulpade.f90:36): scrcoul_g(ig, :, :) *= factor(ig); factor = truncate(q + G_ig) is host code */
int sgw_coulpade(sgw_ctx *ctx, int ngc, int nfreq, const double *factor /* ngc */, sgw_cplx *scrcoul_g);
/* analytic_coeff (analytic.f90:50): scrcoul_g(ngc, ngc, num_freq()) in place -> coefficients of the model
 * (Godby-Needs godby_needs.f90:34, Pade pade.f90 pade_coeff incl. the mirrored frequencies of freqbins_symm, AAA aaa.f90
 * with max_point = num_freq() / 3 and the relative threshold `thres`; the AAA weights are defined up to a common phase;
 * 'aaa pole': the same fit followed by aaa_pole_residual + pole_correction, poles in no particular order) */
int sgw_analytic_coeff(sgw_ctx *ctx, int model_coul, double thres, const sgw_freqbins *freq, int ngc, sgw_cplx *scrcoul_g);
/* pade_robust (algo/analytic/src/pade_robust.f90:177; the routine the reference's unit test algo/analytic/test/pade.pf
 * exercises): func(num_point) sampled on the circle radius * exp(2 pi i j / num_point); deg_num / deg_den in: requested, out:
 * found; coeff_num / coeff_den need room for the requested degrees + 1; tol_coeff, tol_fft <= 0 select the reference's defaults
 * (1e-14, tol_coeff). */
int sgw_pade_robust(sgw_ctx *ctx, double radius, int num_point, const sgw_cplx *func, int *deg_num, int *deg_den,
                    sgw_cplx *coeff_num, sgw_cplx *coeff_den, double tol_coeff, double tol_fft);
/* aaa_pole_residual (vendor/analytic/src/aaa.f90:93; the routine vendor/analytic/test/testAAA.pf:45-74,387-422 exercises) on a
 * GIVEN barycentric approximant of m support points: the m - 1 finite poles (the reference: ZGGEV on the arrowhead pencil
 * find_pole:388; here Aberth-Ehrlich on the same polynomial) and their four-point residues (calculate_residual:464).
 * pole / residual need room for m - 1 entries; *num_pole out.  Poles come in no particular order. */
int sgw_aaa_pole_residual(sgw_ctx *ctx, int m, const sgw_cplx *position, const sgw_cplx *value, const sgw_cplx *weight,
                          sgw_cplx *pole, sgw_cplx *residual, int *num_pole);
/* analytic_eval (analytic.f90:211) at nout frequencies at once: scrcoul(ig, igp, iout) =
 * model(coeff(gmapsym(ig), gmapsym(igp), :), freq%symmetrize(freq_out(iout))) -- the G-space block the reference
 * stores in the corner of scrcoul(nnr_c, nnr_c'). */
int sgw_analytic_eval(sgw_ctx *ctx, int model_coul, const sgw_freqbins *freq, int ngc, const int32_t *gmapsym /* ngc, 1-based */,
                      const sgw_cplx *scrcoul_coeff /* ngc x ngc x num_freq() */, int nout, const sgw_cplx *freq_out,
                      sgw_cplx *scrcoul /* ngc x ngc x nout */);

/* ---- Sigma_c = G W (phys/corr/src/sigma.f90, data/fft/src/fft6.f90) ----
 * grid%corr_fft (algo/grid/src/sigma_grid.f90): box of the correlation cutoff and nl of its ngm_c G vectors (1-based).
 * One image: corr_par_fft = corr_fft (the reference's image distribution of G'/r' is replaced by sharding the
 * (k, q) configurations over GPUs). */
int sgw_set_corr_grid(sgw_ctx *ctx, int nr1, int nr2, int nr3, int ngm_c, const int32_t *nl_c);
/* invfft6 / fwfft6 (fft6.f90:231 / :84) in place on f(nnr_c, nnr_c): invfft6 reads f(:ngm_c, :ngm_c) and fills all of
 * f; fwfft6 reads all of f and writes f(:ngm_c, :ngm_c) (the rest is left unchanged; the reference leaves scratch). */
int sgw_invfft6(sgw_ctx *ctx, double omega, sgw_cplx *f);
int sgw_fwfft6(sgw_ctx *ctx, double omega, sgw_cplx *f);
/* sigma_correlation (sigma.f90:528): Green's function at freq%green(mu) for the operator in `slot` (green_prepare data:
 * map), its 6-D transform, and for every (omega_sigma, omega_green) pair W(omega_sigma - omega_green) by analytic_eval,
 * the real-space product (sigma_prod, :417) and the transform back; sigma(ngm_c, ngm_c, num_sigma) += result.
 * The sum over omega_green is taken in real space before ONE forward transform per omega_sigma (linear, same result).
 * ierr_out: solver code of the Green's function (0/1/2/3). */
int sgw_sigma_correlation(sgw_ctx *ctx, int slot, const sgw_solver_cfg *cfg_green, double omega_cell, double mu,
                          sgw_cplx alpha, int model_coul, const sgw_freqbins *freq, int ngm_c, const int32_t *map /* ngm_c */,
                          const int32_t *gmapsym /* ngm_c, 1-based */, const sgw_cplx *coulomb /* ngm_c x ngm_c x num_freq() */,
                          sgw_cplx *sigma /* ngm_c x ngm_c x num_sigma */, int32_t *ierr_out);

/* ---- data/parallel/src/parallel.f90:80 parallel_task: contiguous blocks, remainder to the LAST ranks.
 * rank 0-based; first/last 1-based; num_task[nproc]. ---- */
int sgw_parallel_task(int nproc, int rank, int num_task_total, int32_t *first_task, int32_t *last_task,
                      int32_t *num_task);

/* micro-benchmark hooks (device-resident, no host copies inside): apply the operator `reps` times to
 * nvec resident random vectors; returns device milliseconds per repetition of each part. */
int sgw_bench_linear_op(sgw_ctx *ctx, int slot, int nvec, int reps, double *ms_total, double *ms_fft,
                        double *ms_gemm);

#ifdef __cplusplus
}
#endif
#endif /* SGW_B200_H */
