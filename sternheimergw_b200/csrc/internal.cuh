// internal.cuh -- context, device descriptors and helpers of libsgw_b200 (not part of the C ABI).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include <map>
#include <string>
#include <vector>

#include "../../include/sgw_b200.h"
#include "fft_core.h"

namespace sgw {

typedef double2 cplx;

// ---------------------------------------------------------------- complex helpers
__host__ __device__ __forceinline__ cplx cmake(double a, double b) { cplx r; r.x = a; r.y = b; return r; }
__host__ __device__ __forceinline__ cplx cadd(cplx a, cplx b) { return cmake(a.x + b.x, a.y + b.y); }
__host__ __device__ __forceinline__ cplx csub(cplx a, cplx b) { return cmake(a.x - b.x, a.y - b.y); }
__host__ __device__ __forceinline__ cplx cmul(cplx a, cplx b) { return cmake(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }
__host__ __device__ __forceinline__ cplx cneg(cplx a) { return cmake(-a.x, -a.y); }
__host__ __device__ __forceinline__ cplx cconj(cplx a) { return cmake(a.x, -a.y); }
__host__ __device__ __forceinline__ cplx cscale(double s, cplx a) { return cmake(s * a.x, s * a.y); }
// y + a*x
__host__ __device__ __forceinline__ cplx cfma(cplx a, cplx x, cplx y) {
  return cmake(y.x + (a.x * x.x - a.y * x.y), y.y + (a.x * x.y + a.y * x.x));
}
// Smith's algorithm is what gfortran emits for complex division; plain formula is within 1-2 ulp here
__host__ __device__ __forceinline__ cplx cdiv(cplx a, cplx b) {
  double r, d;
  if (fabs(b.x) >= fabs(b.y)) {
    r = b.y / b.x; d = b.x + b.y * r;
    return cmake((a.x + a.y * r) / d, (a.y - a.x * r) / d);
  }
  r = b.x / b.y; d = b.y + b.x * r;
  return cmake((a.x * r + a.y) / d, (a.y * r - a.x) / d);
}

#ifdef __CUDACC__
// FP64 tensor-core MMA and asynchronous global -> shared staging shared by gemm.cu and sigma.cu
__device__ __forceinline__ void dmma(double &c0, double &c1, double a, double b) {
  asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
      : "+d"(c0), "+d"(c1)
      : "d"(a), "d"(b));
}
// 16-byte asynchronous global -> shared copy; bytes = 0 zero-fills the destination (tile edges)
__device__ __forceinline__ void cp_async16(void *smem, const void *gmem, int bytes) {
  const unsigned sa = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(sa), "l"(gmem), "r"(bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory"); }

#endif

// ---------------------------------------------------------------- device descriptors
struct GridDev {
  int nx, ny, nz;
  int rx1, rx2, ry1, ry2, rz1, rz2;
  int pitchx;              // plane row pitch in smem (odd)
  const cplx *twx, *twy, *twz;
};

// Column structure of a plane-wave sphere on the FFT box.  Coefficient vectors that the library keeps on
// the device are stored in "column order": entry p belongs to column colof[p] (columns sorted by (y,x)),
// z index zof[p]; col_ptr delimits columns.
struct SphereDev {
  int npw, ncol, nxs;
  const int *col_x, *col_y, *col_ptr, *colof, *zof, *xs;
  const int *col_off;      // col_y * (nx | 1) + col_x: offset of the column inside a shared-memory plane
};

// An FFT box other than the context's main one (used for the alias-free coarse grid of the Delta-rho accumulation)
struct FftGrid {
  int n1 = 0, n2 = 0, n3 = 0;
  Plan1D px, py, pz;
  cplx *d_twx = nullptr, *d_twy = nullptr, *d_twz = nullptr;
};

struct Sphere {
  int npw = 0, ncol = 0, nxs = 0;
  std::vector<int> perm;   // perm[p] = caller's 0-based index of internal entry p
  std::vector<int> h_col_x, h_col_y, h_col_ptr, h_zof;   // host copies (box coordinates of the grid it was built on)
  int *d_col_x = nullptr, *d_col_y = nullptr, *d_col_ptr = nullptr, *d_colof = nullptr, *d_zof = nullptr,
      *d_xs = nullptr, *d_perm = nullptr, *d_col_off = nullptr;
  short *d_ytab = nullptr;         // k_plane_vloc index table (fft.cu build_sphere), valid for the y plan (ytab_ry1, ytab_ry2)
  int ytab_ry1 = 0, ytab_ry2 = 0;
  short *d_ztab = nullptr;         // TMA z-pass gather table (fft.cu build_ztab), valid for the z plan (ztab_rz1, ztab_rz2)
  int ztab_rz1 = 0, ztab_rz2 = 0, zmaxlen = 0;
  SphereDev dev() const {
    SphereDev s; s.npw = npw; s.ncol = ncol; s.nxs = nxs; s.col_x = d_col_x; s.col_y = d_col_y;
    s.col_ptr = d_col_ptr; s.colof = d_colof; s.zof = d_zof; s.xs = d_xs; s.col_off = d_col_off; return s;
  }
};

struct KSlot {
  bool set = false, dense = false;
  int npw = 0, npwx = 0, nkb = 0, nbnd = 0;
  double alpha_pv = 0.0;
  Sphere sph;
  double *d_g2kin = nullptr;   // npwx (column order, zero padded)
  cplx *d_P = nullptr;         // npwx x (nkb + nbnd): [vkb | evq] rows in column order
  double *d_dion = nullptr;    // nkb x nkb
  // the same matrix by rows, exact zeros dropped (QE's deeq is block diagonal per atom): row i = entries d_dion_ptr[i] .. [i+1]
  int *d_dion_ptr = nullptr, *d_dion_col = nullptr;
  double *d_dion_val = nullptr;
  cplx *d_A = nullptr;         // dense backend n x n
};

struct KPair {
  bool set = false;
  int slot = -1, npw_k = 0, nbnd = 0;
  double wk = 0.0;
  Sphere sph_k;
  cplx *d_evc = nullptr;       // npwx x nbnd rows in sph_k column order
  std::vector<double> et;
  // metals (sgw_set_kpair_metal): all bands of evq with their energies, nbnd_occ(ikk), wg(:, ikk) / wk(ikk)
  int nbnd_all = 0, nocc_k = 0;
  cplx *d_evq_all = nullptr;   // npwx x nbnd_all rows in the k+q slot's column order
  std::vector<double> et_q, wg_over_wk;
};

// kernel classes timed separately (CUDA events on the library's stream) when profiling is on
enum ProfClass { PC_FFT_Z = 0, PC_FFT_PLANE, PC_GEMM_PROJ, PC_GEMM_OUT, PC_SHIFT, PC_SEED, PC_RHO_PLANE, PC_OTHER, PC_GW_PROD, PC_SHIFT_GEMM, PC_N };
struct ProfRec { int cls; cudaEvent_t a, b; };

// grid%corr_fft of the correlation cutoff (sgw_set_corr_grid): the 6-D transforms of fft6.f90 act on ngm_c <= ~100 G vectors
// of a box of a few hundred points, so they are applied as sphere-pruned DFT matrices on the FP64 tensor path:
//   Ec(r, G) = exp(-i G r)  (nnr x ngm),   ET(G, r) = exp(+i G r)  (ngm x nnr)
struct CorrGrid {
  bool set = false;
  int n1 = 0, n2 = 0, n3 = 0, nnr = 0, ngm = 0;
  cplx *d_Ec = nullptr, *d_ET = nullptr;
};

struct Workspace {
  std::map<std::string, std::pair<void *, size_t>> bufs;
};

}  // namespace sgw

struct sgw_ctx {
  int device = 0;
  cudaStream_t stream = nullptr;
  cudaStream_t own_stream = nullptr;             // created by sgw_create; `stream` may point at a caller's stream (sgw_set_stream)
  std::string err;
  // grid
  bool grid_set = false, vloc_set = false, system_set = false;
  int nr1 = 0, nr2 = 0, nr3 = 0;
  int nr1x = 0, nr2x = 0, nr3x = 0;              // physical dimensions of the caller's real-space arrays (dffts%nr1x ...; >= nr)
  sgw::Plan1D px, py, pz;
  sgw::cplx *d_twx = nullptr, *d_twy = nullptr, *d_twz = nullptr;
  double *d_vperm = nullptr;    // local potential in permuted real-space order [pz][py][px]
  double *d_vperm_t = nullptr;  // the same with y fastest: [pz][px][py]
  std::vector<int> permx, permy, permz;  // position -> natural index
  std::vector<sgw::KSlot> slots;
  std::vector<sgw::KPair> pairs;
  // system
  double omega_cell = 0.0, tpiba2 = 0.0, xq[3] = {0, 0, 0};
  int ngm = 0;
  std::vector<double> g;        // 3 x ngm
  std::vector<int32_t> nl;      // ngm, 1-based
  std::map<int, sgw::Sphere> rho_spheres;  // density-sphere column structures keyed by ngc
  sgw::Workspace ws;
  sgw_stats stats;
  bool profiling = false;
  int prof_z_class = 0;                          // class the z passes are billed to (PC_FFT_Z = 0 for H.psi; the Delta-rho stage sets its own)
  int64_t launches = 0;
  std::vector<cudaEvent_t> ev_pool;
  std::vector<sgw::ProfRec> prof_recs;
  double prof_ms[sgw::PC_N] = {0};
  int64_t prof_n[sgw::PC_N] = {0};
  cudaEvent_t ev0 = nullptr, ev1 = nullptr, ev2 = nullptr, ev3 = nullptr;
  // alias-free coarse grid for the Delta-rho accumulation of sgw_coulomb (coulomb.cu: rho_grid_prepare)
  sgw::FftGrid rho_grid;
  bool rho_grid_on = false, rho_last_coarse = false;
  long tables_version = 0, rho_grid_version = -1;
  int rho_grid_ngc = -1;
  sgw::Sphere rho_sph_c;
  std::vector<sgw::Sphere> pair_k_c, pair_kq_c;
  // control_gw globals of the self-consistent branch (sgw_set_mixing): niter_gw, alpha_mix(:), tr2_gw, nmix_gw
  int mix_niter = 0, mix_nmix = 0, last_scf_iter = 0;
  bool solve_direct = true;                      // control_gw solve_direct (coulomb.f90:104)
  bool lgauss = false;                           // [QE] klist lgauss / degauss / ngauss, ener ef (sgw_set_smearing)
  double ef = 0.0, degauss = 0.0;
  int ngauss = 0;
  std::vector<double> mix_alpha;
  double mix_tr2 = 0.0;
  cudaEvent_t ev_iter[2] = {nullptr, nullptr};   // solver look-ahead (bicgstab.cu)
  int *h_flags = nullptr;                        // pinned host flags read back by the solvers
  int gemm_cta_per_sm = 0;                       // resident k_zgemm CTAs per SM (0 = attributes not set yet)
  int sm_count = 148;
  size_t smem_optin = 0;
  sgw_message_fn msg_fn = nullptr;               // sgw_set_message_callback: the reference's stdout warnings
  void *msg_user = nullptr;
  sgw::CorrGrid corr;                            // sigma.cu
  bool gw_attr_set = false;
  // k-point lanes of sgw_coulomb (coulomb.cu: klanes_prepare): shallow copies of this context that share its device tables
  // but own a stream, a workspace, events and statistics, so that independent k-points run concurrently on small systems
  std::vector<sgw_ctx *> lanes;
  bool is_lane = false;
};

namespace sgw {

#define SGW_CUDA(call)                                                                     \
  do {                                                                                     \
    cudaError_t e_ = (call);                                                               \
    if (e_ != cudaSuccess) {                                                               \
      char buf_[512];                                                                      \
      snprintf(buf_, sizeof buf_, "%s:%d: %s: %s", __FILE__, __LINE__, #call, cudaGetErrorString(e_)); \
      ctx->err = buf_;                                                                     \
      return SGW_E_CUDA;                                                                   \
    }                                                                                      \
  } while (0)

#define SGW_CHECK(expr)                    \
  do {                                     \
    int r_ = (expr);                       \
    if (r_ < 0) return r_;                 \
  } while (0)

#define SGW_ARG(cond, msg)                 \
  do {                                     \
    if (!(cond)) {                         \
      ctx->err = std::string("invalid argument: ") + msg; \
      return SGW_E_ARG;                    \
    }                                      \
  } while (0)

#define SGW_LAUNCH_CHECK()                 \
  do {                                     \
    ctx->launches++;                       \
    SGW_CUDA(cudaGetLastError());          \
  } while (0)

// Table buffers (operator data, spheres, twiddles) come from a small size-keyed cache: a Fortran host re-installs the same
// tables for every q-point, and cudaFree / cudaMalloc on a process that holds tens of GB of solver workspace were measured
// at 7..350 ms per re-installation (they synchronise the device and edit its page tables).  dev_free parks the block,
// dev_malloc hands out a parked block of the same size; sgw_destroy of the last context returns everything to the driver.
cudaError_t dev_malloc(void **p, size_t bytes);
void dev_free(void *p);
void dev_pool_trim();

// Large host <-> device copies of CALLER-OWNED (pageable) arrays.  cudaMemcpyAsync from pageable memory is staged by the driver
// on one thread (measured 8.6 GB/s for the 206 MB of Si64 tables, 9-10 GB/s for the 1.85 GB dielectric matrices); here the
// array is cut into chunks that NL host threads copy through their own page-locked buffers and streams, so that the memcpy
// into / out of the staging area runs on several cores and overlaps the DMA.  Synchronous with respect to ctx->stream: on return
// of h2d_large the data is ordered before later work on ctx->stream; d2h_large returns when the host array is complete.
// Small arrays take the plain path.  SGW_COPY_THREADS=0 disables the staging (A/B).
int h2d_large(sgw_ctx *ctx, void *d_dst, const void *h_src, size_t bytes);
int d2h_large(sgw_ctx *ctx, void *h_dst, const void *d_src, size_t bytes);

// workspace: named growable device buffers owned by the context
int ws_get(sgw_ctx *ctx, const char *name, size_t bytes, void **out);
template <typename T>
inline int ws(sgw_ctx *ctx, const char *name, size_t count, T **out) {
  void *p = nullptr;
  int r = ws_get(ctx, name, count * sizeof(T), &p);
  *out = (T *)p;
  return r;
}
void ws_free_all(sgw_ctx *ctx);

template <typename T>
inline int upload(sgw_ctx *ctx, T **dptr, const T *h, size_t count) {
  if (*dptr) { dev_free(*dptr); *dptr = nullptr; }
  if (count == 0) return SGW_OK;
  SGW_CUDA(dev_malloc((void **)dptr, count * sizeof(T)));
  SGW_CUDA(cudaMemcpyAsync(*dptr, h, count * sizeof(T), cudaMemcpyHostToDevice, ctx->stream));
  SGW_CUDA(cudaStreamSynchronize(ctx->stream));
  return SGW_OK;
}

void prof_begin(sgw_ctx *ctx, int cls);
void prof_end(sgw_ctx *ctx);
struct ProfScope {
  sgw_ctx *c;
  ProfScope(sgw_ctx *ctx, int cls) : c(ctx) { if (c->profiling) prof_begin(c, cls); }
  ~ProfScope() { if (c->profiling) prof_end(c); }
};
void begin_call(sgw_ctx *ctx);
void end_call(sgw_ctx *ctx);
// Padded boxes (nr1x > nr1): the caller's 1-based linear FFT indices and real-space arrays use the physical dimensions; the
// library works on the compact box, so indices are converted once when they come in
inline bool grid_padded(const sgw_ctx *ctx) { return ctx->nr1x != ctx->nr1 || ctx->nr2x != ctx->nr2 || ctx->nr3x != ctx->nr3; }
inline int32_t unpad_index(const sgw_ctx *ctx, int32_t idx1) {
  if (!grid_padded(ctx)) return idx1;
  const long i = (long)idx1 - 1;
  const long x = i % ctx->nr1x, y = (i / ctx->nr1x) % ctx->nr2x, z = i / ((long)ctx->nr1x * ctx->nr2x);
  return (int32_t)(x + (long)ctx->nr1 * (y + (long)ctx->nr2 * z) + 1);
}
void lanes_destroy(sgw_ctx *ctx);              // api.cu: frees what the lanes own (streams, workspaces, events)
GridDev grid_dev(const sgw_ctx *ctx, const FftGrid *gr = nullptr);
int build_sphere(sgw_ctx *ctx, int npw, const int32_t *nl_1based, Sphere *sph);
// same entries, same order and same column partition as `fine` (built on the context's grid), re-expressed in the box `gr`
int remap_sphere(sgw_ctx *ctx, const Sphere &fine, const FftGrid &gr, Sphere *out);
int make_fft_grid(sgw_ctx *ctx, int n1, int n2, int n3, FftGrid *gr);
void free_fft_grid(FftGrid *gr);
void free_sphere(Sphere *s);

// ---- fft.cu : batched local-potential pipeline ----
enum PlaneMode { PLANE_VLOC = 0, PLANE_FIELD = 1, PLANE_TO_R = 2, PLANE_FROM_R = 3 };
// G (column order, nvec vectors with leading dim ld) -> T1[vec][pz][col]
int fft_zpass_g2r(sgw_ctx *ctx, const Sphere &s, int nvec, const cplx *in, long ld, cplx *T, const int *active,
                  const FftGrid *gr = nullptr);
// plane stage: T_in (sphere sin) -> 2-D inverse -> (x v | x field | store R) ; (load R) -> 2-D forward -> T_out (sphere sout)
int fft_plane(sgw_ctx *ctx, PlaneMode mode, const Sphere *sin, const Sphere *sout, int nvec, const cplx *Tin, cplx *Tout,
              const cplx *field, int vec_per_field, cplx *R, const int *active, int in_mod = 0, const FftGrid *gr = nullptr);
// incdrhoscf ([QE], solve_linter.f90:489-497) for npf (perturbation, frequency) pairs: per z-plane, loop over the nocc
// bands: 2-D inverse of dpsi, acc += conj(psi_v(r)) dpsi(r); then wgt*acc -> 2-D forward -> columns of the density
// sphere `sout` (Tout[pf][pz][col], += if accumulate) or, when Rout != null, the real-space planes Rout[pf][pz][nxy]
int fft_plane_rho(sgw_ctx *ctx, const Sphere &sin, const Sphere &sout, int npf, int nocc, const cplx *Tin, const cplx *psir,
                  double wgt, cplx *Tout, int accumulate, const FftGrid *gr = nullptr, const cplx *psir_t = nullptr);
// out[p][x][y] = in[p][y][x]: the y-fastest copy of psi_v(r) that k_plane_rho_v2 reads (psir_t above)
int fft_transpose_planes(sgw_ctx *ctx, const FftGrid *gr, long nplanes, const cplx *in, cplx *out);
// epilogue modes of the final z pass
struct ZEpilogue {
  int mode;              // 0: out = val ; 1: out = out*keep + g2kin*psi + sigma*psi + val (H.psi) ; 2: out += val
  const double *g2kin;   // npwx
  const cplx *psi;       // nvec x ld
  const cplx *sigma;     // per vector shift (device) or null
  long sigma_stride;
  int keep_out;          // 1: add to existing out (non-local part already there)
};
int fft_zpass_r2g(sgw_ctx *ctx, const Sphere &s, int nvec, const cplx *T, cplx *out, long ld, const ZEpilogue &epi,
                  const int *active, const FftGrid *gr = nullptr);

// ---- gemm.cu : non-local projectors and valence projector (complex FP64, DMMA) ----
// out[:, v] = P * (W .* (P^H psi[:, v]))  with W = blockdiag(dion, alpha_pv * I_nbnd); out overwritten (npwx rows)
int nonlocal_apply(sgw_ctx *ctx, const KSlot &k, double alpha_pv, int nvec, const cplx *psi, long ldpsi, cplx *out,
                   long ldout, const int *active);
// generic helpers used by the Coulomb pipeline
int gemm_ch_n(sgw_ctx *ctx, int m, int n, int k, const cplx *A, long lda, const cplx *B, long ldb, cplx *C, long ldc);
int gemm_n_n(sgw_ctx *ctx, int m, int n, int k, cplx alpha, const cplx *A, long lda, const cplx *B, long ldb, cplx beta,
             cplx *C, long ldc);
int gemm_n_n_batched(sgw_ctx *ctx, int m, int n, int k, cplx alpha, const cplx *A, long lda, long bsA, const cplx *B, long ldb,
                     long bsB, cplx beta, cplx *C, long ldc, long bsC, int nbatch, const int *list, const int *count);
int gemm_ch_n_batched(sgw_ctx *ctx, int m, int n, int k, const cplx *A, long lda, long bsA, const cplx *B, long ldb, long bsB, cplx *C,
                      long ldc, long bsC, int nbatch);
int dense_apply(sgw_ctx *ctx, const KSlot &k, int nvec, const cplx *psi, long ldpsi, const cplx *sigma, long sigma_stride,
                cplx *out, long ldout, const int *active);

// ---- operator.cu ----
// apsi[:, v] = (H + sigma_v + alpha_pv P_v) psi[:, v] for vectors in column order (device resident)
int apply_operator(sgw_ctx *ctx, int slot, double alpha_pv, int nvec, const cplx *psi, long ldpsi, const cplx *sigma,
                   long sigma_stride, cplx *apsi, long ldapsi, const int *active);
int permute_in(sgw_ctx *ctx, const Sphere &s, int nvec, const cplx *src, long lds, cplx *dst, long ldd, int npwx);
int permute_out(sgw_ctx *ctx, const Sphere &s, int nvec, const cplx *src, long lds, cplx *dst, long ldd);

// ---- bicgstab.cu / subspace.cu / select.cu ----
// Optional +-omega average of solve_linter.f90:464-480 done inside the solver: for right-hand side b and frequency f
//   y(b, f) = x_f                                   (f = 0 when the list starts with the static frequency)
//           = 1/2 x_f + 1/2 x_{nfreq + f - first}   otherwise (first = zero_freq)
// written to d_y + (((b / group) * nfreq + f) * group + b % group) * n.  The multishift solver in lazy mode never forms the
// individual x^sigma then: it combines the coefficient columns and materialises nfreq vectors instead of nshift - 1.
struct AvgSpec {
  cplx *d_y = nullptr;
  int nfreq = 0, zero_freq = 0, group = 1;
  int *d_done = nullptr;     // [nrhs] 1: y written by the solver; 0: the caller-side average from x still has to run
};

struct SolveBatch {
  AvgSpec avg;
  int slot;
  double alpha_pv;
  int nrhs, nshift, n;       // n = vector length used (npwx of the slot or dense n)
  const cplx *d_b; long ldb; // device, column order
  const cplx *d_sigma;       // device nshift x nrhs
  cplx *d_x;                 // device: x[(irhs*nshift + ishift)*n + ig]
  int *d_ierr;               // device nrhs
};
int bicgstab_batched(sgw_ctx *ctx, const SolveBatch &sb, int lmax, double threshold, int max_iter, const int *d_todo);
size_t bicgstab_bytes_per_rhs(int n, int lmax, int nshift);   // solver state per right-hand side (without x and b)
int subspace_batched(sgw_ctx *ctx, const SolveBatch &sb, double threshold, int max_iter, const int *d_todo);
int select_solver_batched(sgw_ctx *ctx, const SolveBatch &sb, const sgw_solver_cfg *cfg);

// ---- coulomb.cu ----
int green_function_core(sgw_ctx *ctx, int slot, const sgw_solver_cfg *cfg, int ngc, const int32_t *map, int ngp,
                        const int32_t *fft_map, int nfreq, const sgw_cplx *omega, sgw_cplx *green, int32_t *ierr_out,
                        cplx **d_green_out);

}  // namespace sgw
