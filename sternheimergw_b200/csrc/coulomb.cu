// coulomb.cu -- the screened-Coulomb and Green's-function pipelines around the batched solver:
//   solve_linter   phys/coul/src/solve_linter.f90:55-624 (direct branch)   -> sgw_solve_linter
//   coulomb        phys/coul/src/coulomb.f90:29-176                        -> sgw_coulomb
//   coulomb_q0G0   phys/coul/src/coulomb_q0G0.f90:31-158                   -> sgw_coulomb_q0G0
//   unfold_w       algo/symmetry/src/unfold_w.f90:84 (identity symmetry)   -> sgw_unfold_w
//   invert_epsilon phys/coul/src/invert_epsilon.f90:23-90                  -> sgw_invert_epsilon (invert.cu)
//   green_function phys/green/src/green.f90:105-226                        -> sgw_green_function
// One call handles a whole block of perturbations: dV_bare psi (dvqpsi_us.f90:99-130), -P_c^+ ([QE] orthogonalize),
// the multishift solves batched over perturbations x bands, the +-omega average (solve_linter.f90:464-480),
// the Delta-rho accumulation ([QE] incdrhoscf) and the Hartree kernel ([QE] dv_of_drho, lrpa) all stay on the
// device; only wavefunction tables go in and scrcoul comes out.
#include "internal.cuh"

#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <cmath>
#include <thread>

using namespace sgw;

namespace sgw {

__device__ __forceinline__ int perm_dev(int r1, int r2, int pos) { return pos / r2 + r1 * (pos % r2); }

// dvbare(r) of a delta perturbation at G (coulomb.f90:129-134: invfft of a unit coefficient) in the library's
// permuted real-space order: exp(+i G.r) as a product of three twiddle-table entries.
__global__ void k_delta_field(GridDev g, int np, const int *__restrict__ mill /* 3 x np, wrapped to 0..n-1 */,
                              cplx *__restrict__ field) {
  const long nnr = (long)g.nx * g.ny * g.nz;
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  const int p = blockIdx.y;
  if (i >= nnr || p >= np) return;
  const int px = (int)(i % g.nx), py = (int)((i / g.nx) % g.ny), pz = (int)(i / ((long)g.nx * g.ny));
  const int x = perm_dev(g.rx1, g.rx2, px), y = perm_dev(g.ry1, g.ry2, py), z = perm_dev(g.rz1, g.rz2, pz);
  const cplx a = cconj(g.twx[(int)(((long)mill[3 * p] * x) % g.nx)]);
  const cplx b = cconj(g.twy[(int)(((long)mill[3 * p + 1] * y) % g.ny)]);
  const cplx c = cconj(g.twz[(int)(((long)mill[3 * p + 2] * z) % g.nz)]);
  field[(long)p * nnr + i] = cmul(cmul(a, b), c);
}

// natural (column-major nr1,nr2,nr3) <-> permuted [pz][py][px] real-space order
__global__ void k_nat2perm(GridDev g, int nvec, const cplx *__restrict__ nat, cplx *__restrict__ prm) {
  const long nnr = (long)g.nx * g.ny * g.nz;
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  const int v = blockIdx.y;
  if (i >= nnr) return;
  const int px = (int)(i % g.nx), py = (int)((i / g.nx) % g.ny), pz = (int)(i / ((long)g.nx * g.ny));
  const long j = perm_dev(g.rx1, g.rx2, px) + (long)g.nx * (perm_dev(g.ry1, g.ry2, py) + (long)g.ny * perm_dev(g.rz1, g.rz2, pz));
  prm[(long)v * nnr + i] = nat[(long)v * nnr + j];
}
__global__ void k_perm2nat(GridDev g, int nvec, const cplx *__restrict__ prm, cplx *__restrict__ nat, double scale) {
  const long nnr = (long)g.nx * g.ny * g.nz;
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  const int v = blockIdx.y;
  if (i >= nnr) return;
  const int px = (int)(i % g.nx), py = (int)((i / g.nx) % g.ny), pz = (int)(i / ((long)g.nx * g.ny));
  const long j = perm_dev(g.rx1, g.rx2, px) + (long)g.nx * (perm_dev(g.ry1, g.ry2, py) + (long)g.ny * perm_dev(g.rz1, g.rz2, pz));
  nat[(long)v * nnr + j] = cscale(scale, prm[(long)v * nnr + i]);
}

// +-omega average (solve_linter.f90:464-480) fused with the reordering the Delta-rho stage wants:
// dst[((pf - pf0) * nocc + ib) * n + e],  pf = p * nfreq + ifreq,  from x[((p * nocc + ib) * nshift + is) * n + e]
__global__ void k_average(int n, int nocc, int nfreq, int nshift, int zero_freq, int pf0, const cplx *__restrict__ x,
                          cplx *__restrict__ dst) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n) return;
  const int ib = blockIdx.y, pfl = blockIdx.z;
  const int pf = pf0 + pfl, p = pf / nfreq, ifreq = pf % nfreq;
  const cplx *xr = x + ((long)(p * nocc + ib) * nshift) * n;
  cplx v = xr[(long)ifreq * n + e];
  const int first = zero_freq ? 1 : 0;
  if (ifreq >= first) {
    const cplx w = xr[(long)(nfreq + ifreq - first) * n + e];
    v = cscale(0.5, v);                          // ZSCAL 0.5 then ZAXPY 0.5 (:469-478)
    v = cmake(v.x + 0.5 * w.x, v.y + 0.5 * w.y);
  }
  dst[((long)pfl * nocc + ib) * n + e] = v;
}

// [QE] dv_of_drho with lrpa: dv(G) = e2 fpi drho(G) / (tpiba2 |q+G|^2); out = -dv (solve_linter.f90:598)
__global__ void k_hartree(int npw, int nvec, const double *__restrict__ fac /* column order */, cplx *__restrict__ drho,
                          int zero_pos, double sign = -1.0) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  const int v = blockIdx.y;
  if (e >= npw) return;
  cplx d = drho[(long)v * npw + e];
  if (e == zero_pos) d = cmake(0.0, 0.0);       // zero-mean fix (:544-550)
  drho[(long)v * npw + e] = cscale(sign * fac[e], d);
}

// coulomb.f90:143-157: scrcoul(igp, iw, indx) = -dV_H(G_igp) + delta(igp, ig)
__global__ void k_scr_extract(int ngc, int nfs, int np, const int *__restrict__ perm, const double *__restrict__ fac,
                              const int *__restrict__ ig0 /* 0-based perturbation G per task */,
                              const cplx *__restrict__ drho /* [p][iw][pos] */, cplx *__restrict__ scr /* ngc x nfs x np */,
                              int direct = 1) {
  const int pos = blockIdx.x * blockDim.x + threadIdx.x;
  const int iw = blockIdx.y, p = blockIdx.z;
  if (pos >= ngc) return;
  const int igp = perm[pos];
  // direct: eps = delta - v drho (:149-157); self-consistent: the input already is dV_scf(G) and is copied (:149-151)
  cplx v = direct ? cscale(-fac[pos], drho[((long)p * nfs + iw) * ngc + pos]) : drho[((long)p * nfs + iw) * ngc + pos];
  if (direct && igp == ig0[p]) v.x += 1.0;
  scr[(long)igp + (long)ngc * (iw + (long)nfs * p)] = v;
}

// unfold_w.f90:84: out(ig_unique(ig), igp, iw) = CONJG(in(igp, iw, ig))
__global__ void k_unfold(int ngc, int nfs, int nuniq, const int *__restrict__ ig_unique, const cplx *__restrict__ in,
                         cplx *__restrict__ out) {
  const int igp = blockIdx.x * blockDim.x + threadIdx.x;
  const int iw = blockIdx.y, ig = blockIdx.z;
  if (igp >= ngc) return;
  const int row = ig_unique[ig] - 1;
  if (row < 0 || row >= ngc) return;
  out[(long)row + (long)ngc * (igp + (long)ngc * iw)] = cconj(in[(long)igp + (long)ngc * (iw + (long)nfs * ig)]);
}

// unfold_w.f90:104-129: row ig that is not symmetry-unique is the row of its sym_friend, rotated by R = sym_ig(ig):
//   out(ig, gmapsym(igp, invs(R)), iw) = out(sym_friend(ig), igp, iw) * eigv(sym_friend(ig), R) * CONJG(eigv(igp, R))
// sym_friend(ig) is always a unique G (stern_symm.f90:94), whose row k_unfold has already filled, so the rows are independent.
__global__ void k_unfold_symm(int ngc, int nfs, const int *__restrict__ is_unique, const int *__restrict__ sym_ig,
                              const int *__restrict__ sym_friend, const int *__restrict__ gmapsym, const cplx *__restrict__ eigv,
                              const int *__restrict__ invs, cplx *__restrict__ out) {
  const int igp = blockIdx.x * blockDim.x + threadIdx.x;
  const int iw = blockIdx.y, ig = blockIdx.z;
  if (igp >= ngc || is_unique[ig]) return;
  const int fr = sym_friend[ig] - 1, isym = sym_ig[ig] - 1;
  const int ism1 = invs[isym] - 1;
  const int col = gmapsym[igp + (long)ngc * ism1] - 1;
  const cplx phase = cmul(eigv[fr + (long)ngc * isym], cconj(eigv[igp + (long)ngc * isym]));
  out[(long)ig + (long)ngc * (col + (long)ngc * iw)] = cmul(out[(long)fr + (long)ngc * (igp + (long)ngc * iw)], phase);
}

// ---------------------------------------------------------------- green_function helpers
__global__ void k_green_rhs(int n, int nrhs, const int *__restrict__ pos /* column-order position of e_G', -1 = skip */,
                            cplx *__restrict__ b) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  const int r = blockIdx.y;
  if (e >= n) return;
  b[(long)r * n + e] = (e == pos[r]) ? cmake(-1.0, 0.0) : cmake(0.0, 0.0);    // green.f90:203-204
}
// green(k, igp, ifreq) = green_part(map(k), ifreq) WHERE map > 0 .AND. map < num_g (green.f90:211-213)
__global__ void k_green_scatter(int ngc, int ngp, int nfreq, int num_g, int n, const int *__restrict__ map,
                                const int *__restrict__ invperm, const int *__restrict__ col /* igp of each RHS */,
                                const cplx *__restrict__ x, cplx *__restrict__ green) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  const int ifreq = blockIdx.y, r = blockIdx.z;
  if (k >= ngc) return;
  const int mk = map[k];
  if (mk > 0 && mk < num_g)
    green[(long)k + (long)ngc * (col[r] + (long)ngp * ifreq)] = x[((long)r * nfreq + ifreq) * n + invperm[mk - 1]];
}

static int get_rho_sphere(sgw_ctx *ctx, int ngc, Sphere **out) {
  auto it = ctx->rho_spheres.find(ngc);
  if (it == ctx->rho_spheres.end()) {
    Sphere s;
    SGW_CHECK(build_sphere(ctx, ngc, ctx->nl.data(), &s));
    it = ctx->rho_spheres.emplace(ngc, s).first;
  }
  *out = &it->second;
  return SGW_OK;
}

// e2 fpi / (tpiba2 |q+G|^2) for the first ngc G vectors, in the column order of `sph` ([QE] dv_of_drho, lrpa)
static int hartree_factor(sgw_ctx *ctx, const Sphere &sph, int ngc, const char *wsname, double **d_fac) {
  std::vector<double> fac(ngc);
  const double e2 = 2.0, fpi = 4.0 * M_PI;
  for (int pos = 0; pos < ngc; ++pos) {
    const int ig = sph.perm[pos];
    const double q0 = ctx->g[3 * ig] + ctx->xq[0], q1 = ctx->g[3 * ig + 1] + ctx->xq[1], q2 = ctx->g[3 * ig + 2] + ctx->xq[2];
    const double qg2 = q0 * q0 + q1 * q1 + q2 * q2;
    fac[pos] = qg2 > 1e-8 ? e2 * fpi / (ctx->tpiba2 * qg2) : 0.0;
  }
  SGW_CHECK(ws(ctx, wsname, (size_t)ngc, d_fac));
  SGW_CUDA(cudaMemcpyAsync(*d_fac, fac.data(), sizeof(double) * ngc, cudaMemcpyHostToDevice, ctx->stream));
  SGW_CUDA(cudaStreamSynchronize(ctx->stream));
  return SGW_OK;
}

struct FreqList {
  int nfreq, num_omega, zero_freq;
  std::vector<cplx> omega;
};
static FreqList make_omega(int nfreq, const sgw_cplx *freq) {   // solve_linter.f90:217-252
  FreqList f;
  f.nfreq = nfreq;
  f.zero_freq = std::hypot(freq[0].re, freq[0].im) < 1e-14;
  f.num_omega = f.zero_freq ? 2 * nfreq - 1 : 2 * nfreq;
  f.omega.resize(f.num_omega);
  for (int i = 0; i < nfreq; ++i) f.omega[i] = cmake(freq[i].re, freq[i].im);
  if (f.zero_freq) for (int i = 1; i < nfreq; ++i) f.omega[nfreq + i - 1] = cmake(-freq[i].re, -freq[i].im);
  else for (int i = 0; i < nfreq; ++i) f.omega[nfreq + i] = cmake(-freq[i].re, -freq[i].im);
  return f;
}

static size_t solver_bytes_per_rhs(const sgw_ctx *ctx, const KSlot &ks, int lmax, int nshift) {
  const size_t n = ks.npwx;
  size_t v = bicgstab_bytes_per_rhs((int)n, lmax, nshift) / sizeof(cplx) + (size_t)nshift * n + n;   // bicgstab state + sv_x + rhs
  v += (size_t)((nshift + 1) / 2) * n;                                                    // +-omega averages (co_davg_all)
  v += 2 * (size_t)ctx->nr3 * ks.sph.ncol;                                                // H.psi column buffers
  v += 4 * (size_t)(ks.nkb + ks.nbnd);                                                    // projector coefficients
  return v * sizeof(cplx);
}

// ---- alias-free coarse grid for the Delta-rho accumulation of sgw_coulomb ------------------------------------
// coulomb.f90:143-157 keeps only the first ngc G vectors of Delta-rho.  Delta-rho(G) = sum_G' conj(psi(G'-G)) dpsi(G') is a
// finite convolution: with M_k, M_kq the largest |Miller index| of the k and k+q spheres along an axis and M_out that of
// the ngc retained G vectors, an n-point axis gives those components EXACTLY (no aliasing) as soon as
// n >= M_k + M_kq + M_out + 1.  [QE] incdrhoscf uses the full dffts box (72 at Si64) where 45 suffices, i.e. 4.1x more
// points than the retained components need; the result differs from the full-box one by rounding only.
// SGW_RHO_GRID=fine switches this off (A/B testing); sgw_solve_linter, which returns drho(r) on the full box, never uses it.
static void rho_grid_release(sgw_ctx *ctx) {
  free_sphere(&ctx->rho_sph_c);
  for (auto &s : ctx->pair_k_c) free_sphere(&s);
  for (auto &s : ctx->pair_kq_c) free_sphere(&s);
  ctx->pair_k_c.clear();
  ctx->pair_kq_c.clear();
  free_fft_grid(&ctx->rho_grid);
  ctx->rho_grid_on = false;
}

static void miller_extent(const sgw_ctx *ctx, const Sphere &s, int ext[3]) {
  const int nf[3] = {ctx->nr1, ctx->nr2, ctx->nr3};
  auto mil = [&](int c, int d) { return std::abs(c <= (nf[d] - 1) / 2 ? c : c - nf[d]); };
  ext[0] = ext[1] = ext[2] = 0;
  for (int c = 0; c < s.ncol; ++c) {
    ext[0] = std::max(ext[0], mil(s.h_col_x[c], 0));
    ext[1] = std::max(ext[1], mil(s.h_col_y[c], 1));
  }
  for (int p = 0; p < s.npw; ++p) ext[2] = std::max(ext[2], mil(s.h_zof[p], 2));
}

static int rho_grid_prepare(sgw_ctx *ctx, int ngc, const Sphere &rho_fine, bool *use) {
  *use = false;
  const char *e = getenv("SGW_RHO_GRID");
  if (e && strcmp(e, "fine") == 0) return SGW_OK;
  if (ctx->rho_grid_version == ctx->tables_version && ctx->rho_grid_ngc == ngc) {
    *use = ctx->rho_grid_on;
    return SGW_OK;
  }
  rho_grid_release(ctx);
  ctx->rho_grid_version = ctx->tables_version;
  ctx->rho_grid_ngc = ngc;
  int mo[3], need[3] = {0, 0, 0};
  miller_extent(ctx, rho_fine, mo);
  for (auto &kp : ctx->pairs) {
    if (!kp.set || kp.slot < 0 || kp.slot >= (int)ctx->slots.size() || !ctx->slots[kp.slot].set) return SGW_OK;
    int mk[3], mq[3];
    miller_extent(ctx, kp.sph_k, mk);
    miller_extent(ctx, ctx->slots[kp.slot].sph, mq);
    // alias-free product, and every sphere must itself fit the box without wrap-around collisions (the scatter
    // into the box overwrites, it does not sum)
    for (int d = 0; d < 3; ++d)
      need[d] = std::max(need[d], std::max(mk[d] + mq[d] + mo[d] + 1, 2 * std::max(mk[d], std::max(mq[d], mo[d])) + 1));
  }
  const int nf[3] = {ctx->nr1, ctx->nr2, ctx->nr3};
  int nc[3];
  for (int d = 0; d < 3; ++d) {
    nc[d] = nf[d];
    for (int n = need[d]; n < nf[d]; ++n) {
      Plan1D p;
      if (make_plan(n, &p) && plan_is_fast(p)) { nc[d] = n; break; }      // 45 = 5 x 9 rather than 44 = 4 x 11
    }
  }
  if ((double)nc[0] * nc[1] * nc[2] > 0.7 * (double)nf[0] * nf[1] * nf[2]) return SGW_OK;   // not worth a second grid
  SGW_CHECK(make_fft_grid(ctx, nc[0], nc[1], nc[2], &ctx->rho_grid));
  SGW_CHECK(remap_sphere(ctx, rho_fine, ctx->rho_grid, &ctx->rho_sph_c));
  ctx->pair_k_c.resize(ctx->pairs.size());
  ctx->pair_kq_c.resize(ctx->pairs.size());
  for (size_t ik = 0; ik < ctx->pairs.size(); ++ik) {
    SGW_CHECK(remap_sphere(ctx, ctx->pairs[ik].sph_k, ctx->rho_grid, &ctx->pair_k_c[ik]));
    SGW_CHECK(remap_sphere(ctx, ctx->slots[ctx->pairs[ik].slot].sph, ctx->rho_grid, &ctx->pair_kq_c[ik]));
  }
  ctx->rho_grid_on = true;
  *use = true;
  return SGW_OK;
}

// ---- metals: the smeared projector of [QE] LR_Modules/orthogonalize.f90 (lgauss branch) -------------------------------
// occupation function theta~(x) and its derivative ([QE] Modules/wgauss.f90, w0gauss.f90): -99 Fermi-Dirac, -1 cold smearing,
// 0 Gaussian, n > 0 Methfessel-Paxton
static double wgauss(double x, int n) {
  if (n == -99) return x < -200.0 ? 0.0 : (x > 200.0 ? 1.0 : 1.0 / (1.0 + exp(-x)));
  if (n == -1) {
    const double xp = x - 1.0 / sqrt(2.0), arg = std::min(200.0, xp * xp);
    return 0.5 * erf(xp) + 1.0 / sqrt(2.0 * M_PI) * exp(-arg) + 0.5;
  }
  double w = 0.5 * erfc(-x);
  double hd = 0.0, hp = exp(-std::min(200.0, x * x)), a = 1.0 / sqrt(M_PI);
  int ni = 0;
  for (int i = 1; i <= n; ++i) {
    hd = 2.0 * x * hp - 2.0 * ni * hd; ++ni;
    a = -a / (i * 4.0);
    w -= a * hd;
    hp = 2.0 * x * hd - 2.0 * ni * hp; ++ni;
  }
  return w;
}
static double w0gauss(double x, int n) {
  const double sqrtpm1 = 1.0 / sqrt(M_PI);
  if (n == -99) return fabs(x) <= 36.0 ? 1.0 / (2.0 + exp(-x) + exp(x)) : 0.0;
  if (n == -1) {
    const double xp = x - 1.0 / sqrt(2.0), arg = std::min(200.0, xp * xp);
    return sqrtpm1 * exp(-arg) * (2.0 - sqrt(2.0) * x);
  }
  const double arg = std::min(200.0, x * x);
  double w = exp(-arg) * sqrtpm1, hd = 0.0, hp = exp(-arg), a = sqrtpm1;
  int ni = 0;
  for (int i = 1; i <= n; ++i) {
    hd = 2.0 * x * hp - 2.0 * ni * hd; ++ni;
    a = -a / (i * 4.0);
    hp = 2.0 * x * hd - 2.0 * ni * hp; ++ni;
    w += a * hp;
  }
  return w;
}
// ps(j, r) *= wwg(j, band of r) ; dvpsi(:, r) *= theta~_F(band of r)   (band of r = r % nocc)
__global__ void k_metal_scale(int nb, int nocc, int nrhs, int n, const double *__restrict__ wwg, const double *__restrict__ wg1,
                              cplx *__restrict__ ps, cplx *__restrict__ dvpsi) {
  const int r = blockIdx.y, ib = r % nocc;
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e < nb) ps[(long)r * nb + e] = cscale(wwg[e + (long)nb * ib], ps[(long)r * nb + e]);
  if (e < n) dvpsi[(long)r * n + e] = cscale(wg1[ib], dvpsi[(long)r * n + e]);
}
__global__ void k_band_scale(int n, int nocc, const double *__restrict__ f, cplx *__restrict__ v) {   // v(:, r) *= f(r % nocc)
  const int r = blockIdx.y;
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e < n) v[(long)r * n + e] = cscale(f[r % nocc], v[(long)r * n + e]);
}

// [QE] orthogonalize (solve_linter.f90:337,409) for `nrhs` vectors whose band is (column % nocc): insulators
// dvpsi <- evq (evq^H dvpsi) - dvpsi = -P_c^+ dvpsi; metals (lgauss) the smeared projector.  *d_wgk (metals): wg/wk of the bands
static int orthogonalize_dev(sgw_ctx *ctx, const KPair &kp, const KSlot &ks, int nocc, int nrhs, cplx *dvpsi, double **d_wgk_out) {
  cudaStream_t st = ctx->stream;
  const bool metal = ctx->lgauss;
  const int n = ks.npwx;
  cplx *ps = nullptr;
  double *d_wgk = nullptr;
  if (!metal) {
    const cplx *evq = ks.d_P + (size_t)ks.nkb * n;
    SGW_CHECK(ws(ctx, "co_ps", (size_t)nocc * nrhs, &ps));
    SGW_CHECK(gemm_ch_n(ctx, nocc, nrhs, ks.npw, evq, n, dvpsi, n, ps, nocc));
    SGW_CHECK(gemm_n_n(ctx, n, nrhs, nocc, cmake(1.0, 0.0), evq, n, ps, nocc, cmake(-1.0, 0.0), dvpsi, n));
  } else {
    // orthogonalize.f90, lgauss: ps(j, i) = wwg(j, i) <evq_j|dvpsi_i> over all nbnd bands at k+q, dvpsi_i *= theta~_F,i
    const int nb = kp.nbnd_all;
    std::vector<double> wwg((size_t)nb * nocc), wg1(nocc);
    for (int ib = 0; ib < nocc; ++ib) {
      wg1[ib] = wgauss((ctx->ef - kp.et[ib]) / ctx->degauss, ctx->ngauss);
      const double w0g = w0gauss((ctx->ef - kp.et[ib]) / ctx->degauss, ctx->ngauss) / ctx->degauss;
      for (int jb = 0; jb < nb; ++jb) {
        const double wgp = wgauss((ctx->ef - kp.et_q[jb]) / ctx->degauss, ctx->ngauss);
        const double deltae = kp.et_q[jb] - kp.et[ib];
        const double theta = wgauss(deltae / ctx->degauss, 0);
        double w = wg1[ib] * (1.0 - theta) + wgp * theta;
        if (jb < ks.nbnd) w += fabs(deltae) > 1.0e-5 ? ks.alpha_pv * theta * (wgp - wg1[ib]) / deltae : -ks.alpha_pv * theta * w0g;
        wwg[jb + (size_t)nb * ib] = w;
      }
    }
    double *d_wwg = nullptr, *d_wg1 = nullptr;
    SGW_CHECK(ws(ctx, "co_wwg", (size_t)nb * nocc, &d_wwg));
    SGW_CHECK(ws(ctx, "co_wg1", (size_t)nocc, &d_wg1));
    SGW_CHECK(ws(ctx, "co_wgk", (size_t)nocc, &d_wgk));
    SGW_CUDA(cudaMemcpyAsync(d_wwg, wwg.data(), sizeof(double) * wwg.size(), cudaMemcpyHostToDevice, st));
    SGW_CUDA(cudaMemcpyAsync(d_wg1, wg1.data(), sizeof(double) * nocc, cudaMemcpyHostToDevice, st));
    SGW_CUDA(cudaMemcpyAsync(d_wgk, kp.wg_over_wk.data(), sizeof(double) * nocc, cudaMemcpyHostToDevice, st));
    SGW_CUDA(cudaStreamSynchronize(st));                                                   // host vectors go out of scope
    SGW_CHECK(ws(ctx, "co_ps", (size_t)nb * nrhs, &ps));
    SGW_CHECK(gemm_ch_n(ctx, nb, nrhs, ks.npw, kp.d_evq_all, n, dvpsi, n, ps, nb));
    for (int r0 = 0; r0 < nrhs; r0 += 65535 / nocc * nocc) {
      const int c = std::min(65535 / nocc * nocc, nrhs - r0);
      dim3 gm((unsigned)((std::max(n, nb) + 255) / 256), (unsigned)c);
      k_metal_scale<<<gm, 256, 0, st>>>(nb, nocc, c, n, d_wwg, d_wg1, ps + (size_t)r0 * nb, dvpsi + (size_t)r0 * n);
      SGW_LAUNCH_CHECK();
    }
    SGW_CHECK(gemm_n_n(ctx, n, nrhs, nb, cmake(1.0, 0.0), kp.d_evq_all, n, ps, nb, cmake(-1.0, 0.0), dvpsi, n));
  }
  if (d_wgk_out) *d_wgk_out = d_wgk;
  return SGW_OK;
}

// ---- k-point lanes (see drho_block) ---------------------------------------------------------------------------
__global__ void k_vadd(long n, cplx *__restrict__ a, const cplx *__restrict__ b) {            // a += b
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) a[i] = cadd(a[i], b[i]);
}

// how many k-points of one sgw_coulomb block run concurrently: only where one k-point cannot fill the GPU (few, short vectors)
// and where a workspace per lane is cheap.  SGW_KLANES overrides (1 = off).
static int klanes_wanted(sgw_ctx *ctx, const sgw_solver_cfg *cfg, int np, int nshift, int nfreq, size_t trho_bytes) {
  const size_t nk = ctx->pairs.size();
  if (nk < 2 || ctx->profiling || ctx->is_lane) return 1;      // per-class event timing assumes one stream
  const char *env = getenv("SGW_KLANES");
  const int forced = env ? atoi(env) : 0;
  int want = forced > 0 ? forced : 8;
  want = (int)std::min<size_t>((size_t)want, nk);
  if (want <= 1) return 1;
  size_t work = 0, bytes = 0;
  const size_t nnr = (size_t)ctx->nr1 * ctx->nr2 * ctx->nr3;
  for (auto &kp : ctx->pairs) {
    if (!kp.set || kp.slot < 0 || kp.slot >= (int)ctx->slots.size() || !ctx->slots[kp.slot].set) return 1;
    const KSlot &ks = ctx->slots[kp.slot];
    const size_t nrhs = (size_t)np * ks.nbnd;
    work = std::max(work, nrhs * (size_t)ks.npwx);
    size_t b = nrhs * solver_bytes_per_rhs(ctx, ks, cfg->bicg_lmax, nshift);
    if (cfg->npriority > 0 && cfg->priority[0] == 3) b += nrhs * (size_t)ks.npwx * sizeof(cplx) * 2 * 128;   // subspace bases
    b += 4 * (size_t)ks.nbnd * nnr * sizeof(cplx) + trho_bytes * 2;
    bytes = std::max(bytes, b);
  }
  if (forced <= 0 && work > ((size_t)1 << 22)) return 1;      // big batches fill the machine on their own
  size_t free_b = 0, total_b = 0;
  if (cudaMemGetInfo(&free_b, &total_b) != cudaSuccess) return 1;
  size_t held = 0;
  for (sgw_ctx *l : ctx->lanes) for (auto &kv : l->ws.bufs) held += kv.second.second;
  while (want > 1 && (double)bytes * (want - 1) > 0.5 * (double)(free_b + held)) --want;
  return want;
}

// make lanes 1 .. nlanes-1 exist and mirror the parent's tables (lane 0 is the parent itself)
static int klanes_prepare(sgw_ctx *ctx, int nlanes) {
  while ((int)ctx->lanes.size() < nlanes - 1) {
    sgw_ctx *l = new sgw_ctx();
    l->device = ctx->device;
    l->is_lane = true;
    if (cudaStreamCreateWithFlags(&l->own_stream, cudaStreamNonBlocking) != cudaSuccess) { delete l; ctx->err = "lane stream"; return SGW_E_CUDA; }
    cudaEventCreate(&l->ev0); cudaEventCreate(&l->ev1); cudaEventCreate(&l->ev2); cudaEventCreate(&l->ev3);
    ctx->lanes.push_back(l);
  }
  for (int i = 0; i < nlanes - 1; ++i) {
    sgw_ctx *l = ctx->lanes[i];
    // what the lane owns survives the copy of the parent's state
    cudaStream_t st = l->own_stream;
    Workspace w = std::move(l->ws);
    cudaEvent_t e0 = l->ev0, e1 = l->ev1, e2 = l->ev2, e3 = l->ev3, ei0 = l->ev_iter[0], ei1 = l->ev_iter[1];
    int *hf = l->h_flags;
    std::vector<cudaEvent_t> pool = std::move(l->ev_pool);
    *l = *ctx;
    l->is_lane = true;
    l->lanes.clear();
    l->stream = l->own_stream = st;
    l->ws = std::move(w);
    l->ev0 = e0; l->ev1 = e1; l->ev2 = e2; l->ev3 = e3; l->ev_iter[0] = ei0; l->ev_iter[1] = ei1;
    l->h_flags = hf;
    l->ev_pool = std::move(pool);
    l->prof_recs.clear();
    l->profiling = false;
    l->err.clear();
    l->launches = 0;
    memset(&l->stats, 0, sizeof(l->stats));
  }
  return SGW_OK;
}

static void klane_merge(sgw_ctx *ctx, sgw_ctx *l) {
  ctx->launches += l->launches;
  ctx->stats.n_linear_op += l->stats.n_linear_op;
  ctx->stats.n_fallback += l->stats.n_fallback;
  ctx->stats.n_outer_max = std::max(ctx->stats.n_outer_max, l->stats.n_outer_max);
  ctx->stats.ms_solver += l->stats.ms_solver;           // summed over lanes: concurrent lanes overlap in wall time
  l->launches = 0;
  memset(&l->stats, 0, sizeof(l->stats));
}

// Delta-rho of `np` perturbations whose dvbare(r) sit in d_field (permuted order): d_drhoG[(p*nfreq+ifreq)*rho.npw + pos]
// = fwfft(drho)(G) in the column order of `rho`, summed over the k-points of this pool.
static int drho_block(sgw_ctx *ctx, const sgw_solver_cfg *cfg, int np, const cplx *d_field, const FreqList &fl,
                      const Sphere &rho_fine, cplx *d_drhoG, int *ierr_any, bool coarse = false) {
  const int nfreq = fl.nfreq, nshift = fl.num_omega;
  const long nnr = (long)ctx->nr1 * ctx->nr2 * ctx->nr3;
  const int npf = np * nfreq;
  // grid of the Delta-rho accumulation: the context's box, or the alias-free coarse one (rho_grid_prepare)
  const FftGrid *rg = coarse ? &ctx->rho_grid : nullptr;
  const Sphere &rho = coarse ? ctx->rho_sph_c : rho_fine;
  const int rnz = coarse ? ctx->rho_grid.n3 : ctx->nr3;
  const long rnnr = coarse ? (long)ctx->rho_grid.n1 * ctx->rho_grid.n2 * ctx->rho_grid.n3 : nnr;
  cplx *Trho = nullptr;
  SGW_CHECK(ws(ctx, "co_Trho", (size_t)npf * rnz * rho.ncol, &Trho));
  // one k-point (solve_linter.f90:288 loop body) on the context `c` (the caller's, or one of its lanes): accumulates into Tacc
  auto one_k = [&](sgw_ctx *ctx, size_t ik, cplx *Tacc, bool accumulate, int *ierr_any) -> int {
    cudaStream_t st = ctx->stream;
    const FftGrid *rg = coarse ? &ctx->rho_grid : nullptr;
    const Sphere &rho = coarse ? ctx->rho_sph_c : rho_fine;
    const KPair &kp = ctx->pairs[ik];
    if (!kp.set || kp.slot < 0 || kp.slot >= (int)ctx->slots.size() || !ctx->slots[kp.slot].set) {
      ctx->err = "k-point pair not set (sgw_set_kpair / sgw_set_kpoint)";
      return SGW_E_STATE;
    }
    const KSlot &ks = ctx->slots[kp.slot];
    // bands of the solver loop :367 = nbnd_occ(ikk); for insulators that is the nbnd_occ(ikq) of the projector panel
    const bool metal = ctx->lgauss;
    if (metal && (kp.nbnd_all <= 0 || !kp.d_evq_all)) { ctx->err = "lgauss is set but sgw_set_kpair_metal was not called for this pair"; return SGW_E_STATE; }
    const int n = ks.npwx, nocc = metal ? kp.nocc_k : ks.nbnd;
    if (nocc > kp.nbnd) { ctx->err = "evc holds fewer bands than nbnd_occ"; return SGW_E_ARG; }
    const int nrhs = np * nocc;
    // psi_v(r) for the occupied bands: used by dV psi and again by the Delta-rho accumulation
    cplx *Tk = nullptr, *psir = nullptr, *Tq = nullptr, *dvpsi = nullptr, *d_sig = nullptr, *d_x = nullptr;
    int *d_ierr = nullptr;
    SGW_CHECK(ws(ctx, "co_Tk", (size_t)nocc * ctx->nr3 * kp.sph_k.ncol, &Tk));
    SGW_CHECK(ws(ctx, "co_psir", (size_t)nocc * nnr, &psir));
    struct ZClass {                              // z passes outside H.psi are billed to the stage they belong to
      sgw_ctx *c;
      ZClass(sgw_ctx *cc, int cls) : c(cc) { c->prof_z_class = cls; }
      ~ZClass() { c->prof_z_class = PC_FFT_Z; }
    };
    ZClass zc_setup(ctx, PC_OTHER);
    SGW_CHECK(fft_zpass_g2r(ctx, kp.sph_k, nocc, kp.d_evc, n, Tk, nullptr));
    SGW_CHECK(fft_plane(ctx, PLANE_TO_R, &kp.sph_k, nullptr, nocc, Tk, nullptr, nullptr, 1, psir, nullptr));
    // psi_v(r) once more on the Delta-rho grid when that is a different box
    const Sphere &sq = coarse ? ctx->pair_kq_c[ik] : ks.sph;
    cplx *psir_rho = psir;
    if (coarse) {
      const Sphere &sk = ctx->pair_k_c[ik];
      SGW_CHECK(ws(ctx, "co_psir_c", (size_t)nocc * rnnr, &psir_rho));
      SGW_CHECK(fft_zpass_g2r(ctx, sk, nocc, kp.d_evc, n, Tk, nullptr, rg));
      SGW_CHECK(fft_plane(ctx, PLANE_TO_R, &sk, nullptr, nocc, Tk, nullptr, nullptr, 1, psir_rho, nullptr, 0, rg));
    }
    // y-fastest copy of psi_v(r) on the Delta-rho grid for the accumulation stage of k_plane_rho_v2
    cplx *psir_t = nullptr;
    SGW_CHECK(ws(ctx, "co_psir_t", (size_t)nocc * rnnr, &psir_t));
    SGW_CHECK(fft_transpose_planes(ctx, rg, (long)nocc * rnz, psir_rho, psir_t));
    // dvqpsi_us.f90:99-130: dvpsi = fwfft(dvbare(r) psi(r)) on the k+q sphere (only the bands the solver uses)
    SGW_CHECK(ws(ctx, "co_Tq", (size_t)nrhs * ctx->nr3 * ks.sph.ncol, &Tq));
    SGW_CHECK(ws(ctx, "co_dvpsi", (size_t)nrhs * n, &dvpsi));
    SGW_CHECK(fft_plane(ctx, PLANE_FROM_R, nullptr, &ks.sph, nrhs, nullptr, Tq, d_field, nocc, psir, nullptr, nocc));
    SGW_CUDA(cudaMemsetAsync(dvpsi, 0, sizeof(cplx) * (size_t)nrhs * n, st));
    ZEpilogue epi;
    epi.mode = 0; epi.g2kin = nullptr; epi.psi = nullptr; epi.sigma = nullptr; epi.sigma_stride = 0; epi.keep_out = 0;
    SGW_CHECK(fft_zpass_r2g(ctx, ks.sph, nrhs, Tq, dvpsi, n, epi, nullptr));
    // [QE] orthogonalize (insulator): ps = evq^H dvpsi ; dvpsi <- evq ps - dvpsi   (solve_linter.f90:337)
    double *d_wgk = nullptr;
    SGW_CHECK(orthogonalize_dev(ctx, kp, ks, nocc, nrhs, dvpsi, &d_wgk));                  // solve_linter.f90:337
    // band loop :367-374 -> one batch; sigma = -(et + omega) :369
    {
      std::vector<cplx> sig((size_t)nshift * nrhs);
      for (int p = 0; p < np; ++p)
        for (int ib = 0; ib < nocc; ++ib)
          for (int is = 0; is < nshift; ++is)
            sig[(size_t)(p * nocc + ib) * nshift + is] = cmake(-(kp.et[ib] + fl.omega[is].x), -fl.omega[is].y);
      SGW_CHECK(ws(ctx, "sv_sig", (size_t)nshift * nrhs, &d_sig));
      SGW_CUDA(cudaMemcpyAsync(d_sig, sig.data(), sizeof(cplx) * sig.size(), cudaMemcpyHostToDevice, st));
      SGW_CUDA(cudaStreamSynchronize(st));
    }
    SGW_CHECK(ws(ctx, "sv_x", (size_t)n * nshift * nrhs, &d_x));
    SGW_CHECK(ws(ctx, "sv_ierr", (size_t)nrhs, &d_ierr));
    // the +-omega average (:464-480) is formed by the solver: dpsi_avg[(p * nfreq + ifreq) * nocc + ib], the order the
    // Delta-rho stage consumes
    cplx *davg_all = nullptr;
    int *d_done = nullptr;
    SGW_CHECK(ws(ctx, "co_davg_all", (size_t)npf * nocc * n, &davg_all));
    SGW_CHECK(ws(ctx, "co_done", (size_t)nrhs, &d_done));
    SolveBatch sb;
    sb.slot = kp.slot; sb.alpha_pv = ks.alpha_pv; sb.nrhs = nrhs; sb.nshift = nshift; sb.n = n;
    sb.d_b = dvpsi; sb.ldb = n; sb.d_sigma = d_sig; sb.d_x = d_x; sb.d_ierr = d_ierr;
    sb.avg.d_y = davg_all; sb.avg.nfreq = nfreq; sb.avg.zero_freq = fl.zero_freq; sb.avg.group = nocc; sb.avg.d_done = d_done;
    cudaEventRecord(ctx->ev2, st);
    ctx->prof_z_class = PC_FFT_Z;
    SGW_CHECK(select_solver_batched(ctx, sb, cfg));
    ctx->prof_z_class = PC_RHO_PLANE;
    cudaEventRecord(ctx->ev3, st);
    {
      std::vector<int> ie(nrhs);
      SGW_CUDA(cudaMemcpyAsync(ie.data(), d_ierr, sizeof(int) * nrhs, cudaMemcpyDeviceToHost, st));
      SGW_CUDA(cudaStreamSynchronize(st));
      for (int r = 0; r < nrhs; ++r) if (ie[r] != 0) *ierr_any = ie[r];                     // :370
      float ms = 0.f;
      cudaEventElapsedTime(&ms, ctx->ev2, ctx->ev3);
      ctx->stats.ms_solver += ms;
    }
    // dpsi *= wg/wk (= 1 for fully occupied bands, :373); average +-omega; incdrhoscf with weight 2 wk / omega
    if (metal) {
      const long nv = (long)npf * nocc;
      const int step = 65535 / nocc * nocc;
      for (long v0 = 0; v0 < nv; v0 += step) {
        const int c = (int)std::min<long>(step, nv - v0);
        dim3 gs((unsigned)((n + 255) / 256), (unsigned)c);
        k_band_scale<<<gs, 256, 0, st>>>(n, nocc, d_wgk, davg_all + (size_t)v0 * n);
        SGW_LAUNCH_CHECK();
      }
    }
    const double wgt = 2.0 * kp.wk / ctx->omega_cell;
    const size_t per_pf = (size_t)nocc * ((size_t)rnz * sq.ncol + n) * sizeof(cplx);
    size_t budget = (size_t)2 << 30;
    int pfc = (int)std::max<size_t>(1, std::min<size_t>(npf, budget / per_pf));
    pfc = std::min(pfc, std::max(1, 65535 / nocc));
    cplx *Td = nullptr;
    SGW_CHECK(ws(ctx, "co_Td", (size_t)pfc * nocc * rnz * sq.ncol, &Td));
    for (int pf0 = 0; pf0 < npf; pf0 += pfc) {
      const int c = std::min(pfc, npf - pf0);
      const cplx *davg = davg_all + (size_t)pf0 * nocc * n;
      SGW_CHECK(fft_zpass_g2r(ctx, sq, c * nocc, davg, n, Td, nullptr, rg));
      SGW_CHECK(fft_plane_rho(ctx, sq, rho, c, nocc, Td, psir_rho, wgt, Tacc + (size_t)pf0 * rnz * rho.ncol, accumulate, rg, psir_t));
    }
    return SGW_OK;
  };
  cudaStream_t st = ctx->stream;
  const size_t nk = ctx->pairs.size();
  const int nlanes = klanes_wanted(ctx, cfg, np, nshift, nfreq, (size_t)npf * rnz * rho.ncol * sizeof(cplx));
  if (nlanes <= 1) {
    for (size_t ik = 0; ik < nk; ++ik) SGW_CHECK(one_k(ctx, ik, Trho, ik > 0, ierr_any));          // solve_linter.f90:288
  } else {
    // Small systems (a few hundred plane waves, a handful of bands): the kernels of one k-point fill a fraction of the GPU and
    // the solve is a chain of short dependent launches.  The k-points of the loop :288 are independent until the Delta-rho
    // sum, so they run concurrently: lane l (own stream, own workspace, own host thread) takes k = l, l + L, ... and
    // accumulates its own Delta-rho; the lanes' sums are added in lane order afterwards (deterministic).
    SGW_CUDA(cudaStreamSynchronize(st));                                                    // d_field is ready
    SGW_CHECK(klanes_prepare(ctx, nlanes));
    std::vector<int> rcs(nlanes, SGW_OK), ierrs(nlanes, 0);
    std::vector<cplx *> Tl(nlanes, nullptr);
    Tl[0] = Trho;
    for (int l = 1; l < nlanes; ++l) SGW_CHECK(ws(ctx->lanes[l - 1], "co_Trho", (size_t)npf * rnz * rho.ncol, &Tl[l]));
    auto run_lane = [&](int l) {
      sgw_ctx *c = l == 0 ? ctx : ctx->lanes[l - 1];
      cudaSetDevice(c->device);
      bool first = true;
      for (size_t ik = l; ik < nk && rcs[l] == SGW_OK; ik += nlanes) {
        rcs[l] = one_k(c, ik, Tl[l], !first, &ierrs[l]);
        first = false;
      }
      if (cudaStreamSynchronize(c->stream) != cudaSuccess && rcs[l] == SGW_OK) { c->err = "lane stream failed"; rcs[l] = SGW_E_CUDA; }
    };
    {
      std::vector<std::thread> th;
      for (int l = 1; l < nlanes; ++l) th.emplace_back(run_lane, l);
      run_lane(0);
      for (auto &t : th) t.join();
    }
    for (int l = 0; l < nlanes; ++l) {
      if (l > 0) klane_merge(ctx, ctx->lanes[l - 1]);
      if (ierrs[l]) *ierr_any = ierrs[l];
    }
    for (int l = 0; l < nlanes; ++l)
      if (rcs[l] != SGW_OK) { if (l > 0) ctx->err = ctx->lanes[l - 1]->err; return rcs[l]; }
    const long tot = (long)npf * rnz * rho.ncol;
    for (int l = 1; l < nlanes; ++l) {
      k_vadd<<<(unsigned)std::min<long>((tot + 255) / 256, 65535L * 16), 256, 0, st>>>(tot, Trho, Tl[l]);
      SGW_LAUNCH_CHECK();
    }
  }
  // mp_sum over pools (:521) is the caller's (one pool per context); fwfft of drho on the density sphere
  SGW_CUDA(cudaMemsetAsync(d_drhoG, 0, sizeof(cplx) * (size_t)npf * rho.npw, st));
  ZEpilogue epi;
  epi.mode = 0; epi.g2kin = nullptr; epi.psi = nullptr; epi.sigma = nullptr; epi.sigma_stride = 0; epi.keep_out = 0;
  ctx->prof_z_class = PC_RHO_PLANE;
  for (int v0 = 0; v0 < npf; v0 += 32768) {
    const int c = std::min(32768, npf - v0);
    SGW_CHECK(fft_zpass_r2g(ctx, rho, c, Trho + (size_t)v0 * rnz * rho.ncol, d_drhoG + (size_t)v0 * rho.npw, rho.npw, epi, nullptr, rg));
  }
  ctx->prof_z_class = PC_FFT_Z;
  return SGW_OK;
}

// how many perturbations fit next to each other on the device
static int perturbation_chunk(sgw_ctx *ctx, const sgw_solver_cfg *cfg, int nshift, int nfs, const Sphere &rho, int want) {
  size_t free_b = 0, total_b = 0;
  if (cudaMemGetInfo(&free_b, &total_b) != cudaSuccess) return 1;
  size_t held = 0;
  for (auto &kv : ctx->ws.bufs) {                                   // workspace this pipeline re-uses counts as available
    const std::string &nm = kv.first;
    if (nm.rfind("ie_", 0) == 0 || nm.rfind("sg_", 0) == 0 || nm.rfind("an_", 0) == 0 || nm.rfind("gr_", 0) == 0 || nm.rfind("uf_", 0) == 0)
      continue;                                                     // other entry points' buffers stay allocated next to it
    held += kv.second.second;
  }
  const long nnr = (long)ctx->nr1 * ctx->nr2 * ctx->nr3;
  size_t per = 0, fixed = (size_t)3 << 30;
  for (auto &kp : ctx->pairs) {
    if (!kp.set || kp.slot < 0 || kp.slot >= (int)ctx->slots.size()) continue;
    const KSlot &ks = ctx->slots[kp.slot];
    size_t p = (size_t)ks.nbnd * solver_bytes_per_rhs(ctx, ks, cfg->bicg_lmax, nshift);
    p += (size_t)ks.nbnd * ctx->nr3 * ks.sph.ncol * sizeof(cplx);   // co_Tq
    per = std::max(per, p);
    fixed = std::max(fixed, ((size_t)3 << 30) + (size_t)ks.nbnd * (nnr + (size_t)ctx->nr3 * kp.sph_k.ncol) * sizeof(cplx));
  }
  per += (size_t)nnr * sizeof(cplx) + (size_t)nfs * ((size_t)ctx->nr3 * rho.ncol + rho.npw) * sizeof(cplx);
  const double avail = 0.9 * (double)(free_b + held) - (double)fixed;
  int c = avail > 0 ? (int)(avail / (double)per) : 1;
  c = std::max(1, std::min(c, want));
  // grid.y / grid.z limits of the batched kernels (vectors = perturbations x bands)
  int maxb = 1;
  for (auto &s : ctx->slots) if (s.set) maxb = std::max(maxb, s.nbnd);
  c = std::min(c, std::max(1, 65535 / (maxb * std::max(1, nfs))));
  return c;
}

// ================================================================ self-consistent branch of solve_linter
// (solve_linter.f90:376-460, :564-582) and mix_potential_c (mix_pot_c.f90:25-198) -- SURVEY section 8 row f1.
// Small BLAS-1 style kernels for the mixing (ndim = nnr * nfreq, a handful of calls per iteration):
__global__ void k_vsub(long n, cplx *__restrict__ a, const cplx *__restrict__ b) {            // a -= b
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) a[i] = csub(a[i], b[i]);
}
__global__ void k_vrdiff(long n, cplx *__restrict__ d, const cplx *__restrict__ a) {          // d = a - d
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) d[i] = csub(a[i], d[i]);
}
__global__ void k_vscale2(long n, double sc, cplx *__restrict__ a, cplx *__restrict__ b) {    // a *= sc ; b *= sc
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) { a[i] = cscale(sc, a[i]); b[i] = cscale(sc, b[i]); }
}
__global__ void k_vaxpy_r(long n, double al, const cplx *__restrict__ x, cplx *__restrict__ y) {   // y += al x
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) y[i] = cmake(y[i].x + al * x[i].x, y[i].y + al * x[i].y);
}
// vin -= gamma * (alphamix * df + dv)      (mix_pot_c.f90:179-181, w(i) = 1)
__global__ void k_vbroyden(long n, cplx gamma, double al, const cplx *__restrict__ df, const cplx *__restrict__ dv,
                           cplx *__restrict__ vin) {
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const cplx t = cmake(al * df[i].x + dv[i].x, al * df[i].y + dv[i].y);
  vin[i] = csub(vin[i], cmul(gamma, t));
}
// ZDOTC partials: part[block] = sum conj(x) y over the block's slice (summed on the host in block order: deterministic)
__global__ void __launch_bounds__(256) k_zdotc_part(long n, const cplx *__restrict__ x, const cplx *__restrict__ y,
                                                     cplx *__restrict__ part) {
  __shared__ cplx sm[256];
  cplx acc = cmake(0.0, 0.0);
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x)
    acc = cfma(cconj(x[i]), y[i], acc);
  sm[threadIdx.x] = acc;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if ((int)threadIdx.x < o) sm[threadIdx.x] = cadd(sm[threadIdx.x], sm[threadIdx.x + o]);
    __syncthreads();
  }
  if (threadIdx.x == 0) part[blockIdx.x] = sm[0];
}
// right-hand sides of the per-frequency solves: b2[(ib * num_omega + io)] = rhs[(ifreq(io) * nocc + ib)]   (:434-456)
__global__ void k_iter_rhs(int n, int nocc, int nfreq, int num_omega, int zero_freq, const cplx *__restrict__ rhs,
                           cplx *__restrict__ b2) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n) return;
  const int ib = blockIdx.y, io = blockIdx.z;
  const int ifreq = io < nfreq ? io : (zero_freq ? io - nfreq + 1 : io - nfreq);
  b2[((long)ib * num_omega + io) * n + e] = rhs[((long)ifreq * nocc + ib) * n + e];
}

static int zdotc_dev(sgw_ctx *ctx, long n, const cplx *x, const cplx *y, cplx *out) {
  const int nb = (int)std::min<long>(512, (n + 255) / 256);
  cplx *part = nullptr;
  SGW_CHECK(ws(ctx, "it_part", (size_t)512, &part));
  k_zdotc_part<<<nb, 256, 0, ctx->stream>>>(n, x, y, part);
  SGW_LAUNCH_CHECK();
  cplx h[512];
  SGW_CUDA(cudaMemcpyAsync(h, part, sizeof(cplx) * nb, cudaMemcpyDeviceToHost, ctx->stream));
  SGW_CUDA(cudaStreamSynchronize(ctx->stream));
  cplx s = cmake(0.0, 0.0);
  for (int i = 0; i < nb; ++i) s = cadd(s, h[i]);
  *out = s;
  return SGW_OK;
}

// inverse of a small complex matrix (<= 8 x 8), Gauss-Jordan with partial pivoting; column-major, ld = n
static bool small_inverse(int n, cplx *a) {
  cplx inv[64];
  for (int i = 0; i < n * n; ++i) inv[i] = cmake(0.0, 0.0);
  for (int i = 0; i < n; ++i) inv[i + n * i] = cmake(1.0, 0.0);
  for (int c = 0; c < n; ++c) {
    int p = c;
    double best = std::fabs(a[c + n * c].x) + std::fabs(a[c + n * c].y);
    for (int r = c + 1; r < n; ++r) {
      const double v = std::fabs(a[r + n * c].x) + std::fabs(a[r + n * c].y);
      if (v > best) { best = v; p = r; }
    }
    if (best == 0.0) return false;
    if (p != c)
      for (int k = 0; k < n; ++k) { std::swap(a[c + n * k], a[p + n * k]); std::swap(inv[c + n * k], inv[p + n * k]); }
    const cplx d = cdiv(cmake(1.0, 0.0), a[c + n * c]);
    for (int k = 0; k < n; ++k) { a[c + n * k] = cmul(d, a[c + n * k]); inv[c + n * k] = cmul(d, inv[c + n * k]); }
    for (int r = 0; r < n; ++r) {
      if (r == c) continue;
      const cplx f = a[r + n * c];
      for (int k = 0; k < n; ++k) {
        a[r + n * k] = csub(a[r + n * k], cmul(f, a[c + n * k]));
        inv[r + n * k] = csub(inv[r + n * k], cmul(f, inv[c + n * k]));
      }
    }
  }
  for (int i = 0; i < n * n; ++i) a[i] = inv[i];
  return true;
}

struct MixState {          // df, dv of mix_pot_c.f90:83 (device), allocated at iter == 1
  cplx *df = nullptr, *dv = nullptr, *vinsave = nullptr;
};

// mix_potential_c(ndim, vout, vin, alphamix, dr2, tr2, iter, n_iter, conv) on device vectors
static int mix_potential_c_dev(sgw_ctx *ctx, MixState &m, long ndim, cplx *vout, cplx *vin, double alphamix, double *dr2,
                               double tr2, int iter, int n_iter, bool *conv) {
  cudaStream_t st = ctx->stream;
  const unsigned gb = (unsigned)((ndim + 255) / 256);
  k_vsub<<<gb, 256, 0, st>>>(ndim, vout, vin);                                              // :97-99
  SGW_LAUNCH_CHECK();
  cplx d;
  SGW_CHECK(zdotc_dev(ctx, ndim, vout, vout, &d));
  *dr2 = (std::sqrt(d.x) / (double)ndim) * (std::sqrt(d.x) / (double)ndim);                 // :100-106
  *conv = *dr2 < tr2;                                                                      // :108
  if (*conv) return SGW_OK;                                                                // :114-118
  const int iter_used = std::min(iter - 1, n_iter);                                        // :124
  const int ipos = iter - 1 - ((iter - 2) / n_iter) * n_iter;                              // :129
  if (iter > 1) {                                                                          // :131-140
    cplx *dfp = m.df + (size_t)ndim * (ipos - 1), *dvp = m.dv + (size_t)ndim * (ipos - 1);
    k_vrdiff<<<gb, 256, 0, st>>>(ndim, dfp, vout);
    SGW_LAUNCH_CHECK();
    k_vrdiff<<<gb, 256, 0, st>>>(ndim, dvp, vin);
    SGW_LAUNCH_CHECK();
    SGW_CHECK(zdotc_dev(ctx, ndim, dfp, dfp, &d));
    k_vscale2<<<gb, 256, 0, st>>>(ndim, 1.0 / std::sqrt(d.x), dfp, dvp);
    SGW_LAUNCH_CHECK();
  }
  SGW_CUDA(cudaMemcpyAsync(m.vinsave, vin, sizeof(cplx) * ndim, cudaMemcpyDeviceToDevice, st));   // :142
  cplx beta[64], work[8];
  const double w0 = 0.01;
  for (int i = 0; i < iter_used; ++i) {                                                    // :144-149
    for (int j = i + 1; j < iter_used; ++j) {
      SGW_CHECK(zdotc_dev(ctx, ndim, m.df + (size_t)ndim * j, m.df + (size_t)ndim * i, &d));
      beta[i + iter_used * j] = d;
      beta[j + iter_used * i] = cconj(d);
    }
    beta[i + iter_used * i] = cmake(w0 * w0 + 1.0, 0.0);
  }
  if (iter_used > 0 && !small_inverse(iter_used, beta)) {                                  // :153-157
    ctx->err = "broyden: factorization of the mixing matrix failed";
    return SGW_E_ARG;
  }
  for (int i = 0; i < iter_used; ++i)                                                      // :159-163
    for (int j = i + 1; j < iter_used; ++j) beta[j + iter_used * i] = cconj(beta[i + iter_used * j]);
  for (int i = 0; i < iter_used; ++i) SGW_CHECK(zdotc_dev(ctx, ndim, m.df + (size_t)ndim * i, vout, &work[i]));   // :165-167
  k_vaxpy_r<<<gb, 256, 0, st>>>(ndim, alphamix, vout, vin);                                // :169-171
  SGW_LAUNCH_CHECK();
  for (int i = 0; i < iter_used; ++i) {                                                    // :173-182
    cplx gamma = cmake(0.0, 0.0);
    for (int j = 0; j < iter_used; ++j) gamma = cfma(beta[j + iter_used * i], work[j], gamma);
    k_vbroyden<<<gb, 256, 0, st>>>(ndim, gamma, alphamix, m.df + (size_t)ndim * i, m.dv + (size_t)ndim * i, vin);
    SGW_LAUNCH_CHECK();
  }
  const int inext = iter - ((iter - 1) / n_iter) * n_iter;                                 // :184
  SGW_CUDA(cudaMemcpyAsync(m.df + (size_t)ndim * (inext - 1), vout, sizeof(cplx) * ndim, cudaMemcpyDeviceToDevice, st));
  SGW_CUDA(cudaMemcpyAsync(m.dv + (size_t)ndim * (inext - 1), m.vinsave, sizeof(cplx) * ndim, cudaMemcpyDeviceToDevice, st));
  return SGW_OK;
}

// One perturbation, num_iter > 1.  d_field: dvbare(r) in permuted order; on success d_dvscfin (nfreq x nnr, permuted
// order) holds the self-consistent dV_scf(r, omega).  ierr: solver code, or 10 if not converged within num_iter.
static int solve_linter_iter_core(sgw_ctx *ctx, const sgw_solver_cfg *cfg_global, int num_iter, const cplx *d_field,
                                  double meandvb, const FreqList &fl, const Sphere &rho, cplx *d_dvscfin, int *ierr_out,
                                  int *iter_done) {
  const int nfreq = fl.nfreq, nshift = fl.num_omega;
  const long nnr = (long)ctx->nr1 * ctx->nr2 * ctx->nr3;
  const long ndim = nnr * nfreq;
  cudaStream_t st = ctx->stream;
  SGW_ARG(ctx->mix_niter >= num_iter && ctx->mix_nmix >= 1 && ctx->mix_nmix <= 8,
          "sgw_set_mixing must provide alpha_mix for every iteration and 1 <= nmix_gw <= 8");
  sgw_solver_cfg config = *cfg_global;                                                     // :223
  cplx *dvout = nullptr, *Trho = nullptr, *d_drhoG = nullptr, *Tr = nullptr;
  MixState mix;
  SGW_CHECK(ws(ctx, "it_dvout", (size_t)ndim, &dvout));
  SGW_CHECK(ws(ctx, "it_df", (size_t)ndim * ctx->mix_nmix, &mix.df));
  SGW_CHECK(ws(ctx, "it_dv", (size_t)ndim * ctx->mix_nmix, &mix.dv));
  SGW_CHECK(ws(ctx, "it_vinsave", (size_t)ndim, &mix.vinsave));
  SGW_CHECK(ws(ctx, "co_Trho", (size_t)nfreq * ctx->nr3 * rho.ncol, &Trho));
  SGW_CHECK(ws(ctx, "co_drhoG", (size_t)nfreq * rho.npw, &d_drhoG));
  SGW_CHECK(ws(ctx, "sl_Tr", (size_t)nfreq * ctx->nr3 * rho.ncol, &Tr));
  SGW_CUDA(cudaMemsetAsync(d_dvscfin, 0, sizeof(cplx) * ndim, st));                        // :343
  SGW_CUDA(cudaMemsetAsync(mix.df, 0, sizeof(cplx) * ndim * ctx->mix_nmix, st));
  SGW_CUDA(cudaMemsetAsync(mix.dv, 0, sizeof(cplx) * ndim * ctx->mix_nmix, st));
  double *d_fac = nullptr;
  SGW_CHECK(hartree_factor(ctx, rho, rho.npw, "co_fac", &d_fac));
  int zero_pos = -1;
  if (meandvb < 1e-10)
    for (int pos = 0; pos < rho.npw; ++pos) if (rho.perm[pos] == 0) zero_pos = pos;         // :544-550
  // storage for dvbare psi of every k-point (buffer iubar, :332)
  size_t bare_tot = 0;
  std::vector<size_t> bare_off(ctx->pairs.size());
  for (size_t ik = 0; ik < ctx->pairs.size(); ++ik) {
    const KPair &kp = ctx->pairs[ik];
    if (!kp.set || kp.slot < 0 || kp.slot >= (int)ctx->slots.size() || !ctx->slots[kp.slot].set) {
      ctx->err = "k-point pair not set (sgw_set_kpair / sgw_set_kpoint)";
      return SGW_E_STATE;
    }
    bare_off[ik] = bare_tot;
    bare_tot += (size_t)std::max(ctx->slots[kp.slot].nbnd, kp.nocc_k) * ctx->slots[kp.slot].npwx;   // nbnd_occ(ikk) bands (metals: may exceed nbnd_occ(ikq))
  }
  cplx *bare = nullptr;
  SGW_CHECK(ws(ctx, "it_bare", bare_tot, &bare));
  double dr2 = 0.0;
  bool convt = false;
  int ierr_any = 0, iter;
  ZEpilogue epi0;
  epi0.mode = 0; epi0.g2kin = nullptr; epi0.psi = nullptr; epi0.sigma = nullptr; epi0.sigma_stride = 0; epi0.keep_out = 0;
  for (iter = 1; iter <= num_iter; ++iter) {                                               // :279
    const bool first = iter == 1;
    for (size_t ik = 0; ik < ctx->pairs.size(); ++ik) {                                    // :288
      const KPair &kp = ctx->pairs[ik];
      const KSlot &ks = ctx->slots[kp.slot];
      const bool metal = ctx->lgauss;
      if (metal && (kp.nbnd_all <= 0 || !kp.d_evq_all)) { ctx->err = "lgauss is set but sgw_set_kpair_metal was not called for this pair"; return SGW_E_STATE; }
      const int n = ks.npwx, nocc = metal ? kp.nocc_k : ks.nbnd;                           // bands of the loop :367 = nbnd_occ(ikk)
      if (nocc > kp.nbnd) { ctx->err = "evc holds fewer bands than nbnd_occ"; return SGW_E_ARG; }
      double *d_wgk = nullptr;
      cplx *Tk = nullptr, *psir = nullptr, *Tq = nullptr, *rhs = nullptr, *d_sig = nullptr, *d_x = nullptr,
           *b2 = nullptr, *davg = nullptr, *Td = nullptr;
      int *d_ierr = nullptr;
      SGW_CHECK(ws(ctx, "co_Tk", (size_t)nocc * ctx->nr3 * kp.sph_k.ncol, &Tk));
      SGW_CHECK(ws(ctx, "co_psir", (size_t)nocc * nnr, &psir));
      SGW_CHECK(fft_zpass_g2r(ctx, kp.sph_k, nocc, kp.d_evc, n, Tk, nullptr));
      SGW_CHECK(fft_plane(ctx, PLANE_TO_R, &kp.sph_k, nullptr, nocc, Tk, nullptr, nullptr, 1, psir, nullptr));
      cplx *bare_k = bare + bare_off[ik];
      const int nrhs = first ? nocc : nocc * nshift;
      SGW_CHECK(ws(ctx, "sv_sig", (size_t)nocc * nshift, &d_sig));
      SGW_CHECK(ws(ctx, "sv_x", (size_t)n * nshift * nocc, &d_x));
      SGW_CHECK(ws(ctx, "sv_ierr", (size_t)nocc * nshift, &d_ierr));
      {
        std::vector<cplx> sig((size_t)nshift * nocc);                                      // sigma = -(et + omega) :369
        for (int ib = 0; ib < nocc; ++ib)
          for (int is = 0; is < nshift; ++is) sig[(size_t)ib * nshift + is] = cmake(-(kp.et[ib] + fl.omega[is].x), -fl.omega[is].y);
        SGW_CUDA(cudaMemcpyAsync(d_sig, sig.data(), sizeof(cplx) * sig.size(), cudaMemcpyHostToDevice, st));
        SGW_CUDA(cudaStreamSynchronize(st));
      }
      SolveBatch sb;
      sb.slot = kp.slot; sb.alpha_pv = ks.alpha_pv; sb.n = n; sb.ldb = n; sb.d_sigma = d_sig; sb.d_x = d_x; sb.d_ierr = d_ierr;
      if (first) {
        // dvqpsi_us (:331) -> buffer iubar (:332) -> orthogonalize (:337) -> multishift solves at thresh 1e-2 (:362)
        SGW_CHECK(ws(ctx, "co_Tq", (size_t)nocc * ctx->nr3 * ks.sph.ncol, &Tq));
        SGW_CHECK(fft_plane(ctx, PLANE_FROM_R, nullptr, &ks.sph, nocc, nullptr, Tq, d_field, nocc, psir, nullptr, nocc));
        SGW_CUDA(cudaMemsetAsync(bare_k, 0, sizeof(cplx) * (size_t)nocc * n, st));
        SGW_CHECK(fft_zpass_r2g(ctx, ks.sph, nocc, Tq, bare_k, n, epi0, nullptr));
        SGW_CHECK(ws(ctx, "co_dvpsi", (size_t)nocc * n, &rhs));
        SGW_CUDA(cudaMemcpyAsync(rhs, bare_k, sizeof(cplx) * (size_t)nocc * n, cudaMemcpyDeviceToDevice, st));
        SGW_CHECK(orthogonalize_dev(ctx, kp, ks, nocc, nocc, rhs, &d_wgk));                 // :337
        config.threshold = 1.0e-2;
        sb.nrhs = nocc; sb.nshift = nshift; sb.d_b = rhs;
        SGW_CHECK(select_solver_batched(ctx, sb, &config));
      } else {
        // dvpsi(ifreq) = dvbare psi + fwfft(dvscfin(r, ifreq) psi(r))  (:389-399), orthogonalize (:409), then one
        // single-shift solve per (band, +-omega) (:434-456) -- the bands x frequencies batch
        const int nv = nocc * nfreq;
        SGW_CHECK(ws(ctx, "co_Tq", (size_t)nv * ctx->nr3 * ks.sph.ncol, &Tq));
        SGW_CHECK(ws(ctx, "co_dvpsi", (size_t)nv * n, &rhs));
        for (int f = 0; f < nfreq; ++f)
          SGW_CUDA(cudaMemcpyAsync(rhs + (size_t)f * nocc * n, bare_k, sizeof(cplx) * (size_t)nocc * n, cudaMemcpyDeviceToDevice, st));
        SGW_CHECK(fft_plane(ctx, PLANE_FROM_R, nullptr, &ks.sph, nv, nullptr, Tq, d_dvscfin, nocc, psir, nullptr, nocc));
        ZEpilogue epi2 = epi0;
        epi2.mode = 2;                                                                     // cft_wave(-1) ADDS to dvpsi
        SGW_CHECK(fft_zpass_r2g(ctx, ks.sph, nv, Tq, rhs, n, epi2, nullptr));
        SGW_CHECK(orthogonalize_dev(ctx, kp, ks, nocc, nv, rhs, &d_wgk));                   // :409 (vector f * nocc + ib: band = column % nocc)
        SGW_CHECK(ws(ctx, "it_b2", (size_t)nocc * nshift * n, &b2));
        {
          dim3 gr((n + 255) / 256, nocc, nshift);
          k_iter_rhs<<<gr, 256, 0, st>>>(n, nocc, nfreq, nshift, fl.zero_freq, rhs, b2);
          SGW_LAUNCH_CHECK();
        }
        config.threshold = std::min(1.0e-1 * std::sqrt(dr2), 1.0e-2);                      // :417
        sb.nrhs = nrhs; sb.nshift = 1; sb.d_b = b2;
        SGW_CHECK(select_solver_batched(ctx, sb, &config));
      }
      {
        std::vector<int> ie(nrhs);
        SGW_CUDA(cudaMemcpyAsync(ie.data(), d_ierr, sizeof(int) * nrhs, cudaMemcpyDeviceToHost, st));
        SGW_CUDA(cudaStreamSynchronize(st));
        for (int r = 0; r < nrhs; ++r) if (ie[r] != 0) ierr_any = ie[r];                     // :370
      }
      // +-omega average (:464-480) and incdrhoscf (:489-497); x is laid out [(ib * nshift + is) * n] in both branches
      const double wgt = 2.0 * kp.wk / ctx->omega_cell;
      SGW_CHECK(ws(ctx, "co_davg", (size_t)nfreq * nocc * n, &davg));
      SGW_CHECK(ws(ctx, "co_Td", (size_t)nfreq * nocc * ctx->nr3 * ks.sph.ncol, &Td));
      dim3 ga((n + 255) / 256, nocc, nfreq);
      k_average<<<ga, 256, 0, st>>>(n, nocc, nfreq, nshift, fl.zero_freq, 0, d_x, davg);
      SGW_LAUNCH_CHECK();
      if (metal) {                                                                         // dpsi *= wg / wk (:373), linear: after the average
        dim3 gs((unsigned)((n + 255) / 256), (unsigned)(nfreq * nocc));
        k_band_scale<<<gs, 256, 0, st>>>(n, nocc, d_wgk, davg);
        SGW_LAUNCH_CHECK();
      }
      SGW_CHECK(fft_zpass_g2r(ctx, ks.sph, nfreq * nocc, davg, n, Td, nullptr));
      SGW_CHECK(fft_plane_rho(ctx, ks.sph, rho, nfreq, nocc, Td, psir, wgt, Trho, ik > 0));
    }
    if (ierr_any) break;                                                                   // errore aborts the reference
    // dvscfout = dv_of_drho(drho) (:534-558), real space, permuted order
    SGW_CUDA(cudaMemsetAsync(d_drhoG, 0, sizeof(cplx) * (size_t)nfreq * rho.npw, st));
    SGW_CHECK(fft_zpass_r2g(ctx, rho, nfreq, Trho, d_drhoG, rho.npw, epi0, nullptr));
    {
      dim3 gr((rho.npw + 255) / 256, nfreq);
      k_hartree<<<gr, 256, 0, st>>>(rho.npw, nfreq, d_fac, d_drhoG, zero_pos, +1.0);
      SGW_LAUNCH_CHECK();
    }
    SGW_CHECK(fft_zpass_g2r(ctx, rho, nfreq, d_drhoG, rho.npw, Tr, nullptr));
    SGW_CHECK(fft_plane(ctx, PLANE_TO_R, &rho, nullptr, nfreq, Tr, nullptr, nullptr, 1, dvout, nullptr));
    // mix with the old potential (:566-568)
    SGW_CHECK(mix_potential_c_dev(ctx, mix, ndim, dvout, d_dvscfin, ctx->mix_alpha[iter - 1], &dr2, ctx->mix_tr2 * nfreq, iter,
                                  ctx->mix_nmix, &convt));
    if (convt) break;                                                                      // :582
  }
  *iter_done = std::min(iter, num_iter);
  if (!ierr_any && !convt) ierr_any = 10;                                                  // :588-591
  *ierr_out = ierr_any;
  return SGW_OK;
}

static int check_pipeline_state(sgw_ctx *ctx) {
  if (!ctx->grid_set || !ctx->vloc_set) { ctx->err = "grid / local potential not set"; return SGW_E_STATE; }
  if (!ctx->system_set) { ctx->err = "sgw_set_system must be called first"; return SGW_E_STATE; }
  if (ctx->pairs.empty()) { ctx->err = "no k-point pairs (sgw_set_nksq / sgw_set_kpair)"; return SGW_E_STATE; }
  return SGW_OK;
}

}  // namespace sgw

extern "C" {

int sgw_set_system(sgw_ctx *ctx, double omega_cell, double tpiba2, int ngm, const double *g, const int32_t *nl) {
  if (!ctx) return SGW_E_ARG;
  cudaSetDevice(ctx->device);
  if (!ctx->grid_set) { ctx->err = "sgw_set_grid must be called first"; return SGW_E_STATE; }
  SGW_ARG(omega_cell > 0 && tpiba2 > 0 && ngm > 0 && g && nl, "bad cell / G-vector data");
  ctx->tables_version++;
  ctx->omega_cell = omega_cell;
  ctx->tpiba2 = tpiba2;
  ctx->ngm = ngm;
  ctx->g.assign(g, g + 3 * (size_t)ngm);
  ctx->nl.assign(nl, nl + ngm);
  if (grid_padded(ctx)) for (auto &v : ctx->nl) v = unpad_index(ctx, v);
  for (auto &kv : ctx->rho_spheres) free_sphere(&kv.second);
  ctx->rho_spheres.clear();
  ctx->system_set = true;
  return SGW_OK;
}

int sgw_set_q(sgw_ctx *ctx, const double *xq) {
  if (!ctx || !xq) return SGW_E_ARG;
  for (int i = 0; i < 3; ++i) ctx->xq[i] = xq[i];
  return SGW_OK;
}

int sgw_set_nksq(sgw_ctx *ctx, int nksq) {
  if (!ctx) return SGW_E_ARG;
  cudaSetDevice(ctx->device);
  SGW_ARG(nksq >= 0 && nksq < (1 << 20), "bad nksq");
  ctx->tables_version++;
  for (auto &p : ctx->pairs) {
    free_sphere(&p.sph_k);
    if (p.d_evc) dev_free(p.d_evc);
    if (p.d_evq_all) dev_free(p.d_evq_all);
  }
  ctx->pairs.assign(nksq, KPair());
  return SGW_OK;
}

int sgw_set_kpair(sgw_ctx *ctx, int ik, int slot_kq, int npw_k, const int32_t *nl_igk_k, int nbnd, const sgw_cplx *evc,
                  const double *et, double wk) {
  if (!ctx) return SGW_E_ARG;
  cudaSetDevice(ctx->device);
  SGW_ARG(ik >= 0 && ik < (int)ctx->pairs.size(), "ik outside 0..nksq-1 (sgw_set_nksq)");
  if (slot_kq < 0 || slot_kq >= (int)ctx->slots.size() || !ctx->slots[slot_kq].set || ctx->slots[slot_kq].dense) {
    ctx->err = "slot_kq must be a plane-wave slot set with sgw_set_kpoint";
    return SGW_E_STATE;
  }
  const KSlot &ks = ctx->slots[slot_kq];
  SGW_ARG(npw_k > 0 && npw_k <= ks.npwx && nl_igk_k && nbnd > 0 && evc && et, "bad k-point data");
  KPair &kp = ctx->pairs[ik];
  ctx->tables_version++;
  free_sphere(&kp.sph_k);
  if (kp.d_evc) { dev_free(kp.d_evc); kp.d_evc = nullptr; }
  if (kp.d_evq_all) { dev_free(kp.d_evq_all); kp.d_evq_all = nullptr; }       // metal data belongs to the pair it was set for
  kp.nbnd_all = kp.nocc_k = 0;
  {
    std::vector<int32_t> nlc(nl_igk_k, nl_igk_k + npw_k);
    if (grid_padded(ctx)) for (auto &v : nlc) v = unpad_index(ctx, v);
    SGW_CHECK(build_sphere(ctx, npw_k, nlc.data(), &kp.sph_k));
  }
  const int npwx = ks.npwx;
  {
    cplx *stage = nullptr;
    SGW_CHECK(ws(ctx, "io_in", (size_t)npwx * nbnd, &stage));
    SGW_CUDA(dev_malloc((void **)&kp.d_evc, sizeof(cplx) * (size_t)npwx * nbnd));
    SGW_CHECK(h2d_large(ctx, stage, evc, sizeof(cplx) * (size_t)npwx * nbnd));
    SGW_CHECK(permute_in(ctx, kp.sph_k, nbnd, stage, npwx, kp.d_evc, npwx, npwx));     // rows >= npw_k are zeroed
    SGW_CUDA(cudaStreamSynchronize(ctx->stream));
  }
  kp.slot = slot_kq; kp.npw_k = npw_k; kp.nbnd = nbnd; kp.wk = wk;
  kp.et.assign(et, et + nbnd);
  kp.set = true;
  return SGW_OK;
}

int sgw_set_smearing(sgw_ctx *ctx, int lgauss, double ef, double degauss, int ngauss) {
  if (!ctx) return SGW_E_ARG;
  SGW_ARG(!lgauss || degauss > 0.0, "degauss must be positive for a metal");
  SGW_ARG(ngauss == -99 || ngauss == -1 || (ngauss >= 0 && ngauss <= 10), "ngauss must be -99, -1 or 0..10");
  ctx->lgauss = lgauss != 0; ctx->ef = ef; ctx->degauss = degauss; ctx->ngauss = ngauss;
  return SGW_OK;
}

int sgw_set_kpair_metal(sgw_ctx *ctx, int ik, int nbnd, const sgw_cplx *evq_all, const double *et_q, int nbnd_occ_k,
                        const double *wg_over_wk) {
  if (!ctx) return SGW_E_ARG;
  cudaSetDevice(ctx->device);
  SGW_ARG(ik >= 0 && ik < (int)ctx->pairs.size() && ctx->pairs[ik].set, "pair not set (sgw_set_kpair comes first)");
  KPair &kp = ctx->pairs[ik];
  const KSlot &ks = ctx->slots[kp.slot];
  SGW_ARG(nbnd >= ks.nbnd && evq_all && et_q && nbnd_occ_k > 0 && nbnd_occ_k <= kp.nbnd && wg_over_wk, "bad metal data for the pair");
  ctx->tables_version++;
  if (kp.d_evq_all) { dev_free(kp.d_evq_all); kp.d_evq_all = nullptr; }
  const int npwx = ks.npwx;
  cplx *stage = nullptr;
  SGW_CHECK(ws(ctx, "io_in", (size_t)npwx * nbnd, &stage));
  SGW_CUDA(dev_malloc((void **)&kp.d_evq_all, sizeof(cplx) * (size_t)npwx * nbnd));
  SGW_CUDA(cudaMemcpyAsync(stage, evq_all, sizeof(cplx) * (size_t)npwx * nbnd, cudaMemcpyHostToDevice, ctx->stream));
  SGW_CHECK(permute_in(ctx, ks.sph, nbnd, stage, npwx, kp.d_evq_all, npwx, npwx));
  SGW_CUDA(cudaStreamSynchronize(ctx->stream));
  kp.nbnd_all = nbnd; kp.nocc_k = nbnd_occ_k;
  kp.et_q.assign(et_q, et_q + nbnd);
  kp.wg_over_wk.assign(wg_over_wk, wg_over_wk + nbnd_occ_k);
  return SGW_OK;
}

int sgw_set_mixing(sgw_ctx *ctx, int niter_gw, const double *alpha_mix, double tr2_gw, int nmix_gw) {
  if (!ctx) return SGW_E_ARG;
  SGW_ARG(niter_gw >= 1 && alpha_mix && tr2_gw > 0.0 && nmix_gw >= 1 && nmix_gw <= 8, "bad mixing parameters (1 <= nmix_gw <= 8)");
  ctx->mix_niter = niter_gw;
  ctx->mix_alpha.assign(alpha_mix, alpha_mix + niter_gw);
  ctx->mix_tr2 = tr2_gw;
  ctx->mix_nmix = nmix_gw;
  return SGW_OK;
}

int sgw_set_solve_direct(sgw_ctx *ctx, int solve_direct) {
  if (!ctx) return SGW_E_ARG;
  ctx->solve_direct = solve_direct != 0;
  return SGW_OK;
}

int sgw_get_scf_iterations(const sgw_ctx *ctx) { return ctx ? ctx->last_scf_iter : SGW_E_ARG; }

int sgw_solve_linter(sgw_ctx *ctx, const sgw_solver_cfg *cfg, int num_iter, const sgw_cplx *dvbarein, int nfreq,
                     const sgw_cplx *freq, sgw_cplx *drhoscf, int32_t *ierr_out) {
  if (!ctx) return SGW_E_ARG;
  cudaSetDevice(ctx->device);
  SGW_ARG(cfg && dvbarein && freq && drhoscf && ierr_out && nfreq > 0, "null argument");
  SGW_ARG(cfg->npriority >= 1 && cfg->npriority <= 4, "priority of the solvers not specified");
  SGW_ARG(num_iter >= 1, "num_iter must be >= 1");
  SGW_CHECK(check_pipeline_state(ctx));
  begin_call(ctx);
  const long nnr = (long)ctx->nr1 * ctx->nr2 * ctx->nr3;
  const FreqList fl = make_omega(nfreq, freq);
  Sphere *rho = nullptr;
  SGW_CHECK(get_rho_sphere(ctx, ctx->ngm, &rho));
  cudaStream_t st = ctx->stream;
  GridDev g = grid_dev(ctx);
  cplx *d_nat = nullptr, *d_field = nullptr, *d_drhoG = nullptr, *Tr = nullptr;
  SGW_CHECK(ws(ctx, "sl_nat", (size_t)nnr * nfreq, &d_nat));
  SGW_CHECK(ws(ctx, "co_field", (size_t)nnr, &d_field));
  // padded boxes: the caller's arrays have nr1x * nr2x * nr3x entries per frequency; the library's are compact
  const bool pad = grid_padded(ctx);
  const long nnrx = (long)ctx->nr1x * ctx->nr2x * ctx->nr3x;
  std::vector<sgw_cplx> hbuf;
  auto compact_index = [&](long i) {
    const long x = i % ctx->nr1, y = (i / ctx->nr1) % ctx->nr2, z = i / ((long)ctx->nr1 * ctx->nr2);
    return x + (long)ctx->nr1x * (y + (long)ctx->nr2x * z);
  };
  auto download = [&](const cplx *d_src) -> int {                  // d_src(nnr, nfreq) -> drhoscf(nnrx, nfreq)
    if (!pad) {
      SGW_CUDA(cudaMemcpyAsync(drhoscf, d_src, sizeof(cplx) * (size_t)nnr * nfreq, cudaMemcpyDeviceToHost, st));
      SGW_CUDA(cudaStreamSynchronize(st));
      return SGW_OK;
    }
    hbuf.resize((size_t)nnr * nfreq);
    SGW_CUDA(cudaMemcpyAsync(hbuf.data(), d_src, sizeof(cplx) * (size_t)nnr * nfreq, cudaMemcpyDeviceToHost, st));
    SGW_CUDA(cudaStreamSynchronize(st));
    memset(drhoscf, 0, sizeof(sgw_cplx) * (size_t)nnrx * nfreq);
    for (int f = 0; f < nfreq; ++f)
      for (long i = 0; i < nnr; ++i) drhoscf[(size_t)f * nnrx + compact_index(i)] = hbuf[(size_t)f * nnr + i];
    return SGW_OK;
  };
  if (pad) {
    hbuf.resize((size_t)nnr);
    for (long i = 0; i < nnr; ++i) hbuf[i] = dvbarein[compact_index(i)];
    SGW_CUDA(cudaMemcpyAsync(d_nat, hbuf.data(), sizeof(cplx) * nnr, cudaMemcpyHostToDevice, st));
    SGW_CUDA(cudaStreamSynchronize(st));
  } else {
    SGW_CUDA(cudaMemcpyAsync(d_nat, dvbarein, sizeof(cplx) * nnr, cudaMemcpyHostToDevice, st));
  }
  {
    dim3 gr((unsigned)((nnr + 255) / 256), 1);
    k_nat2perm<<<gr, 256, 0, st>>>(g, 1, d_nat, d_field);
    SGW_LAUNCH_CHECK();
  }
  double s2 = 0.0;                                                                         // solve_linter.f90:532
  for (long i = 0; i < nnrx; ++i) s2 += dvbarein[i].re * dvbarein[i].re + dvbarein[i].im * dvbarein[i].im;
  const double meandvb = std::sqrt(s2) / (double)nnrx;
  if (num_iter > 1) {
    // self-consistent branch: drhoscf = dvscfin (solve_linter.f90:610)
    cplx *d_dvin = nullptr;
    SGW_CHECK(ws(ctx, "it_dvin", (size_t)nnr * nfreq, &d_dvin));
    int ierr_it = 0, iters = 0;
    SGW_CHECK(solve_linter_iter_core(ctx, cfg, num_iter, d_field, meandvb, fl, *rho, d_dvin, &ierr_it, &iters));
    ctx->last_scf_iter = iters;
    dim3 gr((unsigned)((nnr + 255) / 256), nfreq);
    k_perm2nat<<<gr, 256, 0, st>>>(g, nfreq, d_dvin, d_nat, 1.0);
    SGW_LAUNCH_CHECK();
    SGW_CHECK(download(d_nat));
    *ierr_out = ierr_it;
    end_call(ctx);
    return SGW_OK;
  }
  SGW_CHECK(ws(ctx, "co_drhoG", (size_t)nfreq * rho->npw, &d_drhoG));
  int ierr_any = 0;
  SGW_CHECK(drho_block(ctx, cfg, 1, d_field, fl, *rho, d_drhoG, &ierr_any));
  double *d_fac = nullptr;
  SGW_CHECK(hartree_factor(ctx, *rho, rho->npw, "co_fac", &d_fac));
  int zero_pos = -1;
  if (meandvb < 1e-10)
    for (int pos = 0; pos < rho->npw; ++pos) if (rho->perm[pos] == 0) zero_pos = pos;       // :544-550 (G = 0 is ig = 1)
  {
    dim3 gr((rho->npw + 255) / 256, nfreq);
    k_hartree<<<gr, 256, 0, st>>>(rho->npw, nfreq, d_fac, d_drhoG, zero_pos);
    SGW_LAUNCH_CHECK();
  }
  // back to real space (:556 dv_of_drho ends with invfft) and out = -dvscfout (:598, sign already applied)
  SGW_CHECK(ws(ctx, "sl_Tr", (size_t)nfreq * ctx->nr3 * rho->ncol, &Tr));
  SGW_CHECK(fft_zpass_g2r(ctx, *rho, nfreq, d_drhoG, rho->npw, Tr, nullptr));
  cplx *d_R = nullptr;
  SGW_CHECK(ws(ctx, "sl_R", (size_t)nnr * nfreq, &d_R));
  SGW_CHECK(fft_plane(ctx, PLANE_TO_R, rho, nullptr, nfreq, Tr, nullptr, nullptr, 1, d_R, nullptr));
  {
    dim3 gr((unsigned)((nnr + 255) / 256), nfreq);
    k_perm2nat<<<gr, 256, 0, st>>>(g, nfreq, d_R, d_nat, 1.0);
    SGW_LAUNCH_CHECK();
  }
  SGW_CHECK(download(d_nat));
  *ierr_out = ierr_any;
  end_call(ctx);
  return SGW_OK;
}

int sgw_coulomb(sgw_ctx *ctx, const sgw_solver_cfg *cfg, int igstart, int ngc, int ntask, const int32_t *ig_unique,
                int nfs, const sgw_cplx *fiu, sgw_cplx *scrcoul, int32_t *ierr_out) {
  if (!ctx) return SGW_E_ARG;
  cudaSetDevice(ctx->device);
  SGW_ARG(cfg && ig_unique && fiu && scrcoul && ierr_out, "null argument");
  SGW_ARG(cfg->npriority >= 1 && cfg->npriority <= 4, "priority of the solvers not specified");
  SGW_ARG(igstart >= 1 && ntask >= 0 && nfs > 0, "bad task range");
  SGW_CHECK(check_pipeline_state(ctx));
  SGW_ARG(ngc >= 1 && ngc <= ctx->ngm, "num_g_corr outside 1..ngm");
  begin_call(ctx);
  memset(scrcoul, 0, sizeof(sgw_cplx) * (size_t)ngc * nfs * ntask);                        // coulomb.f90:98
  *ierr_out = 0;
  const FreqList fl = make_omega(nfs, fiu);
  Sphere *rho = nullptr;
  SGW_CHECK(get_rho_sphere(ctx, ngc, &rho));
  // tasks that are not skipped by the |q+G|^2 < 1e-8 rule (:126)
  std::vector<int> task_indx, task_ig;
  for (int indx = 0; indx < ntask; ++indx) {
    const int ig = ig_unique[igstart - 1 + indx];
    SGW_ARG(ig >= 1 && ig <= ctx->ngm, "ig_unique entry outside 1..ngm");
    const double q0 = ctx->g[3 * (ig - 1)] + ctx->xq[0], q1 = ctx->g[3 * (ig - 1) + 1] + ctx->xq[1],
                 q2 = ctx->g[3 * (ig - 1) + 2] + ctx->xq[2];
    if (q0 * q0 + q1 * q1 + q2 * q2 < 1e-8) continue;
    task_indx.push_back(indx);
    task_ig.push_back(ig);
  }
  const int nt = (int)task_ig.size();
  if (nt == 0) { end_call(ctx); return SGW_OK; }
  double *d_fac = nullptr;
  SGW_CHECK(hartree_factor(ctx, *rho, ngc, "co_fac", &d_fac));
  const long nnr = (long)ctx->nr1 * ctx->nr2 * ctx->nr3;
  if (!ctx->solve_direct) {
    // coulomb.f90:104-110 with solve_direct = .FALSE.: every perturbation runs the self-consistent solve_linter with
    // num_iter = niter_gw and scrcoul(igp, iw, indx) = fwfft(dV_scf)(G_igp) -- no delta on the diagonal (:153)
    SGW_ARG(ctx->mix_niter > 1, "self-consistent coulomb needs sgw_set_mixing with niter_gw > 1 (coulomb.f90:108)");
    Sphere *rho_full = nullptr;
    SGW_CHECK(get_rho_sphere(ctx, ctx->ngm, &rho_full));
    cudaStream_t st = ctx->stream;
    GridDev g = grid_dev(ctx);
    int *d_mill = nullptr, *d_ig0 = nullptr;
    cplx *d_field = nullptr, *d_dvin = nullptr, *Tg = nullptr, *d_dvG = nullptr, *d_scr = nullptr;
    SGW_CHECK(ws(ctx, "co_mill", (size_t)3, &d_mill));
    SGW_CHECK(ws(ctx, "co_ig0", (size_t)1, &d_ig0));
    SGW_CHECK(ws(ctx, "co_field", (size_t)nnr, &d_field));
    SGW_CHECK(ws(ctx, "it_dvin", (size_t)nnr * nfs, &d_dvin));
    SGW_CHECK(ws(ctx, "it_Tg", (size_t)nfs * ctx->nr3 * rho->ncol, &Tg));
    SGW_CHECK(ws(ctx, "it_dvG", (size_t)nfs * ngc, &d_dvG));
    SGW_CHECK(ws(ctx, "co_scr", (size_t)nfs * ngc, &d_scr));
    ctx->rho_last_coarse = false;
    int ierr_any = 0, iters_max = 0;
    std::vector<cplx> hscr((size_t)ngc * nfs);
    ZEpilogue epi;
    epi.mode = 0; epi.g2kin = nullptr; epi.psi = nullptr; epi.sigma = nullptr; epi.sigma_stride = 0; epi.keep_out = 0;
    for (int t = 0; t < nt && !ierr_any; ++t) {
      const long idx = (long)ctx->nl[task_ig[t] - 1] - 1;
      const int mill[3] = {(int)(idx % ctx->nr1), (int)((idx / ctx->nr1) % ctx->nr2), (int)(idx / ((long)ctx->nr1 * ctx->nr2))};
      const int ig0 = task_ig[t] - 1;
      SGW_CUDA(cudaMemcpyAsync(d_mill, mill, sizeof(mill), cudaMemcpyHostToDevice, st));
      SGW_CUDA(cudaMemcpyAsync(d_ig0, &ig0, sizeof(int), cudaMemcpyHostToDevice, st));
      SGW_CUDA(cudaStreamSynchronize(st));
      {
        dim3 gr((unsigned)((nnr + 255) / 256), 1);
        k_delta_field<<<gr, 256, 0, st>>>(g, 1, d_mill, d_field);                          // coulomb.f90:129-134
        SGW_LAUNCH_CHECK();
      }
      int ierr_it = 0, iters = 0;
      // |dvbare(r)| = 1 everywhere for a delta perturbation: meandvb = 1/sqrt(nnr) (solve_linter.f90:532)
      SGW_CHECK(solve_linter_iter_core(ctx, cfg, ctx->mix_niter, d_field, 1.0 / std::sqrt((double)nnr), fl, *rho_full, d_dvin,
                                       &ierr_it, &iters));
      iters_max = std::max(iters_max, iters);
      if (ierr_it) { ierr_any = ierr_it; break; }
      // fwfft('Rho') of dV_scf and the first ngc components (:146-151)
      SGW_CHECK(fft_plane(ctx, PLANE_FROM_R, nullptr, rho, nfs, nullptr, Tg, nullptr, 1, d_dvin, nullptr));
      SGW_CUDA(cudaMemsetAsync(d_dvG, 0, sizeof(cplx) * (size_t)nfs * ngc, st));
      SGW_CHECK(fft_zpass_r2g(ctx, *rho, nfs, Tg, d_dvG, ngc, epi, nullptr));
      {
        dim3 gr((ngc + 127) / 128, nfs, 1);
        k_scr_extract<<<gr, 128, 0, st>>>(ngc, nfs, 1, rho->d_perm, d_fac, d_ig0, d_dvG, d_scr, 0);
        SGW_LAUNCH_CHECK();
      }
      SGW_CUDA(cudaMemcpyAsync(hscr.data(), d_scr, sizeof(cplx) * (size_t)ngc * nfs, cudaMemcpyDeviceToHost, st));
      SGW_CUDA(cudaStreamSynchronize(st));
      memcpy(scrcoul + (size_t)ngc * nfs * task_indx[t], hscr.data(), sizeof(cplx) * (size_t)ngc * nfs);
    }
    ctx->last_scf_iter = iters_max;
    *ierr_out = ierr_any;
    end_call(ctx);
    return SGW_OK;
  }
  int chunk = perturbation_chunk(ctx, cfg, fl.num_omega, nfs, *rho, nt);
  chunk = (nt + (nt + chunk - 1) / chunk - 1) / ((nt + chunk - 1) / chunk);      // equal-sized blocks: no short tail block
  bool coarse = false;
  SGW_CHECK(rho_grid_prepare(ctx, ngc, *rho, &coarse));
  ctx->rho_last_coarse = coarse;
  cudaStream_t st = ctx->stream;
  GridDev g = grid_dev(ctx);
  std::vector<cplx> hscr((size_t)ngc * nfs * chunk);
  int ierr_any = 0;
  for (int t0 = 0; t0 < nt; t0 += chunk) {
    const int np = std::min(chunk, nt - t0);
    std::vector<int> mill(3 * np), ig0(np);
    for (int p = 0; p < np; ++p) {
      const long idx = (long)ctx->nl[task_ig[t0 + p] - 1] - 1;
      mill[3 * p] = (int)(idx % ctx->nr1);
      mill[3 * p + 1] = (int)((idx / ctx->nr1) % ctx->nr2);
      mill[3 * p + 2] = (int)(idx / ((long)ctx->nr1 * ctx->nr2));
      ig0[p] = task_ig[t0 + p] - 1;
    }
    int *d_mill = nullptr, *d_ig0 = nullptr;
    cplx *d_field = nullptr, *d_drhoG = nullptr, *d_scr = nullptr;
    SGW_CHECK(ws(ctx, "co_mill", (size_t)3 * np, &d_mill));
    SGW_CHECK(ws(ctx, "co_ig0", (size_t)np, &d_ig0));
    SGW_CHECK(ws(ctx, "co_field", (size_t)nnr * np, &d_field));
    SGW_CHECK(ws(ctx, "co_drhoG", (size_t)np * nfs * ngc, &d_drhoG));
    SGW_CHECK(ws(ctx, "co_scr", (size_t)np * nfs * ngc, &d_scr));
    SGW_CUDA(cudaMemcpyAsync(d_mill, mill.data(), sizeof(int) * 3 * np, cudaMemcpyHostToDevice, st));
    SGW_CUDA(cudaMemcpyAsync(d_ig0, ig0.data(), sizeof(int) * np, cudaMemcpyHostToDevice, st));
    {
      dim3 gr((unsigned)((nnr + 255) / 256), np);
      k_delta_field<<<gr, 256, 0, st>>>(g, np, d_mill, d_field);                           // coulomb.f90:129-134
      SGW_LAUNCH_CHECK();
    }
    SGW_CHECK(drho_block(ctx, cfg, np, d_field, fl, *rho, d_drhoG, &ierr_any, coarse));    // :137
    {
      dim3 gr((ngc + 127) / 128, nfs, np);
      k_scr_extract<<<gr, 128, 0, st>>>(ngc, nfs, np, rho->d_perm, d_fac, d_ig0, d_drhoG, d_scr);   // :143-157
      SGW_LAUNCH_CHECK();
    }
    SGW_CUDA(cudaMemcpyAsync(hscr.data(), d_scr, sizeof(cplx) * (size_t)ngc * nfs * np, cudaMemcpyDeviceToHost, st));
    SGW_CUDA(cudaStreamSynchronize(st));
    for (int p = 0; p < np; ++p)
      memcpy(scrcoul + (size_t)ngc * nfs * task_indx[t0 + p], &hscr[(size_t)ngc * nfs * p], sizeof(cplx) * (size_t)ngc * nfs);
  }
  *ierr_out = ierr_any;
  end_call(ctx);
  return SGW_OK;
}

int sgw_get_rho_grid(const sgw_ctx *ctx, int *dims) {
  if (!ctx || !dims) return SGW_E_ARG;
  const bool on = ctx->rho_last_coarse && ctx->rho_grid_on && ctx->rho_grid_version == ctx->tables_version;
  dims[0] = on ? ctx->rho_grid.n1 : ctx->nr1;
  dims[1] = on ? ctx->rho_grid.n2 : ctx->nr2;
  dims[2] = on ? ctx->rho_grid.n3 : ctx->nr3;
  return on ? 1 : 0;
}

int sgw_coulomb_q0G0(sgw_ctx *ctx, const sgw_solver_cfg *cfg, int nfs, const sgw_cplx *fiu, sgw_cplx *eps_m,
                     int32_t *ierr_out) {
  if (ctx && ctx->lgauss) {
    ctx->err = "coulomb_q0G0 is the insulator treatment of the q -> 0 head (coulomb_q0G0.f90); not defined for lgauss";
    return SGW_E_UNSUPPORTED;
  }
  // coulomb_q0G0.f90:31-158: the G = G' = 0 element at the (shifted) q currently installed
  const int32_t one = 1;
  return sgw_coulomb(ctx, cfg, 1, 1, 1, &one, nfs, fiu, eps_m, ierr_out);
}

int sgw_unfold_w(sgw_ctx *ctx, int ngc, int nfs, int ngmunique, const int32_t *ig_unique, const sgw_cplx *scrcoul_in,
                 sgw_cplx *scrcoul_out) {
  if (!ctx) return SGW_E_ARG;
  cudaSetDevice(ctx->device);
  SGW_ARG(ngc > 0 && nfs > 0 && ngmunique > 0 && ig_unique && scrcoul_in && scrcoul_out, "bad argument");
  for (int i = 0; i < ngmunique; ++i) SGW_ARG(ig_unique[i] >= 1 && ig_unique[i] <= ngc, "ig_unique outside 1..num_g_corr");
  begin_call(ctx);
  cudaStream_t st = ctx->stream;
  cplx *d_in = nullptr, *d_out = nullptr;
  int *d_iu = nullptr;
  const size_t nin = (size_t)ngc * nfs * ngmunique, nout = (size_t)ngc * ngc * nfs;
  SGW_CHECK(ws(ctx, "uf_in", nin, &d_in));
  SGW_CHECK(ws(ctx, "uf_out", nout, &d_out));
  SGW_CHECK(ws(ctx, "uf_iu", (size_t)ngmunique, &d_iu));
  SGW_CHECK(h2d_large(ctx, d_in, scrcoul_in, sizeof(cplx) * nin));
  SGW_CHECK(h2d_large(ctx, d_out, scrcoul_out, sizeof(cplx) * nout));                                 // INTENT(INOUT)-like
  SGW_CUDA(cudaMemcpyAsync(d_iu, ig_unique, sizeof(int) * ngmunique, cudaMemcpyHostToDevice, st));
  dim3 gr((ngc + 127) / 128, nfs, ngmunique);
  k_unfold<<<gr, 128, 0, st>>>(ngc, nfs, ngmunique, d_iu, d_in, d_out);
  SGW_LAUNCH_CHECK();
  SGW_CHECK(d2h_large(ctx, scrcoul_out, d_out, sizeof(cplx) * nout));
  end_call(ctx);
  return SGW_OK;
}

int sgw_unfold_w_symm(sgw_ctx *ctx, int ngc, int nfs, int ngmunique, const int32_t *ig_unique, int nsym, const int32_t *sym_ig,
                      const int32_t *sym_friend, const int32_t *gmapsym, const sgw_cplx *eigv, const int32_t *invs,
                      const sgw_cplx *scrcoul_in, sgw_cplx *scrcoul_out) {
  if (!ctx) return SGW_E_ARG;
  cudaSetDevice(ctx->device);
  SGW_ARG(ngc > 0 && nfs > 0 && ngmunique > 0 && nsym > 0 && ig_unique && sym_ig && sym_friend && gmapsym && eigv && invs && scrcoul_in &&
              scrcoul_out, "bad argument");
  std::vector<int> uniq(ngc, 0);
  for (int i = 0; i < ngmunique; ++i) {
    SGW_ARG(ig_unique[i] >= 1 && ig_unique[i] <= ngc, "ig_unique outside 1..num_g_corr");
    uniq[ig_unique[i] - 1] = 1;
  }
  for (int i = 0; i < nsym; ++i) SGW_ARG(invs[i] >= 1 && invs[i] <= nsym, "invs outside 1..nsym");
  for (int ig = 0; ig < ngc; ++ig) {
    if (uniq[ig]) continue;
    SGW_ARG(sym_ig[ig] >= 1 && sym_ig[ig] <= nsym, "sym_ig of a non-unique G outside 1..nsym");
    SGW_ARG(sym_friend[ig] >= 1 && sym_friend[ig] <= ngc && uniq[sym_friend[ig] - 1], "sym_friend of a non-unique G is not a unique G");
    const int ism1 = invs[sym_ig[ig] - 1] - 1;
    for (int igp = 0; igp < ngc; ++igp) {
      const int c = gmapsym[igp + (size_t)ngc * ism1];
      SGW_ARG(c >= 1 && c <= ngc, "gmapsym maps a G vector outside the correlation list (not closed under the small group of q)");
    }
  }
  begin_call(ctx);
  cudaStream_t st = ctx->stream;
  cplx *d_in = nullptr, *d_out = nullptr, *d_eig = nullptr;
  int *d_iu = nullptr, *d_uq = nullptr, *d_sig = nullptr, *d_sfr = nullptr, *d_gm = nullptr, *d_inv = nullptr;
  const size_t nin = (size_t)ngc * nfs * ngmunique, nout = (size_t)ngc * ngc * nfs;
  SGW_CHECK(ws(ctx, "uf_in", nin, &d_in));
  SGW_CHECK(ws(ctx, "uf_out", nout, &d_out));
  SGW_CHECK(ws(ctx, "uf_iu", (size_t)ngmunique, &d_iu));
  SGW_CHECK(ws(ctx, "uf_uq", (size_t)ngc, &d_uq));
  SGW_CHECK(ws(ctx, "uf_sig", (size_t)ngc, &d_sig));
  SGW_CHECK(ws(ctx, "uf_sfr", (size_t)ngc, &d_sfr));
  SGW_CHECK(ws(ctx, "uf_gm", (size_t)ngc * nsym, &d_gm));
  SGW_CHECK(ws(ctx, "uf_eig", (size_t)ngc * nsym, &d_eig));
  SGW_CHECK(ws(ctx, "uf_inv", (size_t)nsym, &d_inv));
  SGW_CUDA(cudaMemcpyAsync(d_in, scrcoul_in, sizeof(cplx) * nin, cudaMemcpyHostToDevice, st));
  SGW_CUDA(cudaMemsetAsync(d_out, 0, sizeof(cplx) * nout, st));
  SGW_CUDA(cudaMemcpyAsync(d_iu, ig_unique, sizeof(int) * ngmunique, cudaMemcpyHostToDevice, st));
  SGW_CUDA(cudaMemcpyAsync(d_uq, uniq.data(), sizeof(int) * ngc, cudaMemcpyHostToDevice, st));
  SGW_CUDA(cudaMemcpyAsync(d_sig, sym_ig, sizeof(int) * ngc, cudaMemcpyHostToDevice, st));
  SGW_CUDA(cudaMemcpyAsync(d_sfr, sym_friend, sizeof(int) * ngc, cudaMemcpyHostToDevice, st));
  SGW_CUDA(cudaMemcpyAsync(d_gm, gmapsym, sizeof(int) * (size_t)ngc * nsym, cudaMemcpyHostToDevice, st));
  SGW_CUDA(cudaMemcpyAsync(d_eig, eigv, sizeof(cplx) * (size_t)ngc * nsym, cudaMemcpyHostToDevice, st));
  SGW_CUDA(cudaMemcpyAsync(d_inv, invs, sizeof(int) * nsym, cudaMemcpyHostToDevice, st));
  {
    dim3 gr((ngc + 127) / 128, nfs, ngmunique);
    k_unfold<<<gr, 128, 0, st>>>(ngc, nfs, ngmunique, d_iu, d_in, d_out);
    SGW_LAUNCH_CHECK();
  }
  for (int g0 = 0; g0 < ngc; g0 += 65535) {       // grid.z limit
    const int ng = std::min(65535, ngc - g0);
    dim3 gr((ngc + 127) / 128, nfs, ng);
    k_unfold_symm<<<gr, 128, 0, st>>>(ngc, nfs, d_uq + g0, d_sig + g0, d_sfr + g0, d_gm, d_eig, d_inv, d_out + g0);
    SGW_LAUNCH_CHECK();
  }
  SGW_CUDA(cudaMemcpyAsync(scrcoul_out, d_out, sizeof(cplx) * nout, cudaMemcpyDeviceToHost, st));
  SGW_CUDA(cudaStreamSynchronize(st));
  end_call(ctx);
  return SGW_OK;
}

int sgw_green_function(sgw_ctx *ctx, int slot, const sgw_solver_cfg *cfg, int ngc, const int32_t *map, int ngp,
                       const int32_t *fft_map, int nfreq, const sgw_cplx *omega, sgw_cplx *green, int32_t *ierr_out) {
  if (!ctx) return SGW_E_ARG;
  SGW_ARG(green != nullptr, "null argument");
  return sgw::green_function_core(ctx, slot, cfg, ngc, map, ngp, fft_map, nfreq, omega, green, ierr_out, nullptr);
}

}  // extern "C"

// green_function with the result left on the device (workspace "gr_green", layout green(ngc, ngp, nfreq)) for
// sgw_sigma_correlation; `green` may be null (no download).
int sgw::green_function_core(sgw_ctx *ctx, int slot, const sgw_solver_cfg *cfg, int ngc, const int32_t *map, int ngp,
                             const int32_t *fft_map, int nfreq, const sgw_cplx *omega, sgw_cplx *green, int32_t *ierr_out,
                             cplx **d_green_out) {
  if (!ctx) return SGW_E_ARG;
  cudaSetDevice(ctx->device);
  SGW_ARG(cfg && map && fft_map && omega && ierr_out, "null argument");
  SGW_ARG(cfg->npriority >= 1 && cfg->npriority <= 4, "priority of the solvers not specified");
  SGW_ARG(ngc > 0 && ngp > 0 && nfreq > 0, "bad sizes");
  if (slot < 0 || slot >= (int)ctx->slots.size() || !ctx->slots[slot].set) { ctx->err = "operator slot not set"; return SGW_E_STATE; }
  const KSlot &ks = ctx->slots[slot];
  const int n = ks.npwx, num_g = ks.npw;
  for (int k = 0; k < ngc; ++k) SGW_ARG(map[k] >= 0 && map[k] <= num_g, "map entry outside 0..num_g");
  for (int i = 0; i < ngp; ++i) SGW_ARG(fft_map[i] >= 1 && fft_map[i] <= ngc, "fft_map entry outside 1..num_g_corr");
  begin_call(ctx);
  if (green) memset(green, 0, sizeof(sgw_cplx) * (size_t)ngc * ngp * nfreq);                // green.f90:184
  *ierr_out = 0;
  cudaStream_t st = ctx->stream;
  std::vector<int> invperm(num_g);
  for (int p = 0; p < num_g; ++p) invperm[ks.sph.perm[p]] = p;
  std::vector<int> pos, col;            // right-hand sides that exist: G' inside the k sphere (:199-200 CYCLE otherwise)
  for (int i = 0; i < ngp; ++i) {
    const int ig = map[fft_map[i] - 1];                                                     // :199
    if (ig == 0) continue;
    pos.push_back(invperm[ig - 1]);
    col.push_back(i);
  }
  const int nlist = (int)pos.size();
  std::vector<cplx> msig(nfreq);
  for (int i = 0; i < nfreq; ++i) msig[i] = cmake(-omega[i].re, -omega[i].im);              // :207
  // RHS chunk that fits on the device
  size_t free_b = 0, total_b = 0, held = 0;
  cudaMemGetInfo(&free_b, &total_b);
  for (auto &kv : ctx->ws.bufs) held += kv.second.second;
  const size_t per = solver_bytes_per_rhs(ctx, ks, cfg->bicg_lmax, nfreq) + (size_t)nfreq * sizeof(cplx);
  const size_t gbytes = (size_t)ngc * ngp * nfreq * sizeof(cplx);
  double avail = 0.85 * (double)(free_b + held) - (double)gbytes - (double)((size_t)1 << 30);
  int chunk = avail > 0 ? (int)std::min<double>(avail / (double)per, 32768.0) : 1;
  chunk = std::max(1, std::min(chunk, nlist));
  cplx *d_green = nullptr, *d_b = nullptr, *d_sig = nullptr, *d_x = nullptr;
  int *d_map = nullptr, *d_inv = nullptr, *d_pos = nullptr, *d_col = nullptr, *d_ierr = nullptr;
  SGW_CHECK(ws(ctx, "gr_green", (size_t)ngc * ngp * nfreq, &d_green));
  SGW_CHECK(ws(ctx, "gr_map", (size_t)ngc, &d_map));
  SGW_CHECK(ws(ctx, "gr_inv", (size_t)num_g, &d_inv));
  SGW_CHECK(ws(ctx, "gr_pos", (size_t)nlist, &d_pos));
  SGW_CHECK(ws(ctx, "gr_col", (size_t)nlist, &d_col));
  SGW_CUDA(cudaMemsetAsync(d_green, 0, gbytes, st));
  SGW_CUDA(cudaMemcpyAsync(d_map, map, sizeof(int) * ngc, cudaMemcpyHostToDevice, st));
  SGW_CUDA(cudaMemcpyAsync(d_inv, invperm.data(), sizeof(int) * num_g, cudaMemcpyHostToDevice, st));
  SGW_CUDA(cudaMemcpyAsync(d_pos, pos.data(), sizeof(int) * nlist, cudaMemcpyHostToDevice, st));
  SGW_CUDA(cudaMemcpyAsync(d_col, col.data(), sizeof(int) * nlist, cudaMemcpyHostToDevice, st));
  SGW_CHECK(ws(ctx, "sv_b", (size_t)n * chunk, &d_b));
  SGW_CHECK(ws(ctx, "sv_sig", (size_t)nfreq * chunk, &d_sig));
  SGW_CHECK(ws(ctx, "sv_x", (size_t)n * nfreq * chunk, &d_x));
  SGW_CHECK(ws(ctx, "sv_ierr", (size_t)chunk, &d_ierr));
  std::vector<cplx> sig((size_t)nfreq * chunk);
  for (int r = 0; r < chunk; ++r) memcpy(&sig[(size_t)r * nfreq], msig.data(), sizeof(cplx) * nfreq);
  SGW_CUDA(cudaMemcpyAsync(d_sig, sig.data(), sizeof(cplx) * sig.size(), cudaMemcpyHostToDevice, st));
  SGW_CUDA(cudaStreamSynchronize(st));
  int ierr_any = 0;
  for (int r0 = 0; r0 < nlist; r0 += chunk) {                                               // :196
    const int nr = std::min(chunk, nlist - r0);
    dim3 gb((n + 255) / 256, nr);
    k_green_rhs<<<gb, 256, 0, st>>>(n, nr, d_pos + r0, d_b);
    SGW_LAUNCH_CHECK();
    SolveBatch sb;
    sb.slot = slot; sb.alpha_pv = 0.0;                                                      // green_operator :283
    sb.nrhs = nr; sb.nshift = nfreq; sb.n = n;
    sb.d_b = d_b; sb.ldb = n; sb.d_sigma = d_sig; sb.d_x = d_x; sb.d_ierr = d_ierr;
    cudaEventRecord(ctx->ev2, st);
    ctx->prof_z_class = PC_FFT_Z;
    SGW_CHECK(select_solver_batched(ctx, sb, cfg));
    ctx->prof_z_class = PC_RHO_PLANE;
    cudaEventRecord(ctx->ev3, st);
    std::vector<int> ie(nr);
    SGW_CUDA(cudaMemcpyAsync(ie.data(), d_ierr, sizeof(int) * nr, cudaMemcpyDeviceToHost, st));
    SGW_CUDA(cudaStreamSynchronize(st));
    for (int r = 0; r < nr; ++r) if (ie[r] != 0) ierr_any = ie[r];                           // :208
    float ms = 0.f;
    cudaEventElapsedTime(&ms, ctx->ev2, ctx->ev3);
    ctx->stats.ms_solver += ms;
    dim3 gs((ngc + 127) / 128, nfreq, nr);
    k_green_scatter<<<gs, 128, 0, st>>>(ngc, ngp, nfreq, num_g, n, d_map, d_inv, d_col + r0, d_x, d_green);
    SGW_LAUNCH_CHECK();
  }
  if (green) SGW_CUDA(cudaMemcpyAsync(green, d_green, gbytes, cudaMemcpyDeviceToHost, st));
  SGW_CUDA(cudaStreamSynchronize(st));
  if (d_green_out) *d_green_out = d_green;
  *ierr_out = ierr_any;
  end_call(ctx);
  return SGW_OK;
}
