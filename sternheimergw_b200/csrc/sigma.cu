// sigma.cu -- SURVEY.md section 8 rows f2 and f3 on the device:
//   f3  analytic continuation of W    algo/analytic/src/analytic.f90:50 (analytic_coeff), :211 (analytic_eval),
//                                     pade.f90 (pade_coeff / pade_eval), godby_needs.f90, freqbins.f90:243 (freqbins_symm),
//                                     phys/coul/src/coulpade.f90
//   f2  Sigma_c = G W                 phys/corr/src/sigma.f90:417 (sigma_prod), :528 (sigma_correlation),
//                                     data/fft/src/fft6.f90:84 (fwfft6), :231 (invfft6)
//
// Design.  The correlation box holds a few hundred points and <= ~100 G vectors (5^3..9^3 boxes, 15..59 G at the
// BASELINE configs), so the reference's 2 x nnr tiny 3-D FFTs per 6-D transform are replaced by sphere-pruned DFT matrices
// applied on the FP64 tensor path (DMMA):   f(r,r') = 1/Omega  Ec f(G,G') ET ,  f(G,G') = Omega/nnr^2  ET f(r,r') Ec
// with Ec(r,G) = exp(-iGr), ET(G,r) = exp(+iGr).  The reference evaluates, for every (omega_sigma, omega_green) pair,
// analytic_eval -> invfft6 -> product with G(r,r') -> fwfft6 -> accumulate.  Here one omega_sigma is ONE pass:
//   k_analytic_eval   W(G,G') for all omega_green at once
//   k_zgemm           X_b = Ec W_b / Omega                              (one GEMM over all omega_green)
//   k_gw_product      acc(r,r') = sum_b alpha w_b G_b(r,r') (X_b ET)(r,r')   -- W(r,r') is never written to memory: the
//                     second half of invfft6 is the main loop of the kernel and the product with G its epilogue
//   k_zgemm x 2       sigma(G,G') += Omega/nnr^2 ET acc Ec               (one fwfft6 per omega_sigma instead of one per pair)
// Algebraically identical to the reference (all steps are linear), re-associated.
#include "internal.cuh"

#include <string.h>

#include <algorithm>
#include <cmath>

using namespace sgw;

namespace sgw {

// ---------------------------------------------------------------- arithmetic of the continuation kernels
// The Pade recurrences amplify rounding differences, so these kernels use exactly the operation order of the oracle
// (numpy: Smith division with a reciprocal, products without FMA contraction); *_rn intrinsics are never fused by nvcc.
__device__ __forceinline__ cplx cmul_nf(cplx a, cplx b) {
  return cmake(__dsub_rn(__dmul_rn(a.x, b.x), __dmul_rn(a.y, b.y)), __dadd_rn(__dmul_rn(a.x, b.y), __dmul_rn(a.y, b.x)));
}
__device__ __forceinline__ cplx cdiv_nf(cplx a, cplx b) {
  if (fabs(b.x) >= fabs(b.y)) {
    if (b.x == 0.0 && b.y == 0.0) return cmake(a.x / fabs(b.x), a.y / fabs(b.y));
    const double rat = __ddiv_rn(b.y, b.x);
    const double scl = __ddiv_rn(1.0, __dadd_rn(b.x, __dmul_rn(b.y, rat)));
    return cmake(__dmul_rn(__dadd_rn(a.x, __dmul_rn(a.y, rat)), scl), __dmul_rn(__dsub_rn(a.y, __dmul_rn(a.x, rat)), scl));
  }
  const double rat = __ddiv_rn(b.x, b.y);
  const double scl = __ddiv_rn(1.0, __dadd_rn(b.y, __dmul_rn(b.x, rat)));
  return cmake(__dmul_rn(__dadd_rn(__dmul_rn(a.x, rat), a.y), scl), __dmul_rn(__dsub_rn(__dmul_rn(a.y, rat), a.x), scl));
}
__device__ __forceinline__ cplx protect(cplx x) {      // pade.f90: |x| <= eps24 -> eps24
  return hypot(x.x, x.y) > 1e-24 ? x : cmake(1e-24, 0.0);
}
__device__ __forceinline__ cplx csqrt_dev(cplx z) {    // principal branch, Re >= 0
  const double m = hypot(z.x, z.y);
  if (m == 0.0) return cmake(0.0, 0.0);
  if (z.x >= 0.0) {
    const double t = sqrt(0.5 * (m + z.x));
    return cmake(t, 0.5 * z.y / t);
  }
  const double t = sqrt(0.5 * (m - z.x));
  return cmake(0.5 * fabs(z.y) / t, z.y >= 0.0 ? t : -t);
}

// coulpade.f90:72-85: scrcoul_g(ig, igp, ifreq) *= factor(ig)
__global__ void k_coulpade(int ngc, long total, const double *__restrict__ factor, cplx *__restrict__ scr) {
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  scr[i] = cscale(factor[(int)(i % ngc)], scr[i]);
}

// freqbins_symm, array part (freqbins.f90:296-302): array(:,:,dst) = array(:,:,src) for the mirrored frequencies
__global__ void k_mirror(long npair, int nmirror, const int *__restrict__ src, const int *__restrict__ dst, cplx *__restrict__ scr) {
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  const int m = blockIdx.y;
  if (i >= npair || m >= nmirror) return;
  scr[i + npair * dst[m]] = scr[i + npair * src[m]];
}

// pade_coeff (pade.f90): one thread per (ig, igp); the g(p, :) row overwrites g(p-1, :) in place inside scrcoul_g itself,
// whose frequency stride is ngc^2 -> coalesced across the pairs.  a(p) = g(p, p) is final once row p is done.
__global__ void k_pade_coeff(long npair, int N, const cplx *__restrict__ z, cplx *__restrict__ scr) {
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= npair) return;
  cplx *g = scr + i;
  for (int f = 0; f < N; ++f) g[npair * f] = protect(g[npair * f]);
  for (int p = 1; p < N; ++p) {
    const cplx prev = g[npair * (p - 1)];
    const cplx zp = z[p - 1];
    for (int f = p; f < N; ++f) {
      const cplx gi = g[npair * f];
      const cplx tmp1 = cdiv_nf(prev, gi);
      const cplx tmp2 = cdiv_nf(gi, gi);
      g[npair * f] = protect(cdiv_nf(csub(tmp1, tmp2), csub(z[f], zp)));
    }
  }
}

// godby_needs_coeffs (godby_needs.f90:34-90)
__global__ void k_gn_coeff(long npair, double omega_p, cplx *__restrict__ scr) {
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= npair) return;
  const cplx c1 = scr[i], c2 = scr[i + npair];
  cplx work = csub(c1, c2);
  bool set_zero = fabs(work.x) < 1e-8;
  if (!set_zero) {
    work = cdiv_nf(c2, work);
    set_zero = work.x < 1e-8;
  }
  if (!set_zero) {
    const cplx b = cscale(omega_p, csqrt_dev(work));
    scr[i + npair] = b;
    scr[i] = cmul(cscale(0.5, c1), b);
  } else {
    scr[i] = cmake(0.0, 0.0);
    scr[i + npair] = cmake(0.0, 0.0);
  }
}

// analytic_eval (analytic.f90:266-300) for nout frequencies: out(ig, igp, io) = model(coeff(gmapsym(ig), gmapsym(igp), :), w_io)
__global__ void k_analytic_eval(int model, int ngc, int N, const int *__restrict__ gmapsym, const cplx *__restrict__ z,
                                const cplx *__restrict__ coeff, const cplx *__restrict__ wout, cplx *__restrict__ out) {
  const int ig = blockIdx.x * blockDim.x + threadIdx.x;
  const int igp = blockIdx.y, io = blockIdx.z;
  if (ig >= ngc) return;
  const long npair = (long)ngc * ngc;
  const cplx *c = coeff + (gmapsym[ig] - 1) + (long)ngc * (gmapsym[igp] - 1);
  const cplx w = wout[io];
  cplx res;
  if (model == SGW_PADE_APPROX) {                      // pade_eval (pade.f90)
    cplx am2 = cmake(0.0, 0.0), am1 = c[0], bm2 = cmake(1.0, 0.0), bm1 = cmake(1.0, 0.0);
    for (int f = 1; f < N; ++f) {
      const cplx fac = cmul_nf(csub(w, z[f - 1]), c[npair * f]);
      const cplx pa = cmul_nf(fac, am2), pb = cmul_nf(fac, bm2);
      const cplx a = cmake(__dadd_rn(am1.x, pa.x), __dadd_rn(am1.y, pa.y));
      const cplx b = cmake(__dadd_rn(bm1.x, pb.x), __dadd_rn(bm1.y, pb.y));
      am2 = am1; am1 = a; bm2 = bm1; bm1 = b;
    }
    res = cdiv_nf(am1, bm1);
  } else if (model == SGW_PADE_ROBUST) {               // pade_eval_robust (pade_robust.f90:38-88): Horner, numerator / denominator
    const int dn = (int)rint(hypot(c[0].x, c[0].y)), dd = (int)rint(hypot(c[npair].x, c[npair].y));
    cplx nu = cmake(0.0, 0.0), de = cmake(0.0, 0.0);
    for (int i = 2 + dn; i >= 2; --i) nu = cadd(c[npair * i], cmul_nf(nu, w));
    for (int i = 3 + dn + dd; i >= 3 + dn; --i) de = cadd(c[npair * i], cmul_nf(de, w));
    res = cdiv_nf(nu, de);
  } else if (model == SGW_AAA_POLE) {                  // aaa_pole_eval (analytic.f90:379-400)
    const int half = N / 2;
    int npl = 0;
    for (int j = 0; j < N - half; ++j) { const cplx rj = c[npair * (half + j)]; npl += (rj.x != 0.0 || rj.y != 0.0); }
    res = cmake(0.0, 0.0);
    for (int j = 0; j < npl; ++j) res = cadd(res, cdiv_nf(c[npair * (half + j)], csub(w, c[npair * j])));
  } else if (model == SGW_AAA_APPROX) {                // aaa_approx_eval (analytic.f90:312-343) + aaa_evaluate (aaa.f90)
    const int mmax = N / 3;
    int mm = 0;
    for (int j = 0; j < mmax; ++j) { const cplx wj = c[npair * (2 * mmax + j)]; mm += hypot(wj.x, wj.y) > 1e-12; }
    cplx num = cmake(0.0, 0.0), den = cmake(0.0, 0.0);
    double dmin = 1e300;
    int jclose = 0;
    for (int j = 0; j < mm; ++j) {
      const cplx pj = c[npair * j], vj = c[npair * (mmax + j)], wj = c[npair * (2 * mmax + j)];
      cplx d = csub(w, pj);
      const double dist = hypot(d.x, d.y);
      if (dist < dmin) { dmin = dist; jclose = j; }          // MINLOC: first minimum
      if (dist <= 1e-14) d = cmake(1e-14, 0.0);
      const cplx cm = cdiv_nf(cmake(1.0, 0.0), d);
      num = cadd(num, cmul(cm, cmul(wj, vj)));
      den = cadd(den, cmul(cm, wj));
    }
    res = cdiv_nf(num, den);
    if (mm > 0 && dmin < 1e-14) res = c[npair * (mmax + jclose)];
  } else {                                             // godby_needs_model (godby_needs.f90:108-135)
    const cplx c1 = c[0], c2 = c[npair];
    if (hypot(c1.x, c1.y) > 1e-8) {
      const cplx one = cmake(1.0, 0.0);
      res = cmul(c1, cadd(cdiv_nf(one, cadd(c2, w)), cdiv_nf(one, csub(c2, w))));
    } else {
      res = cmake(0.0, 0.0);
    }
  }
  out[ig + (long)ngc * (igp + (long)ngc * io)] = res;
}

// ---------------------------------------------------------------- AAA ('aaa', vendor/analytic/src/aaa.f90 via analytic.f90:150-172)
__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// One-sided (Hestenes) Jacobi on the columns of the R x m matrix A (column-major, shared memory), executed by one warp:
// on return the columns are mutually orthogonal (their norms are the singular values, zero columns span the null space) and
// V (m x m, must hold the identity on entry) accumulates the rotations: A_in V = A_out.  Returns 1 if 60 sweeps were not enough.
__device__ int warp_jacobi(cplx *A, int R, int m, cplx *V, int lane) {
  // columns that have become numerically zero (norm below 1e-14 of the largest column: the null space of a rank-deficient or
  // wide matrix) are left alone: their inner products with the other columns are rounding noise of the same relative size as
  // their norm, so the relative criterion below would rotate them for ever
  double s2max = 0.0;
  for (int c = 0; c < m; ++c) {
    double nn = 0.0;
    for (int r = lane; r < R; r += 32) { const cplx x = A[r + (long)R * c]; nn += x.x * x.x + x.y * x.y; }
    s2max = fmax(s2max, warp_sum_d(nn));
  }
  const double floor2 = 1e-28 * s2max;
  for (int sweep = 0; sweep < 60; ++sweep) {
    int rotated = 0;
    for (int p = 0; p < m - 1; ++p)
      for (int q = p + 1; q < m; ++q) {
        cplx *ap = A + (long)R * p, *aq = A + (long)R * q;
        double al = 0.0, be = 0.0, gr = 0.0, gi = 0.0;
        for (int r = lane; r < R; r += 32) {
          const cplx x = ap[r], y = aq[r];
          al += x.x * x.x + x.y * x.y;
          be += y.x * y.x + y.y * y.y;
          gr += x.x * y.x + x.y * y.y;          // conj(x) * y
          gi += x.x * y.y - x.y * y.x;
        }
        al = warp_sum_d(al); be = warp_sum_d(be); gr = warp_sum_d(gr); gi = warp_sum_d(gi);
        const double g2 = gr * gr + gi * gi;
        if (!(g2 > 1e-30 * al * be) || g2 == 0.0 || al <= floor2 || be <= floor2) continue;
        rotated = 1;
        const double gabs = sqrt(g2);
        const cplx phc = cmake(gr / gabs, -gi / gabs);            // conj(gamma / |gamma|)
        const double zeta = (be - al) / (2.0 * gabs);
        const double t = (zeta >= 0.0 ? 1.0 : -1.0) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
        const double c = 1.0 / sqrt(1.0 + t * t), sn = c * t;
        for (int r = lane; r < R; r += 32) {
          const cplx x = ap[r], y = cmul(aq[r], phc);
          ap[r] = cmake(c * x.x - sn * y.x, c * x.y - sn * y.y);
          aq[r] = cmake(sn * x.x + c * y.x, sn * x.y + c * y.y);
        }
        cplx *vp = V + (long)m * p, *vq = V + (long)m * q;
        for (int r = lane; r < m; r += 32) {
          const cplx x = vp[r], y = cmul(vq[r], phc);
          vp[r] = cmake(c * x.x - sn * y.x, c * x.y - sn * y.y);
          vq[r] = cmake(sn * x.x + c * y.x, sn * x.y + c * y.y);
        }
        __syncwarp();
      }
    if (!rotated) return 0;
  }
  return 1;
}

// Greedy AAA fit of one (G, G') pair per warp (aaa_generate): support point = first maximum of |f - fit|, weights = right
// singular vector of the smallest singular value of the Loewner submatrix (rows: non-support points, columns: support
// points).  The reference calls LAPACK's SVD; here the R x m submatrix (R = N - m >= 2 m) lives in shared memory and a
// one-sided (Hestenes) Jacobi iteration orthogonalises its columns while accumulating V -- it delivers the small singular
// vectors to high relative accuracy, which is what the barycentric weights need.  The weights are defined up to a
// common phase (the approximant does not depend on it), so coefficient arrays are not comparable entry by entry with a
// LAPACK-based fit; positions, values and the evaluated approximant are.
// Shared memory: zz, ff, fit [N] | A [N x mmax] | V [mmax x mmax] | w [mmax] | sup [N] (int) | supidx [mmax] | rowidx [N]
// pole_mode ('aaa pole', analytic.f90:345-377 + aaa.f90 aaa_pole_residual): after the fit, the m - 1 poles of the barycentric
// form -- the roots of d(x) = sum_j w_j / (x - z_j), which the reference obtains as the finite eigenvalues of an arrowhead
// pencil with ZGGEV -- are found by Aberth-Ehrlich iteration on p(x) = d(x) prod_j (x - z_j) (all roots at once, one lane per
// root, p'/p = d'/d + sum_j 1/(x - z_j)); residues by the reference's four-point average around each pole; poles with
// |residue| > thres are stored as [pole | residue] in halves of N / 2.
__global__ void __launch_bounds__(32) k_aaa_coeff(long npair, int N, int mmax, double thres, const cplx *__restrict__ z,
                                                  cplx *__restrict__ scr, int *__restrict__ info, int pole_mode) {
  const long pair = blockIdx.x;
  if (pair >= npair) return;
  const int lane = threadIdx.x;
  extern __shared__ cplx aaa_sm[];
  cplx *zz = aaa_sm, *ff = zz + N, *fit = ff + N;
  cplx *A = fit + N;
  cplx *V = A + (long)N * mmax;
  cplx *w = V + (long)mmax * mmax;
  int *sup = (int *)(w + mmax), *supidx = sup + N, *rowidx = supidx + mmax;
  for (int i = lane; i < N; i += 32) { zz[i] = z[i]; ff[i] = scr[pair + npair * i]; sup[i] = 0; }
  __syncwarp();
  // average and absolute threshold (setup_work_type)
  double ffmax = 0.0;
  for (int i = lane; i < N; i += 32) ffmax = ffmax > hypot(ff[i].x, ff[i].y) ? ffmax : hypot(ff[i].x, ff[i].y);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) { const double t = __shfl_xor_sync(0xffffffffu, ffmax, o); ffmax = ffmax > t ? ffmax : t; }
  const double thr = thres * ffmax;
  cplx avg = cmake(0.0, 0.0);
  for (int i = 0; i < N; ++i) avg = cadd(avg, ff[i]);
  avg = cmake(avg.x / N, avg.y / N);
  for (int i = lane; i < N; i += 32) fit[i] = avg;
  __syncwarp();
  int m = 0, bad = 0;
  if (pole_mode == 2) {
    // aaa_pole_residual on a GIVEN approximant (aaa.f90:93): the input holds [position | value | weight] in blocks of mmax
    // (the 'aaa' coefficient layout, analytic.f90:160-167); no fit
    for (int c = 0; c < mmax; ++c) {
      const cplx wc = ff[2 * mmax + c];
      if (hypot(wc.x, wc.y) > 1e-12) m = c + 1;                   // analytic.f90:330 counts the weights above eps12
    }
    __syncwarp();
    cplx pz = cmake(0, 0), pv = cmake(0, 0), pw = cmake(0, 0);
    if (lane < m) { pz = ff[lane]; pv = ff[mmax + lane]; pw = ff[2 * mmax + lane]; }     // mmax <= 32 in this mode (host checks)
    __syncwarp();
    if (lane < m) { zz[lane] = pz; ff[lane] = pv; w[lane] = pw; supidx[lane] = lane; }
    __syncwarp();
  }
  for (; pole_mode != 2;) {
    // ---- update_support_point: first maximum of |ff - fit|
    double best = -1.0;
    int bidx = 0x7fffffff;
    for (int i = lane; i < N; i += 32) {
      const cplx d = csub(ff[i], fit[i]);
      const double a = hypot(d.x, d.y);
      if (a > best) { best = a; bidx = i; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const double ob = __shfl_xor_sync(0xffffffffu, best, o);
      const int oi = __shfl_xor_sync(0xffffffffu, bidx, o);
      if (ob > best || (ob == best && oi < bidx)) { best = ob; bidx = oi; }
    }
    if (lane == 0) { fit[bidx] = ff[bidx]; sup[bidx] = 1; }
    ++m;
    __syncwarp();
    if (lane == 0) {                         // support points and the remaining rows in mesh order (PACK)
      int a = 0, b = 0;
      for (int i = 0; i < N; ++i) { if (sup[i]) supidx[a++] = i; else rowidx[b++] = i; }
    }
    __syncwarp();
    const int R = N - m;
    // ---- Loewner submatrix (construct_Loewner_matrix + extract_submatrix) and V = 1
    for (int c = 0; c < m; ++c) {
      const int jc = supidx[c];
      for (int r = lane; r < R; r += 32) {
        const int ir = rowidx[r];
        A[r + (long)R * c] = cdiv_nf(csub(ff[ir], ff[jc]), csub(zz[ir], zz[jc]));
      }
    }
    for (int i = lane; i < m * m; i += 32) V[i] = (i % m == i / m) ? cmake(1.0, 0.0) : cmake(0.0, 0.0);
    __syncwarp();
    // ---- one-sided Jacobi on the columns of A
    if (warp_jacobi(A, R, m, V, lane)) bad = 1;
    // ---- weights = column of V that belongs to the smallest column norm
    int jmin = 0;
    double smin = 1e300;
    for (int c = 0; c < m; ++c) {
      double nn = 0.0;
      for (int r = lane; r < R; r += 32) { const cplx x = A[r + (long)R * c]; nn += x.x * x.x + x.y * x.y; }
      nn = warp_sum_d(nn);
      if (nn < smin) { smin = nn; jmin = c; }
    }
    for (int r = lane; r < m; r += 32) w[r] = V[r + (long)m * jmin];
    __syncwarp();
    // ---- update_fit: barycentric value at the non-support points (Cauchy-matrix products, support order)
    int notconv = 0;
    for (int r = lane; r < R; r += 32) {
      const int ir = rowidx[r];
      cplx num = cmake(0.0, 0.0), den = cmake(0.0, 0.0);
      for (int c = 0; c < m; ++c) {
        const int jc = supidx[c];
        cplx d = csub(zz[ir], zz[jc]);
        if (hypot(d.x, d.y) <= 1e-14) d = cmake(1e-14, 0.0);
        const cplx cm = cdiv_nf(cmake(1.0, 0.0), d);
        num = cadd(num, cmul(cm, cmul(w[c], ff[jc])));
        den = cadd(den, cmul(cm, w[c]));
      }
      const cplx fv = cdiv_nf(num, den);
      fit[ir] = fv;
      const cplx e = csub(fv, ff[ir]);
      if (!(hypot(e.x, e.y) <= thr)) notconv = 1;
    }
    notconv = __any_sync(0xffffffffu, notconv);
    __syncwarp();
    if (!notconv) break;
    if (m >= mmax) {
      if (pole_mode) bad = 2;        // the reference would go on up to N support points; the device fit stops at mmax = N / 2
      break;
    }
  }
  if (pole_mode) {
    cplx *xk = A, *xn = A + mmax, *rs = A + 2 * mmax;      // the Loewner workspace is free now (N * mmax >= 3 * mmax entries)
    const int npole = m - 1;
    cplx cen = cmake(0.0, 0.0);
    for (int c = 0; c < m; ++c) cen = cadd(cen, zz[supidx[c]]);
    cen = cmake(cen.x / m, cen.y / m);
    double rad = 0.0;
    for (int c = 0; c < m; ++c) { const cplx d = csub(zz[supidx[c]], cen); rad = fmax(rad, hypot(d.x, d.y)); }
    if (rad == 0.0) rad = 1.0;
    for (int k = lane; k < npole; k += 32) {
      double sn, cs;
      sincos(6.283185307179586 * k / (npole > 0 ? npole : 1) + 0.35, &sn, &cs);
      xk[k] = cmake(cen.x + 0.7 * rad * cs, cen.y + 0.7 * rad * sn);
    }
    __syncwarp();
    const double scale = hypot(cen.x, cen.y) + rad;
    for (int it = 0; it < 400 && npole > 0; ++it) {
      double dmax = 0.0;
      for (int k = lane; k < npole; k += 32) {
        const cplx x = xk[k];
        cplx d = cmake(0.0, 0.0), dp = cmake(0.0, 0.0), sz = cmake(0.0, 0.0);
        for (int c = 0; c < m; ++c) {
          cplx dz = csub(x, zz[supidx[c]]);
          if (dz.x == 0.0 && dz.y == 0.0) dz = cmake(1e-300, 0.0);
          const cplx t = cdiv_nf(cmake(1.0, 0.0), dz);
          const cplx wt = cmul(w[c], t);
          d = cadd(d, wt);
          dp = csub(dp, cmul(wt, t));
          sz = cadd(sz, t);
        }
        cplx rep = cmake(0.0, 0.0);
        for (int l = 0; l < npole; ++l)
          if (l != k) {
            cplx dx = csub(x, xk[l]);
            if (dx.x == 0.0 && dx.y == 0.0) dx = cmake(1e-300, 0.0);
            rep = cadd(rep, cdiv_nf(cmake(1.0, 0.0), dx));
          }
        // Newton step p / p' = d / (d' + d sum_j 1/(x - z_j)), Aberth correction with the other roots
        const cplx pden = cadd(dp, cmul(d, sz));
        cplx delta = cmake(0.0, 0.0);
        if (pden.x != 0.0 || pden.y != 0.0) {
          const cplx nw = cdiv_nf(d, pden);
          const cplx one_m = csub(cmake(1.0, 0.0), cmul(nw, rep));
          delta = (one_m.x != 0.0 || one_m.y != 0.0) ? cdiv_nf(nw, one_m) : nw;
        }
        xn[k] = csub(x, delta);
        dmax = fmax(dmax, hypot(delta.x, delta.y));
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) dmax = fmax(dmax, __shfl_xor_sync(0xffffffffu, dmax, o));
      __syncwarp();
      for (int k = lane; k < npole; k += 32) xk[k] = xn[k];
      __syncwarp();
      // converged to rounding level; ill-conditioned roots (pole-zero doublets with residues far below any threshold) keep
      // moving by ~1e-12 of the scale, which is irrelevant for the approximant: accept 1e-12 at once, anything below 1e-8 at the end
      if (!(dmax > 1e-12 * scale)) break;
      if (it == 399 && dmax > 1e-8 * scale) bad = 1;
    }
    // residues: (1/4) sum_k f(pole + d_k) d_k, d = 1e-6 (1, i, -1, -i) (aaa.f90 calculate_residual / average_residual)
    for (int k = lane; k < npole; k += 32) {
      const cplx sh[4] = {cmake(1e-6, 0.0), cmake(0.0, 1e-6), cmake(-1e-6, 0.0), cmake(0.0, -1e-6)};
      cplx acc = cmake(0.0, 0.0);
      for (int q = 0; q < 4; ++q) {
        const cplx x = cadd(xk[k], sh[q]);
        cplx num = cmake(0.0, 0.0), den = cmake(0.0, 0.0);
        double dmin = 1e300;
        int jc = 0;
        for (int c = 0; c < m; ++c) {
          cplx dz = csub(x, zz[supidx[c]]);
          const double dist = hypot(dz.x, dz.y);
          if (dist < dmin) { dmin = dist; jc = c; }
          if (dist <= 1e-14) dz = cmake(1e-14, 0.0);
          const cplx cm = cdiv_nf(cmake(1.0, 0.0), dz);
          num = cadd(num, cmul(cm, cmul(w[c], ff[supidx[c]])));
          den = cadd(den, cmul(cm, w[c]));
        }
        cplx fv = cdiv_nf(num, den);
        if (dmin < 1e-14) fv = ff[supidx[jc]];
        acc = cadd(acc, cmul(fv, sh[q]));
      }
      rs[k] = cmake(0.25 * acc.x, 0.25 * acc.y);
    }
    __syncwarp();
    for (int i = lane; i < N; i += 32) scr[pair + npair * i] = cmake(0.0, 0.0);
    __syncwarp();
    if (lane == 0) {
      const int half = N / 2;
      int idx = 0;
      for (int k = 0; k < npole; ++k) {
        const cplx x = xk[k], r = rs[k];
        const cplx dc = csub(x, cen);
        const bool finite = (x.x == x.x) && (x.y == x.y) && hypot(dc.x, dc.y) < 1e8 * rad;   // ZGGEV: |denominator| > eps14
        if (finite && hypot(r.x, r.y) > thres) {
          if (idx < half) { scr[pair + npair * idx] = x; scr[pair + npair * (half + idx)] = r; }
          ++idx;
        }
      }
      if (idx > half) bad = 3;        // analytic.f90:362: two many relevant poles
      if (bad) atomicMax(info, bad);
    }
    return;
  }
  // ---- analytic.f90:160-167: [position | value | weight], each block mmax long, zero padded
  for (int i = lane; i < N; i += 32) scr[pair + npair * i] = cmake(0.0, 0.0);
  __syncwarp();
  for (int c = lane; c < m; c += 32) {
    scr[pair + npair * c] = zz[supidx[c]];
    scr[pair + npair * (mmax + c)] = ff[supidx[c]];
    scr[pair + npair * (2 * mmax + c)] = w[c];
  }
  if (bad && lane == 0) atomicMax(info, bad);
}

// ---------------------------------------------------------------- robust Pade ('pade robust', algo/analytic/src/pade_robust.f90)
// One warp per function (Gonnet-Guettel-Trefethen as coded in pade_robust.f90:177-443): Taylor coefficients from the samples on
// the circle (pade_derivative :448, a DFT here), the rank of the Toeplitz block by the singular values (column norms after
// warp_jacobi) with the degree reduction loop (:343-372), the null vector of C and of C diag(|b| + sqrt(eps)) (the reference's
// SVD + QR of the transposed weighted matrix :386-397: its Q(:, n+1) is the conjugate of that null vector up to a phase, which
// the final division by coeff_den(1) removes), trimming of leading / trailing zeros, normalisation.
// Shared memory: fs [N] | cf [K] | C [dmax x (dmax+1)] | V [(dmax+1)^2] | den [dmax+1] | num [K] | dm (double) [dmax+1]
// Output per function: out[0] = deg_num, out[1] = deg_den, out[2 ..] = numerator then denominator (pade_coeff_robust :150-152).
__global__ void __launch_bounds__(32) k_pade_robust(long nfun, int N, double radius, int deg_num0, int deg_den0, double rel_tol,
                                                    double rel_tol_fft, cplx *__restrict__ scr, long fstride, long estride,
                                                    int *__restrict__ info) {
  const long fun = blockIdx.x;
  if (fun >= nfun) return;
  const int lane = threadIdx.x;
  const int K = deg_num0 + deg_den0 + 1, dmx = deg_den0;
  extern __shared__ cplx pr_sm[];
  cplx *fs = pr_sm, *cf = fs + N, *C = cf + K, *V = C + (long)dmx * (dmx + 1), *den = V + (long)(dmx + 1) * (dmx + 1), *num = den + dmx + 1;
  double *dm = (double *)(num + K);
  cplx *f = scr + fun * fstride;
  for (int i = lane; i < N; i += 32) fs[i] = f[(long)i * estride];
  __syncwarp();
  // ---- pade_derivative: cf_k = (1/N) sum_j f_j exp(-2 pi i j k / N), rescaled by radius^-k
  for (int k = lane; k < K; k += 32) {
    cplx acc = cmake(0.0, 0.0);
    if (k < N)
      for (int j = 0; j < N; ++j) {
        double sn, cs;
        sincospi(-2.0 * (double)(((long)j * k) % N) / N, &sn, &cs);
        acc = cfma(fs[j], cmake(cs, sn), acc);
      }
    cf[k] = cmake(acc.x / N, acc.y / N);
  }
  __syncwarp();
  if (lane == 0 && fabs(radius - 1.0) > 1e-14) {
    double rescale = 1.0;
    for (int k = 1; k < min(K, N); ++k) { rescale = rescale / radius; cf[k] = cscale(rescale, cf[k]); }
  }
  __syncwarp();
  double n2 = 0.0;
  for (int k = lane; k < K; k += 32) n2 += cf[k].x * cf[k].x + cf[k].y * cf[k].y;
  n2 = warp_sum_d(n2);
  const double tol_fft_abs = rel_tol_fft * sqrt(n2);
  for (int k = lane; k < K; k += 32) if (hypot(cf[k].x, cf[k].y) < tol_fft_abs) cf[k] = cmake(0.0, 0.0);
  __syncwarp();
  double imax = 0.0;
  for (int k = lane; k < K; k += 32) imax = fmax(imax, fabs(cf[k].y));
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) imax = fmax(imax, __shfl_xor_sync(0xffffffffu, imax, o));
  if (imax < tol_fft_abs) for (int k = lane; k < K; k += 32) cf[k].y = 0.0;
  __syncwarp();
  // ---- tolerances of the coefficient stage
  n2 = 0.0;
  double amax = 0.0, amax_num = 0.0;
  for (int k = lane; k < K; k += 32) {
    const double a = hypot(cf[k].x, cf[k].y);
    n2 += cf[k].x * cf[k].x + cf[k].y * cf[k].y;
    amax = fmax(amax, a);
    if (k <= deg_num0) amax_num = fmax(amax_num, a);
  }
  n2 = warp_sum_d(n2);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) { amax = fmax(amax, __shfl_xor_sync(0xffffffffu, amax, o)); amax_num = fmax(amax_num, __shfl_xor_sync(0xffffffffu, amax_num, o)); }
  const double abs_tol = rel_tol * sqrt(n2);
  int dn = deg_num0, dd = deg_den0, trivial = 0, have_c = 0, bad = 0;
  if (amax_num <= rel_tol * amax) { dn = 0; dd = 0; trivial = 1; }          // :326-334
  // ---- degree reduction by the rank of the Toeplitz block (:343-372)
  while (!trivial && dd > 0) {
    for (int i = lane; i < dd * (dd + 1); i += 32) {
      const int a = i % dd, j = i / dd, idx = dn + 1 + a - j;
      C[i] = idx >= 0 ? cf[idx] : cmake(0.0, 0.0);
    }
    for (int i = lane; i < (dd + 1) * (dd + 1); i += 32) V[i] = (i % (dd + 1) == i / (dd + 1)) ? cmake(1.0, 0.0) : cmake(0.0, 0.0);
    __syncwarp();
    if (warp_jacobi(C, dd, dd + 1, V, lane)) bad = 1;
    int rho = 0;
    for (int c = 0; c <= dd; ++c) {
      double nn = 0.0;
      for (int r = lane; r < dd; r += 32) { const cplx x = C[r + (long)dd * c]; nn += x.x * x.x + x.y * x.y; }
      nn = warp_sum_d(nn);
      rho += sqrt(nn) > abs_tol;
    }
    have_c = 1;
    if (rho >= dd) break;
    dn -= dd - rho;
    dd = rho;
    have_c = 0;
    if (dn < 0) { bad = 2; dn = 0; break; }
  }
  if (trivial) {
    if (lane == 0) { num[0] = cmake(0.0, 0.0); den[0] = cmake(1.0, 0.0); }
  } else if (dd == 0) {
    for (int k = lane; k <= dn; k += 32) num[k] = cf[k];
    if (lane == 0) den[0] = cmake(1.0, 0.0);
  } else if (dn > 1 && have_c) {
    // null vector of C (column of V with the smallest column norm) -> weights d = |b| + sqrt(eps)
    int jmin = 0;
    double smin = 1e300;
    for (int c = 0; c <= dd; ++c) {
      double nn = 0.0;
      for (int r = lane; r < dd; r += 32) { const cplx x = C[r + (long)dd * c]; nn += x.x * x.x + x.y * x.y; }
      nn = warp_sum_d(nn);
      if (nn < smin) { smin = nn; jmin = c; }
    }
    for (int j = lane; j <= dd; j += 32) { const cplx b = V[j + (long)(dd + 1) * jmin]; dm[j] = hypot(b.x, b.y) + 1.4901161193847656e-08; }
    __syncwarp();
    // null vector of C diag(d)
    for (int i = lane; i < dd * (dd + 1); i += 32) {
      const int a = i % dd, j = i / dd, idx = dn + 1 + a - j;
      C[i] = idx >= 0 ? cscale(dm[j], cf[idx]) : cmake(0.0, 0.0);
    }
    for (int i = lane; i < (dd + 1) * (dd + 1); i += 32) V[i] = (i % (dd + 1) == i / (dd + 1)) ? cmake(1.0, 0.0) : cmake(0.0, 0.0);
    __syncwarp();
    if (warp_jacobi(C, dd, dd + 1, V, lane)) bad = 1;
    jmin = 0; smin = 1e300;
    for (int c = 0; c <= dd; ++c) {
      double nn = 0.0;
      for (int r = lane; r < dd; r += 32) { const cplx x = C[r + (long)dd * c]; nn += x.x * x.x + x.y * x.y; }
      nn = warp_sum_d(nn);
      if (nn < smin) { smin = nn; jmin = c; }
    }
    double dn2 = 0.0;
    for (int j = lane; j <= dd; j += 32) {                                 // coeff_den = d * Q(:, n+1), Q(:, n+1) = conj(null vector)
      den[j] = cscale(dm[j], cconj(V[j + (long)(dd + 1) * jmin]));
      dn2 += den[j].x * den[j].x + den[j].y * den[j].y;
    }
    dn2 = warp_sum_d(dn2);
    const double inv = 1.0 / sqrt(dn2);
    for (int j = lane; j <= dd; j += 32) den[j] = cscale(inv, den[j]);
    __syncwarp();
    for (int i = lane; i <= dn; i += 32) {                                  // coeff_num = zmat(1:deg_num+1, 1:deg_den+1) coeff_den
      cplx acc = cmake(0.0, 0.0);
      for (int j = 0; j <= min(i, dd); ++j) acc = cfma(cf[i - j], den[j], acc);
      num[i] = acc;
    }
    __syncwarp();
    if (lane == 0) {                                                         // :402-424 leading / trailing zeros of the denominator
      int lam = 0;
      while (lam <= dd && !(hypot(den[lam].x, den[lam].y) > rel_tol)) ++lam;
      if (lam > dd) { bad = 2; lam = 0; }
      if (lam > 0) {
        dn -= lam; dd -= lam;
        for (int i = 0; i <= dn; ++i) num[i] = num[i + lam];
        for (int i = 0; i <= dd; ++i) den[i] = den[i + lam];
      }
      int last = dd;
      while (last > 0 && !(hypot(den[last].x, den[last].y) > rel_tol)) --last;
      dd = last;
      if (dn < 0) { bad = 2; dn = 0; }
    }
    dn = __shfl_sync(0xffffffffu, dn, 0); dd = __shfl_sync(0xffffffffu, dd, 0); bad = __shfl_sync(0xffffffffu, bad, 0);
  } else {
    bad = 2;      // deg_num <= 1 with deg_den > 0: coeff_num / coeff_den are never assigned in the reference (:386)
  }
  __syncwarp();
  if (lane == 0 && bad != 2) {
    if (!trivial) {                                                          // :426-432 trailing zeros of the numerator
      int last = dn;
      while (last >= 0 && !(hypot(num[last].x, num[last].y) > abs_tol)) --last;
      if (last < 0) last = 0;
      dn = last;
    }
    const cplx d0 = den[0];                                                  // :434-435
    for (int i = 0; i <= dn; ++i) num[i] = cdiv_nf(num[i], d0);
    for (int i = 0; i <= dd; ++i) den[i] = cdiv_nf(den[i], d0);
    for (int i = 0; i < N; ++i) f[(long)i * estride] = cmake(0.0, 0.0);
    f[0] = cmake((double)dn, 0.0);
    f[estride] = cmake((double)dd, 0.0);
    for (int i = 0; i <= dn; ++i) f[(long)(2 + i) * estride] = num[i];
    for (int i = 0; i <= dd; ++i) f[(long)(3 + dn + i) * estride] = den[i];
  }
  if (bad && lane == 0) atomicMax(info, bad);
}

// Ec(r, G) = exp(-i G r) (nnr x ngm) and ET(G, r) = exp(+i G r) (ngm x nnr); r = i1 + n1 (i2 + n2 i3) (QE column-major box),
// G at box position p: G r = 2 pi (p1 i1 / n1 + p2 i2 / n2 + p3 i3 / n3), each term reduced exactly in integers.
__global__ void k_corr_tables(int n1, int n2, int n3, int ngm, const int *__restrict__ nl, cplx *__restrict__ Ec,
                              cplx *__restrict__ ET) {
  const int nnr = n1 * n2 * n3;
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  const int ig = blockIdx.y;
  if (r >= nnr || ig >= ngm) return;
  const int p = nl[ig] - 1;
  const int p1 = p % n1, p2 = (p / n1) % n2, p3 = p / (n1 * n2);
  const int i1 = r % n1, i2 = (r / n1) % n2, i3 = r / (n1 * n2);
  const double ph = 2.0 * ((double)((p1 * i1) % n1) / n1 + (double)((p2 * i2) % n2) / n2 + (double)((p3 * i3) % n3) / n3);
  double s, c;
  sincospi(ph, &s, &c);
  Ec[(long)r + (long)nnr * ig] = cmake(c, -s);
  ET[(long)ig + (long)ngm * r] = cmake(c, s);
}

// strided copy of the G block f(:ngm, :ngm) of a host-layout f(nnr, nnr)
__global__ void k_block_copy(int ngm, long ld_src, long ld_dst, const cplx *__restrict__ src, cplx *__restrict__ dst) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int j = blockIdx.y;
  if (i < ngm) dst[(long)i + ld_dst * j] = src[(long)i + ld_src * j];
}

// ---------------------------------------------------------------- the G W product
// part_z(r, r') = sum_{b in chunk z} coef_b * Gr_b(r, r') * (X_b ET)(r, r'),  X_b: M x K (column-major, batch stride M*K),
// ET: K x N, Gr_b: M x N (batch stride M*N).  Tiling, staging and the 3M complex DMMA product are those of k_zgemm
// (gemm.cu); the batch loop sits outside the K loop and the product with G is the epilogue of every batch entry, so the
// real-space W never exists in memory.  The G tile of entry b is prefetched into registers before the K loop of b.
constexpr int GW_BM = 64, GW_BN = 32, GW_BK = 16, GW_T = 256;
constexpr int GW_PK = GW_BK + 4, GW_PM = GW_BM + 2;
constexpr int GW_CTA_PER_SM = 1;   // __launch_bounds__(GW_T, 1): 240 registers, no spills (2 per SM at 128 registers measured equal)
constexpr size_t GW_SMEM = (size_t)2 * (GW_BK * GW_PM + GW_BN * GW_PK) * sizeof(cplx);

__global__ void __launch_bounds__(GW_T, 1) k_gw_product(int M, int N, int K, int nb, int bchunk, const cplx *__restrict__ X,
                                                      const cplx *__restrict__ ET, const cplx *__restrict__ Gr,
                                                      const cplx *__restrict__ coef, cplx *__restrict__ part) {
  const int m0 = blockIdx.x * GW_BM, n0 = blockIdx.y * GW_BN;
  const int b0 = blockIdx.z * bchunk, b1 = min(nb, b0 + bchunk);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  constexpr int ASZ = GW_BK * GW_PM, BSZ = GW_BN * GW_PK;
  extern __shared__ cplx gwsm[];
  cplx *As = gwsm, *Bs = gwsm + 2 * ASZ;
  const int wm = (warp & 3) * 16, wn = (warp >> 2) * 16;
  const int g = lane >> 2, t = lane & 3;
  const long MK = (long)M * K, MN = (long)M * N;

  cplx sacc[2][2][2];
#pragma unroll
  for (int i = 0; i < 2; ++i)
#pragma unroll
    for (int j = 0; j < 2; ++j)
#pragma unroll
      for (int c = 0; c < 2; ++c) sacc[i][j][c] = cmake(0.0, 0.0);

  auto stage_load = [&](int st, const cplx *A, int k0) {
    cplx *as = As + st * ASZ, *bs = Bs + st * BSZ;
#pragma unroll
    for (int r = 0; r < GW_BM * GW_BK / GW_T; ++r) {
      const int i = tid + r * GW_T;
      const int m = i % GW_BM, k = i / GW_BM;
      const bool ok = (k0 + k < K) && (m0 + m < M);
      cp_async16(as + k * GW_PM + m, ok ? A + (long)(m0 + m) + (long)(k0 + k) * M : A, ok ? 16 : 0);
    }
#pragma unroll
    for (int r = 0; r < GW_BN * GW_BK / GW_T; ++r) {
      const int i = tid + r * GW_T;
      const int k = i % GW_BK, n = i / GW_BK;
      const bool ok = (k0 + k < K) && (n0 + n < N);
      cp_async16(bs + n * GW_PK + k, ok ? ET + (long)(k0 + k) + (long)(n0 + n) * K : ET, ok ? 16 : 0);
    }
  };

  for (int b = b0; b < b1; ++b) {
    const cplx *A = X + MK * b;
    const cplx *Gb = Gr + MN * b;
    stage_load(0, A, 0);
    cp_async_commit();
    // G tile of this entry -> registers (latency hidden behind the K loop)
    cplx gv[2][2][2];
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
      for (int j = 0; j < 2; ++j)
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          const int row = m0 + wm + i * 8 + g, col = n0 + wn + j * 8 + 2 * t + c;
          gv[i][j][c] = (row < M && col < N) ? __ldg(Gb + (long)row + (long)col * M) : cmake(0.0, 0.0);
        }
    double p1[2][2][2], p2[2][2][2], p3[2][2][2];
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
      for (int j = 0; j < 2; ++j)
#pragma unroll
        for (int c = 0; c < 2; ++c) p1[i][j][c] = p2[i][j][c] = p3[i][j][c] = 0.0;
    int it = 0;
    for (int k0 = 0; k0 < K; k0 += GW_BK, ++it) {
      if (k0 + GW_BK < K) stage_load((it + 1) & 1, A, k0 + GW_BK);
      cp_async_commit();
      cp_async_wait<1>();
      __syncthreads();
      const cplx *as = As + (it & 1) * ASZ, *bs = Bs + (it & 1) * BSZ;
#pragma unroll
      for (int kk = 0; kk < GW_BK; kk += 4) {
        cplx a[2], bb[2];
#pragma unroll
        for (int i = 0; i < 2; ++i) a[i] = as[(kk + t) * GW_PM + wm + i * 8 + g];
#pragma unroll
        for (int j = 0; j < 2; ++j) bb[j] = bs[(wn + j * 8 + g) * GW_PK + kk + t];
        double asum[2], bsum[2];
#pragma unroll
        for (int i = 0; i < 2; ++i) asum[i] = a[i].x + a[i].y;
#pragma unroll
        for (int j = 0; j < 2; ++j) bsum[j] = bb[j].x + bb[j].y;
#pragma unroll
        for (int i = 0; i < 2; ++i)
#pragma unroll
          for (int j = 0; j < 2; ++j) {
            dmma(p1[i][j][0], p1[i][j][1], a[i].x, bb[j].x);
            dmma(p2[i][j][0], p2[i][j][1], a[i].y, bb[j].y);
            dmma(p3[i][j][0], p3[i][j][1], asum[i], bsum[j]);
          }
      }
      __syncthreads();
    }
    const cplx cb = coef[b];
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
      for (int j = 0; j < 2; ++j)
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          const cplx w = cmake(p1[i][j][c] - p2[i][j][c], (p3[i][j][c] - p1[i][j][c]) - p2[i][j][c]);
          sacc[i][j][c] = cfma(cmul(cb, gv[i][j][c]), w, sacc[i][j][c]);
        }
  }
  cplx *P = part + MN * blockIdx.z;
#pragma unroll
  for (int i = 0; i < 2; ++i)
#pragma unroll
    for (int j = 0; j < 2; ++j)
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        const int row = m0 + wm + i * 8 + g, col = n0 + wn + j * 8 + 2 * t + c;
        if (row < M && col < N) P[(long)row + (long)col * M] = sacc[i][j][c];
      }
}

// acc = sum_z part_z in a fixed order (deterministic)
__global__ void k_gw_reduce(long n, int nz, const cplx *__restrict__ part, cplx *__restrict__ acc) {
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  cplx s = part[i];
  for (int z = 1; z < nz; ++z) s = cadd(s, part[i + n * z]);
  acc[i] = s;
}

// ---------------------------------------------------------------- host helpers
static int symm_mesh(const sgw_freqbins *f, std::vector<cplx> *z, std::vector<int> *src, std::vector<int> *dst) {
  // freqbins_symm (freqbins.f90:243-305)
  const int nf = f->num_solver;
  z->clear();
  if (src) { src->clear(); dst->clear(); }
  if (f->freq_symm_coul == 0 || f->freq_symm_coul == 2) {
    for (int i = 0; i < nf; ++i) {
      const cplx s = cmake(f->solver[i].re, f->solver[i].im);
      z->push_back(f->freq_symm_coul == 2 ? cmul(s, s) : s);
    }
    return SGW_OK;
  }
  int num_zero = 0;
  for (int i = 0; i < nf; ++i) num_zero += hypot(f->solver[i].re, f->solver[i].im) < 1e-14;
  if (num_zero > 1) return SGW_E_ARG;
  z->resize(2 * nf - num_zero);
  int ifs = nf;
  for (int i = 0; i < nf; ++i) {
    (*z)[i] = cmake(f->solver[i].re, f->solver[i].im);
    if (hypot(f->solver[i].re, f->solver[i].im) >= 1e-14) {
      (*z)[ifs] = cmake(-f->solver[i].re, -f->solver[i].im);
      if (src) { src->push_back(i); dst->push_back(ifs); }
      ++ifs;
    }
  }
  return SGW_OK;
}

static int check_freq(sgw_ctx *ctx, const sgw_freqbins *f, int model) {
  SGW_ARG(f && f->num_solver > 0 && f->solver, "freqbins: solver frequencies missing");
  SGW_ARG(f->freq_symm_coul >= 0 && f->freq_symm_coul <= 2, "freqbins: freq_symm_coul must be 0, 1 or 2");
  SGW_ARG(model >= SGW_GODBY_NEEDS && model <= SGW_AAA_POLE, "No screening model chosen!");        // analytic.f90:186
  return SGW_OK;
}

// device part of analytic_eval: coefficients and gmapsym already resident
static int analytic_eval_dev(sgw_ctx *ctx, int model, int ngc, int N, const int *d_gmap, const cplx *d_z, const cplx *d_coeff,
                             int nout, const cplx *d_wout, cplx *d_out) {
  dim3 grid((ngc + 63) / 64, ngc, nout);
  ProfScope prof(ctx, PC_OTHER);
  k_analytic_eval<<<grid, 64, 0, ctx->stream>>>(model, ngc, N, d_gmap, d_z, d_coeff, d_wout, d_out);
  SGW_LAUNCH_CHECK();
  return SGW_OK;
}

// f(r, r') = 1/omega Ec fg ET for nb G-space blocks fg_b(ngm, ngm) stored back to back -> fr_b(nnr, nnr)
static int invfft6_dev(sgw_ctx *ctx, double omega, int nb, const cplx *fg, cplx *fr) {
  const CorrGrid &c = ctx->corr;
  cplx *X = nullptr;
  SGW_CHECK(ws(ctx, "sg_x", (size_t)c.nnr * c.ngm * nb, &X));
  SGW_CHECK(gemm_n_n(ctx, c.nnr, c.ngm * nb, c.ngm, cmake(1.0 / omega, 0.0), c.d_Ec, c.nnr, fg, c.ngm, cmake(0, 0), X, c.nnr));
  for (int b = 0; b < nb; ++b)
    SGW_CHECK(gemm_n_n(ctx, c.nnr, c.nnr, c.ngm, cmake(1, 0), X + (size_t)c.nnr * c.ngm * b, c.nnr, c.d_ET, c.ngm, cmake(0, 0),
                       fr + (size_t)c.nnr * c.nnr * b, c.nnr));
  return SGW_OK;
}

// fg(ngm, ngm) = beta fg + scale * omega / nnr^2 ET fr Ec
static int fwfft6_dev(sgw_ctx *ctx, double omega, const cplx *fr, cplx beta, cplx *fg, long ldg) {
  const CorrGrid &c = ctx->corr;
  cplx *U = nullptr;
  SGW_CHECK(ws(ctx, "sg_u", (size_t)c.nnr * c.ngm, &U));
  SGW_CHECK(gemm_n_n(ctx, c.nnr, c.ngm, c.nnr, cmake(1, 0), fr, c.nnr, c.d_Ec, c.nnr, cmake(0, 0), U, c.nnr));
  const double s = omega / ((double)c.nnr * (double)c.nnr);
  SGW_CHECK(gemm_n_n(ctx, c.ngm, c.ngm, c.nnr, cmake(s, 0.0), c.d_ET, c.ngm, U, c.nnr, beta, fg, ldg));
  return SGW_OK;
}

// launch k_pade_robust for nfun functions stored with strides (fstride between functions, estride between samples)
static int pade_robust_dev(sgw_ctx *ctx, long nfun, int N, double radius, int deg_num, int deg_den, double tol, double tol_fft,
                           cplx *d, long fstride, long estride) {
  SGW_ARG(radius > 0.0, "radius in the complex plane must be > 0");                               // pade_robust.f90:290
  SGW_ARG(deg_num >= 0 && deg_den >= 0, "degree of numerator / denominator must be positive");     // :292-295
  SGW_ARG(3 + deg_num + deg_den <= N, "coefficient layout [deg_num, deg_den, num, den] needs 4 + deg_num + deg_den entries");
  const int K = deg_num + deg_den + 1;
  const size_t smem = sizeof(cplx) * ((size_t)N + 2 * (size_t)K + (size_t)deg_den * (deg_den + 1) + (size_t)(deg_den + 1) * (deg_den + 1) +
                                      deg_den + 1) + sizeof(double) * (deg_den + 2);
  if (smem > ctx->smem_optin) { ctx->err = "'pade robust': too many frequencies for the shared-memory fit"; return SGW_E_UNSUPPORTED; }
  if (smem > 48 * 1024) SGW_CUDA(cudaFuncSetAttribute(k_pade_robust, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int *dinfo = nullptr;
  SGW_CHECK(ws(ctx, "an_info", (size_t)1, &dinfo));
  SGW_CUDA(cudaMemsetAsync(dinfo, 0, sizeof(int), ctx->stream));
  k_pade_robust<<<(unsigned)nfun, 32, smem, ctx->stream>>>(nfun, N, radius, deg_num, deg_den, tol, tol_fft, d, fstride, estride, dinfo);
  SGW_LAUNCH_CHECK();
  int hinfo = 0;
  SGW_CUDA(cudaMemcpyAsync(&hinfo, dinfo, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
  SGW_CUDA(cudaStreamSynchronize(ctx->stream));
  if (hinfo == 2) {
    ctx->err = "'pade robust': the degree reduction left a numerator of degree <= 1 with a non-trivial denominator, which "
               "pade_robust.f90:386 leaves undefined";
    return SGW_E_UNSUPPORTED;
  }
  if (hinfo != 0) { ctx->err = "'pade robust': singular value iteration did not converge"; return SGW_E_ARG; }
  return SGW_OK;
}

}  // namespace sgw

extern "C" {

int sgw_aaa_pole_residual(sgw_ctx *ctx, int m, const sgw_cplx *position, const sgw_cplx *value, const sgw_cplx *weight,
                          sgw_cplx *pole, sgw_cplx *residual, int *num_pole) {
  if (!ctx) return SGW_E_ARG;
  cudaSetDevice(ctx->device);
  SGW_ARG(position && value && weight && pole && residual && num_pole, "bad argument");
  SGW_ARG(m >= 1 && m <= 32, "aaa_pole_residual: 1 <= number of support points <= 32 on the device");
  for (int c = 0; c < m; ++c) SGW_ARG(hypot(weight[c].re, weight[c].im) > 1e-12, "aaa_pole_residual: zero weight");
  begin_call(ctx);
  const int mmax = m, N = 3 * m;            // the 'aaa' coefficient layout [position | value | weight] the kernel reads
  std::vector<cplx> h((size_t)N);
  for (int c = 0; c < m; ++c) {
    h[c] = cmake(position[c].re, position[c].im);
    h[mmax + c] = cmake(value[c].re, value[c].im);
    h[2 * mmax + c] = cmake(weight[c].re, weight[c].im);
  }
  cplx *d = nullptr, *dz = nullptr;
  int *dinfo = nullptr;
  SGW_CHECK(ws(ctx, "an_coeff", (size_t)N, &d));
  SGW_CHECK(ws(ctx, "an_z", (size_t)N, &dz));
  SGW_CHECK(ws(ctx, "an_info", (size_t)1, &dinfo));
  SGW_CUDA(cudaMemcpyAsync(d, h.data(), sizeof(cplx) * N, cudaMemcpyHostToDevice, ctx->stream));
  SGW_CUDA(cudaMemsetAsync(dz, 0, sizeof(cplx) * N, ctx->stream));
  SGW_CUDA(cudaMemsetAsync(dinfo, 0, sizeof(int), ctx->stream));
  const size_t smem = sizeof(cplx) * ((size_t)3 * N + (size_t)N * mmax + (size_t)mmax * mmax + mmax) + sizeof(int) * ((size_t)2 * N + mmax);
  if (smem > 48 * 1024) SGW_CUDA(cudaFuncSetAttribute(k_aaa_coeff, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  k_aaa_coeff<<<1, 32, smem, ctx->stream>>>(1, N, mmax, 0.0, dz, d, dinfo, 2);
  SGW_LAUNCH_CHECK();
  int hinfo = 0;
  SGW_CUDA(cudaMemcpyAsync(&hinfo, dinfo, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
  SGW_CUDA(cudaMemcpyAsync(h.data(), d, sizeof(cplx) * N, cudaMemcpyDeviceToHost, ctx->stream));
  SGW_CUDA(cudaStreamSynchronize(ctx->stream));
  end_call(ctx);
  if (hinfo != 0) { ctx->err = "error occured in AAA approximation (the pole iteration did not converge)"; return SGW_E_ARG; }
  const int half = N / 2;
  int np = 0;
  for (int k = 0; k < half && k < m - 1; ++k) {
    const cplx r = h[half + k];
    if (r.x == 0.0 && r.y == 0.0) break;       // pole_correction keeps |residual| > thres = 0: stored entries are contiguous
    pole[np].re = h[k].x; pole[np].im = h[k].y;
    residual[np].re = r.x; residual[np].im = r.y;
    ++np;
  }
  *num_pole = np;
  return SGW_OK;
}

int sgw_pade_robust(sgw_ctx *ctx, double radius, int num_point, const sgw_cplx *func, int *deg_num, int *deg_den,
                    sgw_cplx *coeff_num, sgw_cplx *coeff_den, double tol_coeff, double tol_fft) {
  if (!ctx) return SGW_E_ARG;
  cudaSetDevice(ctx->device);
  SGW_ARG(func && deg_num && deg_den && coeff_num && coeff_den && num_point > 0, "bad argument");
  const double tol = tol_coeff > 0.0 ? tol_coeff : 1e-14;                                         // pade_robust.f90:298-308
  const double tfft = tol_fft > 0.0 ? tol_fft : tol;
  begin_call(ctx);
  // the kernel writes [deg_num, deg_den, numerator, denominator] over the samples: give it room for both
  const int len = std::max(num_point, 4 + *deg_num + *deg_den);
  cplx *d = nullptr;
  SGW_CHECK(ws(ctx, "an_coeff", (size_t)len, &d));
  SGW_CUDA(cudaMemsetAsync(d, 0, sizeof(cplx) * len, ctx->stream));
  SGW_CUDA(cudaMemcpyAsync(d, func, sizeof(cplx) * num_point, cudaMemcpyHostToDevice, ctx->stream));
  {
    const int dn0 = *deg_num, dd0 = *deg_den;
    SGW_ARG(radius > 0.0 && dn0 >= 0 && dd0 >= 0, "radius must be > 0 and the degrees >= 0");
    const int K = dn0 + dd0 + 1;
    const size_t smem = sizeof(cplx) * ((size_t)num_point + 2 * (size_t)K + (size_t)dd0 * (dd0 + 1) + (size_t)(dd0 + 1) * (dd0 + 1) + dd0 + 1) +
                        sizeof(double) * (dd0 + 2);
    if (smem > ctx->smem_optin) { ctx->err = "'pade robust': too many frequencies for the shared-memory fit"; return SGW_E_UNSUPPORTED; }
    if (smem > 48 * 1024) SGW_CUDA(cudaFuncSetAttribute(k_pade_robust, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int *dinfo = nullptr;
    SGW_CHECK(ws(ctx, "an_info", (size_t)1, &dinfo));
    SGW_CUDA(cudaMemsetAsync(dinfo, 0, sizeof(int), ctx->stream));
    // the output needs len entries: run with N = num_point samples but let the kernel clear / write up to len via a padded view
    k_pade_robust<<<1, 32, smem, ctx->stream>>>(1, num_point, radius, dn0, dd0, tol, tfft, d, len, 1, dinfo);
    SGW_LAUNCH_CHECK();
    int hinfo = 0;
    SGW_CUDA(cudaMemcpyAsync(&hinfo, dinfo, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    SGW_CUDA(cudaStreamSynchronize(ctx->stream));
    if (hinfo == 2) { ctx->err = "'pade robust': numerator of degree <= 1 with a non-trivial denominator is undefined in pade_robust.f90:386"; return SGW_E_UNSUPPORTED; }
    if (hinfo != 0) { ctx->err = "'pade robust': singular value iteration did not converge"; return SGW_E_ARG; }
  }
  std::vector<cplx> h(len);
  SGW_CUDA(cudaMemcpyAsync(h.data(), d, sizeof(cplx) * len, cudaMemcpyDeviceToHost, ctx->stream));
  SGW_CUDA(cudaStreamSynchronize(ctx->stream));
  const int dn = (int)std::lround(h[0].x), dd = (int)std::lround(h[1].x);
  for (int i = 0; i <= dn; ++i) { coeff_num[i].re = h[2 + i].x; coeff_num[i].im = h[2 + i].y; }
  for (int i = 0; i <= dd; ++i) { coeff_den[i].re = h[3 + dn + i].x; coeff_den[i].im = h[3 + dn + i].y; }
  *deg_num = dn; *deg_den = dd;
  end_call(ctx);
  return SGW_OK;
}

int sgw_freqbins_num_freq(const sgw_freqbins *freq) {
  if (!freq || freq->num_solver <= 0 || !freq->solver) return SGW_E_ARG;
  std::vector<cplx> z;
  if (symm_mesh(freq, &z, nullptr, nullptr) != SGW_OK) return SGW_E_ARG;
  return (int)z.size();
}

int sgw_coulpade(sgw_ctx *ctx, int ngc, int nfreq, const double *factor, sgw_cplx *scrcoul_g) {
  if (!ctx) return SGW_E_ARG;
  cudaSetDevice(ctx->device);
  SGW_ARG(ngc > 0 && nfreq > 0 && factor && scrcoul_g, "bad argument");
  begin_call(ctx);
  const long total = (long)ngc * ngc * nfreq;
  cplx *d = nullptr;
  double *df = nullptr;
  SGW_CHECK(ws(ctx, "an_coeff", (size_t)total, &d));
  SGW_CHECK(ws(ctx, "an_fac", (size_t)ngc, &df));
  SGW_CUDA(cudaMemcpyAsync(d, scrcoul_g, sizeof(cplx) * total, cudaMemcpyHostToDevice, ctx->stream));
  SGW_CUDA(cudaMemcpyAsync(df, factor, sizeof(double) * ngc, cudaMemcpyHostToDevice, ctx->stream));
  k_coulpade<<<(unsigned)((total + 255) / 256), 256, 0, ctx->stream>>>(ngc, total, df, d);
  SGW_LAUNCH_CHECK();
  SGW_CUDA(cudaMemcpyAsync(scrcoul_g, d, sizeof(cplx) * total, cudaMemcpyDeviceToHost, ctx->stream));
  SGW_CUDA(cudaStreamSynchronize(ctx->stream));
  end_call(ctx);
  return SGW_OK;
}

int sgw_analytic_coeff(sgw_ctx *ctx, int model_coul, double thres, const sgw_freqbins *freq, int ngc, sgw_cplx *scrcoul_g) {
  if (!ctx) return SGW_E_ARG;
  cudaSetDevice(ctx->device);
  SGW_CHECK(check_freq(ctx, freq, model_coul));
  SGW_ARG(ngc > 0 && scrcoul_g, "bad argument");
  std::vector<cplx> z;
  std::vector<int> src, dst;
  if (symm_mesh(freq, &z, &src, &dst) != SGW_OK) {
    ctx->err = "only a single frequency may be smaller than 1e-14";    // freqbins.f90:276
    return SGW_E_ARG;
  }
  const int N = (int)z.size();
  if (model_coul == SGW_GODBY_NEEDS) {
    SGW_ARG(N == 2, "must provide exactly 2 frequencies");                                        // godby_needs.f90:50
    SGW_ARG(freq->solver[1].im >= 0.0, "plasmon frequency must be positive");                     // godby_needs.f90:48
  }
  begin_call(ctx);
  const long npair = (long)ngc * ngc, total = npair * N;
  cplx *d = nullptr, *dz = nullptr;
  SGW_CHECK(ws(ctx, "an_coeff", (size_t)total, &d));
  SGW_CHECK(ws(ctx, "an_z", (size_t)N, &dz));
  SGW_CUDA(cudaMemcpyAsync(d, scrcoul_g, sizeof(cplx) * total, cudaMemcpyHostToDevice, ctx->stream));
  if (model_coul == SGW_PADE_ROBUST) {                                                           // pade_coeff_robust (pade_robust.f90:91-162)
    SGW_ARG(freq->freq_symm_coul == 0, "'pade robust' works on the solver frequencies themselves (no symmetrisation, gwq_readin.f90:371)");
    SGW_ARG(N >= 10, "use at least 10 frequencies to form the circle");
    const double radius = freq->solver[0].re;
    for (int i = 0; i < N; ++i)
      SGW_ARG(fabs(hypot(freq->solver[i].re, freq->solver[i].im) - radius) <= 1e-12, "frequencies must span circle in the complex plane");
    SGW_CHECK(pade_robust_dev(ctx, npair, N, radius, N / 2 - 2, N / 2 - 2, 1e-14, 1e-14, d, 1, npair));
  } else if (model_coul == SGW_GODBY_NEEDS) {
    k_gn_coeff<<<(unsigned)((npair + 127) / 128), 128, 0, ctx->stream>>>(npair, freq->solver[1].im, d);
    SGW_LAUNCH_CHECK();
  } else {
    const bool aaa = model_coul == SGW_AAA_APPROX || model_coul == SGW_AAA_POLE;
    if (aaa && N / 3 < 1) { ctx->err = "'aaa' needs at least 3 frequencies"; return SGW_E_ARG; }
    SGW_CUDA(cudaMemcpyAsync(dz, z.data(), sizeof(cplx) * N, cudaMemcpyHostToDevice, ctx->stream));
    if (!src.empty()) {
      int *dsrc = nullptr, *ddst = nullptr;
      SGW_CHECK(ws(ctx, "an_src", src.size(), &dsrc));
      SGW_CHECK(ws(ctx, "an_dst", dst.size(), &ddst));
      SGW_CUDA(cudaMemcpyAsync(dsrc, src.data(), sizeof(int) * src.size(), cudaMemcpyHostToDevice, ctx->stream));
      SGW_CUDA(cudaMemcpyAsync(ddst, dst.data(), sizeof(int) * dst.size(), cudaMemcpyHostToDevice, ctx->stream));
      dim3 grid((unsigned)((npair + 255) / 256), (unsigned)src.size());
      k_mirror<<<grid, 256, 0, ctx->stream>>>(npair, (int)src.size(), dsrc, ddst, d);
      SGW_LAUNCH_CHECK();
    }
    if (aaa) {
      // analytic.f90:146-148: max_point = num_freq() / 3 for 'aaa'; 'aaa pole' allows num_freq() -- the device fit stops at
      // num_freq() / 2 (beyond that the Loewner submatrix has fewer rows than columns) and reports it
      const int mmax = model_coul == SGW_AAA_APPROX ? N / 3 : N / 2;
      const size_t smem = sizeof(cplx) * ((size_t)3 * N + (size_t)N * mmax + (size_t)mmax * mmax + mmax) +
                          sizeof(int) * ((size_t)2 * N + mmax);
      if (smem > ctx->smem_optin) { ctx->err = "'aaa': frequency mesh too large for the shared-memory fit"; return SGW_E_UNSUPPORTED; }
      if (smem > 48 * 1024) SGW_CUDA(cudaFuncSetAttribute(k_aaa_coeff, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      int *dinfo = nullptr;
      SGW_CHECK(ws(ctx, "an_info", (size_t)1, &dinfo));
      SGW_CUDA(cudaMemsetAsync(dinfo, 0, sizeof(int), ctx->stream));
      k_aaa_coeff<<<(unsigned)npair, 32, smem, ctx->stream>>>(npair, N, mmax, thres, dz, d, dinfo, model_coul == SGW_AAA_POLE ? 1 : 0);
      SGW_LAUNCH_CHECK();
      int hinfo = 0;
      SGW_CUDA(cudaMemcpyAsync(&hinfo, dinfo, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
      SGW_CUDA(cudaStreamSynchronize(ctx->stream));
      if (hinfo == 3) {
        ctx->err = "two many relevant poles, try reducing the coulomb threshold or increasing the number of frequencies";   // analytic.f90:362
        return SGW_E_ARG;
      }
      if (hinfo == 2) {
        ctx->err = "'aaa pole': fit not converged with num_freq() / 2 support points (the reference continues up to num_freq())";
        return SGW_E_UNSUPPORTED;
      }
      if (hinfo != 0) { ctx->err = "error occured in AAA approximation (iteration for the singular vector / the poles did not converge)"; return SGW_E_ARG; }
    } else {
      k_pade_coeff<<<(unsigned)((npair + 63) / 64), 64, 0, ctx->stream>>>(npair, N, dz, d);
      SGW_LAUNCH_CHECK();
    }
  }
  SGW_CUDA(cudaMemcpyAsync(scrcoul_g, d, sizeof(cplx) * total, cudaMemcpyDeviceToHost, ctx->stream));
  SGW_CUDA(cudaStreamSynchronize(ctx->stream));
  end_call(ctx);
  return SGW_OK;
}

int sgw_analytic_eval(sgw_ctx *ctx, int model_coul, const sgw_freqbins *freq, int ngc, const int32_t *gmapsym,
                      const sgw_cplx *scrcoul_coeff, int nout, const sgw_cplx *freq_out, sgw_cplx *scrcoul) {
  if (!ctx) return SGW_E_ARG;
  cudaSetDevice(ctx->device);
  SGW_CHECK(check_freq(ctx, freq, model_coul));
  SGW_ARG(ngc > 0 && nout > 0 && gmapsym && scrcoul_coeff && freq_out && scrcoul, "bad argument");
  for (int i = 0; i < ngc; ++i) SGW_ARG(gmapsym[i] >= 1 && gmapsym[i] <= ngc, "gmapsym entry outside 1..num_g_corr");
  std::vector<cplx> z;
  if (symm_mesh(freq, &z, nullptr, nullptr) != SGW_OK) { ctx->err = "only a single frequency may be smaller than 1e-14"; return SGW_E_ARG; }
  const int N = (int)z.size();
  if (model_coul == SGW_GODBY_NEEDS) SGW_ARG(N == 2, "must provide exactly 2 frequencies");
  std::vector<cplx> w(nout);
  for (int i = 0; i < nout; ++i) {                      // freq_in%symmetrize(freq_out) (analytic.f90:262)
    const cplx f = cmake(freq_out[i].re, freq_out[i].im);
    w[i] = freq->freq_symm_coul == 2 ? cmul(f, f) : f;
  }
  begin_call(ctx);
  const long npair = (long)ngc * ngc;
  cplx *d = nullptr, *dz = nullptr, *dw = nullptr, *dout = nullptr;
  int *dg = nullptr;
  SGW_CHECK(ws(ctx, "an_coeff", (size_t)npair * N, &d));
  SGW_CHECK(ws(ctx, "an_z", (size_t)N, &dz));
  SGW_CHECK(ws(ctx, "an_w", (size_t)nout, &dw));
  SGW_CHECK(ws(ctx, "an_out", (size_t)npair * nout, &dout));
  SGW_CHECK(ws(ctx, "an_gmap", (size_t)ngc, &dg));
  SGW_CUDA(cudaMemcpyAsync(d, scrcoul_coeff, sizeof(cplx) * npair * N, cudaMemcpyHostToDevice, ctx->stream));
  SGW_CUDA(cudaMemcpyAsync(dz, z.data(), sizeof(cplx) * N, cudaMemcpyHostToDevice, ctx->stream));
  SGW_CUDA(cudaMemcpyAsync(dw, w.data(), sizeof(cplx) * nout, cudaMemcpyHostToDevice, ctx->stream));
  SGW_CUDA(cudaMemcpyAsync(dg, gmapsym, sizeof(int) * ngc, cudaMemcpyHostToDevice, ctx->stream));
  SGW_CHECK(analytic_eval_dev(ctx, model_coul, ngc, N, dg, dz, d, nout, dw, dout));
  SGW_CUDA(cudaMemcpyAsync(scrcoul, dout, sizeof(cplx) * npair * nout, cudaMemcpyDeviceToHost, ctx->stream));
  SGW_CUDA(cudaStreamSynchronize(ctx->stream));
  end_call(ctx);
  return SGW_OK;
}

int sgw_set_corr_grid(sgw_ctx *ctx, int nr1, int nr2, int nr3, int ngm_c, const int32_t *nl_c) {
  if (!ctx) return SGW_E_ARG;
  cudaSetDevice(ctx->device);
  SGW_ARG(nr1 > 0 && nr2 > 0 && nr3 > 0 && ngm_c > 0 && nl_c, "bad argument");
  const long nnr = (long)nr1 * nr2 * nr3;
  SGW_ARG(nnr <= 46340, "correlation box too large for the DFT-matrix transforms (nnr_c^2 must fit 31 bits)");
  SGW_ARG(ngm_c <= nnr, "more G vectors than box points");
  for (int i = 0; i < ngm_c; ++i) SGW_ARG(nl_c[i] >= 1 && nl_c[i] <= nnr, "nl entry outside the correlation box");
  CorrGrid &c = ctx->corr;
  if (c.d_Ec) { dev_free(c.d_Ec); c.d_Ec = nullptr; }
  if (c.d_ET) { dev_free(c.d_ET); c.d_ET = nullptr; }
  c.set = false;
  c.n1 = nr1; c.n2 = nr2; c.n3 = nr3; c.nnr = (int)nnr; c.ngm = ngm_c;
  SGW_CUDA(dev_malloc((void **)&c.d_Ec, sizeof(cplx) * nnr * ngm_c));
  SGW_CUDA(dev_malloc((void **)&c.d_ET, sizeof(cplx) * nnr * ngm_c));
  int *dnl = nullptr;
  SGW_CHECK(ws(ctx, "sg_nl", (size_t)ngm_c, &dnl));
  SGW_CUDA(cudaMemcpyAsync(dnl, nl_c, sizeof(int) * ngm_c, cudaMemcpyHostToDevice, ctx->stream));
  dim3 grid((unsigned)((nnr + 127) / 128), ngm_c);
  k_corr_tables<<<grid, 128, 0, ctx->stream>>>(nr1, nr2, nr3, ngm_c, dnl, c.d_Ec, c.d_ET);
  ctx->launches++;
  SGW_CUDA(cudaGetLastError());
  SGW_CUDA(cudaStreamSynchronize(ctx->stream));
  c.set = true;
  return SGW_OK;
}

int sgw_invfft6(sgw_ctx *ctx, double omega, sgw_cplx *f) {
  if (!ctx) return SGW_E_ARG;
  cudaSetDevice(ctx->device);
  SGW_ARG(f && omega != 0.0, "bad argument");
  if (!ctx->corr.set) { ctx->err = "sgw_set_corr_grid missing"; return SGW_E_STATE; }
  const CorrGrid &c = ctx->corr;
  begin_call(ctx);
  cplx *fr = nullptr, *fg = nullptr;
  SGW_CHECK(ws(ctx, "sg_fr", (size_t)c.nnr * c.nnr, &fr));
  SGW_CHECK(ws(ctx, "sg_fg", (size_t)c.ngm * c.ngm, &fg));
  SGW_CUDA(cudaMemcpy2DAsync(fg, sizeof(cplx) * c.ngm, f, sizeof(cplx) * c.nnr, sizeof(cplx) * c.ngm, c.ngm, cudaMemcpyHostToDevice,
                             ctx->stream));
  SGW_CHECK(invfft6_dev(ctx, omega, 1, fg, fr));
  SGW_CUDA(cudaMemcpyAsync(f, fr, sizeof(cplx) * c.nnr * c.nnr, cudaMemcpyDeviceToHost, ctx->stream));
  SGW_CUDA(cudaStreamSynchronize(ctx->stream));
  end_call(ctx);
  return SGW_OK;
}

int sgw_fwfft6(sgw_ctx *ctx, double omega, sgw_cplx *f) {
  if (!ctx) return SGW_E_ARG;
  cudaSetDevice(ctx->device);
  SGW_ARG(f != nullptr, "bad argument");
  if (!ctx->corr.set) { ctx->err = "sgw_set_corr_grid missing"; return SGW_E_STATE; }
  const CorrGrid &c = ctx->corr;
  begin_call(ctx);
  cplx *fr = nullptr, *fg = nullptr;
  SGW_CHECK(ws(ctx, "sg_fr", (size_t)c.nnr * c.nnr, &fr));
  SGW_CHECK(ws(ctx, "sg_fg", (size_t)c.ngm * c.ngm, &fg));
  SGW_CUDA(cudaMemcpyAsync(fr, f, sizeof(cplx) * c.nnr * c.nnr, cudaMemcpyHostToDevice, ctx->stream));
  SGW_CHECK(fwfft6_dev(ctx, omega, fr, cmake(0, 0), fg, c.ngm));
  SGW_CUDA(cudaMemcpy2DAsync(f, sizeof(cplx) * c.nnr, fg, sizeof(cplx) * c.ngm, sizeof(cplx) * c.ngm, c.ngm, cudaMemcpyDeviceToHost,
                             ctx->stream));
  SGW_CUDA(cudaStreamSynchronize(ctx->stream));
  end_call(ctx);
  return SGW_OK;
}

int sgw_sigma_correlation(sgw_ctx *ctx, int slot, const sgw_solver_cfg *cfg_green, double omega_cell, double mu, sgw_cplx alpha,
                          int model_coul, const sgw_freqbins *freq, int ngm_c, const int32_t *map, const int32_t *gmapsym,
                          const sgw_cplx *coulomb, sgw_cplx *sigma, int32_t *ierr_out) {
  if (!ctx) return SGW_E_ARG;
  cudaSetDevice(ctx->device);
  SGW_CHECK(check_freq(ctx, freq, model_coul));
  SGW_ARG(cfg_green && map && gmapsym && coulomb && sigma && ierr_out, "null argument");
  SGW_ARG(freq->num_coul > 0 && freq->coul && freq->weight && freq->num_sigma > 0 && freq->sigma, "freqbins: convolution meshes missing");
  SGW_ARG(omega_cell > 0.0, "cell volume must be positive");
  SGW_ARG(alpha.re == alpha.re && alpha.im == alpha.im, "prefactor of the convolution is NaN");       // sigma.f90:466
  if (!ctx->corr.set) { ctx->err = "sgw_set_corr_grid missing"; return SGW_E_STATE; }
  const CorrGrid &c = ctx->corr;
  SGW_ARG(ngm_c == c.ngm, "screened Coulomb and G-vector FFT type inconsistent");                     // sigma.f90:624
  for (int i = 0; i < ngm_c; ++i) SGW_ARG(gmapsym[i] >= 1 && gmapsym[i] <= ngm_c, "gmapsym and FFT type are inconsistent");
  std::vector<cplx> z;
  if (symm_mesh(freq, &z, nullptr, nullptr) != SGW_OK) { ctx->err = "only a single frequency may be smaller than 1e-14"; return SGW_E_ARG; }
  const int N = (int)z.size();
  if (model_coul == SGW_GODBY_NEEDS) SGW_ARG(N == 2, "must provide exactly 2 frequencies");
  const int ncoul = freq->num_coul, nb = 2 * ncoul, nsig = freq->num_sigma;
  const int nnr = c.nnr, ng = c.ngm;
  const long npair = (long)ng * ng;

  // memory: G(r, r', omega_green) stays resident for all omega_sigma
  {
    size_t free_b = 0, total_b = 0, held = 0;
    cudaMemGetInfo(&free_b, &total_b);
    for (auto &kv : ctx->ws.bufs) held += kv.second.second;
    const double need = 16.0 * ((double)nnr * nnr * (nb + 2 + 16) + (double)nnr * ng * nb * 2);
    if (need > 0.9 * (double)(free_b + held)) {
      ctx->err = "G(r,r',omega) of this correlation box does not fit on the device";
      return SGW_E_UNSUPPORTED;
    }
  }

  // ---- Green's function at freq%green(mu) (sigma.f90:650-655), left on the device
  std::vector<sgw_cplx> fgreen(nb);
  for (int i = 0; i < ncoul; ++i) {                                                                   // freqbins_green
    fgreen[i].re = mu + freq->coul[i].re; fgreen[i].im = freq->coul[i].im;
    fgreen[ncoul + i].re = mu - freq->coul[i].re; fgreen[ncoul + i].im = -freq->coul[i].im;
  }
  std::vector<int32_t> fft_map(ng);
  for (int i = 0; i < ng; ++i) fft_map[i] = i + 1;
  cplx *d_green = nullptr;
  SGW_CHECK(green_function_core(ctx, slot, cfg_green, ng, map, ng, fft_map.data(), nb, fgreen.data(), nullptr, ierr_out, &d_green));
  const sgw_stats st_green = ctx->stats;
  if (*ierr_out != 0) return SGW_OK;                  // the reference aborts here (green.f90:208); sigma is left untouched

  begin_call(ctx);
  cudaStream_t st = ctx->stream;
  cplx *d_gr = nullptr, *d_coeff = nullptr, *d_z = nullptr, *d_w = nullptr, *d_cf = nullptr, *d_wg = nullptr, *d_x = nullptr,
       *d_part = nullptr, *d_acc = nullptr, *d_sigma = nullptr;
  int *d_gmap = nullptr;
  SGW_CHECK(ws(ctx, "sg_gr", (size_t)nnr * nnr * nb, &d_gr));
  SGW_CHECK(invfft6_dev(ctx, omega_cell, nb, d_green, d_gr));                                         // sigma.f90:664-666

  // ---- per (omega_sigma, omega_green): frequency of W and weight of the product (sigma.f90:680-704)
  std::vector<cplx> wtab((size_t)nsig * nb), ctab((size_t)nsig * nb);
  for (int is = 0; is < nsig; ++is)
    for (int b = 0; b < nb; ++b) {
      cplx fc = cmake(mu + freq->sigma[is].re - fgreen[b].re, freq->sigma[is].im - fgreen[b].im);
      if (fc.x * fc.y < 0.0) fc = cconj(fc);                                                          // :688
      wtab[(size_t)is * nb + b] = freq->freq_symm_coul == 2 ? cmul(fc, fc) : fc;                      // symmetrize
      const double wgt = freq->weight[b % ncoul];
      ctab[(size_t)is * nb + b] = cmake(alpha.re * wgt, alpha.im * wgt);
    }
  SGW_CHECK(ws(ctx, "an_coeff", (size_t)npair * N, &d_coeff));
  SGW_CHECK(ws(ctx, "an_z", (size_t)N, &d_z));
  SGW_CHECK(ws(ctx, "an_gmap", (size_t)ng, &d_gmap));
  SGW_CHECK(ws(ctx, "sg_w", (size_t)nsig * nb, &d_w));
  SGW_CHECK(ws(ctx, "sg_cf", (size_t)nsig * nb, &d_cf));
  SGW_CHECK(ws(ctx, "sg_wg", (size_t)npair * nb, &d_wg));
  SGW_CHECK(ws(ctx, "sg_x", (size_t)nnr * ng * nb, &d_x));
  SGW_CHECK(ws(ctx, "sg_sigma", (size_t)npair * nsig, &d_sigma));
  SGW_CUDA(cudaMemcpyAsync(d_coeff, coulomb, sizeof(cplx) * npair * N, cudaMemcpyHostToDevice, st));
  SGW_CUDA(cudaMemcpyAsync(d_z, z.data(), sizeof(cplx) * N, cudaMemcpyHostToDevice, st));
  SGW_CUDA(cudaMemcpyAsync(d_gmap, gmapsym, sizeof(int) * ng, cudaMemcpyHostToDevice, st));
  SGW_CUDA(cudaMemcpyAsync(d_w, wtab.data(), sizeof(cplx) * wtab.size(), cudaMemcpyHostToDevice, st));
  SGW_CUDA(cudaMemcpyAsync(d_cf, ctab.data(), sizeof(cplx) * ctab.size(), cudaMemcpyHostToDevice, st));
  SGW_CUDA(cudaMemcpyAsync(d_sigma, sigma, sizeof(cplx) * npair * nsig, cudaMemcpyHostToDevice, st));

  // batch split of the product kernel in whole waves: CTA count = tiles x nz, GW_CTA_PER_SM resident CTAs per SM;
  // a CTA handles ceil(nb / nz) entries, so the time is ~ waves(nz) x ceil(nb / nz) entry-times: take the minimum
  const int tiles = ((nnr + GW_BM - 1) / GW_BM) * ((nnr + GW_BN - 1) / GW_BN);
  int nz = 1, bchunk = nb;
  {
    long best = -1;
    for (int cand = 1; cand <= std::min(nb, 64); ++cand) {
      const int bc = (nb + cand - 1) / cand;
      const int z = (nb + bc - 1) / bc;
      const long slots = (long)ctx->sm_count * GW_CTA_PER_SM;
      const long waves = ((long)tiles * z + slots - 1) / slots;
      const long cost = waves * bc;
      if (best < 0 || cost < best) { best = cost; nz = z; bchunk = bc; }
    }
  }
  SGW_CHECK(ws(ctx, "sg_part", (size_t)nnr * nnr * nz, &d_part));
  SGW_CHECK(ws(ctx, "sg_acc", (size_t)nnr * nnr, &d_acc));
  if (!ctx->gw_attr_set) {
    SGW_CUDA(cudaFuncSetAttribute(k_gw_product, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)GW_SMEM));
    ctx->gw_attr_set = true;
  }
  for (int is = 0; is < nsig; ++is) {                                                                 // sigma.f90:680
    SGW_CHECK(analytic_eval_dev(ctx, model_coul, ng, N, d_gmap, d_z, d_coeff, nb, d_w + (size_t)is * nb, d_wg));
    // first half of invfft6 for all omega_green: X_b = Ec W_b / Omega
    SGW_CHECK(gemm_n_n(ctx, nnr, ng * nb, ng, cmake(1.0 / omega_cell, 0.0), c.d_Ec, nnr, d_wg, ng, cmake(0, 0), d_x, nnr));
    {
      ProfScope prof(ctx, PC_GW_PROD);
      dim3 grid((nnr + GW_BM - 1) / GW_BM, (nnr + GW_BN - 1) / GW_BN, nz);
      k_gw_product<<<grid, GW_T, GW_SMEM, st>>>(nnr, nnr, ng, nb, bchunk, d_x, c.d_ET, d_gr, d_cf + (size_t)is * nb, d_part);
      SGW_LAUNCH_CHECK();
    }
    const cplx *acc = d_part;
    if (nz > 1) {
      ProfScope prof(ctx, PC_OTHER);
      const long n = (long)nnr * nnr;
      k_gw_reduce<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(n, nz, d_part, d_acc);
      SGW_LAUNCH_CHECK();
      acc = d_acc;
    }
    SGW_CHECK(fwfft6_dev(ctx, omega_cell, acc, cmake(1, 0), d_sigma + (size_t)npair * is, ng));      // :717 sigma += work
  }
  SGW_CUDA(cudaMemcpyAsync(sigma, d_sigma, sizeof(cplx) * npair * nsig, cudaMemcpyDeviceToHost, st));
  SGW_CUDA(cudaStreamSynchronize(st));
  end_call(ctx);
  // one call = G + G*W: merge the counters of the Green's-function part
  ctx->stats.n_linear_op += st_green.n_linear_op;
  ctx->stats.n_kernel_launch += st_green.n_kernel_launch;
  ctx->stats.n_outer_max = st_green.n_outer_max;
  ctx->stats.n_fallback = st_green.n_fallback;
  ctx->stats.ms_solver = st_green.ms_solver;
  ctx->stats.ms_linear_op = st_green.ms_linear_op;
  ctx->stats.ms_total += st_green.ms_total;
  return SGW_OK;
}

}  // extern "C"
