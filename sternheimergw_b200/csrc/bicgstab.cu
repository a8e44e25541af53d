// bicgstab.cu -- batched Frommer multishift BiCGStab(l), restating algo/linear_solver/src/bicgstab.f90
// (init_seed :304, init_shift :370, bicg_part :520, mr_part :708, horner_scheme :941, converged :278)
// for nrhs right-hand sides x nshift shifts at once, entirely on the device:
//   * per-RHS / per-(RHS,shift) recurrence scalars live in device memory and are advanced by tiny scalar
//     kernels (one thread per RHS) -- no host round trip inside an outer iteration;
//   * the BLAS-1 work is fused: every seed vector is read once per step and applied to ALL shifts
//     (k_bicg_update), the u^sigma_{j+1} recurrence of L19/L26 is formed without its intermediate store;
//   * dot products / norms: warp-shuffle + shared-memory block reduction into per-chunk partials that the
//     scalar kernels sum in a fixed order (deterministic, unconjugated ZDOTU semantics);
//   * the MR part keeps the reference's modified Gram-Schmidt order (one CTA per RHS, k_mr_mgs);
//   * convergence is tested on the seed residual only against the ABSOLUTE threshold, twice per outer
//     iteration, exactly as bicgstab.f90:237,245; converged RHS are frozen by a device-side active mask.
// The inverse storage of phi/theta, the position of the alpha_old update (:666-670) and the corrected L28
// (:890-894) follow the reference, not Frommer's paper.
#include "internal.cuh"

namespace sgw {

constexpr int LCAP = 16;                       // linear_solver.pf exercises lmax = 1..15
constexpr int MUCAP = LCAP * (LCAP + 1) / 2;
constexpr int TAUCAP = (LCAP + 1) * (LCAP + 2) / 2 + 2;
constexpr int BT = 256;                        // threads of the streaming kernels
constexpr int DOT_EPT = 4;                     // elements per thread in the dot kernels

struct SeedScal {
  cplx sigma, rho, rho_old, alpha, alpha_old, beta, omega;
  cplx gamma[LCAP], gamma_p[LCAP], gamma_pp[LCAP];
  cplx tau[TAUCAP], nu[LCAP + 1];
};

struct ShiftScal {
  cplx sigma, inv_phi_old, inv_phi, inv_phi_new, inv_theta, alpha, beta;
  cplx f_old;      // inv_theta * inv_phi BEFORE the phi update of this step (L15/L19 factor)
  cplx f_new;      // inv_theta * inv_phi AFTER it (L26 factor, applied with a minus sign)
  cplx inv_alpha;  // 1 / alpha^sigma
  cplx psi, inv_xi;
  cplx mu[MUCAP], gamma[LCAP], gamma_p[LCAP], gamma_pp[LCAP];
};

struct BicgState {
  int n, nrhs, ns, L;      // ns = number of shifted systems (nshift - 1)
  cplx *U, *R, *X, *RT;    // seed: U,R [nrhs][L+1][n] ; X, RT [nrhs][n]
  cplx *US, *XS;           // shifts: US [nrhs][ns][L+1][n] ; XS [nrhs][ns][n]
  SeedScal *seed;          // [nrhs]
  ShiftScal *shift;        // [nrhs][ns]
  cplx *part;              // [3][nrhs][nchunk] dot partials
  int nchunk;
  int *active;             // [nrhs]
  int *iters;              // [nrhs] outer iterations done
  int *nactive;            // [1]
};

__device__ __forceinline__ cplx *seedU(const BicgState &s, int b, int i) { return s.U + ((long)b * (s.L + 1) + i) * s.n; }
__device__ __forceinline__ cplx *seedR(const BicgState &s, int b, int i) { return s.R + ((long)b * (s.L + 1) + i) * s.n; }
__device__ __forceinline__ cplx *shiftU(const BicgState &s, int b, int is, int i) {
  return s.US + (((long)b * s.ns + is) * (s.L + 1) + i) * s.n;
}
__device__ __forceinline__ cplx *shiftX(const BicgState &s, int b, int is) { return s.XS + ((long)b * s.ns + is) * s.n; }

// ---------------------------------------------------------------- reductions
__device__ __forceinline__ cplx block_reduce(cplx v, cplx *sm /* >= 32 */) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    v.x += __shfl_xor_sync(0xffffffffu, v.x, o);
    v.y += __shfl_xor_sync(0xffffffffu, v.y, o);
  }
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  __syncthreads();
  if (lane == 0) sm[w] = v;
  __syncthreads();
  if (w == 0) {
    v = lane < nw ? sm[lane] : cmake(0.0, 0.0);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      v.x += __shfl_xor_sync(0xffffffffu, v.x, o);
      v.y += __shfl_xor_sync(0xffffffffu, v.y, o);
    }
    if (lane == 0) sm[0] = v;
  }
  __syncthreads();
  return sm[0];
}

// up to three unconjugated dots per launch: part[q][b][chunk] = sum_e a_q[e] * b_q[e]
// vector selectors: kind 0 = R[i], 1 = U[i], 2 = RT
struct DotSpec {
  int nd;
  int ka[3], ia[3], kb[3], ib[3];
};
__device__ __forceinline__ const cplx *pick(const BicgState &s, int b, int kind, int i) {
  return kind == 0 ? seedR(s, b, i) : (kind == 1 ? seedU(s, b, i) : s.RT + (long)b * s.n);
}
__global__ void __launch_bounds__(BT) k_dots(BicgState s, DotSpec d) {
  const int b = blockIdx.y;
  if (!s.active[b]) return;
  __shared__ cplx sm[32];
  const int base = blockIdx.x * BT * DOT_EPT;
  for (int q = 0; q < d.nd; ++q) {
    const cplx *x = pick(s, b, d.ka[q], d.ia[q]), *y = pick(s, b, d.kb[q], d.ib[q]);
    cplx acc = cmake(0.0, 0.0);
#pragma unroll
    for (int k = 0; k < DOT_EPT; ++k) {
      const int e = base + k * BT + threadIdx.x;
      if (e < s.n) acc = cfma(x[e], y[e], acc);
    }
    acc = block_reduce(acc, sm);
    if (threadIdx.x == 0) s.part[((long)q * s.nrhs + b) * s.nchunk + blockIdx.x] = acc;
  }
}
__device__ __forceinline__ cplx sum_part(const BicgState &s, int q, int b) {
  cplx r = cmake(0.0, 0.0);
  const cplx *p = s.part + ((long)q * s.nrhs + b) * s.nchunk;
  for (int c = 0; c < s.nchunk; ++c) r = cadd(r, p[c]);
  return r;
}

// ---------------------------------------------------------------- init (init_seed / init_shift)
__global__ void k_init_scal(BicgState s, const cplx *__restrict__ sigma /* nshift x nrhs */, const int *__restrict__ todo) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= s.nrhs) return;
  const int nshift = s.ns + 1;
  SeedScal &sd = s.seed[b];
  sd.sigma = sigma[(long)b * nshift];
  sd.rho = cmake(0, 0);
  sd.rho_old = cmake(1, 0);
  sd.alpha_old = cmake(1, 0);
  sd.alpha = cmake(0, 0);
  sd.omega = cmake(1, 0);
  sd.beta = cmake(0, 0);
  s.active[b] = todo ? (todo[b] != 0) : 1;
  s.iters[b] = 0;
  const int L = s.L;
  // binomials (bicgstab.f90:409-433) and mu_ij = C(j,i) sigma^(j-i) (:467-483)
  double binom[MUCAP];
  for (int jj = 0; jj <= L - 1; ++jj) {
    const int off = jj * (jj + 1) / 2;
    for (int ii = 0; ii <= jj; ++ii) binom[off + ii] = ii == 0 ? 1.0 : (binom[off + ii - 1] * (jj - ii + 1)) / ii;
  }
  for (int is = 0; is < s.ns; ++is) {
    ShiftScal &a = s.shift[(long)b * s.ns + is];
    a.inv_phi_old = a.inv_phi = a.inv_theta = cmake(1, 0);
    a.sigma = csub(sigma[(long)b * nshift + is + 1], sd.sigma);
    cplx spow[LCAP];
    spow[0] = cmake(1, 0);
    for (int ii = 1; ii <= L - 1; ++ii) spow[ii] = cmul(a.sigma, spow[ii - 1]);
    for (int jj = 0; jj <= L - 1; ++jj) {
      const int off = jj * (jj + 1) / 2;
      for (int ii = 0; ii <= jj; ++ii) a.mu[off + ii] = cscale(binom[off + ii], spow[jj - ii]);
    }
  }
}

// r0 = rt0 = b ; u0 = 0 ; x = 0 ; shifted u0 = 0, x = 0
__global__ void __launch_bounds__(BT) k_init_vec(BicgState s, const cplx *__restrict__ bvec, long ldb) {
  const int b = blockIdx.y;
  const int e = blockIdx.x * BT + threadIdx.x;
  if (e >= s.n || !s.active[b]) return;
  const cplx v = bvec[(long)b * ldb + e], z = cmake(0.0, 0.0);
  seedR(s, b, 0)[e] = v;
  s.RT[(long)b * s.n + e] = v;
  seedU(s, b, 0)[e] = z;
  s.X[(long)b * s.n + e] = z;
  for (int is = 0; is < s.ns; ++is) {
    shiftU(s, b, is, 0)[e] = z;
    shiftX(s, b, is)[e] = z;
  }
}

// ---------------------------------------------------------------- bicg_part scalars
// L3 (first step only) + L6: rho = (r_j, rt0) ; beta = alpha rho / rho_old ; rho_old = rho
__global__ void k_scal_beta(BicgState s, int jj) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= s.nrhs || !s.active[b]) return;
  SeedScal &sd = s.seed[b];
  if (jj == 0) sd.rho_old = cneg(cmul(sd.omega, sd.rho_old));        // :587
  sd.rho = sum_part(s, 0, b);                                        // :595
  sd.beta = cdiv(cmul(sd.alpha, sd.rho), sd.rho_old);                // :597
  sd.rho_old = sd.rho;                                               // :599
}

// L11: alpha = rho / (u_{j+1}, rt0); L13 scalars of every shift; L18 alpha_old = alpha (after the shift loop)
__global__ void k_scal_alpha(BicgState s) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= s.nrhs || !s.active[b]) return;
  SeedScal &sd = s.seed[b];
  sd.alpha = cdiv(sd.rho, sum_part(s, 0, b));                        // :614
  const cplx one = cmake(1.0, 0.0);
  for (int is = 0; is < s.ns; ++is) {
    ShiftScal &a = s.shift[(long)b * s.ns + is];
    // :629-631
    const cplx ratio = cdiv(a.inv_phi, a.inv_phi_old);
    cplx den = cadd(one, cmul(sd.alpha, a.sigma));
    den = cadd(den, cmul(cdiv(cmul(sd.alpha, sd.beta), sd.alpha_old), csub(ratio, one)));
    a.inv_phi_new = cdiv(a.inv_phi, den);
    a.beta = cmul(cmul(ratio, ratio), sd.beta);                      // :633
    a.alpha = cmul(cdiv(a.inv_phi_new, a.inv_phi), sd.alpha);        // :635
    a.f_old = cmul(a.inv_theta, a.inv_phi);                          // :638
    a.inv_phi_old = a.inv_phi;                                       // :655
    a.inv_phi = a.inv_phi_new;                                       // :657
    a.f_new = cmul(a.inv_theta, a.inv_phi);                          // :693 (sign applied in the update)
    a.inv_alpha = cdiv(one, a.alpha);                                // :695
  }
  sd.alpha_old = sd.alpha;                                           // :670
}

// ---------------------------------------------------------------- bicg_part vector updates
// L8: u_i = r_i - beta u_i (i <= jj)     [ZSCAL(-beta) then ZAXPY(1, r_i)]
__global__ void __launch_bounds__(BT) k_seed_u(BicgState s, int jj) {
  const int b = blockIdx.y;
  const int e = blockIdx.x * BT + threadIdx.x;
  if (e >= s.n || !s.active[b]) return;
  const cplx mbeta = cneg(s.seed[b].beta);
  for (int i = 0; i <= jj; ++i) {
    cplx *u = seedU(s, b, i);
    u[e] = cadd(cmul(mbeta, u[e]), seedR(s, b, i)[e]);
  }
}

// Everything between the two operator applications of step jj plus the post-operator shifted update:
//   shifts: L15 u^_i = f r_i - beta^ u^_i (i<=jj); L17 x^ += alpha^ u^_0; L19+L26 u^_{j+1} without the store
//   seed  : L22 r_i -= alpha u_{i+1} (i<=jj); L25 x += alpha u_0
// (L26 only needs r_j before/after L22 and the updated u^_j, all local to one vector element.)
__global__ void __launch_bounds__(BT) k_bicg_update(BicgState s, int jj) {
  const int b = blockIdx.y;
  if (!s.active[b]) return;
  extern __shared__ cplx ssc[];   // per shift: beta, f_old, alpha, f_new, inv_alpha, sigma
  for (int i = threadIdx.x; i < s.ns; i += BT) {
    const ShiftScal &a = s.shift[(long)b * s.ns + i];
    ssc[6 * i + 0] = a.beta; ssc[6 * i + 1] = a.f_old; ssc[6 * i + 2] = a.alpha;
    ssc[6 * i + 3] = a.f_new; ssc[6 * i + 4] = a.inv_alpha; ssc[6 * i + 5] = a.sigma;
  }
  __syncthreads();
  const int e = blockIdx.x * BT + threadIdx.x;
  if (e >= s.n) return;
  const cplx alpha = s.seed[b].alpha, malpha = cneg(alpha);
  for (int i = 0; i <= jj; ++i) {
    cplx *r = seedR(s, b, i);
    const cplx r_old = r[e];
    const cplx r_new = cfma(malpha, seedU(s, b, i + 1)[e], r_old);          // :676
    r[e] = r_new;
    for (int is = 0; is < s.ns; ++is) {
      const cplx beta_s = ssc[6 * is + 0], f_old = ssc[6 * is + 1];
      cplx *us = shiftU(s, b, is, i);
      cplx u = cmul(cneg(beta_s), us[e]);                                    // :644
      u = cfma(f_old, r_old, u);                                             // :645
      us[e] = u;
      if (i == 0) {
        cplx *xs = shiftX(s, b, is);
        xs[e] = cfma(ssc[6 * is + 2], u, xs[e]);                             // :650
      }
      if (i == jj) {
        cplx t = cmul(f_old, r_old);                                         // :661-662
        t = cfma(cneg(ssc[6 * is + 3]), r_new, t);                           // :693-694
        t = cmul(ssc[6 * is + 4], t);                                        // :695
        t = cfma(cneg(ssc[6 * is + 5]), u, t);                               // :696
        shiftU(s, b, is, jj + 1)[e] = t;
      }
    }
  }
  cplx *x = s.X + (long)b * s.n;
  x[e] = cfma(alpha, seedU(s, b, 0)[e], x[e]);                               // :685
}

// ---------------------------------------------------------------- convergence (converged :278-301)
__global__ void __launch_bounds__(BT) k_norm_r0(BicgState s) {
  const int b = blockIdx.y;
  if (!s.active[b]) return;
  __shared__ cplx sm[32];
  const cplx *r = seedR(s, b, 0);
  const int base = blockIdx.x * BT * DOT_EPT;
  cplx acc = cmake(0.0, 0.0);
#pragma unroll
  for (int k = 0; k < DOT_EPT; ++k) {
    const int e = base + k * BT + threadIdx.x;
    if (e < s.n) { acc.x += r[e].x * r[e].x; acc.y += r[e].y * r[e].y; }
  }
  acc = block_reduce(acc, sm);
  if (threadIdx.x == 0) s.part[((long)2 * s.nrhs + b) * s.nchunk + blockIdx.x] = acc;
}

__global__ void k_check(BicgState s, double threshold, int outer_iter) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= s.nrhs || !s.active[b]) return;
  const cplx p = sum_part(s, 2, b);
  const double nrm = sqrt(p.x + p.y);
  s.iters[b] = outer_iter;
  if (nrm < threshold) s.active[b] = 0;     // :299 strict '<'
  else atomicAdd(s.nactive, 1);
}

// ---------------------------------------------------------------- mr_part
// Modified Gram-Schmidt of r_1..r_L in the bilinear form (:780-801), one CTA per RHS, then the seed
// gamma / gamma'' recurrences (:804-829) and, per shift, horner_scheme + gamma'^ / gamma''^ (:850-884).
__global__ void __launch_bounds__(1024) k_mr_mgs(BicgState s) {
  const int b = blockIdx.x;
  if (!s.active[b]) return;
  __shared__ cplx sm[32];
  __shared__ cplx coef;
  const int L = s.L, n = s.n;
  SeedScal &sd = s.seed[b];
  for (int jj = 1; jj <= L; ++jj) {
    cplx *rj = seedR(s, b, jj);
    const int off = jj * (jj + 1) / 2 + 1;
    for (int ii = 1; ii <= jj - 1; ++ii) {
      const cplx *ri = seedR(s, b, ii);
      cplx acc = cmake(0.0, 0.0);
      for (int e = threadIdx.x; e < n; e += blockDim.x) acc = cfma(rj[e], ri[e], acc);
      acc = block_reduce(acc, sm);
      if (threadIdx.x == 0) {
        const cplx t = cdiv(acc, sd.nu[ii - 1]);                      // :790
        sd.tau[off + ii - 1] = t;
        coef = cneg(t);
      }
      __syncthreads();
      const cplx c = coef;
      for (int e = threadIdx.x; e < n; e += blockDim.x) rj[e] = cfma(c, ri[e], rj[e]);   // :792
      __syncthreads();
    }
    const cplx *r0 = seedR(s, b, 0);
    cplx a1 = cmake(0.0, 0.0), a2 = cmake(0.0, 0.0);
    for (int e = threadIdx.x; e < n; e += blockDim.x) {
      const cplx v = rj[e];
      a1 = cfma(v, v, a1);                                            // :797
      a2 = cfma(r0[e], v, a2);                                        // :799
    }
    a1 = block_reduce(a1, sm);
    a2 = block_reduce(a2, sm);
    if (threadIdx.x == 0) {
      sd.nu[jj - 1] = a1;
      sd.gamma_p[jj - 1] = cdiv(a2, a1);
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    sd.gamma[L - 1] = sd.gamma_p[L - 1];                              // :804
    sd.omega = sd.gamma[L - 1];                                       // :806
    for (int jj = L - 1; jj >= 1; --jj) {                             // :809-818
      cplx g = sd.gamma_p[jj - 1];
      for (int ii = jj + 1; ii <= L; ++ii) g = csub(g, cmul(sd.tau[ii * (ii + 1) / 2 + jj], sd.gamma[ii - 1]));
      sd.gamma[jj - 1] = g;
    }
    for (int jj = 1; jj <= L - 1; ++jj) {                             // :821-829
      cplx g = sd.gamma[jj];
      for (int ii = jj + 1; ii <= L - 1; ++ii) g = cadd(g, cmul(sd.tau[ii * (ii + 1) / 2 + jj], sd.gamma[ii]));
      sd.gamma_pp[jj - 1] = g;
    }
  }
  __syncthreads();
  for (int is = threadIdx.x; is < s.ns; is += blockDim.x) {
    ShiftScal &a = s.shift[(long)b * s.ns + is];
    const cplx msig = cneg(a.sigma);
    // horner_scheme :969-997
    a.gamma[L - 1] = cneg(sd.gamma[L - 1]);
    for (int jj = L - 1; jj >= 1; --jj) a.gamma[jj - 1] = csub(cmul(msig, a.gamma[jj]), sd.gamma[jj - 1]);
    const cplx psi = cadd(cmul(msig, a.gamma[0]), cmake(1.0, 0.0));
    for (int ii = 1; ii <= L - 1; ++ii)
      for (int jj = L - 1; jj >= ii; --jj) a.gamma[jj - 1] = cadd(cmul(msig, a.gamma[jj]), a.gamma[jj - 1]);
    for (int jj = 1; jj <= L; ++jj) a.gamma[jj - 1] = cdiv(cneg(a.gamma[jj - 1]), psi);
    a.psi = psi;
    a.inv_xi = cmul(a.inv_theta, a.inv_phi);                          // :858
    a.inv_theta = cdiv(a.inv_theta, psi);                             // :860
    for (int jj = 1; jj <= L; ++jj) {                                 // :863-872
      cplx g = cmake(0.0, 0.0);
      for (int ii = jj; ii <= L; ++ii) g = cadd(g, cmul(a.mu[(ii - 1) * ii / 2 + jj - 1], a.gamma[ii - 1]));
      a.gamma_p[jj - 1] = g;
    }
    for (int jj = 1; jj <= L - 1; ++jj) {                             // :875-884
      cplx g = a.gamma_p[jj];
      for (int ii = jj + 1; ii <= L - 1; ++ii) g = cadd(g, cmul(sd.tau[ii * (ii + 1) / 2 + jj], a.gamma_p[ii]));
      a.gamma_pp[jj - 1] = g;
    }
  }
}

// L14-L17 seed, L28-L34 shifts, delayed L32 residual update (:833-925) in one pass over the vectors
__global__ void __launch_bounds__(BT) k_mr_update(BicgState s) {
  const int b = blockIdx.y;
  if (!s.active[b]) return;
  const int L = s.L;
  extern __shared__ cplx ssc[];   // seed: gamma[L], gamma_p[L], gamma_pp[L]; per shift: inv_psi, inv_xi*gamma_p1, inv_xi*gamma_pp[L-1]
  const SeedScal &sd = s.seed[b];
  cplx *sg = ssc, *sgp = ssc + L, *sgpp = ssc + 2 * L, *sh = ssc + 3 * L;
  const int per = L + 1;
  for (int i = threadIdx.x; i < L; i += BT) { sg[i] = sd.gamma[i]; sgp[i] = sd.gamma_p[i]; sgpp[i] = i < L - 1 ? sd.gamma_pp[i] : cmake(0, 0); }
  for (int i = threadIdx.x; i < s.ns; i += BT) {
    const ShiftScal &a = s.shift[(long)b * s.ns + i];
    sh[per * i + 0] = cdiv(cmake(1.0, 0.0), a.psi);                               // :910
    sh[per * i + 1] = cmul(a.gamma_p[0], a.inv_xi);                               // :888
    for (int jj = 1; jj <= L - 1; ++jj) sh[per * i + 1 + jj] = cmul(a.gamma_pp[jj - 1], a.inv_xi);   // :904
  }
  __syncthreads();
  const int e = blockIdx.x * BT + threadIdx.x;
  if (e >= s.n) return;
  // seed x and u0
  cplx r0 = seedR(s, b, 0)[e];
  cplx x = cfma(sg[0], r0, s.X[(long)b * s.n + e]);                               // :833
  cplx u0 = cfma(cneg(sg[L - 1]), seedU(s, b, L)[e], seedU(s, b, 0)[e]);          // :835
  for (int jj = 1; jj <= L - 1; ++jj) {
    u0 = cfma(cneg(sg[jj - 1]), seedU(s, b, jj)[e], u0);                          // :841
    x = cfma(sgpp[jj - 1], seedR(s, b, jj)[e], x);                                // :844
  }
  s.X[(long)b * s.n + e] = x;
  seedU(s, b, 0)[e] = u0;
  // shifted systems
  for (int is = 0; is < s.ns; ++is) {
    const cplx *c = sh + per * is;
    cplx xs = cfma(c[1], r0, shiftX(s, b, is)[e]);                                // :889
    cplx us = cfma(cneg(sg[L - 1]), shiftU(s, b, is, L)[e], shiftU(s, b, is, 0)[e]);   // :893
    for (int jj = 1; jj <= L - 1; ++jj) {
      us = cfma(cneg(sg[jj - 1]), shiftU(s, b, is, jj)[e], us);                   // :900
      xs = cfma(c[1 + jj], seedR(s, b, jj)[e], xs);                               // :905
    }
    shiftU(s, b, is, 0)[e] = cmul(c[0], us);                                      // :911
    shiftX(s, b, is)[e] = xs;
  }
  // delayed residual update :919-925
  for (int jj = 1; jj <= L; ++jj) r0 = cfma(cneg(sgp[jj - 1]), seedR(s, b, jj)[e], r0);
  seedR(s, b, 0)[e] = r0;
}

// ---------------------------------------------------------------- finish: copy out (:258-261), NaN scan (:264-267)
__global__ void __launch_bounds__(BT) k_copy_out(BicgState s, cplx *__restrict__ xout, int *__restrict__ ierr, int max_iter,
                                                  const int *__restrict__ todo) {
  const int b = blockIdx.y;
  if (todo && !todo[b]) return;
  const int e = blockIdx.x * BT + threadIdx.x;
  const int nshift = s.ns + 1;
  bool bad = false;
  if (e < s.n) {
    cplx v = s.X[(long)b * s.n + e];
    xout[((long)b * nshift) * s.n + e] = v;
    bad |= (v.x != v.x) || (v.y != v.y);
    for (int is = 0; is < s.ns; ++is) {
      v = shiftX(s, b, is)[e];
      xout[((long)b * nshift + is + 1) * s.n + e] = v;
      bad |= (v.x != v.x) || (v.y != v.y);
    }
  }
  if (__syncthreads_or(bad) && threadIdx.x == 0) atomicMax(&ierr[b], 2);
}

__global__ void k_set_ierr(BicgState s, int *__restrict__ ierr, const int *__restrict__ todo) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= s.nrhs) return;
  if (todo && !todo[b]) return;
  ierr[b] = s.active[b] ? 1 : 0;     // still active after max_iter outer iterations -> ierr = 1 (:249-253)
}

// ---------------------------------------------------------------- driver
int bicgstab_batched(sgw_ctx *ctx, const SolveBatch &sb, int lmax, double threshold, int max_iter, const int *d_todo) {
  if (lmax < 1 || lmax > LCAP - 1) {
    ctx->err = "bicg_lmax must be in 1..15";
    return SGW_E_ARG;
  }
  if (sb.nrhs <= 0) return SGW_OK;
  BicgState s;
  s.n = sb.n; s.nrhs = sb.nrhs; s.ns = sb.nshift - 1; s.L = lmax;
  const long n = s.n, nr = s.nrhs, L1 = lmax + 1;
  SGW_CHECK(ws(ctx, "bi_U", (size_t)(nr * L1 * n), &s.U));
  SGW_CHECK(ws(ctx, "bi_R", (size_t)(nr * L1 * n), &s.R));
  SGW_CHECK(ws(ctx, "bi_X", (size_t)(nr * n), &s.X));
  SGW_CHECK(ws(ctx, "bi_RT", (size_t)(nr * n), &s.RT));
  SGW_CHECK(ws(ctx, "bi_US", (size_t)(nr * s.ns * L1 * n) + 1, &s.US));
  SGW_CHECK(ws(ctx, "bi_XS", (size_t)(nr * s.ns * n) + 1, &s.XS));
  SGW_CHECK(ws(ctx, "bi_seed", (size_t)nr, &s.seed));
  SGW_CHECK(ws(ctx, "bi_shift", (size_t)(nr * s.ns) + 1, &s.shift));
  s.nchunk = (int)((n + BT * DOT_EPT - 1) / (BT * DOT_EPT));
  SGW_CHECK(ws(ctx, "bi_part", (size_t)(3 * nr * s.nchunk), &s.part));
  SGW_CHECK(ws(ctx, "bi_active", (size_t)nr, &s.active));
  SGW_CHECK(ws(ctx, "bi_iters", (size_t)nr, &s.iters));
  SGW_CHECK(ws(ctx, "bi_nactive", (size_t)1, &s.nactive));
  int *h_nactive = nullptr;
  SGW_CUDA(cudaMallocHost((void **)&h_nactive, sizeof(int)));

  cudaStream_t st = ctx->stream;
  const int gb = (int)((nr + 127) / 128);
  const dim3 gvec((unsigned)((n + BT - 1) / BT), (unsigned)nr);
  const dim3 gdot((unsigned)s.nchunk, (unsigned)nr);
  const size_t sm_bicg = (size_t)(6 * s.ns + 1) * sizeof(cplx);
  const size_t sm_mr = (size_t)(3 * lmax + (lmax + 1) * s.ns + 1) * sizeof(cplx);
  if (sm_bicg > 48 * 1024) SGW_CUDA(cudaFuncSetAttribute(k_bicg_update, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm_bicg));
  if (sm_mr > 48 * 1024) SGW_CUDA(cudaFuncSetAttribute(k_mr_update, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm_mr));

  k_init_scal<<<gb, 128, 0, st>>>(s, sb.d_sigma, d_todo);
  SGW_LAUNCH_CHECK();
  k_init_vec<<<gvec, BT, 0, st>>>(s, sb.d_b, sb.ldb);
  SGW_LAUNCH_CHECK();

  const long ldv = n;   // vectors inside U/R are contiguous with stride (L+1)*n between RHS
  int outer_done = 0;
  int rc = SGW_OK;
  for (int iter = 1; iter <= max_iter && rc == SGW_OK; ++iter) {
    // ---- bicg_part
    for (int jj = 0; jj < lmax && rc == SGW_OK; ++jj) {
      DotSpec d; d.nd = 1; d.ka[0] = 0; d.ia[0] = jj; d.kb[0] = 2; d.ib[0] = 0;     // (r_j, rt0)
      k_dots<<<gdot, BT, 0, st>>>(s, d);
      SGW_LAUNCH_CHECK();
      k_scal_beta<<<gb, 128, 0, st>>>(s, jj);
      SGW_LAUNCH_CHECK();
      k_seed_u<<<gvec, BT, 0, st>>>(s, jj);
      SGW_LAUNCH_CHECK();
      // u_{j+1} = A u_j   (bicgstab.f90:611)
      rc = apply_operator(ctx, sb.slot, sb.alpha_pv, s.nrhs, s.U + (long)jj * n, L1 * ldv, &s.seed[0].sigma,
                          sizeof(SeedScal) / sizeof(cplx), s.U + (long)(jj + 1) * n, L1 * ldv, s.active);
      if (rc != SGW_OK) break;
      d.ka[0] = 1; d.ia[0] = jj + 1;                                                // (u_{j+1}, rt0)
      k_dots<<<gdot, BT, 0, st>>>(s, d);
      SGW_LAUNCH_CHECK();
      k_scal_alpha<<<gb, 128, 0, st>>>(s);
      SGW_LAUNCH_CHECK();
      k_bicg_update<<<gvec, BT, sm_bicg, st>>>(s, jj);
      SGW_LAUNCH_CHECK();
      // r_{j+1} = A r_j   (bicgstab.f90:682)
      rc = apply_operator(ctx, sb.slot, sb.alpha_pv, s.nrhs, s.R + (long)jj * n, L1 * ldv, &s.seed[0].sigma,
                          sizeof(SeedScal) / sizeof(cplx), s.R + (long)(jj + 1) * n, L1 * ldv, s.active);
    }
    if (rc != SGW_OK) break;
    ctx->stats.n_linear_op += 0;   // counted from iteration totals after the loop
    SGW_CUDA(cudaMemsetAsync(s.nactive, 0, sizeof(int), st));
    k_norm_r0<<<gdot, BT, 0, st>>>(s);
    SGW_LAUNCH_CHECK();
    k_check<<<gb, 128, 0, st>>>(s, threshold, iter);                                // :237
    SGW_LAUNCH_CHECK();
    // ---- mr_part (RHS that converged above are masked out)
    k_mr_mgs<<<(unsigned)nr, 1024, 0, st>>>(s);
    SGW_LAUNCH_CHECK();
    k_mr_update<<<gvec, BT, sm_mr, st>>>(s);
    SGW_LAUNCH_CHECK();
    SGW_CUDA(cudaMemsetAsync(s.nactive, 0, sizeof(int), st));
    k_norm_r0<<<gdot, BT, 0, st>>>(s);
    SGW_LAUNCH_CHECK();
    k_check<<<gb, 128, 0, st>>>(s, threshold, iter);                                // :245
    SGW_LAUNCH_CHECK();
    SGW_CUDA(cudaMemcpyAsync(h_nactive, s.nactive, sizeof(int), cudaMemcpyDeviceToHost, st));
    SGW_CUDA(cudaStreamSynchronize(st));
    outer_done = iter;
    if (*h_nactive == 0) break;
  }
  cudaFreeHost(h_nactive);
  if (rc != SGW_OK) return rc;
  k_set_ierr<<<gb, 128, 0, st>>>(s, sb.d_ierr, d_todo);
  SGW_LAUNCH_CHECK();
  k_copy_out<<<gvec, BT, 0, st>>>(s, sb.d_x, sb.d_ierr, max_iter, d_todo);
  SGW_LAUNCH_CHECK();
  // statistics: every RHS did 2*L operator applications per outer iteration it took part in
  {
    std::vector<int> it(nr);
    SGW_CUDA(cudaMemcpyAsync(it.data(), s.iters, sizeof(int) * nr, cudaMemcpyDeviceToHost, st));
    SGW_CUDA(cudaStreamSynchronize(st));
    for (long b = 0; b < nr; ++b) {
      ctx->stats.n_linear_op += 2L * lmax * it[b];
      if (it[b] > ctx->stats.n_outer_max) ctx->stats.n_outer_max = it[b];
    }
  }
  (void)outer_done;
  return SGW_OK;
}

}  // namespace sgw
