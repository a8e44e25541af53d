// bicgstab.cu -- batched Frommer multishift BiCGStab(l), restating algo/linear_solver/src/bicgstab.f90
// (init_seed :304, init_shift :370, bicg_part :520, mr_part :708, horner_scheme :941, converged :278)
// for nrhs right-hand sides x nshift shifts at once, entirely on the device:
//   * per-RHS / per-(RHS,shift) recurrence scalars live in device memory and are advanced by tiny scalar
//     kernels (one thread per RHS) -- no host round trip inside an outer iteration;
//   * the shifted systems never feed back into the seed recurrence, so their BLAS-1 work is DEFERRED: the seed
//     part keeps snapshots of the residuals r_i it had at every step, and one fused kernel per outer iteration
//     (k_shift_fused) replays all L BiCG steps and the MR update of every shift with u^sigma_0..u^sigma_L held in
//     registers.  Per shift and outer iteration that is 4 vector passes (read/write u^sigma_0 and x^sigma)
//     instead of the reference's 136 -- u^sigma_1..L never touch memory; arithmetic order per
//     element is the reference's;
//   * dot products / norms: warp-shuffle + shared-memory block reduction into per-chunk partials that the
//     scalar kernels sum in a fixed order (deterministic, unconjugated ZDOTU semantics);
//   * the MR part keeps the reference's modified Gram-Schmidt order (one CTA per RHS, k_mr_mgs);
//   * convergence is tested on the seed residual only against the ABSOLUTE threshold, twice per outer
//     iteration, exactly as bicgstab.f90:237,245; converged RHS are frozen by a device-side active mask.
// The inverse storage of phi/theta, the position of the alpha_old update (:666-670) and the corrected L28
// (:890-894) follow the reference, not Frommer's paper.
#include "internal.cuh"

#include <stdlib.h>
#include <string.h>

#include <algorithm>

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

struct StepScal {          // per (rhs, shift, step jj): what the deferred shifted update of that step needs
  cplx beta, f_old, alpha, f_new, inv_alpha;
};

struct BicgState {
  int n, nrhs, ns, L;      // ns = number of shifted systems (nshift - 1)
  int nsnap;               // L(L+1)/2 + 1 residual snapshots per RHS
  cplx *U, *R, *RT;        // seed: U,R [nrhs][L+1][n] ; RT [nrhs][n]
  cplx *XO;                // solutions, in the caller's buffer: [(rhs*(ns+1) + is)*n], is = 0 is the seed system
  cplx *US0;               // shifts: u^sigma_0 [nrhs][ns][n]
  cplx *SNAP;              // [nrhs][nsnap][n]: r_i before the update of step jj at jj(jj+1)/2+i, last = r_{L-1} after step L-1
  SeedScal *seed;          // [nrhs]
  ShiftScal *shift;        // [nrhs][ns]
  StepScal *step;          // [nrhs][ns][L]
  int *stage;              // [nrhs] this outer iteration: 0 idle, 1 BiCG part only (converged at :237), 2 BiCG + MR
  cplx *part;              // [3][nrhs][nchunk] dot partials
  int nchunk;
  int *active;             // [nrhs]
  int *iters;              // [nrhs] outer iterations done
  int *nactive;            // [1]
  // ---- lazy (coefficient-space) shifted update, see k_shift_coef / k_shift_gemm
  int lazy;                // 1: the shifted systems are carried as coefficients over the stored basis
  int Tc, slot, base;      // basis slots per RHS, slot of this outer iteration, outer iteration of the last flush
  int nv, kcap;            // vectors per slot (nsnap + L), Tc * nv
  cplx *BAS;               // [nrhs][Tc][nv][n]: per outer iteration the nsnap residual snapshots, then r_0..r_{L-1} of the MR part
  cplx *CX, *CU;           // [nrhs][kcap][ns] (shift index fastest): x^sigma = XB + cxb U0B + sum_k CX_k BAS_k ; u^sigma_0 = cub U0B + sum_k CU_k BAS_k
  cplx *cxb, *cub;         // [nrhs][ns]
  int *flushed;            // [nrhs] XB (= XO) and U0B (= US0) hold a materialised state
  int *last_stage;         // [nrhs] stage of the last outer iteration the RHS took part in
  int *base_b;             // [nrhs] outer iteration of the last flush of THIS right-hand side (only active ones are flushed)
  // ---- +-omega average inside the solver (AvgSpec): y = pair averages of x^sigma, nothing else is materialised
  int avg;                 // 1: on
  int a_nfreq, a_zero, a_group;
  cplx *Y, *CA;            // output (AvgSpec layout) ; combined coefficients [nrhs][kcap][a_nfreq]
  int *ierr;               // [nrhs] for the NaN check of the materialised vectors (bicgstab.f90:264-267)
};

__device__ __forceinline__ cplx *seedU(const BicgState &s, int b, int i) { return s.U + ((long)b * (s.L + 1) + i) * s.n; }
__device__ __forceinline__ cplx *seedR(const BicgState &s, int b, int i) { return s.R + ((long)b * (s.L + 1) + i) * s.n; }
__device__ __forceinline__ cplx *seedX(const BicgState &s, int b) { return s.XO + ((long)b * (s.ns + 1)) * s.n; }
__device__ __forceinline__ cplx *shiftU0(const BicgState &s, int b, int is) { return s.US0 + ((long)b * s.ns + is) * s.n; }
__device__ __forceinline__ cplx *shiftX(const BicgState &s, int b, int is) { return s.XO + ((long)b * (s.ns + 1) + is + 1) * s.n; }
__device__ __forceinline__ cplx *snap(const BicgState &s, int b, int k) {
  if (s.lazy) return s.BAS + (((long)b * s.Tc + s.slot) * s.nv + k) * s.n;
  return s.SNAP + ((long)b * s.nsnap + k) * s.n;
}

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
// init_shift = 0: u^sigma_0 and x^sigma are NOT zeroed here -- the collapsed shifted update treats them as zero in the first
// outer iteration of a right-hand side instead of reading them (saves writing and re-reading 2 ns vectors per RHS)
__global__ void __launch_bounds__(BT) k_init_vec(BicgState s, const cplx *__restrict__ bvec, long ldb, int init_shift) {
  const int b = blockIdx.y;
  const int e = blockIdx.x * BT + threadIdx.x;
  if (e >= s.n || !s.active[b]) return;
  const cplx v = bvec[(long)b * ldb + e], z = cmake(0.0, 0.0);
  seedR(s, b, 0)[e] = v;
  s.RT[(long)b * s.n + e] = v;
  seedU(s, b, 0)[e] = z;
  seedX(s, b)[e] = z;
  if (!init_shift) return;
  for (int is = 0; is < s.ns; ++is) {
    shiftU0(s, b, is)[e] = z;
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

// L11: alpha = rho / (u_{j+1}, rt0); L13 scalars of every shift (kept per step for the deferred shifted update);
// L18 alpha_old = alpha (after the shift loop).  One CTA per RHS, threads over shifts.
__global__ void __launch_bounds__(128) k_scal_alpha(BicgState s, int jj) {
  const int b = blockIdx.x;
  if (!s.active[b]) return;
  SeedScal &sd = s.seed[b];
  __shared__ cplx sh_alpha;
  if (threadIdx.x == 0) sh_alpha = cdiv(sd.rho, sum_part(s, 0, b));   // :614
  __syncthreads();
  const cplx alpha = sh_alpha, beta = sd.beta, alpha_old = sd.alpha_old;
  const cplx one = cmake(1.0, 0.0);
  for (int is = threadIdx.x; is < s.ns; is += blockDim.x) {
    ShiftScal &a = s.shift[(long)b * s.ns + is];
    StepScal &t = s.step[((long)b * s.ns + is) * s.L + jj];
    // :629-631
    const cplx ratio = cdiv(a.inv_phi, a.inv_phi_old);
    cplx den = cadd(one, cmul(alpha, a.sigma));
    den = cadd(den, cmul(cdiv(cmul(alpha, beta), alpha_old), csub(ratio, one)));
    a.inv_phi_new = cdiv(a.inv_phi, den);
    a.beta = cmul(cmul(ratio, ratio), beta);                         // :633
    a.alpha = cmul(cdiv(a.inv_phi_new, a.inv_phi), alpha);           // :635
    a.f_old = cmul(a.inv_theta, a.inv_phi);                          // :638
    a.inv_phi_old = a.inv_phi;                                       // :655
    a.inv_phi = a.inv_phi_new;                                       // :657
    a.f_new = cmul(a.inv_theta, a.inv_phi);                          // :693 (sign applied in the update)
    a.inv_alpha = cdiv(one, a.alpha);                                // :695
    t.beta = a.beta; t.f_old = a.f_old; t.alpha = a.alpha; t.f_new = a.f_new; t.inv_alpha = a.inv_alpha;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    sd.alpha = alpha;
    sd.alpha_old = alpha;                                            // :670
  }
}

// ---------------------------------------------------------------- bicg_part vector updates
// L8: u_i = r_i - beta u_i (i <= jj)     [ZSCAL(-beta) then ZAXPY(1, r_i)]
// JT >= 0: step index known at compile time (L <= 4): the loop unrolls and all loads of a thread are issued before the
// first store (more bytes in flight for these pure HBM streams); JT = -1: generic
template <int JT>
__global__ void __launch_bounds__(BT) k_seed_u(BicgState s, int jj_rt) {
  const int jj = JT >= 0 ? JT : jj_rt;
  const int b = blockIdx.y;
  const int e = blockIdx.x * BT + threadIdx.x;
  if (e >= s.n || !s.active[b]) return;
  const cplx mbeta = cneg(s.seed[b].beta);
  if (JT >= 0) {
    cplx uu[JT >= 0 ? JT + 1 : 1], rr[JT >= 0 ? JT + 1 : 1];
#pragma unroll
    for (int i = 0; i <= JT; ++i) { uu[i] = seedU(s, b, i)[e]; rr[i] = seedR(s, b, i)[e]; }
#pragma unroll
    for (int i = 0; i <= JT; ++i) seedU(s, b, i)[e] = cadd(cmul(mbeta, uu[i]), rr[i]);
    return;
  }
  for (int i = 0; i <= jj; ++i) {
    cplx *u = seedU(s, b, i);
    u[e] = cadd(cmul(mbeta, u[e]), seedR(s, b, i)[e]);
  }
}
template <int JT>
__global__ void __launch_bounds__(BT) k_bicg_update(BicgState s, int jj_rt) {
  const int jj = JT >= 0 ? JT : jj_rt;
  const int b = blockIdx.y;
  if (!s.active[b]) return;
  const int e = blockIdx.x * BT + threadIdx.x;
  if (e >= s.n) return;
  const cplx alpha = s.seed[b].alpha, malpha = cneg(alpha);
  const bool keep = s.ns > 0;
  if (JT >= 0) {
    cplx ro[JT >= 0 ? JT + 1 : 1], un[JT >= 0 ? JT + 2 : 1];
#pragma unroll
    for (int i = 0; i <= JT + 1; ++i) un[i] = seedU(s, b, i)[e];
#pragma unroll
    for (int i = 0; i <= JT; ++i) ro[i] = seedR(s, b, i)[e];
    const cplx x_old = seedX(s, b)[e];
#pragma unroll
    for (int i = 0; i <= JT; ++i) {
      const cplx r_new = cfma(malpha, un[i + 1], ro[i]);                      // :676
      seedR(s, b, i)[e] = r_new;
      if (keep) {
        snap(s, b, JT * (JT + 1) / 2 + i)[e] = ro[i];
        if (JT == s.L - 1 && i == JT) snap(s, b, s.nsnap - 1)[e] = r_new;
      }
    }
    seedX(s, b)[e] = cfma(alpha, un[0], x_old);                                // :685
    return;
  }
  for (int i = 0; i <= jj; ++i) {
    cplx *r = seedR(s, b, i);
    const cplx r_old = r[e];
    const cplx r_new = cfma(malpha, seedU(s, b, i + 1)[e], r_old);          // :676
    r[e] = r_new;
    if (keep) {
      snap(s, b, jj * (jj + 1) / 2 + i)[e] = r_old;
      if (jj == s.L - 1 && i == jj) snap(s, b, s.nsnap - 1)[e] = r_new;
    }
  }
  cplx *x = seedX(s, b);
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

// phase 1 = after bicg_part (:237), phase 2 = after mr_part (:245)
__global__ void k_check(BicgState s, double threshold, int outer_iter, int phase) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= s.nrhs) return;
  if (!s.active[b]) {
    if (phase == 1) s.stage[b] = 0;
    return;
  }
  const cplx p = sum_part(s, 2, b);
  const double nrm = sqrt(p.x + p.y);
  s.iters[b] = outer_iter;
  const bool conv = nrm < threshold;        // :299 strict '<'
  if (phase == 1) {
    s.stage[b] = conv ? 1 : 2;
    if (s.lazy) s.last_stage[b] = conv ? 1 : 2;
  }
  if (conv) s.active[b] = 0;
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

// Deferred update of the shifted systems for one whole outer iteration (bicg_part :621-664,:688-698 for every
// step jj, then mr_part :850-911), one thread per vector element, u^sigma_0..u^sigma_L in registers.
// Must run after k_mr_mgs (needs r_1..r_{L-1} after the Gram-Schmidt) and BEFORE k_mr_seed (needs r_0 before :919-925).
// LT > 0: compile-time L with the residual snapshots preloaded into registers; LT == 0: any L <= LCAP-1.
constexpr int SHIFT_CHUNK = 32;
template <int LT>
__global__ void __launch_bounds__(BT) k_shift_fused(BicgState s, int chunk) {
  const int b = blockIdx.y;
  const int stage = s.stage[b];
  if (stage == 0) return;
  const int L = LT > 0 ? LT : s.L;
  const int is0 = blockIdx.z * chunk, nsl = min(chunk, s.ns - is0);
  extern __shared__ cplx ssc[];   // per local shift: [5*L step scalars][sigma][1/psi][inv_xi*gamma_p1][inv_xi*gamma_pp 1..L-1]
  const int per = 5 * L + 2 + L;
  for (int t = threadIdx.x; t < nsl * L; t += BT) {
    const int il = t / L, jj = t % L;
    const StepScal &q = s.step[((long)b * s.ns + is0 + il) * L + jj];
    cplx *d = ssc + il * per + 5 * jj;
    d[0] = q.beta; d[1] = q.f_old; d[2] = q.alpha; d[3] = q.f_new; d[4] = q.inv_alpha;
  }
  for (int il = threadIdx.x; il < nsl; il += BT) {
    const ShiftScal &a = s.shift[(long)b * s.ns + is0 + il];
    cplx *d = ssc + il * per + 5 * L;
    d[0] = a.sigma;
    if (stage == 2) {
      d[1] = cdiv(cmake(1.0, 0.0), a.psi);                                        // :910
      d[2] = cmul(a.gamma_p[0], a.inv_xi);                                        // :888
      for (int jj = 1; jj <= L - 1; ++jj) d[2 + jj] = cmul(a.gamma_pp[jj - 1], a.inv_xi);   // :904
    }
  }
  __shared__ cplx sgam[LCAP];
  for (int i = threadIdx.x; i < L; i += BT) sgam[i] = s.seed[b].gamma[i];
  __syncthreads();
  const int e = blockIdx.x * BT + threadIdx.x;
  if (e >= s.n) return;
  constexpr int NSN = LT > 0 ? LT * (LT + 1) / 2 + 1 : 1;
  constexpr int NRR = LT > 0 ? LT : 1;
  constexpr int NUS = LT > 0 ? LT + 1 : LCAP;
  cplx sn[NSN], rr[NRR];
  const cplx *snapg = snap(s, b, 0) + e;
  const long n = s.n;
  if (LT > 0) {
#pragma unroll
    for (int k = 0; k < NSN; ++k) sn[k] = snapg[(long)k * n];
    if (stage == 2) {
#pragma unroll
      for (int j = 0; j < NRR; ++j) rr[j] = seedR(s, b, j)[e];
    }
  }
  const int last = LT > 0 ? NSN - 1 : s.nsnap - 1;
  for (int il = 0; il < nsl; ++il) {
    const cplx *sc = ssc + il * per;
    const cplx sigma = sc[5 * L];
    cplx *pu0 = shiftU0(s, b, is0 + il) + e, *px = shiftX(s, b, is0 + il) + e;
    cplx us[NUS];
    us[0] = *pu0;
    cplx xs = *px;
#pragma unroll
    for (int jj = 0; jj < (LT > 0 ? LT : L); ++jj) {
      const cplx beta_s = sc[5 * jj], f_old = sc[5 * jj + 1], alpha_s = sc[5 * jj + 2], f_new = sc[5 * jj + 3],
                 inv_alpha = sc[5 * jj + 4];
#pragma unroll
      for (int i = 0; i <= jj; ++i) {
        const int k = jj * (jj + 1) / 2 + i;
        const cplx r_old = LT > 0 ? sn[LT > 0 ? k : 0] : snapg[(long)k * n];
        cplx u = cmul(cneg(beta_s), us[i]);                                      // :644
        u = cfma(f_old, r_old, u);                                               // :645
        us[i] = u;
        if (i == 0) xs = cfma(alpha_s, u, xs);                                   // :650
        if (i == jj) {
          const int kn = (jj < L - 1) ? (jj + 1) * (jj + 2) / 2 + jj : last;
          const cplx r_new = LT > 0 ? sn[LT > 0 ? kn : 0] : snapg[(long)kn * n];
          cplx t = cmul(f_old, r_old);                                           // :661-662
          t = cfma(cneg(f_new), r_new, t);                                       // :693-694
          t = cmul(inv_alpha, t);                                                // :695
          t = cfma(cneg(sigma), u, t);                                           // :696
          us[jj + 1] = t;
        }
      }
    }
    if (stage == 2) {
      const cplx *c = sc + 5 * L + 1;    // c[0] = 1/psi, c[1] = inv_xi gamma_p1, c[1+jj] = inv_xi gamma_pp_jj
      const cplx r0 = LT > 0 ? rr[0] : seedR(s, b, 0)[e];
      xs = cfma(c[1], r0, xs);                                                   // :889
      cplx u0 = cfma(cneg(sgam[L - 1]), us[L], us[0]);                           // :893
#pragma unroll
      for (int jj = 1; jj <= (LT > 0 ? LT : L) - 1; ++jj) {
        u0 = cfma(cneg(sgam[jj - 1]), us[jj], u0);                               // :900
        const cplx rj = LT > 0 ? rr[LT > 0 ? jj : 0] : seedR(s, b, jj)[e];
        xs = cfma(c[1 + jj], rj, xs);                                            // :905
      }
      *pu0 = cmul(c[0], u0);                                                     // :911
    }
    *px = xs;
  }
}

// ---------------------------------------------------------------- collapsed shifted update (L <= 4)
// The deferred shifted update is LINEAR in (u^sigma_0, the L(L+1)/2+1 residual snapshots, r_0..r_{L-1}): replaying
// the recurrences of k_shift_fused on coefficient vectors instead of on vector elements (k_shift_coef, one thread per
// (rhs, shift)) gives   x^sigma += sum_k X_k basis_k ,  u^sigma_0 <- sum_k C_k basis_k .  Per element and shift that is
// 1 + 2L terms for x and 2 + L(L+1)/2 terms for u_0 (84 FP64 FMAs at L = 4 instead of 196), which turns the update from
// FP64-pipe-bound (ncu: 62 % FP64, 42 % DRAM) into an HBM-bound stream.  Algebraically identical to the reference's
// recurrences (bicgstab.f90:621-664,:688-698,:850-911), re-associated; SGW_SHIFT=faithful keeps the step-by-step kernel.
template <int LT>
struct ShiftCoef {
  static constexpr int NSN = LT * (LT + 1) / 2 + 1;
  cplx xu0;            // x  += xu0 * u0_old
  cplx xs[LT];         //     + xs[jj] * snap[jj(jj+1)/2]        (the i = 0 snapshots)
  cplx xr[LT];         //     + xr[j] * r_j                      (stage 2 only)
  cplx uu0;            // u0 <- uu0 * u0_old + sum_k us[k] * snap[k]   (stage 2 only)
  cplx us[NSN];
};

template <int LT>
__global__ void __launch_bounds__(128) k_shift_coef(BicgState s, ShiftCoef<LT> *__restrict__ out) {
  constexpr int NSN = LT * (LT + 1) / 2 + 1, NB = 1 + NSN;   // basis: u0_old, snap[0..NSN-1]
  const long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long)s.nrhs * s.ns) return;
  const int b = (int)(idx / s.ns), is = (int)(idx % s.ns);
  const int stage = s.stage[b];
  if (stage == 0) return;
  const ShiftScal &a = s.shift[(long)b * s.ns + is];
  const StepScal *q = s.step + ((long)b * s.ns + is) * LT;
  const cplx zero = cmake(0.0, 0.0), sigma = a.sigma;
  cplx US[LT + 1][NB], X[NB];
#pragma unroll
  for (int i = 0; i <= LT; ++i)
#pragma unroll
    for (int k = 0; k < NB; ++k) US[i][k] = zero;
#pragma unroll
  for (int k = 0; k < NB; ++k) X[k] = zero;
  US[0][0] = cmake(1.0, 0.0);
#pragma unroll
  for (int jj = 0; jj < LT; ++jj) {
    const cplx beta_s = q[jj].beta, f_old = q[jj].f_old, alpha_s = q[jj].alpha, f_new = q[jj].f_new, inv_alpha = q[jj].inv_alpha;
#pragma unroll
    for (int i = 0; i <= jj; ++i) {
      const int k = jj * (jj + 1) / 2 + i;
#pragma unroll
      for (int c = 0; c < NB; ++c) US[i][c] = cmul(cneg(beta_s), US[i][c]);              // :644
      US[i][1 + k] = cadd(US[i][1 + k], f_old);                                        // :645
      if (i == 0) {
#pragma unroll
        for (int c = 0; c < NB; ++c) X[c] = cfma(alpha_s, US[0][c], X[c]);              // :650
      }
      if (i == jj) {
        const int kn = (jj < LT - 1) ? (jj + 1) * (jj + 2) / 2 + jj : NSN - 1;
        cplx T[NB];
#pragma unroll
        for (int c = 0; c < NB; ++c) T[c] = zero;
        T[1 + k] = f_old;                                                              // :661-662
        T[1 + kn] = csub(T[1 + kn], f_new);                                            // :693-694
#pragma unroll
        for (int c = 0; c < NB; ++c) {
          T[c] = cmul(inv_alpha, T[c]);                                                // :695
          US[jj + 1][c] = cfma(cneg(sigma), US[i][c], T[c]);                           // :696
        }
      }
    }
  }
  ShiftCoef<LT> &o = out[idx];
  o.xu0 = X[0];
#pragma unroll
  for (int jj = 0; jj < LT; ++jj) o.xs[jj] = X[1 + jj * (jj + 1) / 2];
  if (stage == 2) {
    const SeedScal &sd = s.seed[b];
    const cplx inv_psi = cdiv(cmake(1.0, 0.0), a.psi);                                  // :910
    o.xr[0] = cmul(a.gamma_p[0], a.inv_xi);                                            // :888
#pragma unroll
    for (int jj = 1; jj <= LT - 1; ++jj) o.xr[jj] = cmul(a.gamma_pp[jj - 1], a.inv_xi); // :904
    cplx U0[NB];
#pragma unroll
    for (int c = 0; c < NB; ++c) {
      cplx u0 = cfma(cneg(sd.gamma[LT - 1]), US[LT][c], US[0][c]);                     // :893
#pragma unroll
      for (int jj = 1; jj <= LT - 1; ++jj) u0 = cfma(cneg(sd.gamma[jj - 1]), US[jj][c], u0);   // :900
      U0[c] = cmul(inv_psi, u0);                                                       // :911
    }
    o.uu0 = U0[0];
#pragma unroll
    for (int k = 0; k < NSN; ++k) o.us[k] = U0[1 + k];
  }
  if (!s.lazy) return;
  // ---- lazy mode: fold this outer iteration into the coefficients over the stored basis instead of streaming the vectors
  //   x  += xu0 u0_old + sum_jj xs[jj] snap[jj(jj+1)/2] (+ sum_j xr[j] r_j)        u0 <- uu0 u0_old + sum_k us[k] snap[k]
  // with u0_old = cub U0B + sum_{k < K0} CU_k BAS_k (K0 = slot * nv entries of the earlier slots)
  // coefficient arrays are [rhs][k][shift]: the threads of a warp (consecutive shifts of one RHS) touch consecutive entries
  const long ldc = s.ns;
  cplx *cx = s.CX + (long)b * s.kcap * ldc + is, *cu = s.CU + (long)b * s.kcap * ldc + is;
  const int K0 = s.slot * s.nv;
  const cplx xu0 = o.xu0;
  s.cxb[idx] = cfma(xu0, s.cub[idx], s.cxb[idx]);
  for (int k = 0; k < K0; ++k) cx[k * ldc] = cfma(xu0, cu[k * ldc], cx[k * ldc]);
  for (int k = 0; k < s.nv; ++k) cx[(K0 + k) * ldc] = zero;
#pragma unroll
  for (int jj = 0; jj < LT; ++jj) cx[(K0 + jj * (jj + 1) / 2) * ldc] = o.xs[jj];
  if (stage == 2) {
#pragma unroll
    for (int j = 0; j < LT; ++j) cx[(K0 + NSN + j) * ldc] = o.xr[j];
    const cplx uu0 = o.uu0;
    s.cub[idx] = cmul(uu0, s.cub[idx]);
    for (int k = 0; k < K0; ++k) cu[k * ldc] = cmul(uu0, cu[k * ldc]);
#pragma unroll
    for (int k = 0; k < NSN; ++k) cu[(K0 + k) * ldc] = o.us[k];
#pragma unroll
    for (int j = 0; j < LT; ++j) cu[(K0 + NSN + j) * ldc] = zero;
  } else {
    for (int k = 0; k < s.nv; ++k) cu[(K0 + k) * ldc] = zero;
  }
}

// ---------------------------------------------------------------- lazy shifted update: materialisation
// X_b (n x ns) = BAS_b (n x K_b) CX_b (K_b x ns)  [+ XB + U0B diag(cxb) when the RHS has been flushed before], and at a flush
// also U0_b = BAS_b CU_b + U0B diag(cub): one batched complex DMMA GEMM over the right-hand sides (3M product, cp.async double
// buffer, the tiling of k_zgemm<false,false> in gemm.cu).  K_b = basis vectors of RHS b since the last flush.
constexpr int LG_BM = 64, LG_BN = 32, LG_BK = 16, LG_T = 256, LG_PK = LG_BK + 4, LG_PM = LG_BM + 2;
constexpr size_t LG_SMEM = 2 * (size_t)(LG_BK * LG_PM + LG_BN * LG_PK) * sizeof(cplx);
// what = 0: x^sigma at the end (every right-hand side that has unmaterialised iterations) ; 1: u^sigma_0 and 2: x^sigma at a flush
// (only the right-hand sides that are still active: the others keep their coefficients until the end).  K is short (15 vectors per outer iteration), so one CTA walks LG_TM consecutive 64-row tiles
// times all column tiles of its right-hand side in ONE continuous cp.async pipeline: the first chunk of the next tile is in flight
// while the last chunk of the current one is multiplied, and the fill / drain of the pipeline is paid once per CTA, not per tile.
constexpr int LG_TM = 8;
__device__ __forceinline__ cplx *avg_col(const BicgState &s, int b, int f) {
  return s.Y + (((long)(b / s.a_group) * s.a_nfreq + f) * s.a_group + b % s.a_group) * s.n;
}
// coefficient columns of the pair averages: CA[b][k][f] = 1/2 CX[k][f - 1] + 1/2 CX[k][nfreq + f - first - 1]  (global shift g <-> is = g - 1;
// the seed system g = 0 has no coefficients: its x is added in the epilogue of the final materialisation)
__global__ void k_lazy_combine(BicgState s, int what) {
  const int b = blockIdx.y;
  const int itb = s.iters[b], baseb = s.base_b[b];
  if (itb <= baseb || (what != 0 && !s.active[b])) return;
  const int K = s.nv * (itb - baseb - 1) + (s.last_stage[b] == 2 ? s.nv : s.nsnap);
  const int nf = s.a_nfreq, first = s.a_zero ? 1 : 0;
  const cplx *cx = s.CX + (long)b * s.ns * s.kcap;
  cplx *ca = s.CA + (long)b * nf * s.kcap;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < K * nf; i += gridDim.x * blockDim.x) {
    const int k = i / nf, f = i - k * nf;
    cplx v = cmake(0.0, 0.0);
    if (f >= first) {
      const int g2 = nf + f - first;
      const cplx w = cx[(long)k * s.ns + g2 - 1];
      v = cmake(0.5 * w.x, 0.5 * w.y);
      if (f >= 1) { const cplx u = cx[(long)k * s.ns + f - 1]; v = cmake(v.x + 0.5 * u.x, v.y + 0.5 * u.y); }
    }
    ca[(long)k * nf + f] = v;
  }
}
__global__ void __launch_bounds__(LG_T, 3) k_shift_gemm(BicgState s, int what) {
  const int b = blockIdx.y;
  const int itb = s.iters[b], baseb = s.base_b[b];
  if (itb <= baseb || (what != 0 && !s.active[b])) return;
  const int K = s.nv * (itb - baseb - 1) + (s.last_stage[b] == 2 ? s.nv : s.nsnap);
  const bool avg = s.avg && what != 1;
  const int M = s.n, N = avg ? s.a_nfreq : s.ns;
  const int ntn = (N + LG_BN - 1) / LG_BN, ntm_all = (M + LG_BM - 1) / LG_BM;
  const int mt0 = blockIdx.x * LG_TM, ntm = min(LG_TM, ntm_all - mt0);
  const int nk = (K + LG_BK - 1) / LG_BK;
  const int ntile = ntm * ntn, nstep = ntile * nk;
  const cplx *A = s.BAS + (long)b * s.Tc * s.nv * s.n;                       // M x K, column k at A + k n
  const cplx *B = avg ? s.CA + (long)b * s.a_nfreq * s.kcap                  // K x nfreq stored [k][f]
                      : (what != 1 ? s.CX : s.CU) + (long)b * s.ns * s.kcap; // K x N stored [k][is]
  const long lda = s.n, ldb = N;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  extern __shared__ cplx lgsm[];
  constexpr int ASZ = LG_BK * LG_PM, BSZ = LG_BN * LG_PK;
  cplx *As = lgsm, *Bs = lgsm + 2 * ASZ;
  const int wm = (warp & 3) * 16, wn = (warp >> 2) * 16;
  const int g = lane >> 2, t = lane & 3;
  double p1[2][2][2], p2[2][2][2], p3[2][2][2];
  auto stage_load = [&](int st, int step) {
    const int tile = step / nk, k0 = (step - tile * nk) * LG_BK;
    const int m0 = (mt0 + tile / ntn) * LG_BM, n0 = (tile % ntn) * LG_BN;
    cplx *as = As + st * ASZ, *bs = Bs + st * BSZ;
#pragma unroll
    for (int r = 0; r < LG_BM * LG_BK / LG_T; ++r) {
      const int i = tid + r * LG_T;
      const int m = i % LG_BM, k = i / LG_BM;
      const bool ok = (k0 + k < K) && (m0 + m < M);
      cp_async16(as + k * LG_PM + m, ok ? A + (long)(m0 + m) + (long)(k0 + k) * lda : A, ok ? 16 : 0);
    }
#pragma unroll
    for (int r = 0; r < LG_BN * LG_BK / LG_T; ++r) {
      const int i = tid + r * LG_T;
      const int nn = i % LG_BN, k = i / LG_BN;
      const bool ok = (k0 + k < K) && (n0 + nn < N);
      cp_async16(bs + nn * LG_PK + k, ok ? B + (long)(k0 + k) * ldb + (long)(n0 + nn) : B, ok ? 16 : 0);
    }
  };
  const bool fl = s.flushed[b] != 0;
  cplx *C = what != 1 ? s.XO + ((long)b * (s.ns + 1) + 1) * s.n : s.US0 + (long)b * s.ns * s.n;
  const cplx *U0B = s.US0 + (long)b * s.ns * s.n;
  if (nstep > 0) stage_load(0, 0);
  cp_async_commit();
  for (int step = 0; step < nstep; ++step) {
    const int tile = step / nk, kc = step - tile * nk;
    if (kc == 0) {
#pragma unroll
      for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int j = 0; j < 2; ++j)
#pragma unroll
          for (int c = 0; c < 2; ++c) p1[i][j][c] = p2[i][j][c] = p3[i][j][c] = 0.0;
    }
    if (step + 1 < nstep) stage_load((step + 1) & 1, step + 1);
    cp_async_commit();
    cp_async_wait<1>();
    __syncthreads();
    const cplx *as = As + (step & 1) * ASZ, *bs = Bs + (step & 1) * BSZ;
#pragma unroll
    for (int kk = 0; kk < LG_BK; kk += 4) {
      cplx a[2], bb[2];
#pragma unroll
      for (int i = 0; i < 2; ++i) a[i] = as[(kk + t) * LG_PM + wm + i * 8 + g];
#pragma unroll
      for (int j = 0; j < 2; ++j) bb[j] = bs[(wn + j * 8 + g) * LG_PK + kk + t];
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
    if (kc == nk - 1) {
      const int m0 = (mt0 + tile / ntn) * LG_BM, n0 = (tile % ntn) * LG_BN;
#pragma unroll
      for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int j = 0; j < 2; ++j)
#pragma unroll
          for (int c = 0; c < 2; ++c) {
            const int row = m0 + wm + i * 8 + g;
            const int col = n0 + wn + j * 8 + 2 * t + c;
            if (row < M && col < N) {
              cplx acc = cmake(p1[i][j][c] - p2[i][j][c], (p3[i][j][c] - p1[i][j][c]) - p2[i][j][c]);
              if (avg) {
                // column = frequency f: pair (global shifts f and g2), see AvgSpec
                const int first = s.a_zero ? 1 : 0;
                cplx *yc = avg_col(s, b, col) + row;
                if (fl) {
                  if (col >= first) {
                    const int g2 = s.a_nfreq + col - first;
                    cplx t = cmul(s.cxb[(long)b * s.ns + g2 - 1], U0B[(long)row + (long)(g2 - 1) * s.n]);
                    if (col >= 1) t = cadd(t, cmul(s.cxb[(long)b * s.ns + col - 1], U0B[(long)row + (long)(col - 1) * s.n]));
                    acc = cmake(acc.x + 0.5 * t.x, acc.y + 0.5 * t.y);
                  }
                  acc = cadd(acc, *yc);
                }
                if (what == 0 && col == 0) {       // the seed system's x (global shift 0) enters once, at the end
                  const cplx xs = seedX(s, b)[row];
                  acc = first ? cadd(acc, xs) : cmake(acc.x + 0.5 * xs.x, acc.y + 0.5 * xs.y);
                }
                if (what == 0 && ((acc.x != acc.x) || (acc.y != acc.y))) atomicMax(&s.ierr[b], 2);   // bicgstab.f90:264-267
                *yc = acc;
              } else {
                const long off = (long)row + (long)col * s.n;
                if (fl) {
                  if (what != 1) acc = cadd(cfma(s.cxb[(long)b * s.ns + col], U0B[off], acc), C[off]);
                  else acc = cfma(s.cub[(long)b * s.ns + col], C[off], acc);
                }
                C[off] = acc;
              }
            }
          }
    }
  }
}

// after a flush: the materialised state is the new base of every right-hand side that took part since the last one
__global__ void k_lazy_reset(BicgState s) {
  const long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long)s.nrhs * s.ns) return;
  const int b = (int)(idx / s.ns);
  if (!s.active[b]) return;
  s.cxb[idx] = cmake(0.0, 0.0);
  s.cub[idx] = cmake(1.0, 0.0);
}
__global__ void k_lazy_mark(BicgState s, int new_base) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= s.nrhs) return;
  if (s.active[b]) { s.flushed[b] = 1; s.base_b[b] = new_base; }
}
__global__ void k_lazy_init(BicgState s) {
  const long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx < (long)s.nrhs * s.ns) { s.cxb[idx] = cmake(0.0, 0.0); s.cub[idx] = cmake(1.0, 0.0); }
  if (idx < s.nrhs) { s.flushed[idx] = 0; s.last_stage[idx] = 0; s.base_b[idx] = 0; }
}

constexpr int CSHIFT_CHUNK = 32;
template <int LT>
__global__ void __launch_bounds__(BT, 2) k_shift_apply(BicgState s, const ShiftCoef<LT> *__restrict__ coef, int chunk) {
  constexpr int NSN = LT * (LT + 1) / 2 + 1;
  const int b = blockIdx.y;
  const int stage = s.stage[b];
  if (stage == 0) return;
  const int is0 = blockIdx.z * chunk, nsl = min(chunk, s.ns - is0);
  __shared__ ShiftCoef<LT> sc[CSHIFT_CHUNK];
  {
    const cplx *src = (const cplx *)(coef + (long)b * s.ns + is0);
    cplx *dst = (cplx *)sc;
    const int nc = nsl * (int)(sizeof(ShiftCoef<LT>) / sizeof(cplx));
    for (int t = threadIdx.x; t < nc; t += BT) dst[t] = src[t];
  }
  __syncthreads();
  const int e = blockIdx.x * BT + threadIdx.x;
  if (e >= s.n) return;
  const long n = s.n;
  cplx sn[NSN], rr[LT];
  const cplx *snapg = snap(s, b, 0) + e;
#pragma unroll
  for (int k = 0; k < NSN; ++k) sn[k] = snapg[(long)k * n];
  if (stage == 2) {
#pragma unroll
    for (int j = 0; j < LT; ++j) rr[j] = seedR(s, b, j)[e];
  }
  cplx *pu0 = shiftU0(s, b, is0) + e, *px = shiftX(s, b, is0) + e;
  // first outer iteration of this right-hand side: u^sigma_0 = x^sigma = 0 (bicgstab.f90 init_shift), not read from memory
  const bool first = s.iters[b] == 1;
  const cplx zero = cmake(0.0, 0.0);
  // software pipeline: shifts are processed in groups of G and the loads of the NEXT group are issued before the arithmetic
  // of the current one, so every thread keeps 2 G independent 16-byte loads in flight (the kernel is a pure HBM stream:
  // 2 CTAs x 256 threads per SM need ~4 loads per thread to cover the DRAM latency-bandwidth product)
  constexpr int G = 2;   // measured (Si64 step): G = 1 -> 98.7 ms, G = 2 -> 81.6 ms, G = 4 -> 85.0 ms (spills at the 128-register cap)
  cplx u0n[G], xn[G];
#pragma unroll
  for (int g = 0; g < G; ++g)
    if (g < nsl) { u0n[g] = first ? zero : pu0[(long)g * n]; xn[g] = first ? zero : px[(long)g * n]; }
  for (int il = 0; il < nsl; il += G) {
    cplx u0[G], x0[G];
#pragma unroll
    for (int g = 0; g < G; ++g) { u0[g] = u0n[g]; x0[g] = xn[g]; }
#pragma unroll
    for (int g = 0; g < G; ++g)
      if (il + G + g < nsl) {
        u0n[g] = first ? zero : pu0[(long)(il + G + g) * n];
        xn[g] = first ? zero : px[(long)(il + G + g) * n];
      }
#pragma unroll
    for (int g = 0; g < G; ++g) {
      if (il + g >= nsl) break;
      const ShiftCoef<LT> &c = sc[il + g];
      cplx x = cfma(c.xu0, u0[g], x0[g]);
#pragma unroll
      for (int jj = 0; jj < LT; ++jj) x = cfma(c.xs[jj], sn[jj * (jj + 1) / 2], x);
      if (stage == 2) {
#pragma unroll
        for (int j = 0; j < LT; ++j) x = cfma(c.xr[j], rr[j], x);
        cplx u = cmul(c.uu0, u0[g]);
#pragma unroll
        for (int k = 0; k < NSN; ++k) u = cfma(c.us[k], sn[k], u);
        pu0[(long)(il + g) * n] = u;
      }
      px[(long)(il + g) * n] = x;
    }
  }
}

// L14-L17 of the seed system and the delayed L32 residual update (:833-844, :919-925)
__global__ void __launch_bounds__(BT) k_mr_seed(BicgState s) {
  const int b = blockIdx.y;
  if (!s.active[b]) return;
  const int L = s.L;
  __shared__ cplx sg[LCAP], sgp[LCAP], sgpp[LCAP];
  const SeedScal &sd = s.seed[b];
  for (int i = threadIdx.x; i < L; i += BT) { sg[i] = sd.gamma[i]; sgp[i] = sd.gamma_p[i]; sgpp[i] = i < L - 1 ? sd.gamma_pp[i] : cmake(0, 0); }
  __syncthreads();
  const int e = blockIdx.x * BT + threadIdx.x;
  if (e >= s.n) return;
  cplx r0 = seedR(s, b, 0)[e];
  if (s.lazy) {   // r_0 (before :919-925) and r_1..r_{L-1} (after the Gram-Schmidt) are basis vectors of the shifted x updates (:889,:905)
    snap(s, b, s.nsnap)[e] = r0;
    for (int jj = 1; jj <= L - 1; ++jj) snap(s, b, s.nsnap + jj)[e] = seedR(s, b, jj)[e];
  }
  cplx *px = seedX(s, b) + e;
  cplx x = cfma(sg[0], r0, *px);                                                  // :833
  cplx u0 = cfma(cneg(sg[L - 1]), seedU(s, b, L)[e], seedU(s, b, 0)[e]);          // :835
  for (int jj = 1; jj <= L - 1; ++jj) {
    u0 = cfma(cneg(sg[jj - 1]), seedU(s, b, jj)[e], u0);                          // :841
    x = cfma(sgpp[jj - 1], seedR(s, b, jj)[e], x);                                // :844
  }
  *px = x;
  seedU(s, b, 0)[e] = u0;
  for (int jj = 1; jj <= L; ++jj) r0 = cfma(cneg(sgp[jj - 1]), seedR(s, b, jj)[e], r0);   // :919-925
  seedR(s, b, 0)[e] = r0;
}

// ---------------------------------------------------------------- finish: copy out (:258-261), NaN scan (:264-267)
// the solutions already live in the caller's buffer (:258-261); NaN scan (:264-267)
__global__ void __launch_bounds__(BT) k_nan_scan(BicgState s, int *__restrict__ ierr, const int *__restrict__ todo) {
  const int b = blockIdx.y;
  if (todo && !todo[b]) return;
  const int e = blockIdx.x * BT + threadIdx.x;
  const int nshift = s.avg ? 1 : s.ns + 1;      // average mode: the shifted solutions were checked when they were materialised
  bool bad = false;
  if (e < s.n) {
    for (int is = 0; is < nshift; ++is) {
      const cplx v = s.XO[((long)b * nshift + is) * s.n + e];
      bad |= (v.x != v.x) || (v.y != v.y);
    }
  }
  if (__syncthreads_or(bad) && threadIdx.x == 0) atomicMax(&ierr[b], 2);
}

__global__ void k_avg_done(BicgState s, const int *__restrict__ ierr, const int *__restrict__ todo, int *__restrict__ done) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= s.nrhs) return;
  if (todo && !todo[b]) return;
  done[b] = ierr[b] == 0 ? 1 : 0;
}

__global__ void k_set_ierr(BicgState s, int *__restrict__ ierr, const int *__restrict__ todo) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= s.nrhs) return;
  if (todo && !todo[b]) return;
  ierr[b] = s.active[b] ? 1 : 0;     // still active after max_iter outer iterations -> ierr = 1 (:249-253)
}

// ---------------------------------------------------------------- driver
template <int LT>
static int launch_shift_fused(sgw_ctx *ctx, const BicgState &s, int chunk, size_t smem) {
  if (smem > 48 * 1024) SGW_CUDA(cudaFuncSetAttribute(k_shift_fused<LT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  dim3 grid((unsigned)((s.n + BT - 1) / BT), (unsigned)s.nrhs, (unsigned)((s.ns + chunk - 1) / chunk));
  ProfScope prof(ctx, PC_SHIFT);
  k_shift_fused<LT><<<grid, BT, smem, ctx->stream>>>(s, chunk);
  SGW_LAUNCH_CHECK();
  return SGW_OK;
}

template <int LT>
static int launch_shift_collapsed(sgw_ctx *ctx, const BicgState &s) {
  ShiftCoef<LT> *coef = nullptr;
  SGW_CHECK(ws(ctx, "bi_scoef", (size_t)s.nrhs * s.ns, &coef));
  ProfScope prof(ctx, PC_SHIFT);
  const long tot = (long)s.nrhs * s.ns;
  k_shift_coef<LT><<<(unsigned)((tot + 127) / 128), 128, 0, ctx->stream>>>(s, coef);
  SGW_LAUNCH_CHECK();
  if (s.lazy) return SGW_OK;            // the vectors are touched once, at the flush / at the end (lazy_materialise)
  const int nchunks = (s.ns + CSHIFT_CHUNK - 1) / CSHIFT_CHUNK;     // equal-sized chunks of <= 32 shifts per CTA
  const int chunk = (s.ns + nchunks - 1) / nchunks;
  dim3 grid((unsigned)((s.n + BT - 1) / BT), (unsigned)s.nrhs, (unsigned)((s.ns + chunk - 1) / chunk));
  k_shift_apply<LT><<<grid, BT, 0, ctx->stream>>>(s, coef, chunk);
  SGW_LAUNCH_CHECK();
  return SGW_OK;
}

static bool shift_faithful() {
  const char *e = getenv("SGW_SHIFT");
  return e && strcmp(e, "faithful") == 0;
}
// SGW_SHIFT=stream keeps the collapsed update that streams u^sigma_0 / x^sigma every outer iteration (k_shift_apply)
static bool shift_stream() {
  const char *e = getenv("SGW_SHIFT");
  return e && strcmp(e, "stream") == 0;
}
// basis slots (outer iterations) kept per right-hand side before the lazy shifted update is flushed (SGW_LAZY_TC)
static int lazy_period() {
  const char *e = getenv("SGW_LAZY_TC");
  const int v = e ? atoi(e) : 6;
  return std::max(1, std::min(v, 64));
}
static bool lazy_shift(int lmax, int nshift) { return nshift > 1 && (lmax == 2 || lmax == 4) && !shift_faithful() && !shift_stream(); }

// device bytes of solver state per right-hand side (the batch sizing of coulomb.cu / green uses this)
size_t bicgstab_bytes_per_rhs(int n, int lmax, int nshift) {
  const size_t L1 = lmax + 1, ns = nshift - 1, nsnap = (size_t)lmax * (lmax + 1) / 2 + 1;
  size_t v = 2 * L1 + 1;                                         // U, R, RT
  v += ns;                                                       // u^sigma_0
  if (lazy_shift(lmax, nshift)) v += (size_t)lazy_period() * (nsnap + lmax);
  else if (ns > 0) v += nsnap;
  return v * (size_t)n * sizeof(cplx);
}

// flush_iter > 0: flush of the still active right-hand sides after outer iteration flush_iter; 0: final materialisation
static int lazy_materialise(sgw_ctx *ctx, const BicgState &s, int flush_iter) {
  static bool attr = false;
  if (!attr) {
    SGW_CUDA(cudaFuncSetAttribute(k_shift_gemm, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)LG_SMEM));
    attr = true;
  }
  ProfScope prof(ctx, PC_SHIFT_GEMM);
  const int ntm = (s.n + LG_BM - 1) / LG_BM;
  dim3 grid((unsigned)((ntm + LG_TM - 1) / LG_TM), (unsigned)s.nrhs);
  if (s.avg) {
    dim3 gc((unsigned)std::max(1, std::min(16, (s.kcap * s.a_nfreq + 255) / 256)), (unsigned)s.nrhs);
    k_lazy_combine<<<gc, 256, 0, ctx->stream>>>(s, flush_iter > 0 ? 2 : 0);
    SGW_LAUNCH_CHECK();
  }
  k_shift_gemm<<<grid, LG_T, LG_SMEM, ctx->stream>>>(s, flush_iter > 0 ? 2 : 0);
  SGW_LAUNCH_CHECK();
  if (flush_iter > 0) {
    k_shift_gemm<<<grid, LG_T, LG_SMEM, ctx->stream>>>(s, 1);
    SGW_LAUNCH_CHECK();
    const long tot = (long)s.nrhs * s.ns;
    k_lazy_reset<<<(unsigned)((tot + 127) / 128), 128, 0, ctx->stream>>>(s);
    SGW_LAUNCH_CHECK();
    k_lazy_mark<<<(unsigned)((s.nrhs + 127) / 128), 128, 0, ctx->stream>>>(s, flush_iter);
    SGW_LAUNCH_CHECK();
  }
  return SGW_OK;
}

int bicgstab_batched(sgw_ctx *ctx, const SolveBatch &sb, int lmax, double threshold, int max_iter, const int *d_todo) {
  if (lmax < 1 || lmax > LCAP - 1) {
    ctx->err = "bicg_lmax must be in 1..15";
    return SGW_E_ARG;
  }
  if (sb.nrhs <= 0) return SGW_OK;
  BicgState s;
  s.n = sb.n; s.nrhs = sb.nrhs; s.ns = sb.nshift - 1; s.L = lmax;
  s.nsnap = lmax * (lmax + 1) / 2 + 1;
  s.XO = sb.d_x;
  const long n = s.n, nr = s.nrhs, L1 = lmax + 1;
  SGW_CHECK(ws(ctx, "bi_U", (size_t)(nr * L1 * n), &s.U));
  SGW_CHECK(ws(ctx, "bi_R", (size_t)(nr * L1 * n), &s.R));
  SGW_CHECK(ws(ctx, "bi_RT", (size_t)(nr * n), &s.RT));
  SGW_CHECK(ws(ctx, "bi_US0", (size_t)(nr * s.ns * n) + 1, &s.US0));
  s.lazy = lazy_shift(lmax, sb.nshift) ? 1 : 0;
  s.Tc = lazy_period(); s.slot = 0; s.base = 0;
  s.nv = s.nsnap + lmax; s.kcap = s.Tc * s.nv;
  SGW_CHECK(ws(ctx, "bi_SNAP", (s.ns > 0 && !s.lazy) ? (size_t)(nr * s.nsnap * n) : 1, &s.SNAP));
  SGW_CHECK(ws(ctx, "bi_BAS", s.lazy ? (size_t)(nr * s.Tc * s.nv * n) : 1, &s.BAS));
  SGW_CHECK(ws(ctx, "bi_CX", s.lazy ? (size_t)(nr * s.ns * s.kcap) : 1, &s.CX));
  SGW_CHECK(ws(ctx, "bi_CU", s.lazy ? (size_t)(nr * s.ns * s.kcap) : 1, &s.CU));
  SGW_CHECK(ws(ctx, "bi_cxb", (size_t)(nr * s.ns) + 1, &s.cxb));
  SGW_CHECK(ws(ctx, "bi_cub", (size_t)(nr * s.ns) + 1, &s.cub));
  SGW_CHECK(ws(ctx, "bi_flushed", (size_t)nr, &s.flushed));
  SGW_CHECK(ws(ctx, "bi_lstage", (size_t)nr, &s.last_stage));
  SGW_CHECK(ws(ctx, "bi_baseb", (size_t)nr, &s.base_b));
  s.avg = (s.lazy && sb.avg.d_y && sb.avg.nfreq > 0 && (2 * sb.avg.nfreq - (sb.avg.zero_freq ? 1 : 0)) == sb.nshift) ? 1 : 0;
  s.a_nfreq = sb.avg.nfreq; s.a_zero = sb.avg.zero_freq; s.a_group = std::max(1, sb.avg.group);
  s.Y = sb.avg.d_y; s.ierr = sb.d_ierr;
  SGW_CHECK(ws(ctx, "bi_CA", s.avg ? (size_t)(nr * s.kcap * s.a_nfreq) : 1, &s.CA));
  SGW_CHECK(ws(ctx, "bi_seed", (size_t)nr, &s.seed));
  SGW_CHECK(ws(ctx, "bi_shift", (size_t)(nr * s.ns) + 1, &s.shift));
  SGW_CHECK(ws(ctx, "bi_step", (size_t)(nr * s.ns * lmax) + 1, &s.step));
  s.nchunk = (int)((n + BT * DOT_EPT - 1) / (BT * DOT_EPT));
  SGW_CHECK(ws(ctx, "bi_part", (size_t)(3 * nr * s.nchunk), &s.part));
  SGW_CHECK(ws(ctx, "bi_active", (size_t)nr, &s.active));
  SGW_CHECK(ws(ctx, "bi_stage", (size_t)nr, &s.stage));
  SGW_CHECK(ws(ctx, "bi_iters", (size_t)nr, &s.iters));
  SGW_CHECK(ws(ctx, "bi_nactive", (size_t)1, &s.nactive));
  // The host runs ONE outer iteration ahead of the device: iteration i+1 is enqueued before the active count of
  // iteration i is read back, so the stream never drains while the host waits (a drained stream costs a full
  // launch latency per kernel and makes the solve sensitive to host jitter).  If iteration i turns out to have
  // converged everything, the already enqueued iteration i+1 is a no-op: every kernel returns on the active mask.
  if (!ctx->h_flags) SGW_CUDA(cudaMallocHost((void **)&ctx->h_flags, 4 * sizeof(int)));
  volatile int *h_nactive = ctx->h_flags;
  for (int i = 0; i < 2; ++i)
    if (!ctx->ev_iter[i]) SGW_CUDA(cudaEventCreateWithFlags(&ctx->ev_iter[i], cudaEventDisableTiming));

  cudaStream_t st = ctx->stream;
  const int gb = (int)((nr + 127) / 128);
  const dim3 gvec((unsigned)((n + BT - 1) / BT), (unsigned)nr);
  const dim3 gdot((unsigned)s.nchunk, (unsigned)nr);
  // shifts per CTA of the fused update: scalars of one chunk must fit in 48 KB of shared memory
  const int per_shift = 5 * lmax + 2 + lmax;
  int chunk = std::min(SHIFT_CHUNK, (int)(48 * 1024 / (per_shift * sizeof(cplx))));
  chunk = std::max(1, std::min(chunk, std::max(1, s.ns)));
  const size_t sm_fused = (size_t)chunk * per_shift * sizeof(cplx);

  k_init_scal<<<gb, 128, 0, st>>>(s, sb.d_sigma, d_todo);
  SGW_LAUNCH_CHECK();
  const bool faithful = shift_faithful();
  const bool collapsed = s.ns > 0 && !faithful && (lmax == 2 || lmax == 4);
  k_init_vec<<<gvec, BT, 0, st>>>(s, sb.d_b, sb.ldb, collapsed ? 0 : 1);
  SGW_LAUNCH_CHECK();
  if (s.lazy) {
    const long tot = std::max<long>(nr * s.ns, nr);
    k_lazy_init<<<(unsigned)((tot + 127) / 128), 128, 0, st>>>(s);
    SGW_LAUNCH_CHECK();
  }

  const long ldv = n;   // vectors inside U/R are contiguous with stride (L+1)*n between RHS
  int rc = SGW_OK;
  for (int iter = 1; iter <= max_iter && rc == SGW_OK; ++iter) {
    s.slot = iter - 1 - s.base;
    // ---- bicg_part (seed system; the shifted systems only record their step scalars)
    for (int jj = 0; jj < lmax && rc == SGW_OK; ++jj) {
      DotSpec d; d.nd = 1; d.ka[0] = 0; d.ia[0] = jj; d.kb[0] = 2; d.ib[0] = 0;     // (r_j, rt0)
      {
        ProfScope prof(ctx, PC_SEED);
        k_dots<<<gdot, BT, 0, st>>>(s, d);
        SGW_LAUNCH_CHECK();
        k_scal_beta<<<gb, 128, 0, st>>>(s, jj);
        SGW_LAUNCH_CHECK();
        switch (jj) {
          case 0: k_seed_u<0><<<gvec, BT, 0, st>>>(s, jj); break;
          case 1: k_seed_u<1><<<gvec, BT, 0, st>>>(s, jj); break;
          case 2: k_seed_u<2><<<gvec, BT, 0, st>>>(s, jj); break;
          case 3: k_seed_u<3><<<gvec, BT, 0, st>>>(s, jj); break;
          default: k_seed_u<-1><<<gvec, BT, 0, st>>>(s, jj); break;
        }
        SGW_LAUNCH_CHECK();
      }
      // u_{j+1} = A u_j   (bicgstab.f90:611)
      rc = apply_operator(ctx, sb.slot, sb.alpha_pv, s.nrhs, s.U + (long)jj * n, L1 * ldv, &s.seed[0].sigma,
                          sizeof(SeedScal) / sizeof(cplx), s.U + (long)(jj + 1) * n, L1 * ldv, s.active);
      if (rc != SGW_OK) break;
      d.ka[0] = 1; d.ia[0] = jj + 1;                                                // (u_{j+1}, rt0)
      {
        ProfScope prof(ctx, PC_SEED);
        k_dots<<<gdot, BT, 0, st>>>(s, d);
        SGW_LAUNCH_CHECK();
        k_scal_alpha<<<(unsigned)nr, 128, 0, st>>>(s, jj);
        SGW_LAUNCH_CHECK();
        switch (jj) {
          case 0: k_bicg_update<0><<<gvec, BT, 0, st>>>(s, jj); break;
          case 1: k_bicg_update<1><<<gvec, BT, 0, st>>>(s, jj); break;
          case 2: k_bicg_update<2><<<gvec, BT, 0, st>>>(s, jj); break;
          case 3: k_bicg_update<3><<<gvec, BT, 0, st>>>(s, jj); break;
          default: k_bicg_update<-1><<<gvec, BT, 0, st>>>(s, jj); break;
        }
        SGW_LAUNCH_CHECK();
      }
      // r_{j+1} = A r_j   (bicgstab.f90:682)
      rc = apply_operator(ctx, sb.slot, sb.alpha_pv, s.nrhs, s.R + (long)jj * n, L1 * ldv, &s.seed[0].sigma,
                          sizeof(SeedScal) / sizeof(cplx), s.R + (long)(jj + 1) * n, L1 * ldv, s.active);
    }
    if (rc != SGW_OK) break;
    {
      ProfScope prof(ctx, PC_SEED);
      SGW_CUDA(cudaMemsetAsync(s.nactive, 0, sizeof(int), st));
      k_norm_r0<<<gdot, BT, 0, st>>>(s);
      SGW_LAUNCH_CHECK();
      k_check<<<gb, 128, 0, st>>>(s, threshold, iter, 1);                           // :237
      SGW_LAUNCH_CHECK();
      // ---- mr_part (RHS that converged above are masked out)
      k_mr_mgs<<<(unsigned)nr, 1024, 0, st>>>(s);
      SGW_LAUNCH_CHECK();
    }
    if (collapsed) {
      rc = lmax == 4 ? launch_shift_collapsed<4>(ctx, s) : launch_shift_collapsed<2>(ctx, s);
      if (rc != SGW_OK) break;
    } else if (s.ns > 0) {
      switch (lmax) {
        case 1: rc = launch_shift_fused<1>(ctx, s, chunk, sm_fused); break;
        case 2: rc = launch_shift_fused<2>(ctx, s, chunk, sm_fused); break;
        case 4: rc = launch_shift_fused<4>(ctx, s, chunk, sm_fused); break;
        default: rc = launch_shift_fused<0>(ctx, s, chunk, sm_fused); break;
      }
      if (rc != SGW_OK) break;
    }
    {
      ProfScope prof(ctx, PC_SEED);
      k_mr_seed<<<gvec, BT, 0, st>>>(s);
      SGW_LAUNCH_CHECK();
      SGW_CUDA(cudaMemsetAsync(s.nactive, 0, sizeof(int), st));
      k_norm_r0<<<gdot, BT, 0, st>>>(s);
      SGW_LAUNCH_CHECK();
      k_check<<<gb, 128, 0, st>>>(s, threshold, iter, 2);                           // :245
      SGW_LAUNCH_CHECK();
    }
    if (s.lazy && iter - s.base == s.Tc) {   // basis slots exhausted: materialise x^sigma and u^sigma_0, start a new window
      rc = lazy_materialise(ctx, s, iter);
      if (rc != SGW_OK) break;
      s.base = iter;
    }
    SGW_CUDA(cudaMemcpyAsync((void *)&h_nactive[iter & 1], s.nactive, sizeof(int), cudaMemcpyDeviceToHost, st));
    SGW_CUDA(cudaEventRecord(ctx->ev_iter[iter & 1], st));
    if (iter >= 2) {
      SGW_CUDA(cudaEventSynchronize(ctx->ev_iter[(iter - 1) & 1]));
      if (h_nactive[(iter - 1) & 1] == 0) break;
    }
  }
  if (rc != SGW_OK) return rc;
  k_set_ierr<<<gb, 128, 0, st>>>(s, sb.d_ierr, d_todo);
  SGW_LAUNCH_CHECK();
  if (s.lazy) SGW_CHECK(lazy_materialise(ctx, s, 0));      // average mode: NaN check of the shifted solutions in its epilogue
  k_nan_scan<<<gvec, BT, 0, st>>>(s, sb.d_ierr, d_todo);
  SGW_LAUNCH_CHECK();
  if (s.avg && sb.avg.d_done) {
    k_avg_done<<<gb, 128, 0, st>>>(s, sb.d_ierr, d_todo, sb.avg.d_done);
    SGW_LAUNCH_CHECK();
  }
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
  return SGW_OK;
}

}  // namespace sgw
