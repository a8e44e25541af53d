// subspace.cu -- batched SGW Krylov-subspace solver, restating algo/linear_solver/src/linear_solver.f90:82-503
// with data/algebra/src/gram_schmidt.f90:35-135 (modified Gram-Schmidt carrying the second vector set) and
// select.cu's priority/fallback chain of algo/linear_solver/src/select_solver.f90:67-161.
//
// Shifts are processed sequentially and carry the subspace (V, W = (A + sigma) V) from shift to shift exactly
// like the reference (:153-192); the batch axis is the right-hand side.  One CTA owns one RHS:
//   k_sub_shift    : W += (sigma - sigma_old) V, then gram_schmidt(1, W, V)           (:272-300)
//   k_sub_residual : r = b - sum_i w_i <w_i|b>, convergence on the RELATIVE threshold,  (:312-383)
//                    and on convergence x = sum_i v_i <w_i|b> + NaN scan                (:453-503, :186-190)
//   apply_operator : new = (A + sigma) r  for the RHS that still iterate                (:168)
//   k_sub_expand   : append (new, r), gram_schmidt(m+1, W, V)                           (:391-440)
// Inside the re-orthonormalisation the columns j > i are independent, so each warp takes its own columns
// (dot + two axpys) between two block barriers; the dependency order is the reference's MGS order.
//
// Default path (SGW_SUB=mgs selects the kernels above, which keep the reference's operation order):
//   * re-orthonormalisation at a new shift by Cholesky-QR: G = W^H W from per-row-block partial Gram matrices
//     (k_sub_gram, which also applies W += dsigma V), G = R^H R and T = R^-1 in shared memory (k_sub_chol),
//     W <- W T, V <- V T in place (k_sub_tri); a second sweep runs for the right-hand sides whose first factor was
//     ill conditioned.  QR with a positive diagonal is unique, so this is the factorisation MGS computes, in O(1)
//     dependent steps over the vector length instead of m(m-1)/2 dependent dot/axpy pairs.
//   * k_sub_step: one kernel per iteration appends (A r, r), orthogonalises the new pair against the basis by
//     classical Gram-Schmidt with the "twice is enough" re-orthogonalisation (all m dots at once, warp per column),
//     and updates the residual INCREMENTALLY: the overlaps <w_i|b> of the old columns do not change inside a shift
//     (:367 recomputes them), so r <- r - w_m <w_m|b>.
//   * the host runs one iteration ahead of the device (count of active systems read back with one iteration lag),
//     so there is no stream synchronisation per iteration.
#include "internal.cuh"

#include <stdlib.h>
#include <string.h>

#include <algorithm>

namespace sgw {

constexpr int ST = 1024;  // largest CTA of the one-CTA-per-RHS kernels (see sub_threads)

struct SubState {
  int n, nrhs, nshift, cap, max_iter;
  cplx *V, *W;              // [nrhs][cap][n]
  cplx *Rv, *New;           // [nrhs][n]
  const cplx *b; long ldb;
  const cplx *sigma;        // nshift x nrhs
  cplx *sig_old;            // [nrhs]
  cplx *x;                  // [(rhs*nshift + ishift)*n]
  double *absthr;           // [nrhs]
  int *m, *act, *alive, *done, *iter, *ierr, *count;
  long *nop;                // [1] operator applications
  cplx *c;                  // [nrhs][n] overlaps <w_i|b> of the current shift
  cplx *G;                  // [nrhs][ldg*ldg] Gram matrices W^H W, then T = R^-1 (upper triangle)
  int *again;               // [nrhs] second Cholesky-QR sweep wanted
};

__device__ __forceinline__ cplx *colV(const SubState &s, int b, int i) { return s.V + ((long)b * s.cap + i) * s.n; }
__device__ __forceinline__ cplx *colW(const SubState &s, int b, int i) { return s.W + ((long)b * s.cap + i) * s.n; }

__device__ __forceinline__ cplx warp_sum(cplx v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    v.x += __shfl_xor_sync(0xffffffffu, v.x, o);
    v.y += __shfl_xor_sync(0xffffffffu, v.y, o);
  }
  return v;
}

__device__ __forceinline__ cplx block_sum(cplx v, cplx *sm) {
  v = warp_sum(v);
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
  __syncthreads();
  if (lane == 0) sm[w] = v;
  __syncthreads();
  if (w == 0) {
    v = lane < nw ? sm[lane] : cmake(0.0, 0.0);
    v = warp_sum(v);
    if (lane == 0) sm[0] = v;
  }
  __syncthreads();
  return sm[0];
}

// 2-norm of a column (norm.f90: ZLANGE 'F'); plain sum of squares
__device__ __forceinline__ double block_norm(const cplx *x, int n, cplx *sm) {
  cplx acc = cmake(0.0, 0.0);
  for (int e = threadIdx.x; e < n; e += blockDim.x) { acc.x += x[e].x * x[e].x; acc.y += x[e].y * x[e].y; }
  acc = block_sum(acc, sm);
  return sqrt(acc.x + acc.y);
}

__global__ void __launch_bounds__(ST) k_sub_init(SubState s, double threshold, const int *__restrict__ todo) {
  const int b = blockIdx.x;
  __shared__ cplx sm[32];
  const double nb = block_norm(s.b + (long)b * s.ldb, s.n, sm);
  if (threadIdx.x == 0) {
    s.absthr[b] = threshold * nb;                       // linear_solver.f90:213
    s.m[b] = 0;
    s.sig_old[b] = cmake(0.0, 0.0);                     // :242
    s.alive[b] = todo ? (todo[b] != 0) : 1;
    s.act[b] = 0;
    s.done[b] = 0;
    s.iter[b] = 0;
  }
  const int alive = todo ? (todo[b] != 0) : 1;
  if (alive)
    for (long i = threadIdx.x; i < (long)s.nshift * s.n; i += blockDim.x) s.x[(long)b * s.nshift * s.n + i] = cmake(0.0, 0.0);
}

// gram_schmidt(first = 1): normalise column i, then remove it from all later columns (warp per column)
__device__ void gs_full(const SubState &s, int b, int m, cplx *sm) {
  const int n = s.n, lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
  for (int i = 0; i < m; ++i) {
    cplx *wi = colW(s, b, i), *vi = colV(s, b, i);
    const double inv = 1.0 / block_norm(wi, n, sm);     // gram_schmidt.f90:114
    for (int e = threadIdx.x; e < n; e += blockDim.x) { wi[e] = cscale(inv, wi[e]); vi[e] = cscale(inv, vi[e]); }
    __syncthreads();
    for (int j = i + 1 + w; j < m; j += nw) {           // :121-132
      cplx *wj = colW(s, b, j), *vj = colV(s, b, j);
      cplx acc = cmake(0.0, 0.0);
      for (int e = lane; e < n; e += 32) acc = cfma(cconj(wi[e]), wj[e], acc);
      acc = cneg(warp_sum(acc));
      for (int e = lane; e < n; e += 32) { wj[e] = cfma(acc, wi[e], wj[e]); vj[e] = cfma(acc, vi[e], vj[e]); }
    }
    __syncthreads();
  }
}

__global__ void __launch_bounds__(ST) k_sub_shift(SubState s, int ishift, int only_above) {
  const int b = blockIdx.x;
  if (!s.alive[b]) return;
  __shared__ cplx sm[32];
  const int m = s.m[b], n = s.n;
  if (m <= only_above) return;                          // those bases are re-orthonormalised by Cholesky-QR
  const cplx sg = s.sigma[(long)b * s.nshift + ishift];
  const cplx d = csub(sg, s.sig_old[b]);                // :289
  for (int i = 0; i < m; ++i) {
    cplx *wi = colW(s, b, i);
    const cplx *vi = colV(s, b, i);
    for (int e = threadIdx.x; e < n; e += blockDim.x) wi[e] = cfma(d, vi[e], wi[e]);   // :292
  }
  __syncthreads();
  gs_full(s, b, m, sm);                                 // :295
  if (threadIdx.x == 0) {
    s.sig_old[b] = sg;                                  // :298
    s.done[b] = 0;
    s.iter[b] = 0;
  }
}

__global__ void __launch_bounds__(ST) k_sub_residual(SubState s, int ishift) {
  const int b = blockIdx.x;
  if (threadIdx.x == 0) s.act[b] = 0;
  if (!s.alive[b] || s.done[b]) return;
  extern __shared__ cplx dyn[];       // overlaps c[cap]
  __shared__ cplx sm[32];
  const int m = s.m[b], n = s.n, lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
  const cplx *bb = s.b + (long)b * s.ldb;
  if (s.iter[b] >= s.max_iter) {                        // loop ran out without convergence (:176-180)
    if (threadIdx.x == 0) { s.ierr[b] = 1; s.alive[b] = 0; }
    return;
  }
  for (int i = w; i < m; i += nw) {                     // overlap_i = <w_i | b>  (:367)
    const cplx *wi = colW(s, b, i);
    cplx acc = cmake(0.0, 0.0);
    for (int e = lane; e < n; e += 32) acc = cfma(cconj(wi[e]), bb[e], acc);
    acc = warp_sum(acc);
    if (lane == 0) { dyn[i] = acc; if (s.c) s.c[(long)b * s.n + i] = acc; }
  }
  __syncthreads();
  cplx *r = s.Rv + (long)b * n;
  cplx acc = cmake(0.0, 0.0);
  for (int e = threadIdx.x; e < n; e += blockDim.x) {
    cplx v = bb[e];
    for (int i = 0; i < m; ++i) v = cfma(cneg(dyn[i]), colW(s, b, i)[e], v);   // :368
    r[e] = v;
    acc.x += v.x * v.x; acc.y += v.y * v.y;
  }
  acc = block_sum(acc, sm);
  const double nrm = sqrt(acc.x + acc.y);               // :373
  const bool conv = nrm < s.absthr[b];                  // :374
  if (conv) {
    // obtain_result :486-501 and NaN scan :186-190
    cplx *x = s.x + ((long)b * s.nshift + ishift) * n;
    int bad = 0;
    for (int e = threadIdx.x; e < n; e += blockDim.x) {
      cplx v = cmake(0.0, 0.0);
      for (int i = 0; i < m; ++i) v = cfma(dyn[i], colV(s, b, i)[e], v);
      x[e] = v;
      bad |= (v.x != v.x) || (v.y != v.y);
    }
    bad = __syncthreads_or(bad);
    if (threadIdx.x == 0) {
      s.done[b] = 1;
      if (bad) { s.ierr[b] = 2; s.alive[b] = 0; }
    }
    return;
  }
  if (threadIdx.x == 0) {
    if (n == m) {                                       // :376-378 sets 3, :176-180 overwrites it with 1
      s.ierr[b] = 1; s.alive[b] = 0;
    } else {
      s.act[b] = 1;
      s.iter[b] += 1;
      atomicAdd(s.count, 1);
      atomicAdd((unsigned long long *)s.nop, 1ull);
    }
  }
}

__global__ void __launch_bounds__(ST) k_sub_expand(SubState s) {
  const int b = blockIdx.x;
  if (!s.act[b]) return;
  __shared__ cplx sm[32];
  __shared__ cplx coef;
  const int m = s.m[b], n = s.n;
  cplx *wm = colW(s, b, m), *vm = colV(s, b, m);
  const cplx *nw = s.New + (long)b * n, *r = s.Rv + (long)b * n;
  for (int e = threadIdx.x; e < n; e += blockDim.x) { wm[e] = nw[e]; vm[e] = r[e]; }   // :426,:432
  __syncthreads();
  for (int j = 0; j < m; ++j) {                          // gram_schmidt.f90:94-106 (first = m+1)
    const cplx *wj = colW(s, b, j), *vj = colV(s, b, j);
    cplx acc = cmake(0.0, 0.0);
    for (int e = threadIdx.x; e < n; e += blockDim.x) acc = cfma(cconj(wj[e]), wm[e], acc);
    acc = block_sum(acc, sm);
    if (threadIdx.x == 0) coef = cneg(acc);
    __syncthreads();
    const cplx c = coef;
    for (int e = threadIdx.x; e < n; e += blockDim.x) { wm[e] = cfma(c, wj[e], wm[e]); vm[e] = cfma(c, vj[e], vm[e]); }
    __syncthreads();
  }
  const double inv = 1.0 / block_norm(wm, n, sm);        // :114
  for (int e = threadIdx.x; e < n; e += blockDim.x) { wm[e] = cscale(inv, wm[e]); vm[e] = cscale(inv, vm[e]); }
  if (threadIdx.x == 0) s.m[b] = m + 1;
}


// ---------------------------------------------------------------- default path: Cholesky-QR + fused iteration step
constexpr int MCH = 96;      // largest basis the Cholesky-QR path handles (R lives in shared memory); larger ones use gs_full
constexpr int TRI_R = 32;    // rows per tile of k_sub_tri

__global__ void k_sub_wshift(SubState s, int ishift, int mmax) {          // W += (sigma - sigma_old) V   (:289-292)
  const int b = blockIdx.z, i = blockIdx.y;
  if (!s.alive[b] || i >= s.m[b] || s.m[b] > MCH) return;
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= s.n) return;
  const cplx d = csub(s.sigma[(long)b * s.nshift + ishift], s.sig_old[b]);
  cplx *wi = colW(s, b, i);
  wi[e] = cfma(d, colV(s, b, i)[e], wi[e]);
}

// G = R^H R (upper R) of the leading m x m block of the Gram matrix, then T = R^-1 written over G (upper triangle, leading
// dimension ldg).  pass 0 flags the right-hand sides whose factor is far from the identity for a second sweep.
__global__ void __launch_bounds__(256) k_sub_chol(SubState s, int ldg, int pass) {
  const int b = blockIdx.x;
  if (!s.alive[b]) return;
  const int m = s.m[b];
  if (m == 0 || m > MCH) return;
  if (pass == 1 && !s.again[b]) return;
  extern __shared__ cplx R[];                    // [m][m], column-major: R(i,j) at i + m*j
  __shared__ double s_d;
  cplx *G = s.G + (long)b * ldg * ldg;
  const int tid = threadIdx.x, nt = blockDim.x;
  for (int idx = tid; idx < m * m; idx += nt) { const int i = idx % m, j = idx / m; R[idx] = G[i + (long)ldg * j]; }
  __syncthreads();
  double dmin = 1e300, dmax = 0.0;
  for (int k = 0; k < m; ++k) {
    if (tid == 0) s_d = sqrt(R[k + m * k].x);
    __syncthreads();
    const double d = s_d, inv = 1.0 / d;
    dmin = fmin(dmin, d); dmax = fmax(dmax, d);
    for (int j = k + tid; j < m; j += nt) R[k + m * j] = (j == k) ? cmake(d, 0.0) : cscale(inv, R[k + m * j]);
    __syncthreads();
    const int t = m - k - 1;                     // trailing block: G(i,j) -= conj(R(k,i)) R(k,j), k < i <= j
    for (int idx = tid; idx < t * t; idx += nt) {
      const int i = k + 1 + idx % t, j = k + 1 + idx / t;
      if (i <= j) R[i + m * j] = cfma(cneg(cconj(R[k + m * i])), R[k + m * j], R[i + m * j]);
    }
    __syncthreads();
  }
  // T = R^-1 by columns (independent): T(j,j) = 1/R(j,j); T(i,j) = -sum_{k=i+1..j} R(i,k) T(k,j) / R(i,i); stored in the
  // strictly lower triangle of the shared array (T(i,j) at (j,i)) until every column is done
  for (int j = tid; j < m; j += nt) {
    const double tjj = 1.0 / R[j + m * j].x;
    for (int i = j - 1; i >= 0; --i) {
      cplx acc = cscale(tjj, R[i + m * j]);      // k = j term
      for (int k = i + 1; k < j; ++k) acc = cfma(R[i + m * k], R[j + m * k], acc);   // T(k,j) sits at (j,k)
      R[j + m * i] = cscale(-1.0 / R[i + m * i].x, acc);
    }
  }
  __syncthreads();
  for (int idx = tid; idx < m * m; idx += nt) {
    const int i = idx % m, j = idx / m;
    if (i < j) G[i + (long)ldg * j] = R[j + m * i];
    else if (i == j) G[i + (long)ldg * j] = cmake(1.0 / R[i + m * i].x, 0.0);
  }
  if (tid == 0 && pass == 0) s.again[b] = !(dmin > 0.5 * dmax);   // NaN counts as ill conditioned
}

// X <- X T in place for X = W (blockIdx.z = 0) and X = V (1): rows of a tile are staged in shared memory, T is read through L1
__global__ void __launch_bounds__(256) k_sub_tri(SubState s, int ldg, int pass) {
  const int b = blockIdx.y;
  if (!s.alive[b]) return;
  const int m = s.m[b], n = s.n;
  if (m == 0 || m > MCH) return;
  if (pass == 1 && !s.again[b]) return;
  extern __shared__ cplx tile[];                 // [m][TRI_R]
  cplx *X = blockIdx.z == 0 ? colW(s, b, 0) : colV(s, b, 0);
  const cplx *T = s.G + (long)b * ldg * ldg;
  const int r0 = blockIdx.x * TRI_R, tid = threadIdx.x;
  for (int idx = tid; idx < m * TRI_R; idx += blockDim.x) {
    const int i = idx / TRI_R, e = r0 + idx % TRI_R;
    tile[idx] = e < n ? X[(long)i * n + e] : cmake(0.0, 0.0);
  }
  __syncthreads();
  const int r = tid % TRI_R, e = r0 + r;
  if (e >= n) return;
  for (int j = tid / TRI_R; j < m; j += blockDim.x / TRI_R) {
    const cplx *tj = T + (long)ldg * j;
    cplx acc = cmake(0.0, 0.0);
    for (int i = 0; i <= j; ++i) acc = cfma(tile[i * TRI_R + r], __ldg(&tj[i]), acc);
    X[(long)j * n + e] = acc;
  }
}

__global__ void k_sub_shift_done(SubState s, int ishift) {                // :298 and the per-shift counters
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= s.nrhs || !s.alive[b] || s.m[b] > MCH) return;
  s.sig_old[b] = s.sigma[(long)b * s.nshift + ishift];
  s.done[b] = 0;
  s.iter[b] = 0;
}

// One iteration for the right-hand sides with act != 0: append (new, r) to (W, V) with the Gram-Schmidt step of
// gram_schmidt.f90:94-114 (classical projections, repeated when the norm dropped: "twice is enough"), then the residual of
// linear_solver.f90:361-374 updated by the new column only, the convergence test, and on convergence obtain_result (:486-501).
__global__ void __launch_bounds__(ST) k_sub_step(SubState s, int ishift) {
  const int b = blockIdx.x;
  if (!s.act[b]) return;
  extern __shared__ cplx hs[];                   // [cap] projections of the new column
  __shared__ cplx sm[32];
  const int m = s.m[b], n = s.n, lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
  cplx *wm = colW(s, b, m), *vm = colV(s, b, m);
  cplx *r = s.Rv + (long)b * n;
  const cplx *nv = s.New + (long)b * n, *bb = s.b + (long)b * s.ldb;
  cplx a2 = cmake(0.0, 0.0);
  for (int e = threadIdx.x; e < n; e += blockDim.x) {                     // :426, :432
    const cplx x = nv[e];
    wm[e] = x; vm[e] = r[e];
    a2.x += x.x * x.x; a2.y += x.y * x.y;
  }
  a2 = block_sum(a2, sm);
  double before = a2.x + a2.y;
  for (int pass = 0; pass < 2 && m > 0; ++pass) {
    for (int j = w; j < m; j += nw) {
      const cplx *wj = colW(s, b, j);
      cplx acc = cmake(0.0, 0.0);
      for (int e = lane; e < n; e += 32) acc = cfma(cconj(wj[e]), wm[e], acc);
      acc = warp_sum(acc);
      if (lane == 0) hs[j] = cneg(acc);
    }
    __syncthreads();
    cplx q = cmake(0.0, 0.0);
    for (int e = threadIdx.x; e < n; e += blockDim.x) {
      cplx x = wm[e], y = vm[e];
      for (int j = 0; j < m; ++j) { const cplx h = hs[j]; x = cfma(h, colW(s, b, j)[e], x); y = cfma(h, colV(s, b, j)[e], y); }
      wm[e] = x; vm[e] = y;
      q.x += x.x * x.x; q.y += x.y * x.y;
    }
    q = block_sum(q, sm);                        // (its barriers also order the hs reuse of the next pass)
    const double after = q.x + q.y;
    if (after > 0.5 * before) { before = after; break; }
    before = after;
  }
  const double inv = 1.0 / sqrt(before);                                  // :114
  cplx cm = cmake(0.0, 0.0);
  for (int e = threadIdx.x; e < n; e += blockDim.x) cm = cfma(cconj(wm[e]), bb[e], cm);
  cm = cscale(inv, block_sum(cm, sm));                                    // <w_m | b>  (:367)
  cplx acc = cmake(0.0, 0.0);
  for (int e = threadIdx.x; e < n; e += blockDim.x) {
    const cplx x = cscale(inv, wm[e]);
    wm[e] = x; vm[e] = cscale(inv, vm[e]);
    const cplx v = cfma(cneg(cm), x, r[e]);                                // :368 for the new column
    r[e] = v;
    acc.x += v.x * v.x; acc.y += v.y * v.y;
  }
  acc = block_sum(acc, sm);
  const double nrm = sqrt(acc.x + acc.y);                                  // :373
  cplx *c = s.c + (long)b * s.n;
  if (threadIdx.x == 0) { c[m] = cm; s.m[b] = m + 1; }
  __syncthreads();
  const int m1 = m + 1;
  if (s.iter[b] >= s.max_iter) {                                          // the loop :159 ran out (:176-180): no further test
    if (threadIdx.x == 0) { s.ierr[b] = 1; s.alive[b] = 0; s.act[b] = 0; }
    return;
  }
  if (nrm < s.absthr[b]) {                                                // :374 -> obtain_result, NaN scan (:186-190)
    cplx *x = s.x + ((long)b * s.nshift + ishift) * n;
    int bad = 0;
    for (int e = threadIdx.x; e < n; e += blockDim.x) {
      cplx v = cmake(0.0, 0.0);
      for (int i = 0; i < m1; ++i) v = cfma(c[i], colV(s, b, i)[e], v);
      x[e] = v;
      bad |= (v.x != v.x) || (v.y != v.y);
    }
    bad = __syncthreads_or(bad);
    if (threadIdx.x == 0) {
      s.done[b] = 1; s.act[b] = 0;
      if (bad) { s.ierr[b] = 2; s.alive[b] = 0; }
    }
    return;
  }
  if (threadIdx.x == 0) {
    if (n == m1) {                                                        // :376-378 sets 3, :176-180 overwrites it with 1
      s.ierr[b] = 1; s.alive[b] = 0; s.act[b] = 0;
    } else {
      s.iter[b] += 1;
      atomicAdd(s.count, 1);
      atomicAdd((unsigned long long *)s.nop, 1ull);
    }
  }
}

__global__ void k_sub_finish(SubState s, const int *__restrict__ todo) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= s.nrhs) return;
  if (todo && !todo[b]) return;
  if (s.alive[b]) s.ierr[b] = 0;
}

__global__ void k_copy_cols(int n, int ncol, const cplx *__restrict__ src, long src_rhs_stride, cplx *__restrict__ dst,
                            long dst_rhs_stride) {
  const int b = blockIdx.y;
  const long tot = (long)n * ncol;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < tot; i += (long)gridDim.x * blockDim.x)
    dst[(long)b * dst_rhs_stride + i] = src[(long)b * src_rhs_stride + i];
}

static int subspace_core(sgw_ctx *ctx, const SolveBatch &sb, double threshold, int max_iter, const int *d_todo);

// threads of the one-CTA-per-RHS kernels: the Gram-Schmidt chains are sequences of block reductions, whose latency grows
// with the number of warps, so short vectors get small CTAs (SGW_SUB_THREADS overrides)
static int sub_threads(int n) {
  static int forced = -1;
  if (forced < 0) { const char *e = getenv("SGW_SUB_THREADS"); forced = e ? atoi(e) : 0; }
  if (forced >= 32 && forced <= ST && forced % 32 == 0) return forced;
  (void)n;
  return 512;   // measured on the gw_licl stand-in (n = 544, 102 shifts): 64 -> 4.4 s, 128 -> 2.4 s, 256 -> 2.1 s, 512 -> 1.45 s
}

// dense sub-batch <-> batch (right-hand sides list[0..c) of the caller's batch)
__global__ void k_sub_gather(int n, int nshift, const int *__restrict__ list, const cplx *__restrict__ b, long ldb,
                             const cplx *__restrict__ sigma, cplx *__restrict__ bc, cplx *__restrict__ sc) {
  const int j = blockIdx.y, src = list[j];
  for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < n; e += gridDim.x * blockDim.x) bc[(long)j * n + e] = b[(long)src * ldb + e];
  if (blockIdx.x == 0)
    for (int i = threadIdx.x; i < nshift; i += blockDim.x) sc[(long)j * nshift + i] = sigma[(long)src * nshift + i];
}
__global__ void k_sub_scatter(int n, int nshift, const int *__restrict__ list, const cplx *__restrict__ xc, const int *__restrict__ ic,
                              cplx *__restrict__ x, int *__restrict__ ierr) {
  const int j = blockIdx.y, dst = list[j];
  const long tot = (long)n * nshift;
  for (long e = (long)blockIdx.x * blockDim.x + threadIdx.x; e < tot; e += (long)gridDim.x * blockDim.x)
    x[(long)dst * tot + e] = xc[(long)j * tot + e];
  if (blockIdx.x == 0 && threadIdx.x == 0) ierr[dst] = ic[j];
}

// SGW subspace solver for the right-hand sides flagged in d_todo (all if null).  The basis of every right-hand side costs
// 2 x m x n complex numbers, so the flagged systems -- usually a handful that fell through from the BiCGStab solver
// (select_solver.f90:157) -- are COMPACTED into dense sub-batches whose size is bounded by the free device memory, instead of
// allocating a basis for every right-hand side of the caller's batch.
int subspace_batched(sgw_ctx *ctx, const SolveBatch &sb, double threshold, int max_iter, const int *d_todo) {
  if (sb.nrhs <= 0) return SGW_OK;
  cudaStream_t st = ctx->stream;
  std::vector<int> list;
  if (d_todo) {
    std::vector<int> h(sb.nrhs);
    SGW_CUDA(cudaMemcpyAsync(h.data(), d_todo, sizeof(int) * sb.nrhs, cudaMemcpyDeviceToHost, st));
    SGW_CUDA(cudaStreamSynchronize(st));
    for (int b = 0; b < sb.nrhs; ++b) if (h[b]) list.push_back(b);
  } else {
    list.resize(sb.nrhs);
    for (int b = 0; b < sb.nrhs; ++b) list[b] = b;
  }
  if (list.empty()) return SGW_OK;
  // chunk size: a basis of up to `mguess` vectors per system next to everything else that is already allocated
  size_t free_b = 0, total_b = 0, held = 0;
  if (cudaMemGetInfo(&free_b, &total_b) != cudaSuccess) free_b = (size_t)4 << 30;
  for (const char *nm : {"ss_V", "ss_W", "ss_V2", "ss_W2", "ss_R", "ss_New", "ss_bc", "ss_xc"}) {
    auto it = ctx->ws.bufs.find(nm);
    if (it != ctx->ws.bufs.end()) held += it->second.second;
  }
  const size_t n = sb.n;
  const size_t mguess = std::min<size_t>(n, 96);
  const size_t per = (2 * mguess + 3 + (size_t)sb.nshift) * n * sizeof(cplx);
  const double avail = 0.8 * (double)(free_b + held);
  int chunk = (int)std::max(1.0, std::min((double)list.size(), avail / (double)per));
  chunk = std::min(chunk, 65535);
  const bool whole = !d_todo && chunk >= sb.nrhs;
  if (whole) return subspace_core(ctx, sb, threshold, max_iter, nullptr);     // nothing to compact
  int *d_list = nullptr, *d_ic = nullptr;
  cplx *bc = nullptr, *sc = nullptr, *xc = nullptr;
  SGW_CHECK(ws(ctx, "ss_list", (size_t)chunk, &d_list));
  SGW_CHECK(ws(ctx, "ss_ic", (size_t)chunk, &d_ic));
  SGW_CHECK(ws(ctx, "ss_bc", (size_t)chunk * n, &bc));
  SGW_CHECK(ws(ctx, "ss_sc", (size_t)chunk * sb.nshift, &sc));
  SGW_CHECK(ws(ctx, "ss_xc", (size_t)chunk * sb.nshift * n, &xc));
  for (size_t i0 = 0; i0 < list.size(); i0 += chunk) {
    const int c = (int)std::min<size_t>(chunk, list.size() - i0);
    SGW_CUDA(cudaMemcpyAsync(d_list, list.data() + i0, sizeof(int) * c, cudaMemcpyHostToDevice, st));
    dim3 g(16, (unsigned)c);
    k_sub_gather<<<g, 256, 0, st>>>(sb.n, sb.nshift, d_list, sb.d_b, sb.ldb, sb.d_sigma, bc, sc);
    SGW_LAUNCH_CHECK();
    SolveBatch sub = sb;
    sub.avg = AvgSpec();
    sub.nrhs = c; sub.d_b = bc; sub.ldb = (long)n; sub.d_sigma = sc; sub.d_x = xc; sub.d_ierr = d_ic;
    {
      const std::vector<int> ones(c, 1);           // select_solver.f90:121: ierr = 1 until a solver reports success
      SGW_CUDA(cudaMemcpyAsync(d_ic, ones.data(), sizeof(int) * c, cudaMemcpyHostToDevice, st));
      SGW_CUDA(cudaStreamSynchronize(st));
    }
    SGW_CHECK(subspace_core(ctx, sub, threshold, max_iter, nullptr));
    dim3 g2(64, (unsigned)c);
    k_sub_scatter<<<g2, 256, 0, st>>>(sb.n, sb.nshift, d_list, xc, d_ic, sb.d_x, sb.d_ierr);
    SGW_LAUNCH_CHECK();
    SGW_CUDA(cudaStreamSynchronize(st));          // d_list is rewritten by the next chunk
  }
  return SGW_OK;
}

static int subspace_core_mgs(sgw_ctx *ctx, const SolveBatch &sb, double threshold, int max_iter, const int *d_todo) {
  if (sb.nrhs <= 0) return SGW_OK;
  SubState s;
  s.c = nullptr; s.G = nullptr; s.again = nullptr;
  s.n = sb.n; s.nrhs = sb.nrhs; s.nshift = sb.nshift; s.max_iter = max_iter;
  s.b = sb.d_b; s.ldb = sb.ldb; s.sigma = sb.d_sigma; s.x = sb.d_x; s.ierr = sb.d_ierr;
  const long n = s.n, nr = s.nrhs;
  s.cap = 32;
  if (s.cap > s.n) s.cap = s.n;
  SGW_CHECK(ws(ctx, "ss_V", (size_t)(nr * s.cap * n), &s.V));
  SGW_CHECK(ws(ctx, "ss_W", (size_t)(nr * s.cap * n), &s.W));
  SGW_CHECK(ws(ctx, "ss_R", (size_t)(nr * n), &s.Rv));
  SGW_CHECK(ws(ctx, "ss_New", (size_t)(nr * n), &s.New));
  SGW_CHECK(ws(ctx, "ss_sig", (size_t)nr, &s.sig_old));
  SGW_CHECK(ws(ctx, "ss_thr", (size_t)nr, &s.absthr));
  int *ints = nullptr;
  SGW_CHECK(ws(ctx, "ss_int", (size_t)(5 * nr + 1), &ints));
  s.m = ints; s.act = ints + nr; s.alive = ints + 2 * nr; s.done = ints + 3 * nr; s.iter = ints + 4 * nr; s.count = ints + 5 * nr;
  SGW_CHECK(ws(ctx, "ss_nop", (size_t)1, &s.nop));
  cudaStream_t st = ctx->stream;
  SGW_CUDA(cudaMemsetAsync(s.nop, 0, sizeof(long), st));
  int *h_count = nullptr;
  SGW_CUDA(cudaMallocHost((void **)&h_count, sizeof(int)));
  k_sub_init<<<(unsigned)nr, sub_threads(s.n), 0, st>>>(s, threshold, d_todo);
  SGW_LAUNCH_CHECK();
  int total_cols = 0;   // upper bound of the basis size of any RHS
  int rc = SGW_OK;
  for (int ishift = 0; ishift < s.nshift && rc == SGW_OK; ++ishift) {
    k_sub_shift<<<(unsigned)nr, sub_threads(s.n), 0, st>>>(s, ishift, -1);
    SGW_LAUNCH_CHECK();
    for (int it = 0; it <= max_iter; ++it) {
      SGW_CUDA(cudaMemsetAsync(s.count, 0, sizeof(int), st));
      const size_t dyn = (size_t)(s.cap + 1) * sizeof(cplx);
      if (dyn > 48 * 1024) SGW_CUDA(cudaFuncSetAttribute(k_sub_residual, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn));
      k_sub_residual<<<(unsigned)nr, sub_threads(s.n), dyn, st>>>(s, ishift);
      SGW_LAUNCH_CHECK();
      SGW_CUDA(cudaMemcpyAsync(h_count, s.count, sizeof(int), cudaMemcpyDeviceToHost, st));
      SGW_CUDA(cudaStreamSynchronize(st));
      if (*h_count == 0) break;
      rc = apply_operator(ctx, sb.slot, sb.alpha_pv, s.nrhs, s.Rv, n, s.sigma + ishift, s.nshift, s.New, n, s.act);
      if (rc != SGW_OK) break;
      // grow the subspace storage (the reference reallocates every iteration, :424-433).  total_cols is an upper bound over
      // the batch; no right-hand side can hold more than n basis vectors (k_sub_residual stops it at m == n, :376-378), so
      // a capacity of n columns never needs to grow
      if (total_cols + 1 > s.cap && s.cap < s.n) {
        int ncap = s.cap * 2;
        if (ncap > s.n) ncap = s.n;
        if (ncap <= s.cap || (size_t)(ncap + 1) * sizeof(cplx) > ctx->smem_optin) {
          ctx->err = "subspace solver: basis capacity exhausted";
          rc = SGW_E_UNSUPPORTED;
          break;
        }
        cplx *nV = nullptr, *nW = nullptr;
        const char *nameV = (s.V == (cplx *)ctx->ws.bufs["ss_V"].first) ? "ss_V2" : "ss_V";
        const char *nameW = (s.W == (cplx *)ctx->ws.bufs["ss_W"].first) ? "ss_W2" : "ss_W";
        rc = ws(ctx, nameV, (size_t)(nr * ncap * n), &nV);
        if (rc == SGW_OK) rc = ws(ctx, nameW, (size_t)(nr * ncap * n), &nW);
        if (rc != SGW_OK) break;
        dim3 g(64, (unsigned)nr);
        k_copy_cols<<<g, 256, 0, st>>>(s.n, s.cap, s.V, (long)s.cap * n, nV, (long)ncap * n);
        SGW_LAUNCH_CHECK();
        k_copy_cols<<<g, 256, 0, st>>>(s.n, s.cap, s.W, (long)s.cap * n, nW, (long)ncap * n);
        SGW_LAUNCH_CHECK();
        s.V = nV; s.W = nW; s.cap = ncap;
      }
      k_sub_expand<<<(unsigned)nr, sub_threads(s.n), 0, st>>>(s);
      SGW_LAUNCH_CHECK();
      ++total_cols;
    }
  }
  cudaFreeHost(h_count);
  if (rc != SGW_OK) return rc;
  k_sub_finish<<<(unsigned)((nr + 127) / 128), 128, 0, st>>>(s, d_todo);
  SGW_LAUNCH_CHECK();
  long nop = 0;
  SGW_CUDA(cudaMemcpyAsync(&nop, s.nop, sizeof(long), cudaMemcpyDeviceToHost, st));
  SGW_CUDA(cudaStreamSynchronize(st));
  ctx->stats.n_linear_op += nop;
  return SGW_OK;
}


static bool sub_mgs() {
  const char *e = getenv("SGW_SUB");
  return e && strcmp(e, "mgs") == 0;
}

static int subspace_core(sgw_ctx *ctx, const SolveBatch &sb, double threshold, int max_iter, const int *d_todo) {
  if (sb.nrhs <= 0) return SGW_OK;
  if (sub_mgs()) return subspace_core_mgs(ctx, sb, threshold, max_iter, d_todo);
  SubState s;
  s.n = sb.n; s.nrhs = sb.nrhs; s.nshift = sb.nshift; s.max_iter = max_iter;
  s.b = sb.d_b; s.ldb = sb.ldb; s.sigma = sb.d_sigma; s.x = sb.d_x; s.ierr = sb.d_ierr;
  const long n = s.n, nr = s.nrhs;
  s.cap = std::min(32, s.n);
  cudaStream_t st = ctx->stream;
  SGW_CHECK(ws(ctx, "ss_V", (size_t)(nr * s.cap * n), &s.V));
  SGW_CHECK(ws(ctx, "ss_W", (size_t)(nr * s.cap * n), &s.W));
  // columns beyond a system's basis are read (and ignored) by the batched Gram product: they must hold finite numbers
  SGW_CUDA(cudaMemsetAsync(s.V, 0, sizeof(cplx) * (size_t)(nr * s.cap * n), st));
  SGW_CUDA(cudaMemsetAsync(s.W, 0, sizeof(cplx) * (size_t)(nr * s.cap * n), st));
  SGW_CHECK(ws(ctx, "ss_R", (size_t)(nr * n), &s.Rv));
  SGW_CHECK(ws(ctx, "ss_New", (size_t)(nr * n), &s.New));
  SGW_CHECK(ws(ctx, "ss_c", (size_t)(nr * n), &s.c));
  SGW_CHECK(ws(ctx, "ss_G", (size_t)nr * MCH * MCH, &s.G));
  SGW_CHECK(ws(ctx, "ss_sig", (size_t)nr, &s.sig_old));
  SGW_CHECK(ws(ctx, "ss_thr", (size_t)nr, &s.absthr));
  int *ints = nullptr;
  SGW_CHECK(ws(ctx, "ss_int", (size_t)(6 * nr + 1), &ints));
  s.m = ints; s.act = ints + nr; s.alive = ints + 2 * nr; s.done = ints + 3 * nr; s.iter = ints + 4 * nr; s.again = ints + 5 * nr;
  s.count = ints + 6 * nr;
  SGW_CHECK(ws(ctx, "ss_nop", (size_t)1, &s.nop));
  SGW_CUDA(cudaMemsetAsync(s.nop, 0, sizeof(long), st));
  SGW_CUDA(cudaMemsetAsync(s.again, 0, sizeof(int) * nr, st));
  if (!ctx->h_flags) SGW_CUDA(cudaMallocHost((void **)&ctx->h_flags, 4 * sizeof(int)));
  volatile int *h_count = ctx->h_flags + 2;
  for (int i = 0; i < 2; ++i)
    if (!ctx->ev_iter[i]) SGW_CUDA(cudaEventCreateWithFlags(&ctx->ev_iter[i], cudaEventDisableTiming));
  SGW_CUDA(cudaFuncSetAttribute(k_sub_chol, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(MCH * MCH * sizeof(cplx))));
  SGW_CUDA(cudaFuncSetAttribute(k_sub_tri, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(MCH * TRI_R * sizeof(cplx))));
  const int nt = sub_threads(s.n);
  const unsigned gb = (unsigned)((nr + 127) / 128);
  k_sub_init<<<(unsigned)nr, nt, 0, st>>>(s, threshold, d_todo);
  SGW_LAUNCH_CHECK();
  int total_cols = 0;   // upper bound of the basis size of any RHS
  for (int ishift = 0; ishift < s.nshift; ++ishift) {
    // ---- orthogonal_subspace (:272-300)
    if (total_cols > 0) {
      ProfScope prof(ctx, PC_OTHER);
      const int mg = std::min(total_cols, MCH);
      dim3 gw((unsigned)((n + 255) / 256), (unsigned)mg, (unsigned)nr);
      k_sub_wshift<<<gw, 256, 0, st>>>(s, ishift, mg);
      SGW_LAUNCH_CHECK();
      for (int pass = 0; pass < 2; ++pass) {
        SGW_CHECK(gemm_ch_n_batched(ctx, mg, mg, s.n, s.W, n, (long)s.cap * n, s.W, n, (long)s.cap * n, s.G, mg, (long)mg * mg, (int)nr));
        k_sub_chol<<<(unsigned)nr, 256, (size_t)mg * mg * sizeof(cplx), st>>>(s, mg, pass);
        SGW_LAUNCH_CHECK();
        dim3 gt((unsigned)((n + TRI_R - 1) / TRI_R), (unsigned)nr, 2);
        k_sub_tri<<<gt, 256, (size_t)mg * TRI_R * sizeof(cplx), st>>>(s, mg, pass);
        SGW_LAUNCH_CHECK();
      }
      if (total_cols > MCH) {                      // bases too large for the shared-memory factorisation
        k_sub_shift<<<(unsigned)nr, nt, 0, st>>>(s, ishift, MCH);
        SGW_LAUNCH_CHECK();
      }
    }
    k_sub_shift_done<<<gb, 128, 0, st>>>(s, ishift);
    SGW_LAUNCH_CHECK();
    // ---- first residual of the shift from all overlaps (:312-383); later ones are updated by k_sub_step
    const size_t dyn = (size_t)(s.cap + 1) * sizeof(cplx);
    if (dyn > 48 * 1024) {
      SGW_CUDA(cudaFuncSetAttribute(k_sub_residual, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn));
      SGW_CUDA(cudaFuncSetAttribute(k_sub_step, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn));
    }
    SGW_CUDA(cudaMemsetAsync(s.count, 0, sizeof(int), st));
    k_sub_residual<<<(unsigned)nr, nt, dyn, st>>>(s, ishift);
    SGW_LAUNCH_CHECK();
    int k = 0;                                     // counts read back so far; the host stays one iteration ahead of the device
    SGW_CUDA(cudaMemcpyAsync((void *)&h_count[0], s.count, sizeof(int), cudaMemcpyDeviceToHost, st));
    SGW_CUDA(cudaEventRecord(ctx->ev_iter[0], st));
    for (;;) {
      SGW_CUDA(cudaMemsetAsync(s.count, 0, sizeof(int), st));
      SGW_CHECK(apply_operator(ctx, sb.slot, sb.alpha_pv, s.nrhs, s.Rv, n, s.sigma + ishift, s.nshift, s.New, n, s.act));
      // grow the subspace storage (the reference reallocates every iteration, :424-433).  total_cols is an upper bound over
      // the batch; no right-hand side can hold more than n basis vectors (:376-378), so a capacity of n columns never grows
      if (total_cols + 1 > s.cap && s.cap < s.n) {
        const int ncap = std::min(s.cap * 2, s.n);
        if ((size_t)(ncap + 1) * sizeof(cplx) > ctx->smem_optin) {
          ctx->err = "subspace solver: basis capacity exhausted";
          return SGW_E_UNSUPPORTED;
        }
        cplx *nV = nullptr, *nW = nullptr;
        const char *nameV = (s.V == (cplx *)ctx->ws.bufs["ss_V"].first) ? "ss_V2" : "ss_V";
        const char *nameW = (s.W == (cplx *)ctx->ws.bufs["ss_W"].first) ? "ss_W2" : "ss_W";
        SGW_CHECK(ws(ctx, nameV, (size_t)(nr * ncap * n), &nV));
        SGW_CHECK(ws(ctx, nameW, (size_t)(nr * ncap * n), &nW));
        SGW_CUDA(cudaMemsetAsync(nV, 0, sizeof(cplx) * (size_t)(nr * ncap * n), st));
        SGW_CUDA(cudaMemsetAsync(nW, 0, sizeof(cplx) * (size_t)(nr * ncap * n), st));
        dim3 g(64, (unsigned)nr);
        k_copy_cols<<<g, 256, 0, st>>>(s.n, s.cap, s.V, (long)s.cap * n, nV, (long)ncap * n);
        SGW_LAUNCH_CHECK();
        k_copy_cols<<<g, 256, 0, st>>>(s.n, s.cap, s.W, (long)s.cap * n, nW, (long)ncap * n);
        SGW_LAUNCH_CHECK();
        s.V = nV; s.W = nW; s.cap = ncap;
      }
      const size_t dyn2 = (size_t)(s.cap + 1) * sizeof(cplx);
      if (dyn2 > 48 * 1024) {
        SGW_CUDA(cudaFuncSetAttribute(k_sub_residual, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn2));
        SGW_CUDA(cudaFuncSetAttribute(k_sub_step, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn2));
      }
      {
        ProfScope prof(ctx, PC_OTHER);
        k_sub_step<<<(unsigned)nr, nt, dyn2, st>>>(s, ishift);
        SGW_LAUNCH_CHECK();
      }
      ++total_cols;
      ++k;
      SGW_CUDA(cudaMemcpyAsync((void *)&h_count[k & 1], s.count, sizeof(int), cudaMemcpyDeviceToHost, st));
      SGW_CUDA(cudaEventRecord(ctx->ev_iter[k & 1], st));
      SGW_CUDA(cudaEventSynchronize(ctx->ev_iter[(k - 1) & 1]));
      if (h_count[(k - 1) & 1] == 0) {             // nothing was active in the iteration enqueued last: it was a no-op
        --total_cols;
        break;
      }
      if (k > max_iter + 1) break;
    }
  }
  k_sub_finish<<<gb, 128, 0, st>>>(s, d_todo);
  SGW_LAUNCH_CHECK();
  long nop = 0;
  SGW_CUDA(cudaMemcpyAsync(&nop, s.nop, sizeof(long), cudaMemcpyDeviceToHost, st));
  SGW_CUDA(cudaStreamSynchronize(st));
  ctx->stats.n_linear_op += nop;
  return SGW_OK;
}

// ---------------------------------------------------------------- select_solver.f90:67-161
__global__ void k_sel_init(int nrhs, int *ierr, int *todo) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= nrhs) return;
  ierr[b] = 1;     // :121
  todo[b] = 1;
}
__global__ void k_sel_update(int nrhs, const int *ierr, int *todo, int *nleft) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= nrhs) return;
  todo[b] = ierr[b] != 0;      // :157 only systems that did not converge try the next solver
  if (todo[b]) atomicAdd(nleft, 1);
}

// solve_linter.f90:464-480 for the right-hand sides whose average the solver did not form itself (AvgSpec.d_done == 0)
__global__ void k_average_masked(int n, int nshift, AvgSpec a, const cplx *__restrict__ x) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  const int b = blockIdx.y, f = blockIdx.z;
  if (e >= n || a.d_done[b]) return;
  const cplx *xr = x + (long)b * nshift * n;
  cplx v = xr[(long)f * n + e];
  const int first = a.zero_freq ? 1 : 0;
  if (f >= first) {
    const cplx w = xr[(long)(a.nfreq + f - first) * n + e];
    v = cscale(0.5, v);                          // ZSCAL 0.5 then ZAXPY 0.5 (:469-478)
    v = cmake(v.x + 0.5 * w.x, v.y + 0.5 * w.y);
  }
  a.d_y[(((long)(b / a.group) * a.nfreq + f) * a.group + b % a.group) * n + e] = v;
}

int select_solver_batched(sgw_ctx *ctx, const SolveBatch &sb, const sgw_solver_cfg *cfg) {
  int *todo = nullptr, *nleft = nullptr;
  if (sb.avg.d_y) {
    if (!sb.avg.d_done || sb.avg.group < 1 || 2 * sb.avg.nfreq - (sb.avg.zero_freq ? 1 : 0) != sb.nshift) {
      ctx->err = "select_solver: inconsistent average specification";
      return SGW_E_ARG;
    }
    SGW_CUDA(cudaMemsetAsync(sb.avg.d_done, 0, sizeof(int) * sb.nrhs, ctx->stream));
  }
  SGW_CHECK(ws(ctx, "sel_todo", (size_t)sb.nrhs, &todo));
  SGW_CHECK(ws(ctx, "sel_nleft", (size_t)1, &nleft));
  const int gb = (sb.nrhs + 127) / 128;
  k_sel_init<<<gb, 128, 0, ctx->stream>>>(sb.nrhs, sb.d_ierr, todo);
  SGW_LAUNCH_CHECK();
  for (int is = 0; is < cfg->npriority; ++is) {
    switch (cfg->priority[is]) {
      case 1:   // bicgstab_multi :134-138
      case 2:   // bicgstab_no_multi :140-146 calls the multishift routine with the whole sigma/xx num_shift times:
                // identical inputs -> identical outputs, so one call reproduces its result
        SGW_CHECK(bicgstab_batched(ctx, sb, cfg->bicg_lmax, cfg->threshold, cfg->max_iter, todo));
        break;
      case 3:   // sgw_linear_solver :148-152
        SGW_CHECK(subspace_batched(ctx, sb, cfg->threshold, cfg->max_iter, todo));
        break;
      default:
        ctx->err = "unknown solver in priority list";
        return SGW_E_ARG;
    }
    SGW_CUDA(cudaMemsetAsync(nleft, 0, sizeof(int), ctx->stream));
    k_sel_update<<<gb, 128, 0, ctx->stream>>>(sb.nrhs, sb.d_ierr, todo, nleft);
    SGW_LAUNCH_CHECK();
    int h = 0;
    SGW_CUDA(cudaMemcpyAsync(&h, nleft, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    SGW_CUDA(cudaStreamSynchronize(ctx->stream));
    if (h == 0) break;
    if (ctx->msg_fn) {       // the reference writes these lines to stdout once per right-hand side; here once per batch
      char buf[256];
      if (cfg->priority[is] == 3)
        snprintf(buf, sizeof buf, "WARNING: SternheimerGW linear solver did not converge (%d of %d right-hand sides)", h, sb.nrhs);   // linear_solver.f90:177
      else
        snprintf(buf, sizeof buf, "WARNING: BiCGstab algorithm did not converge in %d iterations. (%d of %d right-hand sides)",
                 cfg->max_iter, h, sb.nrhs);                                                                                          // bicgstab.f90:250
      ctx->msg_fn(buf, ctx->msg_user);
      if (is + 1 < cfg->npriority) ctx->msg_fn("First choice of solver did not converge, try a different one", ctx->msg_user);       // select_solver.f90:126
    }
    if (is + 1 < cfg->npriority) ctx->stats.n_fallback += h;
  }
  if (sb.avg.d_y) {
    const int bstep = std::max(1, 65535 / sb.avg.group) * sb.avg.group;   // grid.y limit, group-aligned so that the layout formula holds per chunk
    for (int b0 = 0; b0 < sb.nrhs; b0 += bstep) {
      const int nb = std::min(bstep, sb.nrhs - b0);
      AvgSpec a = sb.avg;
      a.d_y = sb.avg.d_y + (long)(b0 / a.group) * a.nfreq * a.group * sb.n;
      a.d_done = sb.avg.d_done + b0;
      dim3 grid((unsigned)((sb.n + 255) / 256), (unsigned)nb, (unsigned)a.nfreq);
      k_average_masked<<<grid, 256, 0, ctx->stream>>>(sb.n, sb.nshift, a, sb.d_x + (long)b0 * sb.nshift * sb.n);
      SGW_LAUNCH_CHECK();
    }
  }
  return SGW_OK;
}

}  // namespace sgw
