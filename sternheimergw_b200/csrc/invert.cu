// invert.cu -- invert_epsilon (phys/coul/src/invert_epsilon.f90:23-90) -> sgw_invert_epsilon
//
// The reference inverts eps(G,G',omega) one frequency at a time with ZGETRF + ZGETRI (:59-66).  Here all nfs matrices are
// inverted together, in place, by a two-level BLOCKED Gauss-Jordan elimination with partial pivoting (the IZAMAX criterion
// |re| + |im| of ZGETRF, first maximum wins), so that all but O(n^2 NB) of the 8 n^3 flops per matrix are rank-NB updates on
// the FP64 tensor path (k_zgemm, batched over the frequencies):
//
//   for every panel P of NB columns
//     for every sub-panel S of SB columns of P
//       k_gjb_subpanel   unblocked Gauss-Jordan on the n x SB sub-panel held in shared memory (one CTA per frequency):
//                        pivot search, row swap, scaling and the rank-1 update of the sub-panel's own columns
//       k_gjb_prep       apply the sub-panel's row swaps to the other columns of P, move their pivot rows to R, zero them
//       k_zgemm (batch)  P \ S  +=  S_transformed (n x SB)  *  R (SB x |P \ S|)
//     k_gjb_prep, k_zgemm (batch)   the same for the columns outside P with the transformed panel (K = NB)
//   k_gj_colswap         undo the row interchanges on the columns (ZGETRI's final sweep)
//
// Why that is the same elimination: step k of the in-place algorithm multiplies the matrix by G_k P_k (P_k the interchange,
// G_k the identity with column k replaced by the multipliers) and stores column k of G_k where e_k would appear.  Later
// interchanges only touch rows > k, so G_last P_last ... G_k0 P_k0 = W (P_last ... P_k0) with W the identity outside the
// panel's columns, and the in-place panel sweep leaves exactly W's panel columns T in the panel.  For a column j outside the
// panel  W a_j = (a_j with the panel's rows zeroed) + T a_j[panel rows], after the interchanges -- the update above.
// Matrices whose sub-panel does not fit in shared memory (n > ~7000) take the same path with the sub-panel steps done by
// two kernels per column in global memory (k_gj_pivot / k_gj_elim), which with SB = NB = n is the unblocked algorithm.
#include "internal.cuh"

#include <stdlib.h>
#include <string.h>

#include <algorithm>

using namespace sgw;

namespace {

__global__ void k_eps_wings(int n, cplx *__restrict__ a) {          // invert_epsilon.f90:46-56, :72-81
  cplx *m = a + (long)blockIdx.y * n * n;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n || i == 0) return;
  m[i] = cmake(0.0, 0.0);
  m[(long)n * i] = cmake(0.0, 0.0);
}
__global__ void k_eps_diag(int n, cplx *__restrict__ a) {           // :84-88
  cplx *m = a + (long)blockIdx.y * n * n;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  m[(long)i + (long)n * i].x -= 1.0;
}

// block-wide argmax of (val, lowest index on ties); result valid in every thread of warp 0
__device__ __forceinline__ void argmax_block(double &best, int &bi, int n, double *sval, int *sidx) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
  for (int o = 16; o > 0; o >>= 1) {
    const double ob = __shfl_xor_sync(0xffffffffu, best, o);
    const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
    if (ob > best || (ob == best && oi < bi)) { best = ob; bi = oi; }
  }
  if (lane == 0) { sval[w] = best; sidx[w] = bi; }
  __syncthreads();
  if (w == 0) {
    best = lane < nw ? sval[lane] : -1.0;
    bi = lane < nw ? sidx[lane] : n;
    for (int o = 16; o > 0; o >>= 1) {
      const double ob = __shfl_xor_sync(0xffffffffu, best, o);
      const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
      if (ob > best || (ob == best && oi < bi)) { best = ob; bi = oi; }
    }
  }
}

// Unblocked Gauss-Jordan on the sub-panel of columns [s0, s0 + sb) of every matrix, all n rows, held in shared memory
// (column-major, pitch n).  One CTA per matrix.
__global__ void __launch_bounds__(1024) k_gjb_subpanel(int n, int s0, int sb, cplx *__restrict__ a, int *__restrict__ piv,
                                                        int *__restrict__ info) {
  cplx *m = a + (long)blockIdx.x * n * n + (long)n * s0;
  int *pv = piv + (long)blockIdx.x * n;
  extern __shared__ cplx pan[];                                 // [sb][n]
  __shared__ double sval[32];
  __shared__ int sidx[32];
  const int tid = threadIdx.x, nt = blockDim.x;
  for (int e = tid; e < n * sb; e += nt) pan[e] = m[e];         // the sub-panel's columns are contiguous in memory
  __syncthreads();
  for (int kk = 0; kk < sb; ++kk) {
    const int k = s0 + kk;
    const cplx *ck = pan + (long)kk * n;
    double best = -1.0;
    int bi = n;
    for (int i = k + tid; i < n; i += nt) {
      const double s = fabs(ck[i].x) + fabs(ck[i].y);
      if (s > best) { best = s; bi = i; }
    }
    argmax_block(best, bi, n, sval, sidx);
    if (tid < 32) {                                             // warp 0: interchange rows k <-> p and scale row k
      const int p = bi;
      cplx t = cmake(0.0, 0.0);
      if (tid < sb && best > 0.0) {
        t = pan[(long)tid * n + p];
        if (p != k) pan[(long)tid * n + p] = pan[(long)tid * n + k];
      }
      const double pr = __shfl_sync(0xffffffffu, t.x, kk), pi = __shfl_sync(0xffffffffu, t.y, kk);
      const cplx inv = best > 0.0 ? cdiv(cmake(1.0, 0.0), cmake(pr, pi)) : cmake(0.0, 0.0);
      if (tid < sb && best > 0.0) pan[(long)tid * n + k] = (tid == kk) ? inv : cmul(t, inv);
      if (tid == 0) {
        pv[k] = best > 0.0 ? p : k;
        if (!(best > 0.0)) atomicMax(info, k + 1);
      }
    }
    __syncthreads();
    for (int i = tid; i < n; i += nt) {                         // rank-1 update of the other rows; column k restarts at 0
      if (i == k) continue;
      const cplx f = cneg(pan[(long)kk * n + i]);
      for (int j = 0; j < sb; ++j) {
        const cplx cur = (j == kk) ? cmake(0.0, 0.0) : pan[(long)j * n + i];
        pan[(long)j * n + i] = cfma(f, pan[(long)j * n + k], cur);
      }
    }
    __syncthreads();
  }
  for (int e = tid; e < n * sb; e += nt) m[e] = pan[e];
}

// The same two half-steps on global memory for column k, restricted to the columns [c0, c1) (sub-panel too large for
// shared memory; with c0 = 0, c1 = n it is the unblocked algorithm).
__global__ void __launch_bounds__(1024) k_gj_pivot(int n, int k, int c0, int c1, cplx *__restrict__ a, int *__restrict__ piv,
                                                    cplx *__restrict__ colk, int *__restrict__ info) {
  cplx *m = a + (long)blockIdx.x * n * n;
  int *pv = piv + (long)blockIdx.x * n;
  cplx *ck = colk + (long)blockIdx.x * n;
  __shared__ double sval[32];
  __shared__ int sidx[32];
  __shared__ int s_p;
  __shared__ cplx s_inv;
  const int tid = threadIdx.x;
  double best = -1.0;
  int bi = n;
  for (int i = k + tid; i < n; i += blockDim.x) {
    const cplx v = m[(long)i + (long)n * k];
    const double s = fabs(v.x) + fabs(v.y);
    if (s > best) { best = s; bi = i; }
  }
  argmax_block(best, bi, n, sval, sidx);
  if (tid == 0) {
    if (!(best > 0.0)) { atomicMax(info, k + 1); s_inv = cmake(0.0, 0.0); bi = k; }
    else s_inv = cdiv(cmake(1.0, 0.0), m[(long)bi + (long)n * k]);
    s_p = bi;
    pv[k] = bi;
  }
  __syncthreads();
  const int p = s_p;
  const cplx inv = s_inv;
  for (int j = c0 + tid; j < c1; j += blockDim.x) {           // swap rows k <-> p, scale row k
    cplx akj = m[(long)p + (long)n * j];
    if (p != k) m[(long)p + (long)n * j] = m[(long)k + (long)n * j];
    m[(long)k + (long)n * j] = (j == k) ? inv : cmul(akj, inv);
  }
  __syncthreads();
  for (int i = tid; i < n; i += blockDim.x) {                 // multipliers; column k of the other rows restarts at 0
    if (i == k) { ck[i] = cmake(0.0, 0.0); continue; }
    ck[i] = m[(long)i + (long)n * k];
    m[(long)i + (long)n * k] = cmake(0.0, 0.0);
  }
}
__global__ void __launch_bounds__(256) k_gj_elim(int n, int k, int c0, int c1, cplx *__restrict__ a, const cplx *__restrict__ colk) {
  cplx *m = a + (long)blockIdx.z * n * n;
  const cplx *ck = colk + (long)blockIdx.z * n;
  const int i = blockIdx.x * 64 + (threadIdx.x & 63);
  const int j0 = c0 + (blockIdx.y * 4 + (threadIdx.x >> 6)) * 8;
  if (i >= n || i == k) return;
  const cplx f = cneg(ck[i]);
#pragma unroll
  for (int jj = 0; jj < 8; ++jj) {
    const int j = j0 + jj;
    if (j < c1) m[(long)i + (long)n * j] = cfma(f, m[(long)k + (long)n * j], m[(long)i + (long)n * j]);
  }
}

// For the columns j of [c0, c1) outside [r0, r0 + rb): apply the interchanges recorded for the pivots r0 .. r0 + rb - 1, move
// the rows r0 .. r0 + rb - 1 to R(:, j) (rb x n per matrix, leading dimension rb) and zero them in the matrix.  Matrix 0 also
// writes the list of those columns for the batched update.
__global__ void k_gjb_prep(int n, int r0, int rb, int c0, int c1, cplx *__restrict__ a, const int *__restrict__ piv,
                           cplx *__restrict__ R, int *__restrict__ list, int *__restrict__ count) {
  const int jj = blockIdx.x * blockDim.x + threadIdx.x;
  const int ncol = (c1 - c0) - rb;
  if (jj >= ncol) return;
  const int j = (c0 + jj < r0) ? c0 + jj : c0 + jj + rb;
  if (blockIdx.y == 0) {
    list[jj] = j;
    if (jj == 0) *count = ncol;
  }
  cplx *col = a + (long)blockIdx.y * n * n + (long)n * j;
  const int *pv = piv + (long)blockIdx.y * n;
  cplx *r = R + ((long)blockIdx.y * n + j) * rb;
  for (int kk = 0; kk < rb; ++kk) {
    const int k = r0 + kk, p = pv[k];
    if (p != k) {
      const cplx t = col[k];
      col[k] = col[p];
      col[p] = t;
    }
  }
  for (int kk = 0; kk < rb; ++kk) {
    r[kk] = col[r0 + kk];
    col[r0 + kk] = cmake(0.0, 0.0);
  }
}

__global__ void k_gj_colswap(int n, cplx *__restrict__ a, const int *__restrict__ piv) {
  cplx *m = a + (long)blockIdx.y * n * n;
  const int *pv = piv + (long)blockIdx.y * n;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  for (int k = n - 1; k >= 0; --k) {
    const int p = pv[k];
    if (p != k) {
      const cplx t = m[(long)i + (long)n * k];
      m[(long)i + (long)n * k] = m[(long)i + (long)n * p];
      m[(long)i + (long)n * p] = t;
    }
  }
}

int env_int(const char *name, int dflt) {
  const char *e = getenv(name);
  return e && *e ? atoi(e) : dflt;
}

}  // namespace

// all nfs matrices of d_a (n x n each, column-major, contiguous) are replaced by their inverses; *d_info = 1 + the largest
// column with a zero pivot (0: none)
static int gauss_jordan_batched(sgw_ctx *ctx, int n, int nfs, cplx *d_a, int *d_piv, int *d_info) {
  cudaStream_t st = ctx->stream;
  // panel widths: NB = K of the big updates; SB = the widest power of two whose n x SB sub-panel fits in shared memory
  int NB = std::max(1, env_int("SGW_GJ_NB", 64));
  const size_t cap = ctx->smem_optin > 4096 ? ctx->smem_optin - 2048 : 0;
  int SB = 16;
  while (SB > 1 && (size_t)n * SB * sizeof(cplx) > cap) SB >>= 1;
  bool in_smem = (size_t)n * SB * sizeof(cplx) <= cap && SB >= 2;
  if (const char *e = getenv("SGW_GJ_PANEL")) { if (!strcmp(e, "global")) in_smem = false; }
  if (!in_smem) SB = 16;
  SB = std::min(std::max(1, env_int("SGW_GJ_SB", SB)), in_smem ? SB : n);
  if (n <= 2 * SB || env_int("SGW_GJ_UNBLOCKED", 0)) { NB = n; if (!in_smem) SB = n; }   // small matrices: one panel
  NB = std::max(NB, SB);
  cplx *d_R = nullptr, *d_colk = nullptr;
  int *d_list = nullptr;
  if (NB < n || SB < NB) {
    SGW_CHECK(ws(ctx, "ie_R", (size_t)nfs * n * NB, &d_R));
    SGW_CHECK(ws(ctx, "ie_list", (size_t)n + 1, &d_list));
  }
  if (!in_smem) SGW_CHECK(ws(ctx, "ie_colk", (size_t)n * nfs, &d_colk));
  const int pt = n >= 512 ? 1024 : (n >= 128 ? 256 : 64);
  if (in_smem) SGW_CUDA(cudaFuncSetAttribute(k_gjb_subpanel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)((size_t)n * SB * sizeof(cplx))));
  auto update = [&](int r0, int rb, int c0, int c1) -> int {  // columns [c0, c1) \ [r0, r0 + rb) += T R
    const int ncol = (c1 - c0) - rb;
    if (ncol <= 0) return SGW_OK;
    dim3 gp((ncol + 127) / 128, nfs);
    k_gjb_prep<<<gp, 128, 0, st>>>(n, r0, rb, c0, c1, d_a, d_piv, d_R, d_list, d_list + n);
    SGW_LAUNCH_CHECK();
    return gemm_n_n_batched(ctx, n, ncol, rb, cmake(1, 0), d_a + (long)n * r0, n, (long)n * n, d_R, rb, (long)rb * n, cmake(1, 0), d_a, n,
                            (long)n * n, nfs, d_list, d_list + n);
  };
  for (int P = 0; P < n; P += NB) {
    const int pb = std::min(NB, n - P);
    for (int S = P; S < P + pb; S += SB) {
      const int sb = std::min(SB, P + pb - S);
      if (in_smem) {
        k_gjb_subpanel<<<nfs, pt, (size_t)n * sb * sizeof(cplx), st>>>(n, S, sb, d_a, d_piv, d_info);
        SGW_LAUNCH_CHECK();
      } else {
        const dim3 ge((n + 63) / 64, (sb + 31) / 32, nfs);
        for (int k = S; k < S + sb; ++k) {
          k_gj_pivot<<<nfs, pt, 0, st>>>(n, k, S, S + sb, d_a, d_piv, d_colk, d_info);
          SGW_LAUNCH_CHECK();
          k_gj_elim<<<ge, 256, 0, st>>>(n, k, S, S + sb, d_a, d_colk);
          SGW_LAUNCH_CHECK();
        }
      }
      SGW_CHECK(update(S, sb, P, P + pb));
    }
    SGW_CHECK(update(P, pb, 0, n));
  }
  const dim3 g1((n + 127) / 128, nfs);
  k_gj_colswap<<<g1, 128, 0, st>>>(n, d_a, d_piv);
  SGW_LAUNCH_CHECK();
  return SGW_OK;
}

extern "C" int sgw_invert_epsilon(sgw_ctx *ctx, int ngc, int nfs, sgw_cplx *scrcoul_g, int lgamma) {
  if (!ctx) return SGW_E_ARG;
  cudaSetDevice(ctx->device);
  SGW_ARG(ngc > 0 && nfs > 0 && scrcoul_g, "bad argument");
  SGW_ARG(nfs <= 65535, "invert_epsilon: more than 65535 frequencies");
  begin_call(ctx);
  cudaStream_t st = ctx->stream;
  cplx *d_a = nullptr;
  int *d_piv = nullptr, *d_info = nullptr;
  const size_t tot = (size_t)ngc * ngc * nfs;
  SGW_CHECK(ws(ctx, "ie_a", tot, &d_a));
  SGW_CHECK(ws(ctx, "ie_piv", (size_t)ngc * nfs + 1, &d_piv));
  d_info = d_piv + (size_t)ngc * nfs;
  SGW_CHECK(h2d_large(ctx, d_a, scrcoul_g, sizeof(cplx) * tot));
  SGW_CUDA(cudaMemsetAsync(d_info, 0, sizeof(int), st));
  const dim3 g1((ngc + 127) / 128, nfs);
  if (lgamma) { k_eps_wings<<<g1, 128, 0, st>>>(ngc, d_a); SGW_LAUNCH_CHECK(); }            // :46-56
  cudaEventRecord(ctx->ev2, st);
  SGW_CHECK(gauss_jordan_batched(ctx, ngc, nfs, d_a, d_piv, d_info));                      // :59-66
  cudaEventRecord(ctx->ev3, st);
  if (lgamma) { k_eps_wings<<<g1, 128, 0, st>>>(ngc, d_a); SGW_LAUNCH_CHECK(); }            // :72-81
  k_eps_diag<<<g1, 128, 0, st>>>(ngc, d_a);                                                 // :84-88
  SGW_LAUNCH_CHECK();
  int info = 0;
  SGW_CUDA(cudaMemcpyAsync(&info, d_info, sizeof(int), cudaMemcpyDeviceToHost, st));
  SGW_CUDA(cudaStreamSynchronize(st));
  SGW_CHECK(d2h_large(ctx, scrcoul_g, d_a, sizeof(cplx) * tot));
  end_call(ctx);
  float ms_gj = 0.f;
  if (cudaEventElapsedTime(&ms_gj, ctx->ev2, ctx->ev3) == cudaSuccess) ctx->stats.ms_solver = ms_gj;   // the elimination alone (ms_total includes the copies)
  if (info != 0) {                                                                          // :61,:64 errore
    ctx->err = "invert_epsilon: matrix is singular (zero pivot in column " + std::to_string(info) + ")";
    return SGW_E_ARG;
  }
  return SGW_OK;
}
