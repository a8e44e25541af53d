// gemm.cu -- complex-FP64 GEMMs on the FP64 tensor path (DMMA, mma.sync.m8n8k4.f64) for
//   * the non-local pseudopotential  h psi += vkb (D (vkb^H psi))        ([QE] calbec + add_vuspsi)
//   * the valence projector          A psi += alpha_pv evq (evq^H psi)   (linear_op.f90:147-206, ZGEMM x2)
//   * orthogonalize ([QE]) and the dense test backend (linear_solver.pf:106)
// Both rank-k updates of H.psi share one pass: P = [vkb | evq] is stored as one npwx x (nkb+nbnd) panel,
// coef = P^H psi (split-K, deterministic two-pass reduction), coef' = blockdiag(D, alpha_pv I) coef,
// out = P coef'.  tcgen05 has no f64 kind, so the legacy mma.sync DMMA path is the FP64 tensor route on sm_100a.
#include "internal.cuh"

namespace sgw {

constexpr int BM = 64, BN = 32, BK = 16, GT = 256;
constexpr int PK = BK + 4;   // k-contiguous smem pitch (doubles): 8 rows x 4 k hit 32 distinct bank pairs
constexpr int PM = BM + 8;   // m-contiguous smem pitch

__device__ __forceinline__ void dmma(double &c0, double &c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(c0), "+d"(c1)
               : "d"(a), "d"(b));
}

// C(M x N) = alpha * op(A) * B + beta * C ; op(A) = A^H (A_KCONTIG: A is K x M, column-major) or A (M x K).
// B is K x N column-major.  blockIdx.z selects a K range (split-K): partial results go to C + z * split_stride.
template <bool A_KCONTIG, bool CONJA>
__global__ void __launch_bounds__(GT) k_zgemm(int M, int N, int K, const cplx *__restrict__ A, long lda,
                                               const cplx *__restrict__ B, long ldb, cplx *__restrict__ C, long ldc,
                                               cplx alpha, cplx beta, int kchunk, long split_stride,
                                               const int *__restrict__ active) {
  const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN;
  const int kbeg = blockIdx.z * kchunk, kend = min(K, kbeg + kchunk);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (active) {
    bool any = false;
    for (int n = n0; n < min(N, n0 + BN); ++n) any |= (active[n] != 0);
    if (!any) return;
  }
  constexpr int ASZ = A_KCONTIG ? BM * PK : BK * PM;
  __shared__ double As_re[ASZ], As_im[ASZ], Bs_re[BN * PK], Bs_im[BN * PK];
  const int wm = (warp & 3) * 16, wn = (warp >> 2) * 16;
  const int g = lane >> 2, t = lane & 3;
  double cre[2][2][2], cim[2][2][2];
#pragma unroll
  for (int i = 0; i < 2; ++i)
#pragma unroll
    for (int j = 0; j < 2; ++j) cre[i][j][0] = cre[i][j][1] = cim[i][j][0] = cim[i][j][1] = 0.0;

  for (int k0 = kbeg; k0 < kend; k0 += BK) {
    // ---- stage tiles
    if (A_KCONTIG) {
      for (int i = tid; i < BM * BK; i += GT) {
        const int k = i % BK, m = i / BK;
        cplx v = cmake(0.0, 0.0);
        if (k0 + k < kend && m0 + m < M) v = A[(long)(k0 + k) + (long)(m0 + m) * lda];
        As_re[m * PK + k] = v.x;
        As_im[m * PK + k] = v.y;
      }
    } else {
      for (int i = tid; i < BM * BK; i += GT) {
        const int m = i % BM, k = i / BM;
        cplx v = cmake(0.0, 0.0);
        if (k0 + k < kend && m0 + m < M) v = A[(long)(m0 + m) + (long)(k0 + k) * lda];
        As_re[k * PM + m] = v.x;
        As_im[k * PM + m] = v.y;
      }
    }
    for (int i = tid; i < BN * BK; i += GT) {
      const int k = i % BK, n = i / BK;
      cplx v = cmake(0.0, 0.0);
      if (k0 + k < kend && n0 + n < N) v = B[(long)(k0 + k) + (long)(n0 + n) * ldb];
      Bs_re[n * PK + k] = v.x;
      Bs_im[n * PK + k] = v.y;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < BK; kk += 4) {
      double are[2], aim[2], bre[2], bim[2];
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        const int m = wm + i * 8 + g;
        const int idx = A_KCONTIG ? m * PK + kk + t : (kk + t) * PM + m;
        are[i] = As_re[idx];
        aim[i] = As_im[idx];
      }
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        const int n = wn + j * 8 + g;
        bre[j] = Bs_re[n * PK + kk + t];
        bim[j] = Bs_im[n * PK + kk + t];
      }
#pragma unroll
      for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int j = 0; j < 2; ++j) {
          // op(A) = conj(A)^T: re += ar*br + ai*bi ; im += ar*bi - ai*br.  op(A) = A: re += ar*br - ai*bi ; im += ar*bi + ai*br
          dmma(cre[i][j][0], cre[i][j][1], are[i], bre[j]);
          dmma(cre[i][j][0], cre[i][j][1], CONJA ? aim[i] : -aim[i], bim[j]);
          dmma(cim[i][j][0], cim[i][j][1], are[i], bim[j]);
          dmma(cim[i][j][0], cim[i][j][1], CONJA ? -aim[i] : aim[i], bre[j]);
        }
    }
    __syncthreads();
  }
  cplx *Cz = C + (long)blockIdx.z * split_stride;
  const bool has_beta = (beta.x != 0.0 || beta.y != 0.0);
#pragma unroll
  for (int i = 0; i < 2; ++i)
#pragma unroll
    for (int j = 0; j < 2; ++j)
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        const int row = m0 + wm + i * 8 + g;
        const int col = n0 + wn + j * 8 + 2 * t + c;
        if (row < M && col < N) {
          cplx acc = cmake(cre[i][j][c], cim[i][j][c]);
          cplx r = cmul(alpha, acc);
          if (has_beta) r = cfma(beta, Cz[(long)row + (long)col * ldc], r);
          Cz[(long)row + (long)col * ldc] = r;
        }
      }
}

// coef'(:, v) = blockdiag(D, alpha I) * sum_z partial_z(:, v)
__global__ void k_coef_finish(int m, int nkb, int nvec, int nsplit, const cplx *__restrict__ part, long split_stride,
                              const double *__restrict__ dion, double alpha_pv, cplx *__restrict__ coef,
                              const int *__restrict__ active) {
  const int v = blockIdx.x;
  if (active && !active[v]) return;
  extern __shared__ cplx c[];
  for (int i = threadIdx.x; i < m; i += blockDim.x) {
    cplx s = cmake(0.0, 0.0);
    for (int z = 0; z < nsplit; ++z) s = cadd(s, part[(long)z * split_stride + (long)v * m + i]);
    c[i] = s;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < m; i += blockDim.x) {
    cplx r;
    if (i < nkb) {
      r = cmake(0.0, 0.0);
      for (int j = 0; j < nkb; ++j) {
        const double d = dion[i + (long)nkb * j];
        r.x += d * c[j].x;
        r.y += d * c[j].y;
      }
    } else {
      r = cscale(alpha_pv, c[i]);
    }
    coef[(long)v * m + i] = r;
  }
}

__global__ void k_add_sigma(int n, int nvec, const cplx *__restrict__ psi, long ldpsi, const cplx *__restrict__ sigma,
                            long sigma_stride, cplx *__restrict__ out, long ldout, const int *__restrict__ active) {
  const int v = blockIdx.y;
  if (active && !active[v]) return;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const cplx sg = sigma ? sigma[(long)v * sigma_stride] : cmake(0.0, 0.0);
  out[(long)v * ldout + i] = cfma(sg, psi[(long)v * ldpsi + i], out[(long)v * ldout + i]);
}

int gemm_ch_n(sgw_ctx *ctx, int m, int n, int k, const cplx *A, long lda, const cplx *B, long ldb, cplx *C, long ldc) {
  if (m <= 0 || n <= 0) return SGW_OK;
  dim3 grid((m + BM - 1) / BM, (n + BN - 1) / BN, 1);
  k_zgemm<true, true><<<grid, GT, 0, ctx->stream>>>(m, n, k, A, lda, B, ldb, C, ldc, cmake(1, 0), cmake(0, 0), k > 0 ? k : 1, 0, nullptr);
  SGW_LAUNCH_CHECK();
  return SGW_OK;
}

int gemm_n_n(sgw_ctx *ctx, int m, int n, int k, cplx alpha, const cplx *A, long lda, const cplx *B, long ldb, cplx beta,
             cplx *C, long ldc) {
  if (m <= 0 || n <= 0) return SGW_OK;
  dim3 grid((m + BM - 1) / BM, (n + BN - 1) / BN, 1);
  k_zgemm<false, false><<<grid, GT, 0, ctx->stream>>>(m, n, k, A, lda, B, ldb, C, ldc, alpha, beta, k > 0 ? k : 1, 0, nullptr);
  SGW_LAUNCH_CHECK();
  return SGW_OK;
}

int nonlocal_apply(sgw_ctx *ctx, const KSlot &ks, double alpha_pv, int nvec, const cplx *psi, long ldpsi, cplx *out,
                   long ldout, const int *active) {
  if (nvec <= 0) return SGW_OK;
  const bool use_pv = fabs(alpha_pv) > 1e-14;   // linear_op.f90:131 (eps14)
  const int m = ks.nkb + (use_pv ? ks.nbnd : 0);
  if (m == 0) {  // nothing to add: out must still be defined (zero) on all npwx rows of every vector
    SGW_CUDA(cudaMemset2DAsync(out, (size_t)ldout * sizeof(cplx), 0, (size_t)ks.npwx * sizeof(cplx), nvec, ctx->stream));
    return SGW_OK;
  }
  // split-K so that the projection fills the machine
  const int tiles = ((m + BM - 1) / BM) * ((nvec + BN - 1) / BN);
  int nsplit = (2 * ctx->sm_count + tiles - 1) / tiles;
  const int kblocks = (ks.npw + BK - 1) / BK;
  if (nsplit > kblocks) nsplit = kblocks;
  if (nsplit < 1) nsplit = 1;
  int kchunk = ((kblocks + nsplit - 1) / nsplit) * BK;
  nsplit = (ks.npw + kchunk - 1) / kchunk;
  cplx *part = nullptr, *coef = nullptr;
  const long split_stride = (long)m * nvec;
  SGW_CHECK(ws(ctx, "nl_part", (size_t)split_stride * nsplit, &part));
  SGW_CHECK(ws(ctx, "nl_coef", (size_t)split_stride, &coef));
  {
    dim3 grid((m + BM - 1) / BM, (nvec + BN - 1) / BN, nsplit);
    ProfScope prof(ctx, PC_GEMM_PROJ);
    k_zgemm<true, true><<<grid, GT, 0, ctx->stream>>>(m, nvec, ks.npw, ks.d_P, ks.npwx, psi, ldpsi, part, m, cmake(1, 0),
                                                     cmake(0, 0), kchunk, split_stride, active);
    SGW_LAUNCH_CHECK();
  }
  {
    ProfScope prof(ctx, PC_OTHER);
    k_coef_finish<<<nvec, 128, (size_t)m * sizeof(cplx), ctx->stream>>>(m, ks.nkb, nvec, nsplit, part, split_stride, ks.d_dion,
                                                                      alpha_pv, coef, active);
    SGW_LAUNCH_CHECK();
  }
  {
    dim3 grid((ks.npwx + BM - 1) / BM, (nvec + BN - 1) / BN, 1);
    ProfScope prof(ctx, PC_GEMM_OUT);
    k_zgemm<false, false><<<grid, GT, 0, ctx->stream>>>(ks.npwx, nvec, m, ks.d_P, ks.npwx, coef, m, out, ldout, cmake(1, 0),
                                                       cmake(0, 0), m, 0, active);
    SGW_LAUNCH_CHECK();
  }
  return SGW_OK;
}

int dense_apply(sgw_ctx *ctx, const KSlot &ks, int nvec, const cplx *psi, long ldpsi, const cplx *sigma, long sigma_stride,
                cplx *out, long ldout, const int *active) {
  if (nvec <= 0) return SGW_OK;
  const int n = ks.npw;
  dim3 grid((n + BM - 1) / BM, (nvec + BN - 1) / BN, 1);
  k_zgemm<false, false><<<grid, GT, 0, ctx->stream>>>(n, nvec, n, ks.d_A, n, psi, ldpsi, out, ldout, cmake(1, 0), cmake(0, 0), n, 0,
                                                     active);
  SGW_LAUNCH_CHECK();
  dim3 g2((n + 255) / 256, nvec);
  k_add_sigma<<<g2, 256, 0, ctx->stream>>>(n, nvec, psi, ldpsi, sigma, sigma_stride, out, ldout, active);
  SGW_LAUNCH_CHECK();
  return SGW_OK;
}

}  // namespace sgw
