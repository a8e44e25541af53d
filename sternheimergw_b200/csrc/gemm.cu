// gemm.cu -- complex-FP64 GEMMs on the FP64 tensor path (DMMA, mma.sync.m8n8k4.f64) for
//   * the non-local pseudopotential  h psi += vkb (D (vkb^H psi))        ([QE] calbec + add_vuspsi)
//   * the valence projector          A psi += alpha_pv evq (evq^H psi)   (linear_op.f90:147-206, ZGEMM x2)
//   * orthogonalize ([QE]) and the dense test backend (linear_solver.pf:106)
// Both rank-k updates of H.psi share one pass: P = [vkb | evq] is stored as one npwx x (nkb+nbnd) panel,
// coef = P^H psi (split-K, deterministic two-pass reduction), coef' = blockdiag(D, alpha_pv I) coef,
// out = P coef'.  tcgen05 has no f64 kind, so the legacy mma.sync DMMA path is the FP64 tensor route on sm_100a.
#include "internal.cuh"

#include <stdlib.h>

namespace sgw {

constexpr int BM = 64, BN = 32, BK = 16, GT = 256;
// Shared-memory tiles hold interleaved complex (16 B) elements so that one LDS.128 fetches (re, im) of a fragment
// element and cp.async (LDGSTS) can stage them straight from global memory.  Pitches are chosen so that the 8 lanes
// of a quarter-warp (g = 0..1, t = 0..3) of a 16-byte access hit 8 distinct 16-byte bank groups:
constexpr int PK = BK + 4;   // k-contiguous tiles: (g*PK + t) mod 8 distinct  <=  PK = 4 mod 8
constexpr int PM = BM + 2;   // m-contiguous A tile: (t*PM + g) mod 8 distinct  <=  PM = 2 mod 8
constexpr int NSTAGE = 2;

template <bool A_KCONTIG>
constexpr int a_tile_elems() { return A_KCONTIG ? BM * PK : BK * PM; }
template <bool A_KCONTIG>
constexpr size_t zgemm_smem() { return (size_t)NSTAGE * (a_tile_elems<A_KCONTIG>() + BN * PK) * sizeof(cplx); }

// C(M x N) = alpha * op(A) * B + beta * C ; op(A) = A^H (A_KCONTIG: A is K x M, column-major) or A (M x K).
// B is K x N column-major.  blockIdx.z selects a K range (split-K): partial results go to C + z * split_stride.
// Global -> shared staging is double buffered with cp.async so that the DMMA pipe does not wait on HBM/L2.
// BATCH: blockIdx.z selects one of several independent products instead (A += z * bsA, B += z * bsB, C += z * split_stride,
// the whole K range per CTA); used by the blocked inversion of invert.cu.
template <bool A_KCONTIG, bool CONJA, int MINB, bool BATCH = false>
__global__ void __launch_bounds__(GT, MINB) k_zgemm(int M, int N, int K, const cplx *__restrict__ A, long lda,
                                               const cplx *__restrict__ B, long ldb, cplx *__restrict__ C, long ldc,
                                               cplx alpha, cplx beta, int kchunk, long split_stride,
                                               const int *__restrict__ list, const int *__restrict__ count, int list_mode,
                                               long bsA = 0, long bsB = 0) {
  // Compacted batches: with `list` the problem has *count columns; list_mode bit 0: column j of B is B's column list[j]
  // (gather), list_mode bit 1: column j of C is C's column list[j] (scatter); both bits may be set.  Tiles beyond the
  // count exit at once, so a batch in which most right-hand sides have converged costs what its active columns cost.
  // list_mode bit 2: column tiles run fastest (blockIdx.x = n tile, blockIdx.y = m tile): the CTAs that are resident together
  // then share ONE tile of A, which matters when A (the npwx x m projector panel, 148 MB at Si64) does not fit in L2
  const bool nfast = (list_mode & 4) != 0;
  const int m0 = (nfast ? blockIdx.y : blockIdx.x) * BM, n0 = (nfast ? blockIdx.x : blockIdx.y) * BN;
  if (list) {
    N = min(N, *count);
    if (n0 >= N) return;
  }
  const int kbeg = BATCH ? 0 : blockIdx.z * kchunk, kend = BATCH ? K : min(K, kbeg + kchunk);
  if (BATCH) { A += (long)blockIdx.z * bsA; B += (long)blockIdx.z * bsB; }
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  constexpr int ASZ = a_tile_elems<A_KCONTIG>();
  extern __shared__ cplx gsm[];
  cplx *As = gsm;                      // [NSTAGE][ASZ]
  cplx *Bs = gsm + NSTAGE * ASZ;       // [NSTAGE][BN * PK]
  const int wm = (warp & 3) * 16, wn = (warp >> 2) * 16;
  const int g = lane >> 2, t = lane & 3;
  // 3M complex product (Karatsuba): with a = op(A) element, p1 += ar*br, p2 += ai*bi, p3 += (ar+ai)*(br+bi);
  // re = p1 - p2, im = p3 - p1 - p2.  Three DMMAs per complex tile product instead of four: the kernel is bound by
  // the DMMA pipe (ncu: math_pipe_throttle), so this is a 25 % cut of its critical resource.  The imaginary part
  // carries one extra rounding of size eps*(|p1|+|p2|), i.e. the same norm-wise bound as a 4M product.
  double p1[2][2][2], p2[2][2][2], p3[2][2][2];
#pragma unroll
  for (int i = 0; i < 2; ++i)
#pragma unroll
    for (int j = 0; j < 2; ++j)
#pragma unroll
      for (int c = 0; c < 2; ++c) p1[i][j][c] = p2[i][j][c] = p3[i][j][c] = 0.0;

  auto stage_load = [&](int st, int k0) {
    cplx *as = As + st * ASZ, *bs = Bs + st * (BN * PK);
#pragma unroll
    for (int r = 0; r < BM * BK / GT; ++r) {
      const int i = tid + r * GT;
      if (A_KCONTIG) {
        const int k = i % BK, m = i / BK;
        const bool ok = (k0 + k < kend) && (m0 + m < M);
        cp_async16(as + m * PK + k, ok ? A + (long)(k0 + k) + (long)(m0 + m) * lda : A, ok ? 16 : 0);
      } else {
        const int m = i % BM, k = i / BM;
        const bool ok = (k0 + k < kend) && (m0 + m < M);
        cp_async16(as + k * PM + m, ok ? A + (long)(m0 + m) + (long)(k0 + k) * lda : A, ok ? 16 : 0);
      }
    }
#pragma unroll
    for (int r = 0; r < BN * BK / GT; ++r) {
      const int i = tid + r * GT;
      const int k = i % BK, n = i / BK;
      const bool ok = (k0 + k < kend) && (n0 + n < N);
      const long bcol = ok ? ((list_mode & 1) ? list[n0 + n] : n0 + n) : 0;
      cp_async16(bs + n * PK + k, ok ? B + (long)(k0 + k) + bcol * ldb : B, ok ? 16 : 0);
    }
  };

  if (kbeg < kend) stage_load(0, kbeg);
  cp_async_commit();
  int it = 0;
  for (int k0 = kbeg; k0 < kend; k0 += BK, ++it) {
    if (k0 + BK < kend) stage_load((it + 1) & 1, k0 + BK);
    cp_async_commit();
    cp_async_wait<1>();
    __syncthreads();
    const cplx *as = As + (it & 1) * ASZ, *bs = Bs + (it & 1) * (BN * PK);
#pragma unroll
    for (int kk = 0; kk < BK; kk += 4) {
      cplx a[2], b[2];
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        const int m = wm + i * 8 + g;
        a[i] = as[A_KCONTIG ? m * PK + kk + t : (kk + t) * PM + m];
      }
#pragma unroll
      for (int j = 0; j < 2; ++j) b[j] = bs[(wn + j * 8 + g) * PK + kk + t];
      // op(A) = conj(A)^T uses ai -> -ai
      double ai[2], asum[2], bsum[2];
#pragma unroll
      for (int i = 0; i < 2; ++i) { ai[i] = CONJA ? -a[i].y : a[i].y; asum[i] = a[i].x + ai[i]; }
#pragma unroll
      for (int j = 0; j < 2; ++j) bsum[j] = b[j].x + b[j].y;
#pragma unroll
      for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int j = 0; j < 2; ++j) {
          dmma(p1[i][j][0], p1[i][j][1], a[i].x, b[j].x);
          dmma(p2[i][j][0], p2[i][j][1], ai[i], b[j].y);
          dmma(p3[i][j][0], p3[i][j][1], asum[i], bsum[j]);
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
          cplx acc = cmake(p1[i][j][c] - p2[i][j][c], (p3[i][j][c] - p1[i][j][c]) - p2[i][j][c]);
          cplx r = cmul(alpha, acc);
          const long ccol = (list_mode & 2) ? list[col] : col;
          if (has_beta) r = cfma(beta, Cz[(long)row + ccol * ldc], r);
          Cz[(long)row + ccol * ldc] = r;
        }
      }
}

// opt in to > 48 KB of dynamic shared memory (per device) and record how many CTAs fit on an SM
static int zgemm_init(sgw_ctx *ctx);

template <bool A_KCONTIG, bool CONJA>
static int launch_zgemm(sgw_ctx *ctx, dim3 grid, int M, int N, int K, const cplx *A, long lda, const cplx *B, long ldb, cplx *C,
                        long ldc, cplx alpha, cplx beta, int kchunk, long split_stride, const int *list = nullptr,
                        const int *count = nullptr, int list_mode = 0) {
  constexpr size_t smem = zgemm_smem<A_KCONTIG>();
  SGW_CHECK(zgemm_init(ctx));
  if (!list) list_mode &= 4;
  static int minb = -1;                                         // SGW_GEMM_MINB: resident CTAs per SM the register allocation must allow (2 | 3)
  if (minb < 0) { const char *e = getenv("SGW_GEMM_MINB"); minb = e ? atoi(e) : 3; }
  if (minb == 2)
    k_zgemm<A_KCONTIG, CONJA, 2><<<grid, GT, smem, ctx->stream>>>(M, N, K, A, lda, B, ldb, C, ldc, alpha, beta, kchunk, split_stride, list,
                                                                  count, list_mode);
  else
    k_zgemm<A_KCONTIG, CONJA, 3><<<grid, GT, smem, ctx->stream>>>(M, N, K, A, lda, B, ldb, C, ldc, alpha, beta, kchunk, split_stride, list,
                                                                  count, list_mode);
  SGW_LAUNCH_CHECK();
  return SGW_OK;
}

static int zgemm_init(sgw_ctx *ctx) {
  if (ctx->gemm_cta_per_sm > 0) return SGW_OK;
  SGW_CUDA(cudaFuncSetAttribute(k_zgemm<true, true, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)zgemm_smem<true>()));
  SGW_CUDA(cudaFuncSetAttribute(k_zgemm<false, false, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)zgemm_smem<false>()));
  SGW_CUDA(cudaFuncSetAttribute(k_zgemm<true, true, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)zgemm_smem<true>()));
  SGW_CUDA(cudaFuncSetAttribute(k_zgemm<false, false, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)zgemm_smem<false>()));
  SGW_CUDA(cudaFuncSetAttribute(k_zgemm<false, false, 3, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)zgemm_smem<false>()));
  SGW_CUDA(cudaFuncSetAttribute(k_zgemm<true, true, 3, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)zgemm_smem<true>()));
  int n = 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, k_zgemm<true, true, 3>, GT, zgemm_smem<true>()) != cudaSuccess || n < 1) n = 1;
  ctx->gemm_cta_per_sm = n;
  return SGW_OK;
}

// nbatch independent products C_z = alpha A_z B_z + beta C_z (A_z is M x K, column-major) over the columns list[0 .. *count)
// of B_z and C_z (nullptr list: all N columns)
int gemm_n_n_batched(sgw_ctx *ctx, int m, int n, int k, cplx alpha, const cplx *A, long lda, long bsA, const cplx *B, long ldb,
                     long bsB, cplx beta, cplx *C, long ldc, long bsC, int nbatch, const int *list, const int *count) {
  if (m <= 0 || n <= 0 || nbatch <= 0) return SGW_OK;
  SGW_ARG(nbatch <= 65535, "gemm_n_n_batched: more than 65535 matrices");
  constexpr size_t smem = zgemm_smem<false>();
  SGW_CHECK(zgemm_init(ctx));
  dim3 grid((m + BM - 1) / BM, (n + BN - 1) / BN, nbatch);
  k_zgemm<false, false, 3, true><<<grid, GT, smem, ctx->stream>>>(m, n, k, A, lda, B, ldb, C, ldc, alpha, beta, k, bsC, list, count,
                                                                    list ? 3 : 0, bsA, bsB);
  SGW_LAUNCH_CHECK();
  return SGW_OK;
}

// nbatch independent products C_z = A_z^H B_z (A_z is K x M, B_z is K x N, column-major)
int gemm_ch_n_batched(sgw_ctx *ctx, int m, int n, int k, const cplx *A, long lda, long bsA, const cplx *B, long ldb, long bsB, cplx *C,
                      long ldc, long bsC, int nbatch) {
  if (m <= 0 || n <= 0 || nbatch <= 0) return SGW_OK;
  SGW_ARG(nbatch <= 65535, "gemm_ch_n_batched: more than 65535 matrices");
  constexpr size_t smem = zgemm_smem<true>();
  SGW_CHECK(zgemm_init(ctx));
  dim3 grid((m + BM - 1) / BM, (n + BN - 1) / BN, nbatch);
  k_zgemm<true, true, 3, true><<<grid, GT, smem, ctx->stream>>>(m, n, k, A, lda, B, ldb, C, ldc, cmake(1, 0), cmake(0, 0), k, bsC, nullptr,
                                                                  nullptr, 0, bsA, bsB);
  SGW_LAUNCH_CHECK();
  return SGW_OK;
}

// coef'(:, v) = blockdiag(D, alpha I) * sum_z partial_z(:, v)
__global__ void k_coef_finish(int m, int nkb, int nvec, int nsplit, const cplx *__restrict__ part, long split_stride,
                              const double *__restrict__ dion, const int *__restrict__ dptr, const int *__restrict__ dcol,
                              const double *__restrict__ dval, double alpha_pv, cplx *__restrict__ coef,
                              const int *__restrict__ count) {
  const int v = blockIdx.x;
  if (count && v >= *count) return;
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
      if (dptr) {                      // compressed row: same terms in the same (column) order, exact zeros skipped
        for (int e = dptr[i]; e < dptr[i + 1]; ++e) {
          const double d = dval[e];
          const cplx cj = c[dcol[e]];
          r.x += d * cj.x;
          r.y += d * cj.y;
        }
      } else {
        for (int j = 0; j < nkb; ++j) {
          const double d = dion[i + (long)nkb * j];
          r.x += d * c[j].x;
          r.y += d * c[j].y;
        }
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

// list[j] = index of the j-th active vector (ascending), *count = how many: one CTA, ballot + prefix over warps
__global__ void __launch_bounds__(1024) k_compact_active(int nvec, const int *__restrict__ active, int *__restrict__ list,
                                                         int *__restrict__ count) {
  __shared__ int wsum[32];
  __shared__ int base;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  if (threadIdx.x == 0) base = 0;
  __syncthreads();
  for (int v0 = 0; v0 < nvec; v0 += 1024) {
    const int v = v0 + threadIdx.x;
    const bool on = v < nvec && active[v] != 0;
    const unsigned bal = __ballot_sync(0xffffffffu, on);
    if (lane == 0) wsum[w] = __popc(bal);
    __syncthreads();
    int off = base;
    for (int i = 0; i < w; ++i) off += wsum[i];
    if (on) list[off + __popc(bal & ((1u << lane) - 1u))] = v;
    __syncthreads();
    if (threadIdx.x == 0) { int t = 0; for (int i = 0; i < 32; ++i) t += wsum[i]; base += t; }
    __syncthreads();
  }
  if (threadIdx.x == 0) *count = base;
}

int gemm_ch_n(sgw_ctx *ctx, int m, int n, int k, const cplx *A, long lda, const cplx *B, long ldb, cplx *C, long ldc) {
  if (m <= 0 || n <= 0) return SGW_OK;
  dim3 grid((m + BM - 1) / BM, (n + BN - 1) / BN, 1);
  return launch_zgemm<true, true>(ctx, grid, m, n, k, A, lda, B, ldb, C, ldc, cmake(1, 0), cmake(0, 0), k > 0 ? k : 1, 0);
}

int gemm_n_n(sgw_ctx *ctx, int m, int n, int k, cplx alpha, const cplx *A, long lda, const cplx *B, long ldb, cplx beta,
             cplx *C, long ldc) {
  if (m <= 0 || n <= 0) return SGW_OK;
  dim3 grid((m + BM - 1) / BM, (n + BN - 1) / BN, 1);
  return launch_zgemm<false, false>(ctx, grid, m, n, k, A, lda, B, ldb, C, ldc, alpha, beta, k > 0 ? k : 1, 0);
}

// compact list of the active vectors of a batch (nullptr list = all active)
static int active_list(sgw_ctx *ctx, int nvec, const int *active, int **list, int **count) {
  *list = *count = nullptr;
  if (!active) return SGW_OK;
  SGW_CHECK(ws(ctx, "nl_list", (size_t)nvec, list));
  SGW_CHECK(ws(ctx, "nl_count", (size_t)1, count));
  k_compact_active<<<1, 1024, 0, ctx->stream>>>(nvec, active, *list, *count);
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
  // split-K so that the projection fills the machine in whole waves: pick the split whose CTA count wastes the
  // least of the last wave (slots = SMs x resident CTAs per SM), preferring fewer splits on ties
  const int tiles = ((m + BM - 1) / BM) * ((nvec + BN - 1) / BN);
  const int kblocks = (ks.npw + BK - 1) / BK;
  SGW_CHECK(zgemm_init(ctx));
  const long slots = (long)ctx->sm_count * ctx->gemm_cta_per_sm;
  int nsplit = 1, kchunk = kblocks * BK;
  double best = -1.0;
  for (int cand = 1; cand <= 64 && cand <= kblocks; ++cand) {
    const int kc = ((kblocks + cand - 1) / cand) * BK;
    const int ns = (ks.npw + kc - 1) / kc;
    const long ctas = (long)tiles * ns;
    const long waves = (ctas + slots - 1) / slots;
    // time ~ waves x (work of one CTA) = waves x kc ; lower is better
    const double cost = (double)waves * kc;
    if (best < 0 || cost < best * 0.98) { best = cost; nsplit = ns; kchunk = kc; }
  }
  cplx *part = nullptr, *coef = nullptr;
  int *list = nullptr, *count = nullptr;
  {
    ProfScope prof(ctx, PC_OTHER);
    SGW_CHECK(active_list(ctx, nvec, active, &list, &count));
  }
  const long split_stride = (long)m * nvec;
  SGW_CHECK(ws(ctx, "nl_part", (size_t)split_stride * nsplit, &part));
  SGW_CHECK(ws(ctx, "nl_coef", (size_t)split_stride, &coef));
  {
    dim3 grid((m + BM - 1) / BM, (nvec + BN - 1) / BN, nsplit);
    ProfScope prof(ctx, PC_GEMM_PROJ);
    SGW_CHECK((launch_zgemm<true, true>(ctx, grid, m, nvec, ks.npw, ks.d_P, ks.npwx, psi, ldpsi, part, m, cmake(1, 0),
                                         cmake(0, 0), kchunk, split_stride, list, count, 1)));
  }
  {
    ProfScope prof(ctx, PC_OTHER);
    k_coef_finish<<<nvec, 128, (size_t)m * sizeof(cplx), ctx->stream>>>(m, ks.nkb, nvec, nsplit, part, split_stride, ks.d_dion, ks.d_dion_ptr, ks.d_dion_col, ks.d_dion_val,
                                                                      alpha_pv, coef, count);
    SGW_LAUNCH_CHECK();
  }
  {
    static int nfast = -1;                                      // SGW_GEMM_NFAST=0: row tiles fastest (A/B testing)
    if (nfast < 0) { const char *e = getenv("SGW_GEMM_NFAST"); nfast = e ? atoi(e) : 1; }
    const int mt = (ks.npwx + BM - 1) / BM, nt = (nvec + BN - 1) / BN;
    const bool nf = nfast && mt <= 65535;
    dim3 grid(nf ? nt : mt, nf ? mt : nt, 1);
    ProfScope prof(ctx, PC_GEMM_OUT);
    SGW_CHECK((launch_zgemm<false, false>(ctx, grid, ks.npwx, nvec, m, ks.d_P, ks.npwx, coef, m, out, ldout, cmake(1, 0),
                                           cmake(0, 0), m, 0, list, count, 2 | (nf ? 4 : 0))));
  }
  return SGW_OK;
}

int dense_apply(sgw_ctx *ctx, const KSlot &ks, int nvec, const cplx *psi, long ldpsi, const cplx *sigma, long sigma_stride,
                cplx *out, long ldout, const int *active) {
  if (nvec <= 0) return SGW_OK;
  const int n = ks.npw;
  dim3 grid((n + BM - 1) / BM, (nvec + BN - 1) / BN, 1);
  int *list = nullptr, *count = nullptr;
  SGW_CHECK(active_list(ctx, nvec, active, &list, &count));
  SGW_CHECK((launch_zgemm<false, false>(ctx, grid, n, nvec, n, ks.d_A, n, psi, ldpsi, out, ldout, cmake(1, 0), cmake(0, 0), n, 0,
                                         list, count, 3)));
  dim3 g2((n + 255) / 256, nvec);
  k_add_sigma<<<g2, 256, 0, ctx->stream>>>(n, nvec, psi, ldpsi, sigma, sigma_stride, out, ldout, active);
  SGW_LAUNCH_CHECK();
  return SGW_OK;
}

}  // namespace sgw
