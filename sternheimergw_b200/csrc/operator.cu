// operator.cu -- (H + sigma S + alpha_pv P_v) psi on batches of device-resident vectors
// (algo/linear_solver/src/linear_op.f90:46-144 with [QE] h_psi/s_psi for NC-PP: S = 1), plus the
// caller-order <-> column-order permutations applied at the ABI boundary.
#include "internal.cuh"

namespace sgw {

__global__ void k_permute_in(int npw, int npwx, const int *__restrict__ perm, const cplx *__restrict__ src, long lds,
                             cplx *__restrict__ dst, long ldd) {
  const int v = blockIdx.y;
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= npwx) return;
  dst[(long)v * ldd + p] = p < npw ? src[(long)v * lds + perm[p]] : cmake(0.0, 0.0);
}

__global__ void k_permute_out(int npw, const int *__restrict__ perm, const cplx *__restrict__ src, long lds,
                              cplx *__restrict__ dst, long ldd) {
  const int v = blockIdx.y;
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= npw) return;
  dst[(long)v * ldd + perm[p]] = src[(long)v * lds + p];
}

// dst(:, v) (column order, npwx rows zero padded) = src(perm, v)
int permute_in(sgw_ctx *ctx, const Sphere &s, int nvec, const cplx *src, long lds, cplx *dst, long ldd, int npwx) {
  if (nvec <= 0) return SGW_OK;
  dim3 grid((npwx + 255) / 256, nvec);
  k_permute_in<<<grid, 256, 0, ctx->stream>>>(s.npw, npwx, s.d_perm, src, lds, dst, ldd);
  SGW_LAUNCH_CHECK();
  return SGW_OK;
}

// dst(perm, v) = src(:, v) for the first npw entries (caller order); rows >= npw of dst are left untouched
int permute_out(sgw_ctx *ctx, const Sphere &s, int nvec, const cplx *src, long lds, cplx *dst, long ldd) {
  if (nvec <= 0) return SGW_OK;
  dim3 grid((s.npw + 255) / 256, nvec);
  k_permute_out<<<grid, 256, 0, ctx->stream>>>(s.npw, s.d_perm, src, lds, dst, ldd);
  SGW_LAUNCH_CHECK();
  return SGW_OK;
}

int apply_operator(sgw_ctx *ctx, int slot, double alpha_pv, int nvec, const cplx *psi, long ldpsi, const cplx *sigma,
                   long sigma_stride, cplx *apsi, long ldapsi, const int *active) {
  if (nvec <= 0) return SGW_OK;
  if (slot < 0 || slot >= (int)ctx->slots.size() || !ctx->slots[slot].set) {
    ctx->err = "operator slot not set";
    return SGW_E_STATE;
  }
  const KSlot &ks = ctx->slots[slot];
  if (ks.dense) return dense_apply(ctx, ks, nvec, psi, ldpsi, sigma, sigma_stride, apsi, ldapsi, active);
  if (!ctx->vloc_set) {
    ctx->err = "local potential not set (sgw_set_vloc)";
    return SGW_E_STATE;
  }
  if (ldpsi != ldapsi) {
    ctx->err = "apply_operator: psi and A_psi must share the leading dimension";
    return SGW_E_ARG;
  }
  // non-local + valence projector part first (writes all npwx rows of apsi) ...
  SGW_CHECK(nonlocal_apply(ctx, ks, alpha_pv, nvec, psi, ldpsi, apsi, ldapsi, active));
  // ... then the local part; its epilogue adds kinetic + sigma*psi and keeps what is already in apsi
  cplx *T1 = nullptr, *T2 = nullptr;
  const size_t tsz = (size_t)nvec * ctx->nr3 * ks.sph.ncol;
  SGW_CHECK(ws(ctx, "fft_T1", tsz, &T1));
  SGW_CHECK(ws(ctx, "fft_T2", tsz, &T2));
  SGW_CHECK(fft_zpass_g2r(ctx, ks.sph, nvec, psi, ldpsi, T1, active));
  SGW_CHECK(fft_plane(ctx, PLANE_VLOC, &ks.sph, &ks.sph, nvec, T1, T2, nullptr, 1, nullptr, active));
  ZEpilogue epi;
  epi.mode = 1;
  epi.g2kin = ks.d_g2kin;
  epi.psi = psi;
  epi.sigma = sigma;
  epi.sigma_stride = sigma_stride;
  epi.keep_out = 1;
  SGW_CHECK(fft_zpass_r2g(ctx, ks.sph, nvec, T2, apsi, ldapsi, epi, active));
  return SGW_OK;
}

}  // namespace sgw
