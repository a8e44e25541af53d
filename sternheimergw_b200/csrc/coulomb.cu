// coulomb.cu -- stubs (first GPU bring-up); replaced by the batched Coulomb pipeline
#include "internal.cuh"
extern "C" {
#define STUB { if (ctx) ctx->err = "not implemented yet"; return SGW_E_UNSUPPORTED; }
int sgw_set_system(sgw_ctx *ctx, double, double, int, const double *, const int32_t *) STUB
int sgw_set_q(sgw_ctx *ctx, const double *) STUB
int sgw_set_nksq(sgw_ctx *ctx, int) STUB
int sgw_set_kpair(sgw_ctx *ctx, int, int, int, const int32_t *, int, const sgw_cplx *, const double *, double) STUB
int sgw_solve_linter(sgw_ctx *ctx, const sgw_solver_cfg *, int, const sgw_cplx *, int, const sgw_cplx *, sgw_cplx *, int32_t *) STUB
int sgw_coulomb(sgw_ctx *ctx, const sgw_solver_cfg *, int, int, int, const int32_t *, int, const sgw_cplx *, sgw_cplx *, int32_t *) STUB
int sgw_coulomb_q0G0(sgw_ctx *ctx, const sgw_solver_cfg *, int, const sgw_cplx *, sgw_cplx *, int32_t *) STUB
int sgw_unfold_w(sgw_ctx *ctx, int, int, int, const int32_t *, const sgw_cplx *, sgw_cplx *) STUB
int sgw_invert_epsilon(sgw_ctx *ctx, int, int, sgw_cplx *, int) STUB
int sgw_green_function(sgw_ctx *ctx, int, const sgw_solver_cfg *, int, const int32_t *, int, const int32_t *, int, const sgw_cplx *, sgw_cplx *, int32_t *) STUB
}
