// One kernel family per translation unit: compiled as  -DINST_KIND=<k> -DINST_T=<double|float>
//   1 NeoHook vmult   2 NeoHook diagonal   3 NeoHook cache   4 Laplace (vmult + diagonal)   5 NeoHook residual
#include "launch.cuh"

namespace b200mf {
using TT = INST_T;

#define B2_DISPATCH_NHV(v_, ...)                                                  \
  switch (v_) {                                                                    \
    case NHV_TENSOR2: { constexpr int V = NHV_TENSOR2; __VA_ARGS__; } break;        \
    case NHV_FORM0: { constexpr int V = NHV_FORM0; __VA_ARGS__; } break;            \
    case NHV_TENSOR4: { constexpr int V = NHV_TENSOR4; __VA_ARGS__; } break;        \
    case NHV_TENSOR4_NS: { constexpr int V = NHV_TENSOR4_NS; __VA_ARGS__; } break;  \
    default: { constexpr int V = NHV_SCALAR_REF; __VA_ARGS__; } break;              \
  }

#if INST_KIND == 1
template <>
int run_nh<TT>(const b200mf_ctx *ctx, int variant, const TT *src, TT *dst, const NHArgs<TT> &oa, int64_t g0, int64_t g1,
               cudaStream_t stream, double *dot) {
  int rc = 0;
  B2_DISPATCH_N(ctx->N, TT, { B2_DISPATCH_NHV(variant, { rc = launch_cell_loop<T, N, NHOp<T, N, V>>(ctx, src, dst, oa, g0, g1, stream, dot); }); });
  return rc;
}
#elif INST_KIND == 2
template <>
int run_nh_diag<TT>(const b200mf_ctx *ctx, int variant, TT *diag, const NHArgs<TT> &oa) {
  int rc = 0;
  B2_DISPATCH_N(ctx->N, TT, { B2_DISPATCH_NHV(variant, { rc = launch_cell_loop<T, N, NHOp<T, N, V>, true>(ctx, (const T *)nullptr, diag, oa); }); });
  return rc;
}
#elif INST_KIND == 3
template <>
int run_nh_cache<TT>(const b200mf_ctx *ctx, int variant, const TT *displacement, const NHCacheArgs<TT> &oa) {
  int rc = 0;
  B2_DISPATCH_N(ctx->N, TT, { B2_DISPATCH_NHV(variant, { rc = launch_cell_loop<T, N, NHCacheOp<T, N, V>>(ctx, displacement, (T *)nullptr, oa); }); });
  return rc;
}
#elif INST_KIND == 4
template <>
int run_laplace<TT>(const b200mf_ctx *ctx, bool diag, const TT *src, TT *dst, const LaplaceArgs<TT> &oa, int64_t g0, int64_t g1,
                    cudaStream_t stream, double *dot) {
  int rc = 0;
  const bool affine = ctx->mcell != nullptr;  // every cell affine: per-cell metric, nothing streamed per q-point
  B2_DISPATCH_N(ctx->N, TT, {
    if (diag)
      rc = affine ? launch_cell_loop<T, N, LaplaceAffineOp<T, N>, true>(ctx, (const T *)nullptr, dst, oa)
                  : launch_cell_loop<T, N, LaplaceOp<T, N>, true>(ctx, (const T *)nullptr, dst, oa);
    else if (affine) {
      // N <= 5: one warp per cell group with stacked columns (scalar_kernel.cuh); B200MF_SCALAR_STACK=0 keeps the generic kernel
      static const bool stack = !(getenv("B200MF_SCALAR_STACK") && atoi(getenv("B200MF_SCALAR_STACK")) == 0);
      // measured on B200 (profiles/r02_laplace_stack_ab.jsonl): 1.48x at Q4 FP64, +8 % at Q3 FP32, slower at Q2 / Q3 FP64
      if constexpr (N == 5 || (N == 4 && sizeof(T) == 4)) {
        if (stack)
          rc = launch_scalar_stack<T, N, LaplaceAffineOp<T, N>>(ctx, src, dst, oa, g0, g1, stream, dot);
        else
          rc = launch_cell_loop<T, N, LaplaceAffineOp<T, N>>(ctx, src, dst, oa, g0, g1, stream, dot);
      } else {
        rc = launch_cell_loop<T, N, LaplaceAffineOp<T, N>>(ctx, src, dst, oa, g0, g1, stream, dot);
      }
    } else
      rc = launch_cell_loop<T, N, LaplaceOp<T, N>>(ctx, src, dst, oa, g0, g1, stream, dot);
  });
  return rc;
}
template <>
int run_laplace_rhs<TT>(const b200mf_ctx *ctx, const TT *solution_old, TT *rhs, const LaplaceRhsArgs<TT> &oa) {
  int rc = 0;
  const bool affine = ctx->mcell != nullptr;
  B2_DISPATCH_N(ctx->N, TT, {
    rc = affine ? launch_cell_loop<T, N, LaplaceRhsOp<T, N, true>>(ctx, solution_old, rhs, oa)
                : launch_cell_loop<T, N, LaplaceRhsOp<T, N, false>>(ctx, solution_old, rhs, oa);
  });
  return rc;
}
#elif INST_KIND == 5
template <>
int run_nh_residual<TT>(const b200mf_ctx *ctx, const TT *displacement, TT *rhs, const NHResidualArgs<TT> &oa) {
  int rc = 0;
  B2_DISPATCH_N(ctx->N, TT, { rc = launch_cell_loop<T, N, NHResidualOp<T, N>>(ctx, displacement, rhs, oa); });
  return rc;
}
#else
#error "INST_KIND"
#endif

}  // namespace b200mf
