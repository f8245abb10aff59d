// Entry points of the kernel translation units (inst.cu, compiled once per (kind, dtype) so that the
// library builds in parallel).  Host code (b200mf.cu) sees only these and the argument structs.
#pragma once
#include "common.h"
#include "ops.cuh"

namespace b200mf {

void shape_tables(const double **hd, const float **hf);  // [MAXN+1][3][MAXN*MAXN], b200mf.cu

// cell-group range [g0, g1) on `stream`
template <typename T>
int run_nh(const b200mf_ctx *ctx, int variant, const T *src, T *dst, const NHArgs<T> &oa, int64_t g0, int64_t g1, cudaStream_t stream,
           double *dot = nullptr);  // dot: see CellArgs::dot
template <typename T>
int run_nh_diag(const b200mf_ctx *ctx, int variant, T *diag, const NHArgs<T> &oa);
template <typename T>
int run_nh_cache(const b200mf_ctx *ctx, int variant, const T *displacement, const NHCacheArgs<T> &oa);
template <typename T>
int run_nh_residual(const b200mf_ctx *ctx, const T *displacement, T *rhs, const NHResidualArgs<T> &oa);
template <typename T>
int run_laplace(const b200mf_ctx *ctx, bool diag, const T *src, T *dst, const LaplaceArgs<T> &oa, int64_t g0, int64_t g1,
                cudaStream_t stream, double *dot = nullptr);

template <typename T>
int run_laplace_rhs(const b200mf_ctx *ctx, const T *solution_old, T *rhs, const LaplaceRhsArgs<T> &oa);

}  // namespace b200mf
