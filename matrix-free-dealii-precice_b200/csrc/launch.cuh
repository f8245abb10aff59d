// Launch of the templated cell loop (included by the kernel translation units only).
#pragma once
#include <algorithm>
#include <cstdlib>

#include "kernels.h"
#include "scalar_kernel.cuh"

#ifndef B200MF_MINB5
#define B200MF_MINB5 6
#endif
#ifndef B200MF_MINB6
#define B200MF_MINB6 2
#endif
#ifndef B200MF_MINB7
#define B200MF_MINB7 2
#endif

namespace b200mf {

// every translation unit owns a copy of the __constant__ 1-D matrices (cell_kernel.cuh)
static int ensure_shape_constants() {
  static bool done = false;
  if (done) return 0;
  const double *hd;
  const float *hf;
  shape_tables(&hd, &hf);
  B2_CUDA(cudaMemcpyToSymbol(c_shape_d, hd, sizeof(c_shape_d)));
  B2_CUDA(cudaMemcpyToSymbol(c_shape_f, hf, sizeof(c_shape_f)));
  done = true;
  return 0;
}

// ---------------------------------------------------------------------------
// launch of the templated cell loop

template <typename T, int N, class Op, bool DIAG, int GPB, int MINB, bool DOT = false>
static int launch_variant(const b200mf_ctx *ctx, const T *src, T *dst, const typename Op::Args &oa, int64_t g0, int64_t g1,
                          cudaStream_t stream, double *dot) {
  using Ly = Layout<N, Op::NC, sizeof(T) / 4>;
  if constexpr (!DOT && !DIAG && !Op::PLAIN) {
    if (dot != nullptr) return launch_variant<T, N, Op, DIAG, GPB, MINB, true>(ctx, src, dst, oa, g0, g1, stream, dot);
  }
  auto kern = cell_loop_kernel<T, N, Op, GPB, DIAG, MINB, DOT>;
  const size_t smem = (size_t)GPB * KernelSmem<T, N, Op, DIAG>::GROUP_BYTES;
  static int occ = 0;  // per instantiation
  B2_TRY(ensure_shape_constants());
  if (occ == 0) {
    B2_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    // occupancy the registers allow with the largest carve-out, then give shared memory exactly what
    // those blocks need and leave the rest to L1 (the DoF gather re-reads sectors: L1 hits matter)
    B2_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    int occ_max = 0;
    B2_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ_max, kern, Ly::TPG * GPB, smem));
    if (occ_max < 1) return set_error(B200MF_ERR_CUDA, "cell kernel does not fit on an SM (N=%d, %zu B shared)", N, smem);
    int pct = (int)((occ_max * (smem + 1024) * 100 + 228 * 1024 - 1) / (228 * 1024));
    pct = std::min(100, std::max(1, pct));
    B2_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, pct));
    B2_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, Ly::TPG * GPB, smem));
    if (occ < 1) occ = occ_max;
  }
  if (g1 <= g0) return 0;
  CellArgs<T> A;
  A.group_begin = g0;
  A.n_groups = g1;
  A.nidx = ctx->nidx;
  A.src = src;
  A.dst = dst;
  A.cstride = ctx->cstride;
  A.dot = dot;
  const int64_t want = (g1 - g0 + GPB - 1) / GPB;
  const int64_t cap = (int64_t)ctx->sm_count * occ;
  const unsigned grid = (unsigned)std::max<int64_t>(1, std::min(want, cap));
  B2_TRY(kernel_event(const_cast<b200mf_ctx *>(ctx), stream));
  kern<<<grid, Ly::TPG * GPB, smem, stream>>>(A, oa);
  B2_LAUNCHED();
  B2_TRY(kernel_event(const_cast<b200mf_ctx *>(ctx), stream));
  return 0;
}

template <typename T, int N, class Op, bool DIAG = false>
static int launch_cell_loop(const b200mf_ctx *ctx, const T *src, T *dst, const typename Op::Args &oa, int64_t g0 = 0, int64_t g1 = -1,
                            cudaStream_t stream = nullptr, double *dot = nullptr) {
  using Ly = Layout<N, Op::NC, sizeof(T) / 4>;
  if (g1 < 0) g1 = ctx->n_groups;
  if (!stream) stream = ctx->stream;
  if constexpr (!Ly::WARP_MODE) {
    // scalar Q4 (6 cells per 160-thread group): cap the registers at two resident groups per SM - only one warp of a
    // group works during the plane phase, a second group keeps the FP64 pipe fed (B200: 0.94 -> 0.69 ms at 1.7e7 DoFs)
    // vector operator, N >= 5 (sliced tables, several cells per block): cap the registers so that B200MF_MINB<N> blocks
    // are resident per SM (N = 5: 6 x 64 threads at 168 registers)
    constexpr bool VEC_VMULT = Op::NC == 3 && Op::INTEGRATE && !Op::PLAIN && !DIAG;
    constexpr int MINB = (Op::NC == 1 && N == 5 && !DIAG) ? 2
                         : (VEC_VMULT && N == 5)          ? B200MF_MINB5
                         : (VEC_VMULT && N == 6)          ? B200MF_MINB6
                         : (VEC_VMULT && N == 7)          ? B200MF_MINB7
                                                          : 1;
    return launch_variant<T, N, Op, DIAG, 1, MINB>(ctx, src, dst, oa, g0, g1, stream, dot);
  } else {
    // warps per block, from the B200 sweep tools/gpu_gpb_sweep.sh (profiles/r01_v3_gpb_sweep.jsonl): shared memory
    // (cell buffers + staged table) bounds the resident warps and small blocks pack an SM best; N = 2 has 8 cells
    // per warp and so little shared memory that the 32-blocks-per-SM limit binds first -> 4-warp blocks
    constexpr int GPB = (N == 2 ? 4 : N == 3 ? 2 : (N == 4 && sizeof(T) == 4) ? 2 : 1);
    constexpr bool VEC_VMULT = Op::NC == 3 && Op::INTEGRATE && !Op::PLAIN && !DIAG;
    constexpr int MINBW = (VEC_VMULT && N == 5 && GPB == 1) ? B200MF_MINB5 : 1;  // tuning builds with one N = 5 cell per warp
#ifdef B200MF_GPB_SWEEP  // tuning builds only: B200MF_GPB=1|2|4 picks the block size at run time
    static const int gpb_env = getenv("B200MF_GPB") ? atoi(getenv("B200MF_GPB")) : 0;
    if (gpb_env == 1) return launch_variant<T, N, Op, DIAG, 1, 1>(ctx, src, dst, oa, g0, g1, stream, dot);
    if (gpb_env == 2) return launch_variant<T, N, Op, DIAG, 2, 1>(ctx, src, dst, oa, g0, g1, stream, dot);
    if (gpb_env == 4) return launch_variant<T, N, Op, DIAG, 4, 1>(ctx, src, dst, oa, g0, g1, stream, dot);
#endif
    return launch_variant<T, N, Op, DIAG, GPB, MINBW>(ctx, src, dst, oa, g0, g1, stream, dot);
  }
}

// ---- scalar operator on affine cells, N <= 5: one warp per cell group, stacked columns (scalar_kernel.cuh) ----------
template <typename T, int N, class Op>
static int launch_scalar_stack(const b200mf_ctx *ctx, const T *src, T *dst, const typename Op::Args &oa, int64_t g0, int64_t g1,
                               cudaStream_t stream, double *dot) {
  using G = StackGeom<T, N>;
  constexpr int WPB = 2;
  if (g1 < 0) g1 = ctx->n_groups;
  if (!stream) stream = ctx->stream;
  B2_TRY(ensure_shape_constants());
  const size_t smem = WPB * G::WARP_BYTES;
  static int occ[2] = {0, 0};
  const int di = dot != nullptr;
  auto k0 = scalar_stack_kernel<T, N, Op, WPB, false>;
  auto k1 = scalar_stack_kernel<T, N, Op, WPB, true>;
  if (occ[di] == 0) {
    if (di) {
      B2_CUDA(cudaFuncSetAttribute(k1, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      B2_CUDA(cudaFuncSetAttribute(k1, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
      B2_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ[di], k1, 32 * WPB, smem));
    } else {
      B2_CUDA(cudaFuncSetAttribute(k0, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      B2_CUDA(cudaFuncSetAttribute(k0, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
      B2_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ[di], k0, 32 * WPB, smem));
    }
    if (occ[di] < 1) return set_error(B200MF_ERR_CUDA, "stacked scalar kernel does not fit on an SM (N=%d, %zu B shared)", N, smem);
  }
  if (g1 <= g0) return 0;
  CellArgs<T> A;
  A.group_begin = g0;
  A.n_groups = g1;
  A.nidx = ctx->nidx;
  A.src = src;
  A.dst = dst;
  A.cstride = ctx->cstride;
  A.dot = dot;
  const int64_t n_super = (g1 - g0 + G::GW - 1) / G::GW;
  const int64_t want = (n_super + WPB - 1) / WPB, cap = (int64_t)ctx->sm_count * occ[di];
  const unsigned grid = (unsigned)std::max<int64_t>(1, std::min(want, cap));
  B2_TRY(kernel_event(const_cast<b200mf_ctx *>(ctx), stream));
  if (di)
    k1<<<grid, 32 * WPB, smem, stream>>>(A, oa);
  else
    k0<<<grid, 32 * WPB, smem, stream>>>(A, oa);
  B2_LAUNCHED();
  B2_TRY(kernel_event(const_cast<b200mf_ctx *>(ctx), stream));
  return 0;
}

#ifdef B200MF_ONLY_N  // tuning builds: instantiate one degree only
#define B2_DISPATCH_N(N_, T_, ...)                                                                \
  if ((N_) == B200MF_ONLY_N) { constexpr int N = B200MF_ONLY_N; using T = T_; __VA_ARGS__; }     \
  else rc = set_error(B200MF_ERR_ARG, "tuning build: only degree %d is instantiated", B200MF_ONLY_N - 1);
#else
#define B2_DISPATCH_N(N_, T_, ...)                                                  \
  switch (N_) {                                                                      \
    case 2: { constexpr int N = 2; using T = T_; __VA_ARGS__; } break;                      \
    case 3: { constexpr int N = 3; using T = T_; __VA_ARGS__; } break;                      \
    case 4: { constexpr int N = 4; using T = T_; __VA_ARGS__; } break;                      \
    case 5: { constexpr int N = 5; using T = T_; __VA_ARGS__; } break;                      \
    case 6: { constexpr int N = 6; using T = T_; __VA_ARGS__; } break;                      \
    case 7: { constexpr int N = 7; using T = T_; __VA_ARGS__; } break;                      \
    default: rc = set_error(B200MF_ERR_ARG, "unsupported degree %d", (N_) - 1);      \
  }
#endif

}  // namespace b200mf
