// Scalar (one-component) operator on affine cells, N <= 5: the "stacked" warp kernel.
//
// Replaces cell_loop_kernel<T, N, LaplaceAffineOp> for LaplaceOperator::vmult (include/heat_transfer/heat_transfer.h:
// 170-250, (M + dt c K) u on the heat solver's Cartesian/affine cells).  In the generic kernel a scalar group is a
// multi-warp block (6 cells = 150 column threads at N = 5) whose plane phase - 80 % of the FP64 work of this operator,
// the quadrature-point work being ~10 flops - runs on one warp of five while the other four wait at the barrier, with two
// such blocks per SM (profiles/r01_v3_laplace_q4.jsonl: 0.22 of the FP64 pipe).  Here ONE warp owns the CPG cells of a
// group: a lane (j, i) is the column thread of every cell of the group in turn ("stacked" columns: N^2 lanes, CPG cells
// each), and the same lanes run the CPG*N plane tasks of the group in between - both roles keep the warp's lanes busy,
// nothing ever waits for another warp (__syncwarp only), and 8-10 such warps fit on an SM.  For N <= 4 a warp takes
// GW = floor(32 / N^2) groups side by side.  Tables (index map, per-cell metric, coefficient) are the ones of the generic
// kernel: group g, lane s*N^2 + ji for the column ji of cell s.
//
// The DoF values of the next super-group arrive through cp.async (LDGSTS) while the current one computes.
#pragma once
#include "ops.cuh"

namespace b200mf {

template <int N, int WORDS>
struct StackPads;  // tools: /tmp-free search documented in DESIGN.md (stacked access patterns); conflict-free at N = 4 (FP64), 5
template <> struct StackPads<2, 2> { static constexpr int PL = 8, FQ = 0, CQ = 1; };
template <> struct StackPads<3, 2> { static constexpr int PL = 9, FQ = 3, CQ = 1; };
template <> struct StackPads<4, 2> { static constexpr int PL = 17, FQ = 0, CQ = 0; };
template <> struct StackPads<5, 2> { static constexpr int PL = 25, FQ = 2, CQ = 0; };
template <> struct StackPads<2, 1> { static constexpr int PL = 5, FQ = 0, CQ = 0; };
template <> struct StackPads<3, 1> { static constexpr int PL = 10, FQ = 7, CQ = 0; };
template <> struct StackPads<4, 1> { static constexpr int PL = 17, FQ = 0, CQ = 0; };
template <> struct StackPads<5, 1> { static constexpr int PL = 25, FQ = 2, CQ = 0; };

template <typename T, int N>
struct StackGeom {
  using Ly = Layout<N, 1, sizeof(T) / 4>;
  using P = StackPads<N, sizeof(T) / 4>;
  static constexpr int N2 = N * N, CPG = Ly::CPG, L = Ly::L;
  static constexpr int GW = 32 / N2;          // groups side by side in one warp
  static constexpr int LANES = GW * N2;       // column lanes
  static constexpr int CELLS = GW * CPG;      // cells of a super-group
  static constexpr int NT = CELLS * N;        // plane tasks
  static constexpr int PL = P::PL, QFIELD = N * PL + P::FQ, QCELL = 3 * QFIELD + P::CQ;
  static constexpr int BUF_ELEMS = CELLS * QCELL;
  static constexpr int UBUF_ELEMS = CPG * N * 32;
  static constexpr size_t WARP_BYTES = round_up((size_t)(BUF_ELEMS + UBUF_ELEMS) * sizeof(T), 16);
};

template <typename T, int N, class Op, int WPB, bool DOT>
__global__ void __launch_bounds__(32 * WPB) scalar_stack_kernel(const CellArgs<T> A, const typename Op::Args oa) {
  using G = StackGeom<T, N>;
  using Sh = Shape<T, N>;
  constexpr int N2 = G::N2, CPG = G::CPG, L = G::L, GW = G::GW, PL = G::PL, QFIELD = G::QFIELD, QCELL = G::QCELL;
  static_assert(Op::NC == 1 && Op::NF_IN == 0 && !Op::PRE, "stacked kernel: scalar operator without a staged metric table");
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
  T *bufQ = reinterpret_cast<T *>(smem_raw + (size_t)warp * G::WARP_BYTES);
  T *ubuf = bufQ + G::BUF_ELEMS;
  const int gw = lane / N2, ji = lane % N2;
  const bool col_lane = lane < G::LANES;

  const int64_t n_super = (A.n_groups - A.group_begin + GW - 1) / GW;
  const int64_t sstep = (int64_t)gridDim.x * WPB;
  int64_t sg = (int64_t)blockIdx.x * WPB + warp;

  // group this lane works for inside super-group s (or -1)
  auto my_group = [&](int64_t s) -> int64_t {
    const int64_t g = A.group_begin + s * GW + gw;
    return (col_lane && g < A.n_groups) ? g : -1;
  };
  // asynchronous gather of the nodal values of super-group s into ubuf[(cell_in_group*N + k)*32 + lane]
  auto gather_async = [&](int64_t s) {
    const int64_t g = my_group(s);
    if (g >= 0) {
#pragma unroll
      for (int c = 0; c < CPG; ++c)
#pragma unroll
        for (int k = 0; k < N; ++k) {
          const uint32_t id = __ldg(A.nidx + ((size_t)g * N + k) * L + c * N2 + ji);
          cp_async_elem(ubuf + (c * N + k) * 32 + lane, A.src + (id & NODE_MASK), Op::PLAIN || (id & node_flag(0)) == 0u);
        }
    }
    cp_async_commit();
  };

  double dot_acc = 0.0;
  if (sg < n_super) gather_async(sg);
  for (; sg < n_super; sg += sstep) {
    const int64_t g = my_group(sg);
    const bool active = g >= 0;
    const int groups_here = (int)((n_super - 1 == sg) ? (A.n_groups - A.group_begin - sg * GW) : GW);
    const int nt_here = groups_here * CPG * N;
    // ---- F1: z sweep of every stacked column, planes to field 0
    cp_async_wait_all();
    if (active) {
#pragma unroll
      for (int c = 0; c < CPG; ++c) {
        T u[N], o[N];
#pragma unroll
        for (int k = 0; k < N; ++k) u[k] = ubuf[(c * N + k) * 32 + lane];
        sweep_S<T, N, false>(u, o);
        T *zp = bufQ + (gw * CPG + c) * QCELL + ji;
#pragma unroll
        for (int kq = 0; kq < N; ++kq) zp[kq * PL] = o[kq];
      }
    }
    if (sg + sstep < n_super) gather_async(sg + sstep);  // ubuf entries of this lane have been consumed by this lane
    __syncwarp();
    // ---- F2: plane tasks (cell, k): x, y sweeps, values and x/y derivatives at the quadrature points of the plane
#pragma unroll 1
    for (int t0 = 0; t0 < G::NT; t0 += 32) {
      const int t = t0 + lane;
      if (t < nt_here) {
        T *zp = bufQ + (t / N) * QCELL + (t % N) * PL;
        T a[N][N], b[N][N];
#pragma unroll
        for (int j = 0; j < N; ++j)
#pragma unroll
          for (int i = 0; i < N; ++i) a[j][i] = zp[j * N + i];
#pragma unroll
        for (int j = 0; j < N; ++j) sweep_S<T, N, false>(a[j], b[j]);
#pragma unroll
        for (int qx = 0; qx < N; ++qx) {
          T col[N], o[N];
#pragma unroll
          for (int j = 0; j < N; ++j) col[j] = b[j][qx];
          sweep_S<T, N, false>(col, o);
#pragma unroll
          for (int qy = 0; qy < N; ++qy) a[qy][qx] = o[qy];
        }
#pragma unroll
        for (int qy = 0; qy < N; ++qy) {
          T gx[N];
          sweep_C<T, N>(a[qy], gx);
#pragma unroll
          for (int qx = 0; qx < N; ++qx) {
            zp[qy * N + qx] = a[qy][qx];
            zp[QFIELD + qy * N + qx] = gx[qx];
          }
        }
#pragma unroll
        for (int qx = 0; qx < N; ++qx) {
          T col[N], gy[N];
#pragma unroll
          for (int qy = 0; qy < N; ++qy) col[qy] = a[qy][qx];
          sweep_C<T, N>(col, gy);
#pragma unroll
          for (int qy = 0; qy < N; ++qy) zp[2 * QFIELD + qy * N + qx] = gy[qy];
        }
      }
    }
    __syncwarp();
    // ---- F3: per stacked cell d/dz in registers, quadrature points of the column, z part of the integration
    if (active) {
#pragma unroll 1
      for (int c = 0; c < CPG; ++c) {
        typename Op::CellState st;
        Op::cell_begin(oa, g * CPG + c, true, c * N2 + ji, st);
        T *qp = bufQ + (gw * CPG + c) * QCELL + ji;
        T val[N], gz[N], racc[N];
#pragma unroll
        for (int kq = 0; kq < N; ++kq) {
          val[kq] = qp[kq * PL];
          racc[kq] = T(0);
        }
        sweep_C<T, N>(val, gz);
#pragma unroll
        for (int kq = 0; kq < N; ++kq) {
          T v[1] = {val[kq]}, Gq[1][3] = {{qp[QFIELD + kq * PL], qp[2 * QFIELD + kq * PL], gz[kq]}}, ov[1], oG[1][3];
          const T *q2 = oa.stage2 != nullptr ? oa.stage2 + (size_t)g * Layout<N, 1, sizeof(T) / 4>::table_stride(1) + kq * L + c * N2 + ji : nullptr;
          Op::quad(oa, st, true, nullptr, q2, nullptr, g, kq, c * N2 + ji, v, Gq, ov, oG);
          if constexpr (DOT) dot_acc += (double)(v[0] * ov[0] + Gq[0][0] * oG[0][0] + Gq[0][1] * oG[0][1] + Gq[0][2] * oG[0][2]);
#pragma unroll
          for (int m = 0; m < N; ++m) racc[m] += Sh::C(kq, m) * oG[0][2];
          racc[kq] += ov[0];
          qp[QFIELD + kq * PL] = oG[0][0];
          qp[2 * QFIELD + kq * PL] = oG[0][1];
        }
#pragma unroll
        for (int kq = 0; kq < N; ++kq) qp[kq * PL] = racc[kq];
      }
    }
    __syncwarp();
    // ---- B2: plane tasks: transposed x/y derivative and x, y sweeps
#pragma unroll 1
    for (int t0 = 0; t0 < G::NT; t0 += 32) {
      const int t = t0 + lane;
      if (t < nt_here) {
        T *zp = bufQ + (t / N) * QCELL + (t % N) * PL;
        T a[N][N];
#pragma unroll
        for (int qy = 0; qy < N; ++qy) {
          T gx[N];
#pragma unroll
          for (int qx = 0; qx < N; ++qx) {
            a[qy][qx] = zp[qy * N + qx];
            gx[qx] = zp[QFIELD + qy * N + qx];
          }
#pragma unroll
          for (int qx = 0; qx < N; ++qx)
#pragma unroll
            for (int m = 0; m < N; ++m) a[qy][qx] += Sh::C(m, qx) * gx[m];
        }
#pragma unroll
        for (int qx = 0; qx < N; ++qx) {
          T gy[N];
#pragma unroll
          for (int qy = 0; qy < N; ++qy) gy[qy] = zp[2 * QFIELD + qy * N + qx];
#pragma unroll
          for (int qy = 0; qy < N; ++qy)
#pragma unroll
            for (int m = 0; m < N; ++m) a[qy][qx] += Sh::C(m, qy) * gy[m];
        }
        T b[N][N];
#pragma unroll
        for (int qy = 0; qy < N; ++qy) sweep_S<T, N, true>(a[qy], b[qy]);
#pragma unroll
        for (int i = 0; i < N; ++i) {
          T col[N], o[N];
#pragma unroll
          for (int qy = 0; qy < N; ++qy) col[qy] = b[qy][i];
          sweep_S<T, N, true>(col, o);
#pragma unroll
          for (int j = 0; j < N; ++j) zp[j * N + i] = o[j];
        }
      }
    }
    __syncwarp();
    // ---- B1: z sweep back + scatter-add per stacked cell
    if (active) {
#pragma unroll
      for (int c = 0; c < CPG; ++c) {
        const T *zp = bufQ + (gw * CPG + c) * QCELL + ji;
        T col[N], o[N];
#pragma unroll
        for (int kq = 0; kq < N; ++kq) col[kq] = zp[kq * PL];
        sweep_S<T, N, true>(col, o);
#pragma unroll
        for (int k = 0; k < N; ++k) {
          const uint32_t id = __ldg(A.nidx + ((size_t)g * N + k) * L + c * N2 + ji);
          if (!(id & node_flag(0))) atomicAdd(A.dst + (id & NODE_MASK), o[k]);
        }
      }
    }
    __syncwarp();  // field 0 is rewritten by the next super-group's F1 of other lanes' cells only after this point
  }
  if constexpr (DOT) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) dot_acc += __shfl_down_sync(0xffffffffu, dot_acc, o);
    if (lane == 0 && dot_acc != 0.0) atomicAdd(A.dot, dot_acc);
  }
}

}  // namespace b200mf
