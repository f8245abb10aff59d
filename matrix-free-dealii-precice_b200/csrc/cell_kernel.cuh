// Sum-factorised cell loop for sm_100a — the skeleton shared by every operator
// of the hot path (a1-a9, a13).
//
// Replaces, per cell batch, deal.II's FEEvaluation sequence used by
//   NeoHookOperator::local_apply_cell   (include/solid_mechanics/mf_nh_operator.h:629-647)
//   NeoHookOperator::cache              (:392-523)
//   LaplaceOperator::local_apply        (include/heat_transfer/heat_transfer.h:170-190)
// i.e. reinit / read_dof_values / evaluate / quadrature loop / integrate /
// distribute_local_to_global.
//
// Design (B200-first, not a translation of deal.II's SIMD-over-cells loop):
//  * one thread owns one 1-D "column" of a cell: N = p+1 values per component in
//    registers, N^2 threads per cell.  Sweeps along the register direction use
//    the 1-D matrices as compile-time constant-bank operands (DFMA R,R,c[..],R:
//    no load instruction), the other two directions are reached by transposing
//    through shared memory (2N smem accesses per N^2 FMAs instead of the N^2 of
//    the thread-per-point scheme: B200 has 4x the HBM bytes per SM-clock of
//    V100/A100 but the same 128 B/clk shared-memory port, so the classic scheme
//    is shared-memory bound long before HBM is).
//  * for N^2 <= 32 a cell never straddles a warp: floor(32/N^2) cells per warp
//    ("group") and every barrier is a __syncwarp; N^2 > 32 uses one 64-thread
//    block per cell and __syncthreads.
//  * all per-cell tables are stored per group in the exact lane order the
//    kernel reads them ([group][field][x-index][lane]), so every table load is
//    one contiguous run per warp and DRAM traffic equals the table size.
//  * F, tau and the linearised stress live only in registers.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace b200mf {

constexpr int MAXN = 7;  // p = 1..6
// bits 29..31 of a node index: component c of this node is constrained
constexpr uint32_t NODE_MASK = (1u << 29) - 1u;
__host__ __device__ constexpr uint32_t node_flag(int c) { return 1u << (29 + c); }
constexpr uint32_t NODE_ALL_CONSTRAINED = 7u << 29;

// [N][0: S[q][i] = l_i(xi_q), 1: C[q][r] = collocation derivative on the Gauss points][q*N + i]
static __constant__ double c_shape_d[MAXN + 1][2][MAXN * MAXN];
static __constant__ float c_shape_f[MAXN + 1][2][MAXN * MAXN];

template <typename T, int N>
struct Shape;
template <int N>
struct Shape<double, N> {
  static __device__ __forceinline__ double S(int q, int i) { return c_shape_d[N][0][q * N + i]; }
  static __device__ __forceinline__ double C(int q, int r) { return c_shape_d[N][1][q * N + r]; }
};
template <int N>
struct Shape<float, N> {
  static __device__ __forceinline__ float S(int q, int i) { return c_shape_f[N][0][q * N + i]; }
  static __device__ __forceinline__ float C(int q, int r) { return c_shape_f[N][1][q * N + r]; }
};

// ---------------------------------------------------------------------------
// Thread / shared-memory layout (strides from tools/smem_layout_search.py)
template <int N, int NC>
struct Layout {
  static constexpr int N2 = N * N, N3 = N * N * N;
  static constexpr bool WARP_MODE = (N2 <= 32);
  static constexpr int CPG = WARP_MODE ? 32 / N2 : 1;                 // cells per group
  static constexpr int L = CPG * N2;                                  // dense lanes per group
  static constexpr int TPG = WARP_MODE ? 32 : ((N2 + 31) / 32) * 32;  // threads per group
  static constexpr int JS = (N % 2) ? N : N + 1;                      // j stride (odd => conflict-free x columns)
  static constexpr int KS = N * JS;
  static constexpr int CSTR = N * KS;  // one component of one buffer
  static constexpr int NBUF = 3;
  static constexpr int CELLPAD = (N == 3) ? (NC == 3 ? 8 : 10) : 0;
  static constexpr int CELL = NBUF * NC * CSTR + CELLPAD;
  static constexpr int GROUP_ELEMS = CPG * CELL;
};

// element offset of field f of q-point (iq, lane) in a [group][NF][N][L] table
template <int N, int L>
__host__ __device__ __forceinline__ size_t qoff(int64_t group, int NF, int f, int iq, int lane) {
  return (((size_t)group * NF + f) * N + iq) * L + lane;
}

template <bool WARP>
__device__ __forceinline__ void group_sync() {
  if (WARP)
    __syncwarp();
  else
    __syncthreads();
}

template <typename T>
__device__ __forceinline__ T ld_stream(const T *p) {
  return __ldcs(p);  // q-point tables are read once per vmult: evict-first
}

template <typename T>
struct CellArgs {
  int64_t n_groups;
  const uint32_t *nidx;  // [group][k][lane], flags in bits 29..31
  const T *src;
  T *dst;
};

// 1-D sweeps on a register column -------------------------------------------
// out[r] = sum_k S[r][k] in[k]   (TR: sum_k S[k][r] in[k])
template <typename T, int N, bool TR>
__device__ __forceinline__ void sweep_S(const T (&in)[N], T (&out)[N]) {
#pragma unroll
  for (int r = 0; r < N; ++r) {
    T acc = T(0);
#pragma unroll
    for (int k = 0; k < N; ++k) acc += (TR ? Shape<T, N>::S(k, r) : Shape<T, N>::S(r, k)) * in[k];
    out[r] = acc;
  }
}
template <typename T, int N>
__device__ __forceinline__ void sweep_C(const T (&in)[N], T (&out)[N]) {
#pragma unroll
  for (int r = 0; r < N; ++r) {
    T acc = T(0);
#pragma unroll
    for (int k = 0; k < N; ++k) acc += Shape<T, N>::C(r, k) * in[k];
    out[r] = acc;
  }
}

// ---------------------------------------------------------------------------
// Op concept:
//   static constexpr int  NC;            components
//   static constexpr bool PLAIN;         read_dof_values_plain (ignore constraints)
//   static constexpr bool INTEGRATE;     run the transpose half and scatter
//   struct Args;                         kernel-argument POD
//   struct CellState;                    per-thread per-cell state (e.g. material constants)
//   static void cell_begin(args, cell, valid, CellState&)
//   static void quad(args, state, active, group, iq, lane, v[NC], G[NC][3], ov[NC], oG[NC][3])
//
// q-table addressing: field f (of NF) of quadrature point (iq, lane) of a group lives at
//   table[((group*NF + f)*N + iq)*L + lane]          (see qoff()).
//
// DIAG = true turns the loop into MatrixFreeTools::compute_diagonal (SURVEY.md A.4;
// mf_nh_operator.h:1033, heat_transfer.h:215): the cell operator is applied to every
// local unit vector and entry i of the result is kept, then scatter-added into dst.
template <typename T, int N, class Op, int GPB, bool DIAG = false>
__global__ void __launch_bounds__(Layout<N, Op::NC>::TPG *GPB, Op::MIN_BLOCKS)
    cell_loop_kernel(const CellArgs<T> A, const typename Op::Args oa) {
  constexpr int NC = Op::NC;
  using Ly = Layout<N, NC>;
  using Sh = Shape<T, N>;
  constexpr int N2 = Ly::N2, JS = Ly::JS, KS = Ly::KS, CSTR = Ly::CSTR, L = Ly::L;
  constexpr bool WM = Ly::WARP_MODE;
  static_assert(WM || GPB == 1, "block mode runs one group per block");

  extern __shared__ __align__(16) unsigned char smem_raw[];
  T *smem = reinterpret_cast<T *>(smem_raw);

  const int gib = threadIdx.x / Ly::TPG;
  const int lane = threadIdx.x % Ly::TPG;
  const bool active = lane < L;
  const int cw = active ? lane / N2 : 0;
  const int t = active ? lane % N2 : 0;
  const int a = t % N, b = t / N;

  T *bufA = smem + (size_t)(gib * Ly::CPG + cw) * Ly::CELL;
  T *bufB = bufA + NC * CSTR;
  T *bufC = bufB + NC * CSTR;

  // per-thread rows/columns of the collocation derivative for the two
  // directions that are contracted straight out of shared memory
  T cy[N], cz[N], cyT[N], czT[N];
#pragma unroll
  for (int m = 0; m < N; ++m) {
    cy[m] = Sh::C(a, m);
    cz[m] = Sh::C(b, m);
    cyT[m] = Sh::C(m, a);
    czT[m] = Sh::C(m, b);
  }

  const int oZ = b * JS + a;           // + k*KS : (i=a, j=b) column along z
  const int oY = b * KS + a;           // + j*JS : (i=a, k=b) column along y
  const int oX = b * KS + a * JS;      // + i    : (j=a, k=b) column along x

  for (int64_t group = (int64_t)blockIdx.x * GPB + gib; group < A.n_groups; group += (int64_t)gridDim.x * GPB) {
    typename Op::CellState st;
    Op::cell_begin(oa, group * Ly::CPG + cw, active, st);

    // ---- phase 1 (Z orientation): gather + z sweep -------------------------
    uint32_t nid[N];
#pragma unroll
    for (int k = 0; k < N; ++k)
      nid[k] = active ? __ldg(A.nidx + ((size_t)group * N + k) * L + lane) : NODE_ALL_CONSTRAINED;
    T dacc[NC][N];
    if (DIAG) {
#pragma unroll
      for (int c = 0; c < NC; ++c)
#pragma unroll
        for (int k = 0; k < N; ++k) dacc[c][k] = T(0);
    }
    constexpr int NPASS = DIAG ? NC * Ly::N3 : 1;
#pragma unroll 1
    for (int pass = 0; pass < NPASS; ++pass) {
    const int pc = pass / Ly::N3, pk = (pass / N2) % N, pt = pass % N2;  // unit vector (component, k, column)
    {
      T u[NC][N];
#pragma unroll
      for (int k = 0; k < N; ++k) {
        const size_t base = (size_t)(nid[k] & NODE_MASK) * NC;
#pragma unroll
        for (int c = 0; c < NC; ++c) {
          const bool skip = Op::PLAIN ? !active : (nid[k] & node_flag(c)) != 0u;
          if (DIAG)
            u[c][k] = (!skip && c == pc && k == pk && t == pt) ? T(1) : T(0);
          else
            u[c][k] = skip ? T(0) : __ldg(A.src + base + c);
        }
      }
#pragma unroll
      for (int c = 0; c < NC; ++c) {
        T o[N];
        sweep_S<T, N, false>(u[c], o);
#pragma unroll
        for (int k = 0; k < N; ++k) if (active) bufA[c * CSTR + k * KS + oZ] = o[k];
      }
    }
    group_sync<WM>();
    // ---- phase 2 (Y orientation) -------------------------------------------
#pragma unroll
    for (int c = 0; c < NC; ++c) {
      T col[N], o[N];
#pragma unroll
      for (int j = 0; j < N; ++j) col[j] = bufA[c * CSTR + j * JS + oY];
      sweep_S<T, N, false>(col, o);
#pragma unroll
      for (int j = 0; j < N; ++j) if (active) bufB[c * CSTR + j * JS + oY] = o[j];
    }
    group_sync<WM>();
    // ---- phase 3 (X orientation): values + x derivative at quadrature points
    T val[NC][N], gx[NC][N];
#pragma unroll
    for (int c = 0; c < NC; ++c) {
      T col[N];
#pragma unroll
      for (int i = 0; i < N; ++i) col[i] = bufB[c * CSTR + oX + i];
      sweep_S<T, N, false>(col, val[c]);
      sweep_C<T, N>(val[c], gx[c]);
#pragma unroll
      for (int i = 0; i < N; ++i) if (active) bufA[c * CSTR + oX + i] = val[c][i];
    }
    group_sync<WM>();
    // ---- phase 4: quadrature loop ------------------------------------------
    T acc[NC][N];
#pragma unroll
    for (int c = 0; c < NC; ++c)
#pragma unroll
      for (int i = 0; i < N; ++i) acc[c][i] = T(0);

#pragma unroll
    for (int iq = 0; iq < N; ++iq) {
      T G[NC][3], v[NC], ov[NC], oG[NC][3];
#pragma unroll
      for (int c = 0; c < NC; ++c) {
        T gy = T(0), gz = T(0);
#pragma unroll
        for (int m = 0; m < N; ++m) {
          gy += cy[m] * bufA[c * CSTR + b * KS + m * JS + iq];
          gz += cz[m] * bufA[c * CSTR + m * KS + a * JS + iq];
        }
        G[c][0] = gx[c][iq];
        G[c][1] = gy;
        G[c][2] = gz;
        v[c] = val[c][iq];
      }
      Op::quad(oa, st, active, group, iq, lane, v, G, ov, oG);
      if (Op::INTEGRATE) {
#pragma unroll
        for (int c = 0; c < NC; ++c) {
#pragma unroll
          for (int i = 0; i < N; ++i) acc[c][i] += Sh::C(iq, i) * oG[c][0];
          acc[c][iq] += ov[c];
          if (active) bufB[c * CSTR + oX + iq] = oG[c][1];
          if (active) bufC[c * CSTR + oX + iq] = oG[c][2];
        }
      }
    }
    if (!Op::INTEGRATE) {
      group_sync<WM>();  // bufA is rewritten by the next group's phase 1
      continue;
    }
    group_sync<WM>();
    // ---- phase 5: y/z derivative transposes, x sweep back ------------------
#pragma unroll
    for (int c = 0; c < NC; ++c) {
#pragma unroll
      for (int iq = 0; iq < N; ++iq) {
        T s = acc[c][iq];
#pragma unroll
        for (int m = 0; m < N; ++m) {
          s += cyT[m] * bufB[c * CSTR + b * KS + m * JS + iq];
          s += czT[m] * bufC[c * CSTR + m * KS + a * JS + iq];
        }
        acc[c][iq] = s;
      }
      T o[N];
      sweep_S<T, N, true>(acc[c], o);
#pragma unroll
      for (int i = 0; i < N; ++i) if (active) bufA[c * CSTR + oX + i] = o[i];
    }
    group_sync<WM>();
    // ---- phase 6 (Y) -------------------------------------------------------
#pragma unroll
    for (int c = 0; c < NC; ++c) {
      T col[N], o[N];
#pragma unroll
      for (int j = 0; j < N; ++j) col[j] = bufA[c * CSTR + j * JS + oY];
      sweep_S<T, N, true>(col, o);
#pragma unroll
      for (int j = 0; j < N; ++j) if (active) bufB[c * CSTR + j * JS + oY] = o[j];
    }
    group_sync<WM>();
    // ---- phase 7 (Z): z sweep back + scatter-add ---------------------------
#pragma unroll
    for (int c = 0; c < NC; ++c) {
      T col[N], o[N];
#pragma unroll
      for (int k = 0; k < N; ++k) col[k] = bufB[c * CSTR + k * KS + oZ];
      sweep_S<T, N, true>(col, o);
#pragma unroll
      for (int k = 0; k < N; ++k) {
        if (DIAG)
          dacc[c][k] += (c == pc && k == pk && t == pt) ? o[k] : T(0);
        else if (!(nid[k] & node_flag(c)))
          atomicAdd(A.dst + (size_t)(nid[k] & NODE_MASK) * NC + c, o[k]);
      }
    }
    // the next iteration's first barrier orders these bufB reads before its phase-2 writes
    }  // pass
    if (DIAG) {
#pragma unroll
      for (int c = 0; c < NC; ++c)
#pragma unroll
        for (int k = 0; k < N; ++k)
          if (!(nid[k] & node_flag(c))) atomicAdd(A.dst + (size_t)(nid[k] & NODE_MASK) * NC + c, dacc[c][k]);
    }
  }
}

}  // namespace b200mf
