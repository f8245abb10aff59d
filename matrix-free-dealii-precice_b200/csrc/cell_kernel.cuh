// Sum-factorised cell loop for sm_100a — the skeleton shared by every operator
// of the hot path (a1-a9, a11, a13).
//
// Replaces, per cell batch, deal.II's FEEvaluation sequence used by
//   NeoHookOperator::local_apply_cell   (include/solid_mechanics/mf_nh_operator.h:629-647)
//   NeoHookOperator::cache              (:392-523)
//   LaplaceOperator::local_apply        (include/heat_transfer/heat_transfer.h:170-190)
//   MatrixFreeTools::compute_diagonal   (mf_nh_operator.h:1033, heat_transfer.h:215)
// i.e. reinit / read_dof_values / evaluate / quadrature loop / integrate /
// distribute_local_to_global.
//
// Design (B200-first, not a translation of deal.II's SIMD-over-cells loop).  The first two
// versions of this kernel (profiles/r01_v1_*, r01_v2_*) were bound by the shared-memory / L1
// data pipe (l1tex__data_pipe_lsu_wavefronts 82-87 % at 47-65 % of DRAM): B200 has 4x the HBM
// bytes per SM-clock of V100/A100 but the same 128 B/clk shared-memory port.  This version
// (v3) therefore minimises shared-memory wavefronts per cell:
//  * two thread roles inside a group of cells.  A COLUMN thread (i, j) owns the z-column of
//    one cell: it gathers the nodal values (coalesced along i), applies the z sweep in
//    registers, later runs the quadrature-point physics for its N points and the z part of
//    the integration, and scatters.  A PLANE task (component c, z-plane kq) owns one x-y
//    plane of one component: N^2 values in registers, x and y sweeps and the x/y collocation
//    derivatives without touching shared memory.
//  * per component 16 N^3 element accesses of shared memory per cell instead of (11 + 4N) N^3
//    (N = 4: -41 %, N = 7: -59 %): z-swept nodal planes (column -> plane), values + d/dx + d/dy
//    at the quadrature points (plane -> column), and the same two hops back.  d/dz is a register
//    sweep of the column thread.  All 1-D matrices are compile-time constant-bank operands
//    (DFMA R, R, c[..], R: no load instruction).
//  * for N^2 <= 32 a cell never straddles a warp: floor(32/N^2) cells per warp ("group") and
//    every barrier is a __syncwarp; N^2 > 32 uses one 64-thread block per cell.
//  * all per-cell tables are stored per group in the exact lane order the column threads read
//    them ([group][field][kq][lane], lane = (cell, j, i)); a group's table is one contiguous
//    block that the TMA unit (cp.async.bulk + mbarrier) copies into shared memory one group
//    ahead of the arithmetic, and the DoF gather of the next group is issued into registers
//    before the integration half of the current one: HBM latency is off the critical path.
//  * vectors are stored in component blocks (DoF (node, c) at c*cstride + node): a row of N
//    nodes of one component is one contiguous run (half the sectors per gather / red.global.add
//    of interleaved components).
//  * F, tau and the linearised stress live only in registers.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <type_traits>

namespace b200mf {

constexpr int MAXN = 7;  // p = 1..6
// bits 29..31 of a node index: component c of this node is constrained
constexpr uint32_t NODE_MASK = (1u << 29) - 1u;
__host__ __device__ constexpr uint32_t node_flag(int c) { return 1u << (29 + c); }
constexpr uint32_t NODE_ALL_CONSTRAINED = 7u << 29;

// [N][0: S[q][i] = l_i(xi_q), 1: C[q][r] = collocation derivative on the Gauss points, 2: D = C S][q*N + i],
// [N][3][q] = Gauss weight w_q
constexpr int SHAPE_ROWS = 4;
static __constant__ double c_shape_d[MAXN + 1][SHAPE_ROWS][MAXN * MAXN];
static __constant__ float c_shape_f[MAXN + 1][SHAPE_ROWS][MAXN * MAXN];

template <typename T, int N>
struct Shape;
template <int N>
struct Shape<double, N> {
  static __device__ __forceinline__ double S(int q, int i) { return c_shape_d[N][0][q * N + i]; }
  static __device__ __forceinline__ double C(int q, int r) { return c_shape_d[N][1][q * N + r]; }
  static __device__ __forceinline__ double D(int q, int i) { return c_shape_d[N][2][q * N + i]; }
  static __device__ __forceinline__ double W(int q) { return c_shape_d[N][3][q]; }
};
template <int N>
struct Shape<float, N> {
  static __device__ __forceinline__ float S(int q, int i) { return c_shape_f[N][0][q * N + i]; }
  static __device__ __forceinline__ float C(int q, int r) { return c_shape_f[N][1][q * N + r]; }
  static __device__ __forceinline__ float D(int q, int i) { return c_shape_f[N][2][q * N + i]; }
  static __device__ __forceinline__ float W(int q) { return c_shape_f[N][3][q]; }
};

// ---------------------------------------------------------------------------
// Shared-memory layout.  Two plane buffers per group with nested strides (plane stride PL, element j*N + i):
//   bufZ plane (cw, c, kq)    at cw*ZCELL + c*ZCOMP + kq*PL,             ZCOMP = N*PL + CCZ, ZCELL = NC*ZCOMP + CZ
//   bufQ plane (cw, c, f, kq) at cw*QCELL + c*QCOMP + f*QFIELD + kq*PL,  QFIELD = N*PL + FQ, QCOMP = 3*QFIELD + CCQ, QCELL = NC*QCOMP + CQ
// Pads from tools/smem_layout_search.py (64-bit accesses are served per half-warp over 16 8-byte banks).
template <int WORDS, int NC, int N>
struct SmemPad3;
#define B2_PAD3(W, NC_, N_, PL_, CZ_, CCZ_, CQ_, CCQ_, FQ_, ORD_) \
  template <> struct SmemPad3<W, NC_, N_> { static constexpr int PL = PL_, CZ = CZ_, CCZ = CCZ_, CQ = CQ_, CCQ = CCQ_, FQ = FQ_, ORD = ORD_; };
#include "smem_pads.inc"
#undef B2_PAD3

__host__ __device__ constexpr size_t round_up(size_t x, size_t m) { return (x + m - 1) / m * m; }

// Cells per group and threads per group.  A group is the unit that shares one set of shared-memory buffers and
// barriers: its CPG*N^2 column threads and CPG*NC*N plane tasks are both packed densely from lane 0, so that
// a phase only issues on the warps that hold work (an SM sub-partition spends two cycles on a DFMA whatever
// the number of active lanes).  One warp per group wherever the plane tasks fill most of it (group barriers are
// __syncwarp); the scalar operator (N plane tasks per cell only) and the vector operator at N >= 5 pack several
// cells into a multi-warp group instead (N = 5: 2 cells = 50 columns / 30 plane tasks in 64 threads, N = 6:
// 3 cells = 108 / 54 in 128, N = 7: 2 cells = 98 / 42 in 128).  Host (ctx->cpg, ctx->L) and device use these functions.
#ifndef B200MF_CPG5
#define B200MF_CPG5 2
#endif
#ifndef B200MF_CPG6
#define B200MF_CPG6 3
#endif
#ifndef B200MF_CPG7
#define B200MF_CPG7 2
#endif
__host__ __device__ constexpr int group_cells(int N, int NC) {
#ifdef B200MF_TUNE_CPG51  // tuning builds
  if (NC == 1 && N == 5) return B200MF_TUNE_CPG51;
#endif
  if (NC == 1) return N == 2 ? 8 : N == 3 ? 7 : N == 4 ? 8 : N == 5 ? 6 : N == 6 ? 4 : 2;
  if (N == 5) return B200MF_CPG5;
  if (N == 6) return B200MF_CPG6;
  if (N == 7) return B200MF_CPG7;
  return 32 / (N * N);
}
__host__ __device__ constexpr int group_threads(int N, int NC) { return (group_cells(N, NC) * N * N + 31) / 32 * 32; }

// Staged q-point tables come in two layouts (lane = (cell in group, j, i), L lanes per group):
//   plain  [group][field][kq][lane]   one TMA bulk copy per group stages the whole table of the group;
//   sliced [group][kq][field][lane]   one bulk copy per z-plane of quadrature points into a small ring of
//          slots: the column threads consume plane kq while the planes kq+1.. (of this or the next group) are in
//          flight.  Shared memory per cell drops from NF*N^3 to RING*NF*N^2 elements - at N >= 5 the whole-group
//          table was what held the kernel at 7 warps per SM (profiles/r01_v3_ncu_summary.md).
// A slice is padded to a multiple of 16 bytes (cp.async.bulk granularity).
#ifndef B200MF_SLICED_MINN
#define B200MF_SLICED_MINN 5
#endif
#ifndef B200MF_RING
#define B200MF_RING 2
#endif
#ifndef B200MF_VAL_REGS
#define B200MF_VAL_REGS 0
#endif
// the DoF gather of the next group (sliced kernels; the others prefetch into registers):
//   0  plain loads into registers one group ahead (30-42 registers held across the integration half: under a register
//      cap the compiler spills them right after the loads, i.e. waits for them - profiles/r02_*),
//   1  cp.async into a shared-memory buffer (no registers, + NC*N*L elements of shared memory per group),
//   2  prefetch.global.L2 of the next group's entries during the integration half, plain (L2-hit) loads at the start
//      of the group: neither registers nor shared memory
#ifndef B200MF_ASYNC_GATHER
#define B200MF_ASYNC_GATHER 2
#endif
#ifndef B200MF_PGE
#define B200MF_PGE 0
#endif
#ifndef B200MF_EMPTY_MBAR
#define B200MF_EMPTY_MBAR 0
#endif
// Sliced kernels, lanes without a column (lane >= L): 1 = they skip the plane body and only take part in the slot hand-over
// (compute-sanitizer racecheck reports 0 hazards); 0 = they mirror lane 0's shared-memory reads and discard the values
// (racecheck flags those reads against lane 0's later stores - harmless, and branch-free: Q4 0.551 vs 0.529, Q6 0.292 vs
// 0.270 of HBM; Q5 0.455 vs 0.460); -1 = per degree what measured faster (guard for N = 6 only).  profiles/r02_racecheck.txt
#ifndef B200MF_GUARD_IDLE
#define B200MF_GUARD_IDLE -1
#endif
#ifndef B200MF_PLANE_SPREAD
#define B200MF_PLANE_SPREAD 0
#endif
// sliced kernels run the quadrature loop rolled (one small loop body, no load hoisting across planes: ~70 registers less)
#ifndef B200MF_ROLLQ
#define B200MF_ROLLQ 1
#endif
__host__ __device__ constexpr bool table_sliced(int N, int NC) { return NC == 3 && N >= B200MF_SLICED_MINN; }
// the z-swept nodal planes (column -> plane hand-off) live in field 0 of the q-point buffer: a plane task reads
// its own plane completely before it writes it, so no second buffer is needed
#ifndef B200MF_ALIAS_MINN
#define B200MF_ALIAS_MINN 4
#endif
__host__ __device__ constexpr bool alias_z(int N, int NC) { return NC == 3 && N >= B200MF_ALIAS_MINN; }

struct TableGeom {  // element (group, field, kq, lane) of a table lives at group*gstride + kq*kqstride + field*fstride + lane
  int64_t gstride;
  int kqstride, fstride;
};
__host__ __device__ constexpr TableGeom table_geom(bool sliced, int N, int L, int NF, int elem_bytes) {
  const int al = 16 / elem_bytes;
  if (sliced) {
    const int sl = (int)round_up((size_t)NF * L, al);
    return TableGeom{(int64_t)N * sl, sl, L};
  }
  return TableGeom{(int64_t)round_up((size_t)NF * N * L, al), L, N * L};
}

template <int N, int NC, int WORDS = 2>
struct Layout {
  static constexpr int N2 = N * N, N3 = N * N * N;
  static constexpr int CPG = group_cells(N, NC);   // cells per group
  static constexpr int L = CPG * N2;               // column threads (dense lanes) per group
  static constexpr int TPG = group_threads(N, NC); // threads per group
  static constexpr bool WARP_MODE = (TPG == 32);
  static constexpr int NT = CPG * NC * N;          // plane tasks per group
  static constexpr bool SLICED = table_sliced(N, NC);
  static constexpr bool ALIAS_Z = alias_z(N, NC);
  static constexpr int RING = B200MF_RING;
  static constexpr int GATHER_MODE = SLICED ? B200MF_ASYNC_GATHER : 0;
  static constexpr bool ASYNC_GATHER = GATHER_MODE == 1;
  using P = SmemPad3<WORDS, NC, N>;
  static constexpr int PL = P::PL;
  // plane task (cell, component, z-plane) of a lane: the enumeration is part of the bank-conflict-free layout (smem_pads.inc)
  static constexpr int ORD = P::ORD;
  // slot width of ORD 1: 16 lanes per cell (all tasks of two cells in one warp) or, tuning builds with B200MF_PLANE_SPREAD,
  // 32 (one warp per cell: no warp idles during the plane phases, at 15 of 32 lanes each)
  static constexpr int SLOT = B200MF_PLANE_SPREAD ? 32 : 16;
  static constexpr int PLANE_LANES = ORD == 1 ? CPG * SLOT : NT;  // lanes the tasks are spread over (with gaps for ORD 1)
  static_assert(ORD != 1 || NC * N <= 16, "slot enumeration needs NC*N <= 16");
  __host__ __device__ static constexpr bool task_valid(int t) { return ORD == 1 ? (t % SLOT) < NC * N && t / SLOT < CPG : t < NT; }
  __host__ __device__ static constexpr int task_cell(int t) {
    return ORD == 1 ? t / SLOT : ORD == 2 ? (t / NC) % CPG : ORD == 3 ? t % CPG : t / (NC * N);
  }
  __host__ __device__ static constexpr int task_comp(int t) {
    return ORD == 1 ? (t % SLOT) / N : ORD == 2 ? t % NC : ORD == 3 ? t / (CPG * N) : (t / N) % NC;
  }
  __host__ __device__ static constexpr int task_plane(int t) {
    return ORD == 1 ? (t % SLOT) % N : ORD == 2 ? t / (NC * CPG) : ORD == 3 ? (t / CPG) % N : t % N;
  }
  static constexpr int ZCOMP = N * PL + P::CCZ, ZCELL = NC * ZCOMP + P::CZ;
  static constexpr int QFIELD = N * PL + P::FQ, QCOMP = 3 * QFIELD + P::CCQ, QCELL = NC * QCOMP + P::CQ;
  static constexpr int ZSIZE = ALIAS_Z ? 0 : CPG * ZCELL;
  static constexpr int QSIZE = CPG * QCELL;
  static constexpr int GROUP_ELEMS = ZSIZE + QSIZE;
  __host__ __device__ static constexpr int qplane(int cw, int c, int f, int kq) { return cw * QCELL + c * QCOMP + f * QFIELD + kq * PL; }
  // relative to bufZ (== bufQ when aliased)
  __host__ __device__ static constexpr int zplane(int cw, int c, int kq) {
    return ALIAS_Z ? qplane(cw, c, 0, kq) : cw * ZCELL + c * ZCOMP + kq * PL;
  }
  // staged tables (see table_geom): q-point (kq, lane), field f at  kq*kq_stride(NF) + f*FS + lane  behind the group's base
  static constexpr int FS = SLICED ? L : N * L;
  __host__ __device__ static constexpr size_t slice_stride(int NF) { return round_up((size_t)NF * L, 16 / (4 * WORDS)); }
  __host__ __device__ static constexpr size_t kq_stride(int NF) { return SLICED ? slice_stride(NF) : (size_t)L; }
  __host__ __device__ static constexpr size_t table_stride(int NF) {
    return SLICED ? N * slice_stride(NF) : round_up((size_t)NF * N * L, 16 / (4 * WORDS));
  }
  // tables that are never staged (read or written straight from/to global memory) and the shared-memory scratch: always plain
  static constexpr int FSP = N * L;
  __host__ __device__ static constexpr size_t plain_stride(int NF) { return round_up((size_t)NF * N * L, 16 / (4 * WORDS)); }
};

template <bool WARP>
__device__ __forceinline__ void group_sync() {
  if (WARP)
    __syncwarp();
  else
    __syncthreads();
}

// ---- TMA bulk copy + mbarrier (PTX; SASS: UBLKCP / SYNCS) ------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, unsigned bytes, uint64_t *bar) {
  // the q-point tables are read exactly once per operator application: an L2 evict-first hint keeps them from displacing
  // the src/dst vector lines (re-read by up to 8 cells per node); B200, Q3: 0.583 -> 0.577 ms per launch at 2.1e7 DoFs
  // (-DB200MF_TABLE_NO_EVICT_FIRST restores the plain copy for A/B runs)
#ifndef B200MF_TABLE_NO_EVICT_FIRST
  uint64_t policy;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(policy));
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(
                   smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar)), "l"(policy)
               : "memory");
#else
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
#endif
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, unsigned parity) {
  uint32_t ok;
  do {
    asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}"
                 : "=r"(ok)
                 : "r"(smem_u32(bar)), "r"(parity)
                 : "memory");
  } while (!ok);
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// ---- asynchronous element gather (cp.async, SASS: LDGSTS): global -> shared memory without a register in between;
// valid = false writes a zero (constrained DoFs read as 0) -----------------------------------------------------------
template <typename T>
__device__ __forceinline__ void cp_async_elem(T *smem_dst, const T *gmem_src, bool valid) {
  const unsigned sz = valid ? (unsigned)sizeof(T) : 0u;
  if constexpr (sizeof(T) == 8)
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(smem_u32(smem_dst)), "l"(gmem_src), "r"(sz) : "memory");
  else
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(smem_u32(smem_dst)), "l"(gmem_src), "r"(sz) : "memory");
}
__device__ __forceinline__ void prefetch_l2(const void *p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

template <typename T>
struct CellArgs {
  int64_t group_begin, n_groups;  // groups [group_begin, n_groups)
  const uint32_t *nidx;           // [group][k][lane], flags in bits 29..31
  const T *src;
  T *dst;
  int64_t cstride;  // vector layout: DoF (node, c) lives at c*cstride + node ("component blocks")
  double *dot;      // non-null: += sum over the cells of u_cell . (A_cell u_cell), i.e. this rank's share of src . (A src)
                    // on the unconstrained rows: the CG dot product p . Ap fused into the operator pass
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
//   static constexpr bool INTEGRATE;     run the integration half and scatter
//   static constexpr bool PRE;           one extra evaluation-only pass (pass -1) over a second cell-local field
//                                        (Op::load_pre(..., nid, column)) whose q-point results (Op::quad_pre(..., v, G))
//                                        stay in shared memory
//   static constexpr bool STORE_DOFS;    Op::store_dofs gets the gathered cell-local values
//   static constexpr int  NF_IN, NF_IN2; fields of the (up to two) input tables staged through shared memory
//                                        (NF_IN may be 0: per-cell data only, e.g. the affine Laplace operator)
//   static constexpr int  NF_SCRATCH;    fields of per-group scratch (pre-pass -> main pass)
//   struct Args { const T *stage, *stage2; ... }   stage2 may be null at run time
//   struct CellState;                    per-thread per-cell state (material constants)
//   static void cell_begin(args, cell, valid, lane, CellState&)
//   static void quad(args, state, active, q, q2, scr, group, kq, lane, v[NC], G[NC][3], ov[NC], oG[NC][3])
//       the column thread `lane` = (cell, j, i) handles the q-points (i, j, kq), kq = 0..N-1;
//       q / q2: this q-point's entry of the two staged tables in shared memory, field f at q[f * Layout::FS];
//       scr: the group's scratch, plain layout: field f of q-point (kq, lane) at scr[(f*N + kq)*L + lane];
//       G[c][d] = d u_c / d xi_d (reference gradient), oG likewise for the test-function side
//
// DIAG = true computes what MatrixFreeTools::compute_diagonal computes - entry i of the cell
// operator applied to the local unit vector e_i - without the 3 N^3 operator applications per cell it costs
// in deal.II.  The q-point operation is a linear map (value, d/dxi_1..3) -> (value, d/dxi_1..3); for the unit
// vector of component c only its 4x4 block Q_c(q) matters (obtained generically from four Op::quad calls on
// unit inputs), and
//     diag_i = sum_q sum_{a,b} Q_c(q)[a][b] Phi_a(q,i) Phi_b(q,i),   Phi_0 = S S S, Phi_1 = D S S, Phi_2 = S D S, Phi_3 = S S D
// factorises direction by direction into sweeps with the entry-wise products S.S, S.D, D.D of the 1-D
// matrices: 10 symmetric terms collapse to 6 fields after the z sweep, 3 after y, 1 after x.
template <typename T, int N, class Op, bool DIAG = false>
struct KernelSmem {
  using Ly = Layout<N, Op::NC, sizeof(T) / 4>;
  static constexpr int DIAG_ELEMS = Ly::CPG * 6 * Ly::N3;  // diagonal pass: 6 coefficient fields per cell (see below)
  static constexpr int BUF_ELEMS = (DIAG && DIAG_ELEMS > Ly::GROUP_ELEMS) ? DIAG_ELEMS : Ly::GROUP_ELEMS;
  static constexpr size_t BUF_BYTES = round_up((size_t)BUF_ELEMS * sizeof(T), 16);
  // plain: the whole table of the group; sliced: RING slots of one z-plane of both tables
  static constexpr size_t SL1_ELEMS = Op::NF_IN ? Ly::slice_stride(Op::NF_IN) : 0;
  static constexpr size_t SL2_ELEMS = Op::NF_IN2 ? Ly::slice_stride(Op::NF_IN2) : 0;
  static constexpr size_t SLOT_ELEMS = SL1_ELEMS + SL2_ELEMS;
  static constexpr size_t ST1_BYTES = Ly::SLICED ? Ly::RING * SLOT_ELEMS * sizeof(T) : (Op::NF_IN ? Ly::table_stride(Op::NF_IN) * sizeof(T) : 0);
  static constexpr size_t ST2_BYTES = Ly::SLICED ? 0 : (Op::NF_IN2 ? Ly::table_stride(Op::NF_IN2) * sizeof(T) : 0);
  static constexpr size_t SCR_BYTES = Op::NF_SCRATCH ? Ly::plain_stride(Op::NF_SCRATCH) * sizeof(T) : 0;  // pre-pass -> main pass
  static constexpr size_t MBAR_BYTES = Ly::SLICED ? round_up(8 * 2 * Ly::RING, 16) : 16;  // sliced: "full" and "empty" barrier per slot
  // gathered nodal values of the next group [c][k][lane] (evaluation-only and diagonal passes gather synchronously / not at all)
  static constexpr size_t GATHER_BYTES = (Ly::ASYNC_GATHER && Op::INTEGRATE && !DIAG) ? round_up((size_t)Op::NC * N * Ly::L * sizeof(T), 16) : 0;
  static constexpr size_t GROUP_BYTES = BUF_BYTES + ST1_BYTES + ST2_BYTES + SCR_BYTES + MBAR_BYTES + GATHER_BYTES;
};

template <typename T, int N, class Op, int GPB, bool DIAG = false, int MINB = 1, bool DOT = false>
__global__ void __launch_bounds__(Layout<N, Op::NC>::TPG *GPB, MINB)
    cell_loop_kernel(const CellArgs<T> A, const typename Op::Args oa) {
  constexpr int NC = Op::NC;
  using Ly = Layout<N, NC, sizeof(T) / 4>;
  using Sh = Shape<T, N>;
  using KS_ = KernelSmem<T, N, Op, DIAG>;
  constexpr int N2 = Ly::N2, L = Ly::L, PL = Ly::PL;
  constexpr bool WM = Ly::WARP_MODE;
  constexpr bool SLICED = Ly::SLICED;
  constexpr int RING = Ly::RING;
  static_assert(WM || GPB == 1, "block mode runs one group per block");

  extern __shared__ __align__(128) unsigned char smem_raw[];

  const int gib = threadIdx.x / Ly::TPG;
  const int lane = threadIdx.x % Ly::TPG;
  const bool active = lane < L;           // column thread
  const int cw = active ? lane / N2 : 0;
  const int ji = active ? lane % N2 : 0;  // j*N + i

  unsigned char *gbase = smem_raw + (size_t)gib * KS_::GROUP_BYTES;
  T *bufZ = reinterpret_cast<T *>(gbase);
  T *bufQ = bufZ + Ly::ZSIZE;  // == bufZ when the z-swept planes live in field 0 of the q-point buffer
  T *qs = reinterpret_cast<T *>(gbase + KS_::BUF_BYTES);  // plain: table 1 of the group; sliced: the ring
  T *qs2 = reinterpret_cast<T *>(gbase + KS_::BUF_BYTES + KS_::ST1_BYTES);
  T *scr = reinterpret_cast<T *>(gbase + KS_::BUF_BYTES + KS_::ST1_BYTES + KS_::ST2_BYTES);
  uint64_t *mbar = reinterpret_cast<uint64_t *>(gbase + KS_::BUF_BYTES + KS_::ST1_BYTES + KS_::ST2_BYTES + KS_::SCR_BYTES);
  constexpr bool AG = KS_::GATHER_BYTES != 0;  // next group's nodal values arrive through cp.async
  T *ubuf = reinterpret_cast<T *>(gbase + KS_::BUF_BYTES + KS_::ST1_BYTES + KS_::ST2_BYTES + KS_::SCR_BYTES + KS_::MBAR_BYTES);

  if (lane == 0) {
#pragma unroll
    for (int r = 0; r < (SLICED ? RING : 1); ++r) mbar_init(mbar + r, 1);
    if constexpr (SLICED) {  // "slot r is free again": one arrival per warp of the group
#pragma unroll
      for (int r = 0; r < RING; ++r) mbar_init(mbar + RING + r, Ly::TPG / 32);
    }
  }
  group_sync<WM>();

  const int64_t gstep = (int64_t)gridDim.x * GPB;
  int64_t group = A.group_begin + (int64_t)blockIdx.x * GPB + gib;

  // (Op::NF_IN == 0 and no second table: nothing is staged and the mbarriers are never armed or waited on)
  const bool two = Op::NF_IN2 != 0 && oa.stage2 != nullptr;
  const bool staged = Op::NF_IN != 0 || two;
  // ---- plain layout: one elected lane stages the tables of `group` with one bulk copy per table ------------
  auto stage = [&](int64_t g) {
    if (!SLICED && lane == 0 && staged) {
      mbar_expect_tx(mbar, (unsigned)((Op::NF_IN ? KS_::ST1_BYTES : 0) + (two ? KS_::ST2_BYTES : 0)));
      if (Op::NF_IN) bulk_g2s(qs, oa.stage + (size_t)g * Ly::table_stride(Op::NF_IN ? Op::NF_IN : 1), (unsigned)KS_::ST1_BYTES, mbar);
      if (two) bulk_g2s(qs2, oa.stage2 + (size_t)g * Ly::table_stride(Op::NF_IN2 ? Op::NF_IN2 : 1), (unsigned)KS_::ST2_BYTES, mbar);
    }
  };
  // ---- sliced layout: a ring of z-plane slots.  The slices of a group are consumed in SWEEPS sweeps over
  // kq = 0..N-1 (pre-pass, main pass; the diagonal pass sweeps once per component); the producer cursor runs RING
  // slices ahead of the consumer across sweeps and groups.
  constexpr int SWEEPS = (Op::PRE ? 1 : 0) + (DIAG ? NC : 1);
  int islot = 0, ikq = 0, isw = 0;  // producer cursor (uniform over the group)
  int64_t ig = group;
  int cslot = 0;                    // consumer
  unsigned cpar = 0;
  auto ring_issue = [&]() {
    if (SLICED && staged && ig < A.n_groups) {
      if (lane == 0) {
        uint64_t *mb = mbar + islot;
        T *slot = qs + (size_t)islot * KS_::SLOT_ELEMS;
        mbar_expect_tx(mb, (unsigned)(((Op::NF_IN ? KS_::SL1_ELEMS : 0) + (two ? KS_::SL2_ELEMS : 0)) * sizeof(T)));
        if (Op::NF_IN)
          bulk_g2s(slot, oa.stage + (size_t)ig * Ly::table_stride(Op::NF_IN ? Op::NF_IN : 1) + (size_t)ikq * KS_::SL1_ELEMS,
                   (unsigned)(KS_::SL1_ELEMS * sizeof(T)), mb);
        if (two)
          bulk_g2s(slot + KS_::SL1_ELEMS, oa.stage2 + (size_t)ig * Ly::table_stride(Op::NF_IN2 ? Op::NF_IN2 : 1) + (size_t)ikq * KS_::SL2_ELEMS,
                   (unsigned)(KS_::SL2_ELEMS * sizeof(T)), mb);
      }
      islot = (islot + 1 == RING) ? 0 : islot + 1;
      if (++ikq == N) {
        ikq = 0;
        if (++isw == SWEEPS) {
          isw = 0;
          ig += gstep;
        }
      }
    }
  };
  // this q-plane's entries of the staged tables (field 0, this lane); q_release after the last read of the plane
  const T *q1p = nullptr, *q2p = nullptr;
  auto q_acquire = [&](int kq) {
    if constexpr (SLICED) {
      if (staged) mbar_wait(mbar + cslot, cpar);
      q1p = qs + (size_t)cslot * KS_::SLOT_ELEMS + lane;
      q2p = q1p + KS_::SL1_ELEMS;
    } else {
      q1p = qs + kq * L + lane;
      q2p = qs2 + kq * L + lane;
    }
  };
  auto q_release = [&]() {
    if constexpr (SLICED) {
      if (staged) {
        // every column thread is done with the slot before it is refilled.  (Per-slot "empty" mbarriers - only the
        // producer lane waits - measured the same on B200 (0.542 vs 0.548 of HBM at Q4) and compute-sanitizer's racecheck
        // cannot follow that protocol; the group barrier is the one it verifies: profiles/r02_racecheck.txt.)
#if B200MF_EMPTY_MBAR
        __syncwarp();
        if constexpr (!WM) {
          if ((lane & 31) == 0) mbar_arrive(mbar + RING + cslot);
          if (lane == 0) mbar_wait(mbar + RING + cslot, cpar);
        }
#else
        group_sync<WM>();
#endif
        if (lane == 0) fence_proxy_async();
        ring_issue();
        if (++cslot == RING) {
          cslot = 0;
          cpar ^= 1u;
        }
      }
    }
  };

  auto load_nid = [&](int64_t g, uint32_t(&id)[N]) {
#pragma unroll
    for (int k = 0; k < N; ++k) id[k] = active ? __ldg(A.nidx + ((size_t)g * N + k) * L + lane) : NODE_ALL_CONSTRAINED;
  };
  auto gather = [&](const uint32_t(&id)[N], T(&w)[NC][N]) {
#pragma unroll
    for (int k = 0; k < N; ++k) {
      const T *base = A.src + (id[k] & NODE_MASK);
#pragma unroll
      for (int c = 0; c < NC; ++c) {
        const bool skip = Op::PLAIN ? !active : (id[k] & node_flag(c)) != 0u;
        w[c][k] = skip ? T(0) : __ldg(base + (size_t)c * A.cstride);
      }
    }
  };
  // the same gather, asynchronously into ubuf[c][k][lane]: every thread later reads back only what it requested itself
  // (after cp.async.wait_group), so no barrier is involved
  auto gather_async = [&](const uint32_t(&id)[N]) {
    if (active) {
#pragma unroll
      for (int k = 0; k < N; ++k) {
        const T *base = A.src + (id[k] & NODE_MASK);
#pragma unroll
        for (int c = 0; c < NC; ++c)
          cp_async_elem(ubuf + (c * N + k) * L + lane, base + (size_t)c * A.cstride, Op::PLAIN || (id[k] & node_flag(c)) == 0u);
      }
    }
    cp_async_commit();
  };
  // ... or only warm the L2 with the entries the next group will gather
  constexpr bool PG = Ly::GATHER_MODE == 2 && Op::INTEGRATE && !DIAG;
  // ... and (ops without a pre-pass) run F1 of the next group at the end of the current iteration: the gathered values are
  // loaded before B1, so that the L2 latency hides behind it, and consumed right after it - no registers are carried
  // around the loop
  constexpr bool PGE = PG && !Op::PRE && B200MF_PGE;  // (measured slower on B200 at N = 5: 0.515 vs 0.548 of HBM; off by default)
  auto gather_prefetch = [&](const uint32_t(&id)[N]) {
    if (active) {
#pragma unroll
      for (int k = 0; k < N; ++k) {
        const T *base = A.src + (id[k] & NODE_MASK);
#pragma unroll
        for (int c = 0; c < NC; ++c) prefetch_l2(base + (size_t)c * A.cstride);
      }
    }
  };

  // column-thread offsets into the plane buffers
  const int zoff = Ly::zplane(cw, 0, 0) + ji;
  const int qoff = Ly::qplane(cw, 0, 0, 0) + ji;
  constexpr int ZC = Ly::ALIAS_Z ? Ly::QCOMP : Ly::ZCOMP;
  auto zidx = [&](int c, int kq) { return zoff + c * ZC + kq * PL; };
  auto qidx = [&](int c, int f, int kq) { return qoff + c * Ly::QCOMP + f * Ly::QFIELD + kq * PL; };

  // diagonal pass: this thread's columns of S and D (row index q) for its j (y sweep) and i (x sweep)
  T sy[DIAG ? N : 1], dy[DIAG ? N : 1], sx[DIAG ? N : 1], dx[DIAG ? N : 1];
  if constexpr (DIAG) {
#pragma unroll
    for (int q = 0; q < N; ++q) {
      sy[q] = Sh::S(q, ji / N);
      dy[q] = Sh::D(q, ji / N);
      sx[q] = Sh::S(q, ji % N);
      dx[q] = Sh::D(q, ji % N);
    }
  }

  unsigned parity = 0;
  uint32_t nid[N], nid_next[N];
  T u[NC][N];
#pragma unroll
  for (int k = 0; k < N; ++k) nid[k] = nid_next[k] = NODE_ALL_CONSTRAINED;
  if (group < A.n_groups) {
    if constexpr (SLICED) {
#pragma unroll
      for (int r = 0; r < RING; ++r) ring_issue();
    } else {
      stage(group);
    }
    load_nid(group, nid);
    if constexpr (AG)
      gather_async(nid);
    else if (PG)
      gather_prefetch(nid);
    else if (!DIAG)
      gather(nid, u);
  }

  double dot_acc = 0.0;  // DOT: the fused p . Ap of SolverCG (a separate instantiation: the plain vmult pays nothing)
  bool first_group = true;
  for (; group < A.n_groups; group += gstep) {
    const int64_t gnext = group + gstep;
    const bool has_next = gnext < A.n_groups;
    typename Op::CellState st;
    Op::cell_begin(oa, group * Ly::CPG + cw, active, lane, st);

    if constexpr (Op::STORE_DOFS) Op::store_dofs(oa, group, lane, active, u);
    constexpr int NPASS = 1;
    constexpr int FIRST = Op::PRE ? -1 : 0;
#pragma unroll 1
    for (int pass = FIRST; pass < NPASS; ++pass) {
      const bool pre = Op::PRE && pass < 0;
      if constexpr (DIAG) {
        if (!pre) {
          // ---- diagonal pass ---------------------------------------------------------------------
          if (pass == FIRST) {
            if (has_next) load_nid(gnext, nid_next);
            if (!SLICED && staged) mbar_wait(mbar, parity);
          }
          T *db = bufZ + cw * (6 * Ly::N3);  // [field][k3][j][i] of this thread's cell
          const int tj = ji / N, ti = ji % N;
#pragma unroll 1
          for (int c = 0; c < NC; ++c) {
            T zc[6][N];
#pragma unroll
            for (int f = 0; f < 6; ++f)
#pragma unroll
              for (int k = 0; k < N; ++k) zc[f][k] = T(0);
#pragma unroll 1
            for (int kq = 0; kq < N; ++kq) {
              T Q[4][4];  // Q[a][b]: output a (value, d/dxi_1..3 of the test side) for unit input b
              q_acquire(kq);
#pragma unroll
              for (int b = 0; b < 4; ++b) {
                T v[NC], G[NC][3], ov[NC], oG[NC][3];
#pragma unroll
                for (int cc = 0; cc < NC; ++cc) {
                  v[cc] = (cc == c && b == 0) ? T(1) : T(0);
#pragma unroll
                  for (int d = 0; d < 3; ++d) G[cc][d] = (cc == c && b == d + 1) ? T(1) : T(0);
                }
                Op::quad(oa, st, active, q1p, q2p, scr, group, kq, lane, v, G, ov, oG);
#pragma unroll
                for (int a = 0; a < 4; ++a) {
                  T o = T(0);
#pragma unroll
                  for (int cc = 0; cc < NC; ++cc) o += (cc == c) ? (a == 0 ? ov[cc] : oG[cc][a - 1]) : T(0);
                  Q[a][b] = o;
                }
              }
              q_release();
#pragma unroll
              for (int k = 0; k < N; ++k) {
                const T sv = Sh::S(kq, k), dv = Sh::D(kq, k);
                const T ss = sv * sv, sd = sv * dv, dd = dv * dv;
                zc[0][k] += ss * Q[0][0] + dd * Q[3][3] + sd * (Q[0][3] + Q[3][0]);
                zc[1][k] += ss * (Q[0][1] + Q[1][0]) + sd * (Q[1][3] + Q[3][1]);
                zc[2][k] += ss * (Q[0][2] + Q[2][0]) + sd * (Q[2][3] + Q[3][2]);
                zc[3][k] += ss * Q[1][1];
                zc[4][k] += ss * Q[2][2];
                zc[5][k] += ss * (Q[1][2] + Q[2][1]);
              }
            }
            if (active) {
#pragma unroll
              for (int f = 0; f < 6; ++f)
#pragma unroll
                for (int k = 0; k < N; ++k) db[(f * N + k) * N2 + ji] = zc[f][k];
            }
            group_sync<WM>();
            // y sweep: thread (k2 = tj, i) contracts over q2; x types left: 0 S.S, 1 S.D, 2 D.D
            T yc[3][N];
#pragma unroll
            for (int k = 0; k < N; ++k) yc[0][k] = yc[1][k] = yc[2][k] = T(0);
            if (active) {
#pragma unroll
              for (int q2 = 0; q2 < N; ++q2) {
                const T sv = sy[q2], dv = dy[q2];
                const T ss = sv * sv, sd = sv * dv, dd = dv * dv;
#pragma unroll
                for (int k = 0; k < N; ++k) {
                  const T *z = db + k * N2 + q2 * N + ti;
                  yc[0][k] += ss * z[0 * Ly::N3] + sd * z[2 * Ly::N3] + dd * z[4 * Ly::N3];
                  yc[1][k] += ss * z[1 * Ly::N3] + sd * z[5 * Ly::N3];
                  yc[2][k] += ss * z[3 * Ly::N3];
                }
              }
            }
            group_sync<WM>();
            if (active) {
#pragma unroll
              for (int f = 0; f < 3; ++f)
#pragma unroll
                for (int k = 0; k < N; ++k) db[(f * N + k) * N2 + ji] = yc[f][k];
            }
            group_sync<WM>();
            // x sweep: thread (k2, k1 = ti) contracts over q1, then adds its N diagonal entries of component c
            if (active) {
              T out[N];
#pragma unroll
              for (int k = 0; k < N; ++k) out[k] = T(0);
#pragma unroll
              for (int q1 = 0; q1 < N; ++q1) {
                const T sv = sx[q1], dv = dx[q1];
                const T ss = sv * sv, sd = sv * dv, dd = dv * dv;
#pragma unroll
                for (int k = 0; k < N; ++k) {
                  const T *y = db + k * N2 + tj * N + q1;
                  out[k] += ss * y[0 * Ly::N3] + sd * y[1 * Ly::N3] + dd * y[2 * Ly::N3];
                }
              }
#pragma unroll
              for (int k = 0; k < N; ++k)
                if (!(nid[k] & (1u << (29 + c)))) atomicAdd(A.dst + (size_t)c * A.cstride + (nid[k] & NODE_MASK), out[k]);
            }
            group_sync<WM>();
          }
          if constexpr (!SLICED) {
            parity ^= 1u;
            if (has_next) {
              fence_proxy_async();
              stage(gnext);
            }
          }
          continue;
        }
      }
      // ---- F1 (column): z sweep of the nodal column, planes to bufZ ---------------------------
      // (PGE: done at the end of the previous iteration for every group but the block's first one)
      if (!(PGE && !first_group))
#pragma unroll
      for (int c = 0; c < NC; ++c) {
        T o[N];
        if (pre) {
          T up[N];
          Op::load_pre(oa, group, lane, active, c, nid, up);
          sweep_S<T, N, false>(up, o);
        } else {
          if constexpr (AG) {
            if (c == 0) cp_async_wait_all();
#pragma unroll
            for (int k = 0; k < N; ++k) u[c][k] = active ? ubuf[(c * N + k) * L + lane] : T(0);
          }
          if constexpr (PG) {
            if (c == 0) gather(nid, u);  // L2 hits after the prefetch (PGE: the block's first group only)
          }
          sweep_S<T, N, false>(u[c], o);
        }
#pragma unroll
        for (int kq = 0; kq < N; ++kq)
          if (active) bufZ[zidx(c, kq)] = o[kq];
      }
      if (pass == FIRST && has_next) load_nid(gnext, nid_next);
      group_sync<WM>();
      // ---- F2 (plane tasks): x, y sweeps; values, d/dx, d/dy at the q-points of the plane ------
#pragma unroll 1
      for (int t0 = 0; t0 < Ly::PLANE_LANES; t0 += Ly::TPG) {
        const int t = t0 + lane;
        if (Ly::task_valid(t)) {
          const int tcw = Ly::task_cell(t), tc = Ly::task_comp(t), tk = Ly::task_plane(t);
          const T *zp = bufZ + Ly::zplane(tcw, tc, tk);
          T a[N][N];  // a[j][i]
#pragma unroll
          for (int j = 0; j < N; ++j)
#pragma unroll
            for (int i = 0; i < N; ++i) a[j][i] = zp[j * N + i];
          T b[N][N];  // x sweep: b[j][qx]
#pragma unroll
          for (int j = 0; j < N; ++j) sweep_S<T, N, false>(a[j], b[j]);
          // y sweep: a[qy][qx]
#pragma unroll
          for (int qx = 0; qx < N; ++qx) {
            T col[N], o[N];
#pragma unroll
            for (int j = 0; j < N; ++j) col[j] = b[j][qx];
            sweep_S<T, N, false>(col, o);
#pragma unroll
            for (int qy = 0; qy < N; ++qy) a[qy][qx] = o[qy];
          }
          T *qv = bufQ + Ly::qplane(tcw, tc, 0, tk), *qx_ = bufQ + Ly::qplane(tcw, tc, 1, tk), *qy_ = bufQ + Ly::qplane(tcw, tc, 2, tk);
#pragma unroll
          for (int qy = 0; qy < N; ++qy) {
            T gx[N];
            sweep_C<T, N>(a[qy], gx);
#pragma unroll
            for (int qx = 0; qx < N; ++qx) {
              qv[qy * N + qx] = a[qy][qx];
              qx_[qy * N + qx] = gx[qx];
            }
          }
#pragma unroll
          for (int qx = 0; qx < N; ++qx) {
            T col[N], gy[N];
#pragma unroll
            for (int qy = 0; qy < N; ++qy) col[qy] = a[qy][qx];
            sweep_C<T, N>(col, gy);
#pragma unroll
            for (int qy = 0; qy < N; ++qy) qy_[qy * N + qx] = gy[qy];
          }
        }
      }
      if (!SLICED && pass == FIRST && staged) mbar_wait(mbar, parity);  // this group's tables have landed in shared memory
      group_sync<WM>();
      // ---- F3 (column): d/dz in registers, quadrature loop, z part of the integration ---------
      T racc[NC][N];  // values + z-derivative terms of the test side, at the q-points of the column
      {
        // the values of the column (for d/dz): registers, or - rolled loop - re-read from shared memory in every plane
        // (N^2 instead of N loads per component, 2 N NC fewer live registers)
        constexpr bool VAL_REGS = !(SLICED && B200MF_ROLLQ) || B200MF_VAL_REGS;
        // sliced kernels (multi-warp groups), B200MF_GUARD_IDLE: a lane without a column only takes part in the slot hand-over
        // instead of mirroring lane 0's shared-memory reads
        constexpr bool GUARD_IDLE = SLICED && (B200MF_GUARD_IDLE == 1 || (B200MF_GUARD_IDLE == -1 && N == 6));
        const bool work = !GUARD_IDLE || active;
        T val[NC][N];
#pragma unroll
        for (int c = 0; c < NC; ++c) {
#pragma unroll
          for (int kq = 0; kq < N; ++kq) {
            if (VAL_REGS) val[c][kq] = work ? bufQ[qidx(c, 0, kq)] : T(0);
            racc[c][kq] = T(0);
          }
        }
        // one z-plane of quadrature points; STATIC: kq is a compile-time constant (unrolled loop), else the loop is
        // rolled (sliced tables: one small body, no load hoisting across planes -> far fewer registers) and whatever is
        // indexed by kq comes from shared / constant memory
        auto plane = [&](const int kq, auto STATIC) {
          T G[NC][3], v[NC], ov[NC], oG[NC][3];
          T crow[N];
#pragma unroll
          for (int m = 0; m < N; ++m) crow[m] = Sh::C(kq, m);
          if (work) {
#pragma unroll
            for (int c = 0; c < NC; ++c) {
              if constexpr (decltype(STATIC)::value)
                v[c] = val[c][kq];
              else
                v[c] = bufQ[qidx(c, 0, kq)];
              G[c][0] = bufQ[qidx(c, 1, kq)];
              G[c][1] = bufQ[qidx(c, 2, kq)];
              T acc = T(0);
#pragma unroll
              for (int m = 0; m < N; ++m) acc += crow[m] * (VAL_REGS ? val[c][m] : bufQ[qidx(c, 0, m)]);
              G[c][2] = acc;
            }
          }
          q_acquire(kq);
          if (!work) {
            // hand-over only
          } else if (pre) {
            Op::quad_pre(oa, st, active, q1p, scr, group, kq, lane, v, G);
          } else {
            Op::quad(oa, st, active, q1p, q2p, scr, group, kq, lane, v, G, ov, oG);
            if constexpr (DOT) {
              // src . (A src) = sum over cells and q-points of (value, gradient) . (submitted value, gradient): every
              // operand is live here, so the fused dot product costs no registers beyond its accumulator
              T e = T(0);
#pragma unroll
              for (int c = 0; c < NC; ++c) e += v[c] * ov[c] + G[c][0] * oG[c][0] + G[c][1] * oG[c][1] + G[c][2] * oG[c][2];
              dot_acc += (double)e;
            }
            if (Op::INTEGRATE) {
#pragma unroll
              for (int c = 0; c < NC; ++c) {
#pragma unroll
                for (int m = 0; m < N; ++m) racc[c][m] += crow[m] * oG[c][2];
                if constexpr (decltype(STATIC)::value) racc[c][kq] += ov[c];
                if (active) {
                  bufQ[qidx(c, 1, kq)] = oG[c][0];
                  bufQ[qidx(c, 2, kq)] = oG[c][1];
                }
              }
            }
          }
          if constexpr (!decltype(STATIC)::value) {
            if (Op::INTEGRATE && !pre && work) {  // racc[c][kq] += ov[c] with a run-time kq: one uniform branch instead of N selects per component
              switch (kq) {
#define B2_OV_CASE(M)                                     \
  case M:                                                 \
    if constexpr (M < N) {                                \
      _Pragma("unroll") for (int c = 0; c < NC; ++c) racc[c][M] += ov[c]; \
    }                                                     \
    break;
                B2_OV_CASE(0) B2_OV_CASE(1) B2_OV_CASE(2) B2_OV_CASE(3) B2_OV_CASE(4) B2_OV_CASE(5) B2_OV_CASE(6)
#undef B2_OV_CASE
                default: break;
              }
            }
          }
          q_release();
        };
        if constexpr (SLICED && B200MF_ROLLQ) {
#pragma unroll 1
          for (int kq = 0; kq < N; ++kq) plane(kq, std::false_type{});
        } else {
#pragma unroll
          for (int kq = 0; kq < N; ++kq) plane(kq, std::true_type{});
        }
      }
      if (Op::INTEGRATE && !pre) {
#pragma unroll
        for (int c = 0; c < NC; ++c)
#pragma unroll
          for (int kq = 0; kq < N; ++kq)
            if (active) bufQ[qidx(c, 0, kq)] = racc[c][kq];
      }
      group_sync<WM>();  // q-point results are in bufQ; everyone is done with the staged tables
      if constexpr (!SLICED) {
        if (pass == NPASS - 1) {
          parity ^= 1u;
          if (has_next) {
            fence_proxy_async();
            stage(gnext);
          }
        }
      }
      if (Op::INTEGRATE && !pre) {
        // issue the next group's gather now; it completes under the integration half below
        T u_next[(AG || PG) ? 1 : NC][(AG || PG) ? 1 : N];
        if constexpr (AG) {
          if (has_next) gather_async(nid_next);  // F1 of this group has consumed ubuf (same thread)
        } else if constexpr (PG) {
          if (has_next) gather_prefetch(nid_next);
        } else {
          if (!DIAG && has_next) gather(nid_next, u_next);
        }
        // ---- B2 (plane tasks): transposed x/y derivative and x, y sweeps ------------------------
#pragma unroll 1
        for (int t0 = 0; t0 < Ly::PLANE_LANES; t0 += Ly::TPG) {
          const int t = t0 + lane;
          if (Ly::task_valid(t)) {
            const int tcw = Ly::task_cell(t), tc = Ly::task_comp(t), tk = Ly::task_plane(t);
            const T *qv = bufQ + Ly::qplane(tcw, tc, 0, tk), *qx_ = bufQ + Ly::qplane(tcw, tc, 1, tk), *qy_ = bufQ + Ly::qplane(tcw, tc, 2, tk);
            T a[N][N];  // a[qy][qx]
#pragma unroll
            for (int qy = 0; qy < N; ++qy) {
              T gx[N];
#pragma unroll
              for (int qx = 0; qx < N; ++qx) {
                a[qy][qx] = qv[qy * N + qx];
                gx[qx] = qx_[qy * N + qx];
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
              for (int qy = 0; qy < N; ++qy) gy[qy] = qy_[qy * N + qx];
#pragma unroll
              for (int qy = 0; qy < N; ++qy)
#pragma unroll
                for (int m = 0; m < N; ++m) a[qy][qx] += Sh::C(m, qy) * gy[m];
            }
            T b[N][N];  // x sweep back: b[qy][i]
#pragma unroll
            for (int qy = 0; qy < N; ++qy) sweep_S<T, N, true>(a[qy], b[qy]);
            T *zp = bufZ + Ly::zplane(tcw, tc, tk);
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
        group_sync<WM>();
        // prefetch mode: the next group's entries are (about to be) in the L2 since the start of B2; load them now, so
        // that the L2 latency hides behind B1 and the registers are only held across this light phase
        T u_pg[PGE ? NC : 1][PGE ? N : 1];
        if constexpr (PGE) {
          if (has_next) gather(nid_next, u_pg);
        }
        // ---- B1 (column): z sweep back + scatter-add -------------------------------------------
        if (!SLICED || active)
#pragma unroll
        for (int c = 0; c < NC; ++c) {
          T col[N], o[N];
#pragma unroll
          for (int kq = 0; kq < N; ++kq) col[kq] = bufZ[zidx(c, kq)];
          sweep_S<T, N, true>(col, o);
#pragma unroll
          for (int k = 0; k < N; ++k) {
            if (!(nid[k] & node_flag(c)))
              atomicAdd(A.dst + (size_t)c * A.cstride + (nid[k] & NODE_MASK), o[k]);
          }
        }
        if constexpr (!AG && !PG) {
          if (!DIAG && has_next) {
#pragma unroll
            for (int c = 0; c < NC; ++c)
#pragma unroll
              for (int k = 0; k < N; ++k) u[c][k] = u_next[c][k];
          }
        }
        if constexpr (PGE) {  // F1 of the next group (own columns of bufZ, just read above by this very thread)
          if (has_next) {
#pragma unroll
            for (int c = 0; c < NC; ++c) {
              T o[N];
              sweep_S<T, N, false>(u_pg[c], o);
#pragma unroll
              for (int kq = 0; kq < N; ++kq)
                if (active) bufZ[zidx(c, kq)] = o[kq];
            }
          }
        }
        // the next F1 (next group; a PRE op has no pass after this one) only overwrites the positions of bufZ this very
        // thread has just read: with the z-swept planes aliased into field 0 no barrier is needed before it
        if constexpr (!(Ly::ALIAS_Z && !Op::PRE)) group_sync<WM>();  // bufZ is free for the next pass / group
      }
    }  // pass
    if (!Op::INTEGRATE && has_next) gather(nid_next, u);  // evaluation-only ops (cache): plain prefetch
#pragma unroll
    for (int k = 0; k < N; ++k) nid[k] = nid_next[k];
    first_group = false;
  }
  if constexpr (DOT) {  // one atomic per warp and launch
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) dot_acc += __shfl_down_sync(0xffffffffu, dot_acc, o);
    if ((threadIdx.x & 31) == 0 && dot_acc != 0.0) atomicAdd(A.dot, dot_acc);
  }
}

}  // namespace b200mf
