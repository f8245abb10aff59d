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
//    kernel reads them ([group][field][x-index][lane]); a group's table is one
//    contiguous block that the TMA unit (cp.async.bulk + mbarrier) copies into
//    shared memory one group ahead of the arithmetic, and the DoF gather of the
//    next group is issued into registers before the transpose half of the
//    current one: HBM latency is off the critical path (the first profile,
//    profiles/r01_v1_*, showed 60 % of the stalls on the long scoreboard).
//  * vectors are stored in component blocks (DoF (node,c) at c*cstride+node): a
//    row of N nodes of one component is one contiguous run, which halves the
//    sectors per gather / red.global.add compared with interleaved components.
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
// Thread / shared-memory layout.  Strides from tools/smem_layout_search.py (64-bit accesses are served
// per half-warp over 16 8-byte banks: confirmed by the ncu source page, profiles/r01_*: only the
// Z-orientation accesses of even N pay 2x, 7.8% excess wavefronts in total).
template <int WORDS, int N>
struct SmemPad;  // JS, KPAD (KS = N*JS + KPAD), CELLPAD for 3 components / 1 component
#define B2_PAD(W, N_, JS_, KP_, C3_, C1_) \
  template <> struct SmemPad<W, N_> { static constexpr int JS = JS_, KPAD = KP_, CPAD3 = C3_, CPAD1 = C1_; };
B2_PAD(2, 2, 3, 0, 0, 0)
B2_PAD(2, 3, 3, 0, 8, 10)
B2_PAD(2, 4, 5, 0, 0, 0)
B2_PAD(2, 5, 5, 0, 0, 0)
B2_PAD(2, 6, 7, 0, 0, 0)
B2_PAD(2, 7, 7, 0, 0, 0)
B2_PAD(1, 2, 3, 0, 0, 0)
B2_PAD(1, 3, 3, 0, 8, 10)
B2_PAD(1, 4, 5, 0, 0, 0)
B2_PAD(1, 5, 5, 0, 0, 0)
B2_PAD(1, 6, 7, 0, 0, 0)
B2_PAD(1, 7, 7, 0, 0, 0)
#undef B2_PAD

__host__ __device__ constexpr size_t round_up(size_t x, size_t m) { return (x + m - 1) / m * m; }

template <int N, int NC, int WORDS = 2>
struct Layout {
  static constexpr int N2 = N * N, N3 = N * N * N;
  static constexpr bool WARP_MODE = (N2 <= 32);
  static constexpr int CPG = WARP_MODE ? 32 / N2 : 1;                 // cells per group
  static constexpr int L = CPG * N2;                                  // dense lanes per group
  static constexpr int TPG = WARP_MODE ? 32 : ((N2 + 31) / 32) * 32;  // threads per group
  static constexpr int JS = SmemPad<WORDS, N>::JS;
  static constexpr int KS = N * JS + SmemPad<WORDS, N>::KPAD;
  static constexpr int CSTR = N * KS;  // one component of one buffer
  static constexpr int NBUF = 3;
  static constexpr int CELLPAD = (NC == 3) ? SmemPad<WORDS, N>::CPAD3 : SmemPad<WORDS, N>::CPAD1;
  static constexpr int CELL = NBUF * NC * CSTR + CELLPAD;
  static constexpr int GROUP_ELEMS = CPG * CELL;
  // elements between consecutive groups of a [group][NF][N][L] table (16-byte multiple for cp.async.bulk)
  __host__ __device__ static constexpr size_t table_stride(int NF) { return round_up((size_t)NF * N * L, 16 / (4 * WORDS)); }
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
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
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
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

template <typename T>
struct CellArgs {
  int64_t group_begin, n_groups;  // groups [group_begin, n_groups)
  const uint32_t *nidx;  // [group][k][lane], flags in bits 29..31
  const T *src;
  T *dst;
  int64_t cstride;  // vector layout: DoF (node, c) lives at c*cstride + node ("component blocks")
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
//   static constexpr int  NF_IN, NF_IN2; fields of the (up to two) input tables staged through shared memory
//   struct Args { const T *stage, *stage2; ... }   stage2 may be null at run time
//   struct CellState;                    per-thread per-cell state (e.g. material constants)
//   static void cell_begin(args, cell, valid, CellState&)
//   static void quad(args, state, active, qs, qs2, scr, group, iq, lane, v[NC], G[NC][3], ov[NC], oG[NC][3])
//       qs / qs2: shared-memory copies of this group's tables, field f of q-point (iq, lane) at
//       qs[(f*N + iq)*L + lane]; scr: NF_SCRATCH fields of per-group scratch in the same layout
//   static constexpr bool PRE, STORE_DOFS; int NF_SCRATCH  (see the pass loop)
//
// DIAG = true turns the loop into MatrixFreeTools::compute_diagonal: the cell
// operator is applied to every local unit vector and entry i of the result is kept.
template <typename T, int N, class Op>
struct KernelSmem {
  using Ly = Layout<N, Op::NC, sizeof(T) / 4>;
  static constexpr size_t BUF_BYTES = round_up((size_t)Ly::GROUP_ELEMS * sizeof(T), 16);
  static constexpr size_t ST1_BYTES = Ly::table_stride(Op::NF_IN) * sizeof(T);
  static constexpr size_t ST2_BYTES = Op::NF_IN2 ? Ly::table_stride(Op::NF_IN2) * sizeof(T) : 0;
  static constexpr size_t SCR_BYTES = Op::NF_SCRATCH ? Ly::table_stride(Op::NF_SCRATCH) * sizeof(T) : 0;  // pre-pass -> main pass
  static constexpr size_t GROUP_BYTES = BUF_BYTES + ST1_BYTES + ST2_BYTES + SCR_BYTES + 16;  // + mbarrier
};

template <typename T, int N, class Op, int GPB, bool DIAG = false, int MINB = 1>
__global__ void __launch_bounds__(Layout<N, Op::NC>::TPG *GPB, MINB)
    cell_loop_kernel(const CellArgs<T> A, const typename Op::Args oa) {
  constexpr int NC = Op::NC;
  using Ly = Layout<N, NC, sizeof(T) / 4>;
  using Sh = Shape<T, N>;
  using KS_ = KernelSmem<T, N, Op>;
  constexpr int N2 = Ly::N2, JS = Ly::JS, KS = Ly::KS, CSTR = Ly::CSTR, L = Ly::L;
  constexpr bool WM = Ly::WARP_MODE;
  static_assert(WM || GPB == 1, "block mode runs one group per block");

  extern __shared__ __align__(128) unsigned char smem_raw[];

  const int gib = threadIdx.x / Ly::TPG;
  const int lane = threadIdx.x % Ly::TPG;
  const bool active = lane < L;
  const int cw = active ? lane / N2 : 0;
  const int t = active ? lane % N2 : 0;
  const int a = t % N, b = t / N;

  unsigned char *gbase = smem_raw + (size_t)gib * KS_::GROUP_BYTES;
  T *bufA = reinterpret_cast<T *>(gbase) + (size_t)cw * Ly::CELL;
  T *bufB = bufA + NC * CSTR;
  T *bufC = bufB + NC * CSTR;
  T *qs = reinterpret_cast<T *>(gbase + KS_::BUF_BYTES);
  T *qs2 = reinterpret_cast<T *>(gbase + KS_::BUF_BYTES + KS_::ST1_BYTES);
  T *scr = reinterpret_cast<T *>(gbase + KS_::BUF_BYTES + KS_::ST1_BYTES + KS_::ST2_BYTES);
  uint64_t *mbar = reinterpret_cast<uint64_t *>(gbase + KS_::BUF_BYTES + KS_::ST1_BYTES + KS_::ST2_BYTES + KS_::SCR_BYTES);

  if (lane == 0) mbar_init(mbar, 1);
  group_sync<WM>();

  // stage the tables of `group` into this group's shared buffers (one elected lane drives the TMA unit)
  auto stage = [&](int64_t g) {
    if (lane == 0) {
      const bool two = Op::NF_IN2 != 0 && oa.stage2 != nullptr;
      mbar_expect_tx(mbar, (unsigned)(KS_::ST1_BYTES + (two ? KS_::ST2_BYTES : 0)));
      bulk_g2s(qs, oa.stage + (size_t)g * Ly::table_stride(Op::NF_IN), (unsigned)KS_::ST1_BYTES, mbar);
      if (two) bulk_g2s(qs2, oa.stage2 + (size_t)g * Ly::table_stride(Op::NF_IN2 ? Op::NF_IN2 : 1), (unsigned)KS_::ST2_BYTES, mbar);
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

  const int oZ = b * JS + a;       // + k*KS : (i=a, j=b) column along z
  const int oY = b * KS + a;       // + j*JS : (i=a, k=b) column along y
  const int oX = b * KS + a * JS;  // + i    : (j=a, k=b) column along x

  const int64_t gstep = (int64_t)gridDim.x * GPB;
  int64_t group = A.group_begin + (int64_t)blockIdx.x * GPB + gib;
  unsigned parity = 0;
  uint32_t nid[N], nid_next[N];
  T u[NC][N];
#pragma unroll
  for (int k = 0; k < N; ++k) nid[k] = nid_next[k] = NODE_ALL_CONSTRAINED;
  if (group < A.n_groups) {
    stage(group);
    load_nid(group, nid);
    if (!DIAG) gather(nid, u);
  }

  for (; group < A.n_groups; group += gstep) {
    const int64_t gnext = group + gstep;
    const bool has_next = gnext < A.n_groups;
    typename Op::CellState st;
    Op::cell_begin(oa, group * Ly::CPG + cw, active, st);

    T dacc[NC][N];
    if (DIAG) {
#pragma unroll
      for (int c = 0; c < NC; ++c)
#pragma unroll
        for (int k = 0; k < N; ++k) dacc[c][k] = T(0);
    }
    if constexpr (Op::STORE_DOFS) Op::store_dofs(oa, group, lane, active, u);
    constexpr int NPASS = DIAG ? NC * Ly::N3 : 1;
    // Op::PRE: one extra evaluation-only pass (pass -1) over a second cell-local field (Op::load_pre) whose
    // quadrature-point results (Op::quad_pre) stay in shared memory for the main pass(es)
    constexpr int FIRST = Op::PRE ? -1 : 0;
#pragma unroll 1
    for (int pass = FIRST; pass < NPASS; ++pass) {
      const bool pre = Op::PRE && pass < 0;
      const int pc = pass / Ly::N3, pk = (pass / N2) % N, pt = pass % N2;  // unit vector (component, k, column)
      // ---- phase 1 (Z orientation): z sweep of the gathered values ----------
      if (DIAG) {
#pragma unroll
        for (int k = 0; k < N; ++k)
#pragma unroll
          for (int c = 0; c < NC; ++c) u[c][k] = (!(nid[k] & node_flag(c)) && c == pc && k == pk && t == pt) ? T(1) : T(0);
      }
#pragma unroll
      for (int c = 0; c < NC; ++c) {
        T o[N];
        if (pre) {
          T up[N];
          Op::load_pre(oa, group, lane, active, c, up);
          sweep_S<T, N, false>(up, o);
        } else {
          sweep_S<T, N, false>(u[c], o);
        }
#pragma unroll
        for (int k = 0; k < N; ++k)
          if (active) bufA[c * CSTR + k * KS + oZ] = o[k];
      }
      if (pass == FIRST && has_next) load_nid(gnext, nid_next);
      group_sync<WM>();
      // ---- phase 2 (Y orientation) -------------------------------------------
#pragma unroll
      for (int c = 0; c < NC; ++c) {
        T col[N], o[N];
#pragma unroll
        for (int j = 0; j < N; ++j) col[j] = bufA[c * CSTR + j * JS + oY];
        sweep_S<T, N, false>(col, o);
#pragma unroll
        for (int j = 0; j < N; ++j)
          if (active) bufB[c * CSTR + j * JS + oY] = o[j];
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
        for (int i = 0; i < N; ++i)
          if (active) bufA[c * CSTR + oX + i] = val[c][i];
      }
      if (pass == FIRST) mbar_wait(mbar, parity);  // this group's tables have landed in shared memory
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
        if (pre) {
          Op::quad_pre(oa, st, active, qs, scr, group, iq, lane, G);
          continue;
        }
        Op::quad(oa, st, active, qs, qs2, scr, group, iq, lane, v, G, ov, oG);
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
      group_sync<WM>();  // everyone is done with bufA (values) and with the staged tables
      if (pass == NPASS - 1) {
        parity ^= 1u;
        if (has_next) {
          fence_proxy_async();
          stage(gnext);
        }
      }
      if (Op::INTEGRATE && !pre) {
        // issue the next group's gather now; it completes under the transpose half below
        T u_next[NC][N];
        if (!DIAG && has_next) gather(nid_next, u_next);
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
          for (int i = 0; i < N; ++i)
            if (active) bufA[c * CSTR + oX + i] = o[i];
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
          for (int j = 0; j < N; ++j)
            if (active) bufB[c * CSTR + j * JS + oY] = o[j];
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
              atomicAdd(A.dst + (size_t)c * A.cstride + (nid[k] & NODE_MASK), o[k]);
          }
        }
        if (!DIAG && has_next) {
#pragma unroll
          for (int c = 0; c < NC; ++c)
#pragma unroll
            for (int k = 0; k < N; ++k) u[c][k] = u_next[c][k];
        }
        // the next pass's / group's first barrier orders these bufB reads before its phase-2 writes
      }
    }  // pass
    if (DIAG) {
#pragma unroll
      for (int c = 0; c < NC; ++c)
#pragma unroll
        for (int k = 0; k < N; ++k)
          if (!(nid[k] & node_flag(c))) atomicAdd(A.dst + (size_t)c * A.cstride + (nid[k] & NODE_MASK), dacc[c][k]);
    }
    if (!Op::INTEGRATE && has_next) gather(nid_next, u);  // evaluation-only ops (cache): plain prefetch
#pragma unroll
    for (int k = 0; k < N; ++k) nid[k] = nid_next[k];
  }
}

}  // namespace b200mf
