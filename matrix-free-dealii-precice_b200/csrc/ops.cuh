// Quadrature-point functors plugged into cell_loop_kernel (cell_kernel.cuh).
// Each one restates a branch of the reference's quadrature loops; citations are
// relative to the reference tree.
#pragma once
#include "cell_kernel.cuh"

namespace b200mf {

// Kernel-side variants of NeoHookOperator (MF caching level x material formulation,
// mf_nh_operator.h:215-223, 346-387).
enum NHVariant { NHV_TENSOR2 = 0, NHV_FORM0 = 1, NHV_TENSOR4 = 2, NHV_TENSOR4_NS = 3, NHV_SCALAR_REF = 4 };

// Staged ("small") q-point table of the deformed-frame variants, fields of [group][NF][N][L]:
// K = inverse Jacobian of the *current* configuration, K[e][f] = d xi_f / d x_e
// (MappingQEulerian: J_cur = F J_ref), w = JxW_cur, then per variant
//  tensor2 (18): c1 = 2(mu - 2 lambda ln J)/J (cached_scalar :464-468), c2 = 1/J (cached_second_scalar :469),
//                T = tau/J symmetric 00,11,22,01,02,12 (cached_tensor2 :470-478)
//  form0   (20): cs = J^(-1/3) (cached_scalar :436), F row-major (cached_tensor2 :437)
//  tensor4 (17): c2 = 1/J (:485-486), T = tau/J (:487); the 6x6 cached_tensor4 (:488-493) is a separate
//                36-field table read straight from global memory
// Reference-frame variants stage the reference metric (MF_*) and
//  tensor4_ns : the 81-field cached_tensor4_ns (:497-514) read straight from global memory
//  scalar_ref : c = mu - 2 lambda ln J (cached_scalar :459-461) as second staged table and the cell-local
//               displacement values (cached_second_scalar(cell, i) :416-418) as [group][3][N][L]
enum { QF_K = 0, QF_W = 9, QF_C1 = 10, QF_C2 = 11, QF_T = 12, QF_TENSOR2_COUNT = 18 };
enum { QF0_CS = 10, QF0_F = 11, QF_FORM0_COUNT = 20 };
enum { QF4_C2 = 10, QF4_T = 11, QF_TENSOR4_COUNT = 17 };
// reference metric table: K_ref (9) + JxW_ref
enum { MF_K = 0, MF_W = 9, MF_COUNT = 10 };

__host__ __device__ constexpr int nhv_staged_fields(int v) {
  return v == NHV_TENSOR2 ? (int)QF_TENSOR2_COUNT : v == NHV_FORM0 ? (int)QF_FORM0_COUNT : v == NHV_TENSOR4 ? (int)QF_TENSOR4_COUNT : (int)MF_COUNT;
}
__host__ __device__ constexpr int nhv_big_fields(int v) { return v == NHV_TENSOR4 ? 36 : v == NHV_TENSOR4_NS ? 81 : 0; }
__host__ __device__ constexpr bool nhv_deformed(int v) { return v == NHV_TENSOR2 || v == NHV_FORM0 || v == NHV_TENSOR4; }

template <typename T>
struct NHArgs {
  const T *stage;   // staged table of the variant (deformed) or the reference metric
  const T *stage2;  // scalar_ref: c table; else null
  const T *big;     // tensor4 / tensor4_ns: [group][NB][N][L]
  const T *pos;     // scalar_ref: cell-local displacement [group][3][N][L]
  const uint8_t *matid;
  int64_t n_cells;
  // per material (0: material, 1: material_inclusion), material.h:78-94, 252-258
  T mu[2], lambda[2], rho_alpha[2], kappa[2], c_1[2];
};
template <typename T>
struct NHCell {
  T mu, lambda, rho_alpha, kappa, c_1;
};
template <typename T>
__device__ __forceinline__ void nh_cell_begin(const NHArgs<T> &oa, int64_t cell, bool valid, NHCell<T> &st) {
  int m = 0;
  if (oa.matid != nullptr && valid && cell < oa.n_cells) m = oa.matid[cell] != 0;  // :402-404, :656-658
  // selects, not dynamic indexing: indexing kernel parameters by a run-time value forces a local-memory copy
  st.mu = m ? oa.mu[1] : oa.mu[0];
  st.lambda = m ? oa.lambda[1] : oa.lambda[0];
  st.rho_alpha = m ? oa.rho_alpha[1] : oa.rho_alpha[0];
  st.kappa = m ? oa.kappa[1] : oa.kappa[0];
  st.c_1 = m ? oa.c_1[1] : oa.c_1[0];
}

template <typename T>
__device__ __forceinline__ T det3(const T (&F)[3][3]) {
  return F[0][0] * (F[1][1] * F[2][2] - F[1][2] * F[2][1]) - F[0][1] * (F[1][0] * F[2][2] - F[1][2] * F[2][0]) +
         F[0][2] * (F[1][0] * F[2][1] - F[1][1] * F[2][0]);
}
// Fi = adj(F) (so that F^-1 = Fi / det), returns det F
template <typename T>
__device__ __forceinline__ T adj3(const T (&F)[3][3], T (&Fi)[3][3]) {
  Fi[0][0] = F[1][1] * F[2][2] - F[1][2] * F[2][1];
  Fi[0][1] = F[0][2] * F[2][1] - F[0][1] * F[2][2];
  Fi[0][2] = F[0][1] * F[1][2] - F[0][2] * F[1][1];
  Fi[1][0] = F[1][2] * F[2][0] - F[1][0] * F[2][2];
  Fi[1][1] = F[0][0] * F[2][2] - F[0][2] * F[2][0];
  Fi[1][2] = F[0][2] * F[1][0] - F[0][0] * F[1][2];
  Fi[2][0] = F[1][0] * F[2][1] - F[1][1] * F[2][0];
  Fi[2][1] = F[0][1] * F[2][0] - F[0][0] * F[2][1];
  Fi[2][2] = F[0][0] * F[1][1] - F[0][1] * F[1][0];
  return F[0][0] * Fi[0][0] + F[0][1] * Fi[1][0] + F[0][2] * Fi[2][0];
}

// ---------------------------------------------------------------------------
// a4-a8: the quadrature loops of do_operation_on_cell (mf_nh_operator.h:697-1015)
template <typename T, int N, int V>
struct NHOp {
  static constexpr int NC = 3, NF_IN = nhv_staged_fields(V), NF_IN2 = (V == NHV_SCALAR_REF ? 1 : 0);
  static constexpr int NF_SCRATCH = (V == NHV_SCALAR_REF ? 9 : 0);
  static constexpr bool PLAIN = false, INTEGRATE = true, PRE = (V == NHV_SCALAR_REF), STORE_DOFS = false;
  using Ly = Layout<N, NC, sizeof(T) / 4>;
  using Args = NHArgs<T>;
  using CellState = NHCell<T>;
  static constexpr int FS = Ly::FS;    // field stride of the staged tables
  static constexpr int FSP = Ly::FSP;  // field stride of plain tables (big, pos) and the scratch
  static __device__ __forceinline__ void cell_begin(const Args &oa, int64_t cell, bool valid, int, CellState &st) {
    nh_cell_begin(oa, cell, valid, st);
  }
  static __device__ __forceinline__ void store_dofs(const Args &, int64_t, int, bool, const T (&)[NC][N]) {}

  // scalar_referential: second evaluate() on the cached cell-local displacement (:861-867)
  static __device__ __forceinline__ void load_pre(const Args &oa, int64_t group, int lane, bool active, int c, const uint32_t (&)[N],
                                                  T (&up)[N]) {
    const T *p = oa.pos + ((size_t)group * 3 + c) * FSP + lane;
#pragma unroll
    for (int k = 0; k < N; ++k) up[k] = active ? __ldg(p + (size_t)k * Ly::L) : T(0);
  }
  // F = I + Grad u through the reference inverse Jacobian (:872-899), kept in shared memory for the main pass
  static __device__ __forceinline__ void quad_pre(const Args &, const CellState &, bool active, const T *q, T *scr,
                                                  int64_t, int iq, int lane, const T (&)[NC], const T (&G)[NC][3]) {
    if (!active) return;
    T K[9];
#pragma unroll
    for (int i = 0; i < 9; ++i) K[i] = q[(MF_K + i) * FS];
    T *o = scr + iq * Ly::L + lane;
#pragma unroll
    for (int c = 0; c < 3; ++c)
#pragma unroll
      for (int e = 0; e < 3; ++e)
        o[(c * 3 + e) * FSP] = K[e * 3 + 0] * G[c][0] + K[e * 3 + 1] * G[c][1] + K[e * 3 + 2] * G[c][2] + (c == e ? T(1) : T(0));
  }

  static __device__ __forceinline__ void quad(const Args &oa, const CellState &st, bool active, const T *q, const T *q2,
                                              const T *scr, int64_t group, int iq, int lane, const T (&v)[NC],
                                              const T (&G)[NC][3], T (&ov)[NC], T (&oG)[NC][3]) {
    if (!active) {
#pragma unroll
      for (int c = 0; c < NC; ++c) ov[c] = oG[c][0] = oG[c][1] = oG[c][2] = T(0);
      return;
    }
    T K[9];  // inverse Jacobian of the frame the variant integrates on (QF_K == MF_K == 0)
#pragma unroll
    for (int i = 0; i < 9; ++i) K[i] = q[i * FS];
    const T w = q[9 * FS];
    T g[3][3];  // get_gradient
#pragma unroll
    for (int c = 0; c < 3; ++c)
#pragma unroll
      for (int e = 0; e < 3; ++e) g[c][e] = K[e * 3 + 0] * G[c][0] + K[e * 3 + 1] * G[c][1] + K[e * 3 + 2] * G[c][2];
    T sub[3][3];  // argument of submit_gradient
    T vs;         // factor of submit_value

    if constexpr (V == NHV_TENSOR2) {  // :945-974
      const T c1 = q[QF_C1 * FS], c2 = q[QF_C2 * FS];
      T Ts[6];
#pragma unroll
      for (int i = 0; i < 6; ++i) Ts[i] = q[(QF_T + i) * FS];
      const T Tm[3][3] = {{Ts[0], Ts[3], Ts[4]}, {Ts[3], Ts[1], Ts[5]}, {Ts[4], Ts[5], Ts[2]}};
      const T tmp = T(2) * st.lambda * c2 * (g[0][0] + g[1][1] + g[2][2]);  // :960-964
      const T hc1 = T(0.5) * c1;
#pragma unroll
      for (int c = 0; c < 3; ++c)
#pragma unroll
        for (int e = 0; e < 3; ++e)
          sub[c][e] = hc1 * (g[c][e] + g[e][c]) + (c == e ? tmp : T(0)) + g[c][0] * Tm[0][e] + g[c][1] * Tm[1][e] +
                      g[c][2] * Tm[2][e];  // jc + grad_v * tau/J :967-970
      vs = st.rho_alpha * c2;  // :971
    } else if constexpr (V == NHV_FORM0) {  // :701-840
      const T cs = q[QF0_CS * FS];
      T F[3][3];
#pragma unroll
      for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) F[i][j] = q[(QF0_F + i * 3 + j) * FS];
      const T detF = det3(F);
      const T cs2c = cs * cs * (T(2) * st.c_1);
      T tau_bar[3][3], s[3][3];
#pragma unroll
      for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) {
          tau_bar[i][j] = (F[i][0] * F[j][0] + F[i][1] * F[j][1] + F[i][2] * F[j][2]) * cs2c;  // 2 c_1 b_bar :725-746
          s[i][j] = T(0.5) * (g[i][j] + g[j][i]);
        }
      const T tr_tau_bar = tau_bar[0][0] + tau_bar[1][1] + tau_bar[2][2];
      const T third = T(1) / T(3), two_third = T(2) / T(3);
      const T tr_tau_bar_dim = tr_tau_bar * third;
      const T inv = T(1) / detF;
      const T dPsi = (st.kappa * T(0.5)) * (detF - inv);               // :753-754
      const T dPsiJ = dPsi * detF;
      const T d2Psi = (st.kappa * T(0.5)) * (T(1) + inv * inv);        // :757-758
      const T tr = s[0][0] + s[1][1] + s[2][2];
      const T tmp = detF * (dPsi + detF * d2Psi) * tr;                  // :801-804
      T tis = T(0);
#pragma unroll
      for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) tis += (tau_bar[i][j] - (i == j ? tr_tau_bar_dim : T(0))) * s[i][j];  // tau_iso : s
#pragma unroll
      for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) {
          const T dg = (i == j) ? T(1) : T(0);
          const T tau_iso = tau_bar[i][j] - dg * tr_tau_bar_dim;
          const T dev = s[i][j] - dg * tr * third;
          const T jc = s[i][j] * (-dPsiJ * T(2)) + dg * tmp + (two_third * tr_tau_bar) * dev - (two_third * tr) * tau_iso -
                       dg * two_third * tis;  // :796-822
          T geo = T(0);
#pragma unroll
          for (int k = 0; k < 3; ++k) geo += g[i][k] * (tau_bar[k][j] + (k == j ? dPsiJ - tr_tau_bar_dim : T(0)));  // grad * tau
          sub[i][j] = (jc + geo) * inv;  // :832-838
        }
      vs = st.rho_alpha * inv;  // :839
    } else if constexpr (V == NHV_TENSOR4) {  // :976-994
      const T c2 = q[QF4_C2 * FS];
      T Ts[6];
#pragma unroll
      for (int i = 0; i < 6; ++i) Ts[i] = q[(QF4_T + i) * FS];
      const T Tm[3][3] = {{Ts[0], Ts[3], Ts[4]}, {Ts[3], Ts[1], Ts[5]}, {Ts[4], Ts[5], Ts[2]}};
      constexpr int VI[6] = {0, 1, 2, 0, 0, 1}, VJ[6] = {0, 1, 2, 1, 2, 2};
      T sv[6];
#pragma unroll
      for (int I = 0; I < 6; ++I) sv[I] = (I < 3) ? g[VI[I]][VJ[I]] : (g[VI[I]][VJ[I]] + g[VJ[I]][VI[I]]);  // off-diagonals counted twice
      const T *C4 = oa.big + (size_t)group * Ly::plain_stride(36) + iq * Ly::L + lane;
      T cs6[6];
#pragma unroll
      for (int I = 0; I < 6; ++I) {
        T a = T(0);
#pragma unroll
        for (int J = 0; J < 6; ++J) a += __ldg(C4 + (size_t)(I * 6 + J) * FSP) * sv[J];
        cs6[I] = a;
      }
      const T Cs[3][3] = {{cs6[0], cs6[3], cs6[4]}, {cs6[3], cs6[1], cs6[5]}, {cs6[4], cs6[5], cs6[2]}};
#pragma unroll
      for (int c = 0; c < 3; ++c)
#pragma unroll
        for (int e = 0; e < 3; ++e) sub[c][e] = Cs[c][e] + g[c][0] * Tm[0][e] + g[c][1] * Tm[1][e] + g[c][2] * Tm[2][e];
      vs = st.rho_alpha * c2;
    } else if constexpr (V == NHV_TENSOR4_NS) {  // :996-1010
      const T *A4 = oa.big + (size_t)group * Ly::plain_stride(81) + iq * Ly::L + lane;
#pragma unroll
      for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) {
          T a = T(0);
#pragma unroll
          for (int k = 0; k < 3; ++k)
#pragma unroll
            for (int l = 0; l < 3; ++l) a += __ldg(A4 + (size_t)(((i * 3 + j) * 3 + k) * 3 + l) * FSP) * g[k][l];  // double_contract<2,0,3,1>
          sub[i][j] = a;
        }
      vs = st.rho_alpha;
    } else {  // NHV_SCALAR_REF :848-943
      const T cc = q2[0];
      const T *fp = scr + iq * Ly::L + lane;
      T F[3][3], Fa[3][3];
#pragma unroll
      for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) F[i][j] = fp[(i * 3 + j) * FSP];
      const T idet = T(1) / adj3(F, Fa);
      T gx[3][3];  // Grad v F^-1 :910-922
#pragma unroll
      for (int d = 0; d < 3; ++d)
#pragma unroll
        for (int e = 0; e < 3; ++e) gx[d][e] = (Fa[0][e] * g[d][0] + Fa[1][e] * g[d][1] + Fa[2][e] * g[d][2]) * idet;
      const T trl = T(2) * st.lambda * (gx[0][0] + gx[1][1] + gx[2][2]);
      T queued[3][3];
#pragma unroll
      for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) {
          T a = cc * (gx[i][j] + gx[j][i]) + (i == j ? trl : T(0));  // jc :924-933
#pragma unroll
          for (int k = 0; k < 3; ++k)
            a += gx[i][k] * (st.mu * (F[k][0] * F[j][0] + F[k][1] * F[j][1] + F[k][2] * F[j][2]) - (k == j ? cc : T(0)));  // grad * tau
          queued[i][j] = a;
        }
#pragma unroll
      for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) sub[i][j] = (queued[i][0] * Fa[j][0] + queued[i][1] * Fa[j][1] + queued[i][2] * Fa[j][2]) * idet;  // * F^-T :938
      vs = st.rho_alpha;  // :939
    }
#pragma unroll
    for (int c = 0; c < 3; ++c) {
#pragma unroll
      for (int f = 0; f < 3; ++f)  // submit_gradient: JxW * K^T
        oG[c][f] = w * (K[0 * 3 + f] * sub[c][0] + K[1 * 3 + f] * sub[c][1] + K[2 * 3 + f] * sub[c][2]);
      ov[c] = w * vs * v[c];
    }
  }
};

// ---------------------------------------------------------------------------
// a9: cache() (mf_nh_operator.h:392-523) for every variant, plus the deformed-frame metric
// (replaces MatrixFree::update_mapping(MappingQEulerian), mf_elasticity.h:955)
template <typename T>
struct NHCacheArgs {
  const T *stage;   // reference metric [group][MF_COUNT][N][L]
  const T *stage2;  // unused
  T *qd;            // out: staged table of the variant (deformed variants)
  T *qd2;           // out: scalar_ref c table
  T *big;           // out: tensor4 / tensor4_ns
  T *pos;           // out: scalar_ref cell-local displacement
  const uint8_t *matid;
  int64_t n_cells;
  T mu[2], lambda[2];
  unsigned long long *n_bad;
};

template <typename T, int N, int V>
struct NHCacheOp {
  static constexpr int NC = 3, NF_IN = MF_COUNT, NF_IN2 = 0, NF_SCRATCH = 0;
  static constexpr bool PLAIN = true, INTEGRATE = false, PRE = false, STORE_DOFS = (V == NHV_SCALAR_REF);
  using Ly = Layout<N, NC, sizeof(T) / 4>;
  using Args = NHCacheArgs<T>;
  static constexpr size_t FS = Ly::FS, FSP = Ly::FSP;
  struct CellState {
    T mu, lambda;
    bool real;
  };
  static __device__ __forceinline__ void cell_begin(const Args &oa, int64_t cell, bool valid, int, CellState &st) {
    int m = 0;
    st.real = valid && cell < oa.n_cells;
    if (oa.matid != nullptr && st.real) m = oa.matid[cell] != 0;  // :402-404
    st.mu = m ? oa.mu[1] : oa.mu[0];
    st.lambda = m ? oa.lambda[1] : oa.lambda[0];
  }
  // :416-418: keep the cell-local displacement so that vmult needs no indirect read of it
  static __device__ __forceinline__ void store_dofs(const Args &oa, int64_t group, int lane, bool active, const T (&u)[NC][N]) {
    if (!active) return;
#pragma unroll
    for (int c = 0; c < NC; ++c)
#pragma unroll
      for (int k = 0; k < N; ++k) oa.pos[((size_t)group * 3 + c) * FSP + (size_t)k * Ly::L + lane] = u[c][k];
  }
  static __device__ __forceinline__ void load_pre(const Args &, int64_t, int, bool, int, const uint32_t (&)[N], T (&)[N]) {}
  static __device__ __forceinline__ void quad_pre(const Args &, const CellState &, bool, const T *, T *, int64_t, int, int,
                                                  const T (&)[NC], const T (&)[NC][3]) {}
  static __device__ __forceinline__ void quad(const Args &oa, const CellState &st, bool active, const T *m, const T *, const T *,
                                              int64_t group, int iq, int lane, const T (&)[NC], const T (&G)[NC][3],
                                              T (&)[NC], T (&)[NC][3]) {
    if (!active) return;
    T Kr[9];
#pragma unroll
    for (int i = 0; i < 9; ++i) Kr[i] = m[(MF_K + i) * FS];
    const T wr = m[MF_W * FS];
    T F[3][3];  // F = I + Grad u, :427-428
#pragma unroll
    for (int c = 0; c < 3; ++c)
#pragma unroll
      for (int e = 0; e < 3; ++e)
        F[c][e] = Kr[e * 3 + 0] * G[c][0] + Kr[e * 3 + 1] * G[c][1] + Kr[e * 3 + 2] * G[c][2] + (c == e ? T(1) : T(0));
    T Fi[3][3];
    const T detF = adj3(F, Fi);
    if (st.real && !(detF > T(0))) atomicAdd(oa.n_bad, 1ull);  // :430-435
    const T idet = T(1) / detF;
    const size_t qoff = (size_t)iq * Ly::L + lane;  // plain tables
    if constexpr (nhv_deformed(V)) {
      T *q = oa.qd + (size_t)group * Ly::table_stride(nhv_staged_fields(V)) + (size_t)iq * Ly::kq_stride(nhv_staged_fields(V)) + lane;
      // K_cur[e][f] = sum_m Finv[m][e] K_ref[m][f]   (invjac_cur = F^-T invjac_ref)
#pragma unroll
      for (int e = 0; e < 3; ++e)
#pragma unroll
        for (int f = 0; f < 3; ++f)
          q[(QF_K + e * 3 + f) * FS] = (Fi[0][e] * Kr[0 * 3 + f] + Fi[1][e] * Kr[1 * 3 + f] + Fi[2][e] * Kr[2 * 3 + f]) * idet;
      q[QF_W * FS] = detF * wr;
      if constexpr (V == NHV_FORM0) {  // :419-439
        q[QF0_CS * FS] = pow(detF, T(-1.0 / 3.0));
#pragma unroll
        for (int i = 0; i < 3; ++i)
#pragma unroll
          for (int j = 0; j < 3; ++j) q[(QF0_F + i * 3 + j) * FS] = F[i][j];
      } else {
        const T scalar = st.mu - T(2) * st.lambda * log(detF);  // :455-456
        constexpr int I6[6] = {0, 1, 2, 0, 0, 1}, J6[6] = {0, 1, 2, 1, 2, 2};
        constexpr int QT = (V == NHV_TENSOR2) ? QF_T : QF4_T;
#pragma unroll
        for (int s = 0; s < 6; ++s) {
          const int i = I6[s], j = J6[s];
          const T bij = F[i][0] * F[j][0] + F[i][1] * F[j][1] + F[i][2] * F[j][2];  // b = F F^T
          q[(QT + s) * FS] = (st.mu * bij - (i == j ? scalar : T(0))) * idet;        // tau/J :470-478, :480-487
        }
        if constexpr (V == NHV_TENSOR2) {
          q[QF_C1 * FS] = scalar * T(2) * idet;  // :466
          q[QF_C2 * FS] = idet;                  // :469
        } else {
          q[QF4_C2 * FS] = idet;  // :485-486
          // cached_tensor4 = (2 scalar / J) S + (2 lambda / J) IxI as the 6x6 SymmetricTensor<4> :488-493
          T *C4 = oa.big + (size_t)group * Ly::plain_stride(36) + qoff;
          const T a = scalar * T(2) * idet, b = st.lambda * T(2) * idet;
#pragma unroll
          for (int I = 0; I < 6; ++I)
#pragma unroll
            for (int J = 0; J < 6; ++J) {
              const int i = I6[I], j = J6[I], k = I6[J], l = J6[J];
              const T S4 = T(0.5) * T((i == k && j == l) + (i == l && j == k));
              const T IxI = T(i == j && k == l);
              C4[(size_t)(I * 6 + J) * FSP] = a * S4 + b * IxI;
            }
        }
      }
    } else if constexpr (V == NHV_SCALAR_REF) {
      oa.qd2[(size_t)group * Ly::table_stride(1) + (size_t)iq * Ly::kq_stride(1) + lane] = st.mu - T(2) * st.lambda * log(detF);  // :459-461
    } else {  // NHV_TENSOR4_NS :497-514
      T *A4 = oa.big + (size_t)group * Ly::plain_stride(81) + qoff;
      const T lnJ = log(detF);
      T Fv[3][3];  // F^-1
#pragma unroll
      for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) Fv[i][j] = Fi[i][j] * idet;
      const T l2 = T(2) * st.lambda, sc = st.mu - T(2) * st.lambda * lnJ;
#pragma unroll
      for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j)
#pragma unroll
          for (int k = 0; k < 3; ++k)
#pragma unroll
            for (int l = 0; l < 3; ++l) {
              const T t1 = Fv[j][i] * Fv[l][k];  // outer_product(F^-T, F^-T)
              const T t2 = Fv[l][i] * Fv[j][k];  // outer_product_iljk(F^-T, F^-1), :23-46
              const T t3 = T(i == k && j == l);  // IxI_ikjl
              A4[(size_t)(((i * 3 + j) * 3 + k) * 3 + l) * FSP] = l2 * t1 + sc * t2 + st.mu * t3;
            }
    }
  }
};

// ---------------------------------------------------------------------------
// a13: LaplaceOperator::do_operation_on_cell (heat_transfer.h:234-250): (M + dt*c*K) u
template <typename T>
struct LaplaceArgs {
  const T *stage;   // reference metric [group][MF_COUNT][N][L]  (general cells)
  const T *stage2;  // coefficient [group][1][N][L] or nullptr (Coefficient::value == 1, heat_transfer.h:64-71)
  const T *mcell;   // affine cells: [cell][MC_COUNT] (padded to whole groups)
  T delta_t;
};
// per-cell metric of an affine cell (deal.II stores one Jacobian per Cartesian/affine cell):
// M = det J * K^T K (symmetric, 00 11 22 01 02 12: the Laplacian pulled back to the reference cell), det J
enum { MC_M = 0, MC_DET = 6, MC_COUNT = 8 };
template <typename T, int N>
struct LaplaceOp {
  static constexpr int NC = 1, NF_IN = MF_COUNT, NF_IN2 = 1, NF_SCRATCH = 0;
  static constexpr bool PLAIN = false, INTEGRATE = true, PRE = false, STORE_DOFS = false;
  using Ly = Layout<N, NC, sizeof(T) / 4>;
  using Args = LaplaceArgs<T>;
  struct CellState {};
  static __device__ __forceinline__ void cell_begin(const Args &, int64_t, bool, int, CellState &) {}
  static __device__ __forceinline__ void store_dofs(const Args &, int64_t, int, bool, const T (&)[NC][N]) {}
  static __device__ __forceinline__ void load_pre(const Args &, int64_t, int, bool, int, const uint32_t (&)[N], T (&)[N]) {}
  static __device__ __forceinline__ void quad_pre(const Args &, const CellState &, bool, const T *, T *, int64_t, int, int,
                                                  const T (&)[NC], const T (&)[NC][3]) {}
  static __device__ __forceinline__ void quad(const Args &oa, const CellState &, bool active, const T *m, const T *q2, const T *,
                                              int64_t, int, int, const T (&v)[NC], const T (&G)[NC][3],
                                              T (&ov)[NC], T (&oG)[NC][3]) {
    if (!active) {
      ov[0] = oG[0][0] = oG[0][1] = oG[0][2] = T(0);
      return;
    }
    constexpr int FS = Ly::FS;
    T K[9];
#pragma unroll
    for (int i = 0; i < 9; ++i) K[i] = m[(MF_K + i) * FS];
    const T w = m[MF_W * FS];
    T k = oa.delta_t;
    if (oa.stage2 != nullptr) k *= q2[0];  // :245-247
    T g[3];
#pragma unroll
    for (int e = 0; e < 3; ++e) g[e] = K[e * 3 + 0] * G[0][0] + K[e * 3 + 1] * G[0][1] + K[e * 3 + 2] * G[0][2];
    const T wk = w * k;
#pragma unroll
    for (int f = 0; f < 3; ++f) oG[0][f] = wk * (K[0 * 3 + f] * g[0] + K[1 * 3 + f] * g[1] + K[2 * 3 + f] * g[2]);
    ov[0] = w * v[0];  // submit_value(get_value) :244
  }
};

// The same operator on affine cells: no q-point table at all (the metric is 7 numbers per cell), so the
// kernel streams only the vectors and the index table and is bound by the FP64/FP32 pipe instead of HBM.
template <typename T, int N>
struct LaplaceAffineOp {
  static constexpr int NC = 1, NF_IN = 0, NF_IN2 = 1, NF_SCRATCH = 0;
  static constexpr bool PLAIN = false, INTEGRATE = true, PRE = false, STORE_DOFS = false;
  using Ly = Layout<N, NC, sizeof(T) / 4>;
  using Args = LaplaceArgs<T>;
  struct CellState {
    T M[6], mw;  // delta_t * w_i w_j * M,  det J * w_i w_j
  };
  static __device__ __forceinline__ void cell_begin(const Args &oa, int64_t cell, bool valid, int lane, CellState &st) {
    const int ji = lane % Ly::N2;
    const T wij = valid ? Shape<T, N>::W(ji % N) * Shape<T, N>::W(ji / N) : T(0);
    const T *m = oa.mcell + (size_t)cell * MC_COUNT;
#pragma unroll
    for (int i = 0; i < 6; ++i) st.M[i] = valid ? __ldg(m + MC_M + i) * wij * oa.delta_t : T(0);
    st.mw = valid ? __ldg(m + MC_DET) * wij : T(0);
  }
  static __device__ __forceinline__ void store_dofs(const Args &, int64_t, int, bool, const T (&)[NC][N]) {}
  static __device__ __forceinline__ void load_pre(const Args &, int64_t, int, bool, int, const uint32_t (&)[N], T (&)[N]) {}
  static __device__ __forceinline__ void quad_pre(const Args &, const CellState &, bool, const T *, T *, int64_t, int, int,
                                                  const T (&)[NC], const T (&)[NC][3]) {}
  static __device__ __forceinline__ void quad(const Args &oa, const CellState &st, bool active, const T *, const T *q2, const T *,
                                              int64_t, int iq, int, const T (&v)[NC], const T (&G)[NC][3],
                                              T (&ov)[NC], T (&oG)[NC][3]) {
    const T wk = Shape<T, N>::W(iq);
    T k = wk;
    if (oa.stage2 != nullptr && active) k *= q2[0];  // coefficient(cell, q), heat_transfer.h:245-247
    oG[0][0] = k * (st.M[0] * G[0][0] + st.M[3] * G[0][1] + st.M[4] * G[0][2]);
    oG[0][1] = k * (st.M[3] * G[0][0] + st.M[1] * G[0][1] + st.M[5] * G[0][2]);
    oG[0][2] = k * (st.M[4] * G[0][0] + st.M[5] * G[0][1] + st.M[2] * G[0][2]);
    ov[0] = st.mw * wk * v[0];  // submit_value(get_value) :244
  }
};

// ---------------------------------------------------------------------------
// the cell part of Solid::assemble_residual, fast path (mf_elasticity.h:1764-1846):
//   rhs += int [ -tau(F) F^-T : Grad v  +  rho (b - a) . v ] dV  on the reference configuration,
// total_displacement and acceleration read with read_dof_values_plain, constrained rows not written.
// The acceleration is the extra evaluation-only pass (gathered through the same node indices); its q-point
// values wait in shared memory for the main pass over the displacement.
template <typename T>
struct NHResidualArgs {
  const T *stage;   // reference metric [group][MF_COUNT][N][L]
  const T *stage2;  // unused
  const T *acc;     // acceleration vector (component blocks)
  int64_t cstride;
  const uint8_t *matid;
  int64_t n_cells;
  T mu[2], lambda[2], kappa[2], c_1[2], rho[2];
  int formulation;
  T body_force[3];
  unsigned long long *n_bad;  // quadrature points with det F <= 0 (debug assertion :1812-1817)
};
template <typename T, int N>
struct NHResidualOp {
  static constexpr int NC = 3, NF_IN = MF_COUNT, NF_IN2 = 0, NF_SCRATCH = 3;
  static constexpr bool PLAIN = true, INTEGRATE = true, PRE = true, STORE_DOFS = false;
  using Ly = Layout<N, NC, sizeof(T) / 4>;
  using Args = NHResidualArgs<T>;
  static constexpr int FS = Ly::FS;
  struct CellState {
    T mu, lambda, kappa, c_1, rho;
    bool real;
  };
  static __device__ __forceinline__ void cell_begin(const Args &oa, int64_t cell, bool valid, int, CellState &st) {
    int m = 0;
    st.real = valid && cell < oa.n_cells;
    if (oa.matid != nullptr && st.real) m = oa.matid[cell] != 0;  // :1781-1784
    st.mu = m ? oa.mu[1] : oa.mu[0];
    st.lambda = m ? oa.lambda[1] : oa.lambda[0];
    st.kappa = m ? oa.kappa[1] : oa.kappa[0];
    st.c_1 = m ? oa.c_1[1] : oa.c_1[0];
    st.rho = m ? oa.rho[1] : oa.rho[0];
  }
  static __device__ __forceinline__ void store_dofs(const Args &, int64_t, int, bool, const T (&)[NC][N]) {}
  // phi_acc.read_dof_values_plain(acceleration) :1790-1792
  static __device__ __forceinline__ void load_pre(const Args &oa, int64_t, int, bool active, int c, const uint32_t (&nid)[N], T (&up)[N]) {
#pragma unroll
    for (int k = 0; k < N; ++k) up[k] = active ? __ldg(oa.acc + (size_t)c * oa.cstride + (nid[k] & NODE_MASK)) : T(0);
  }
  static __device__ __forceinline__ void quad_pre(const Args &, const CellState &, bool active, const T *, T *scr, int64_t, int iq,
                                                  int lane, const T (&v)[NC], const T (&)[NC][3]) {
    if (!active) return;
#pragma unroll
    for (int c = 0; c < 3; ++c) scr[(c * N + iq) * Ly::L + lane] = v[c];  // phi_acc.get_value(q)
  }
  static __device__ __forceinline__ void quad(const Args &oa, const CellState &st, bool active, const T *q, const T *, const T *scr,
                                              int64_t, int iq, int lane, const T (&)[NC], const T (&G)[NC][3], T (&ov)[NC],
                                              T (&oG)[NC][3]) {
    if (!active) {
#pragma unroll
      for (int c = 0; c < NC; ++c) ov[c] = oG[c][0] = oG[c][1] = oG[c][2] = T(0);
      return;
    }
    T K[9];
#pragma unroll
    for (int i = 0; i < 9; ++i) K[i] = q[(MF_K + i) * FS];
    const T w = q[MF_W * FS];
    T F[3][3], Fa[3][3];  // F = I + Grad u :1802-1805
#pragma unroll
    for (int c = 0; c < 3; ++c)
#pragma unroll
      for (int e = 0; e < 3; ++e)
        F[c][e] = K[e * 3 + 0] * G[c][0] + K[e * 3 + 1] * G[c][1] + K[e * 3 + 2] * G[c][2] + (c == e ? T(1) : T(0));
    const T detF = adj3(F, Fa);
    if (st.real && !(detF > T(0))) atomicAdd(oa.n_bad, 1ull);
    const T idet = T(1) / detF;
    T tau[3][3];  // get_tau, material.h:116-161
    if (oa.formulation == 0) {
      const T cs = pow(detF, T(-1.0 / 3.0));           // F_iso = J^(-1/3) F
      const T tb = cs * cs * (T(2) * st.c_1);          // tau_bar = 2 c_1 b_bar
      const T tmp = (st.kappa * T(0.5)) * (detF - idet) * detF;  // dPsi_vol/dJ * J
      T tr = T(0);
#pragma unroll
      for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) tau[i][j] = (F[i][0] * F[j][0] + F[i][1] * F[j][1] + F[i][2] * F[j][2]) * tb;
      tr = tau[0][0] + tau[1][1] + tau[2][2];
#pragma unroll
      for (int d = 0; d < 3; ++d) tau[d][d] += tmp - tr / T(3);
    } else {
      const T tmp = st.mu - T(2) * st.lambda * log(detF);
#pragma unroll
      for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j)
          tau[i][j] = st.mu * (F[i][0] * F[j][0] + F[i][1] * F[j][1] + F[i][2] * F[j][2]) - (i == j ? tmp : T(0));
    }
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      T sub[3];  // -(tau F^-T)[c][e] = -sum_k tau[c][k] Finv[e][k] :1830
#pragma unroll
      for (int e = 0; e < 3; ++e) sub[e] = -(tau[c][0] * Fa[e][0] + tau[c][1] * Fa[e][1] + tau[c][2] * Fa[e][2]) * idet;
#pragma unroll
      for (int f = 0; f < 3; ++f) oG[c][f] = w * (K[0 * 3 + f] * sub[0] + K[1 * 3 + f] * sub[1] + K[2 * 3 + f] * sub[2]);
      ov[c] = w * st.rho * (oa.body_force[c] - scr[(c * N + iq) * Ly::L + lane]);  // :1831-1834
    }
  }
};

// ---------------------------------------------------------------------------
// cell loop of LaplaceProblem::assemble_rhs (heat_transfer.h:599-615):
//   rhs += int (u_old + dt f(x_q)) phi   (gather_evaluate(values), submit_value, integrate_scatter(values))
// AFFINE: per-cell metric (JxW = det J w_i w_j w_k), else the staged reference metric.
template <typename T>
struct LaplaceRhsArgs {
  const T *stage;   // reference metric [group][MF_COUNT][N][L] (general cells)
  const T *stage2;  // f at the quadrature points [group][1][N][L] or nullptr (then source_const)
  const T *mcell;   // affine cells
  T delta_t, source_const;
};
template <typename T, int N, bool AFFINE>
struct LaplaceRhsOp {
  static constexpr int NC = 1, NF_IN = AFFINE ? 0 : MF_COUNT, NF_IN2 = 1, NF_SCRATCH = 0;
  static constexpr bool PLAIN = false, INTEGRATE = true, PRE = false, STORE_DOFS = false;
  using Ly = Layout<N, NC, sizeof(T) / 4>;
  using Args = LaplaceRhsArgs<T>;
  struct CellState {
    T mw;  // affine: det J * w_i w_j
  };
  static __device__ __forceinline__ void cell_begin(const Args &oa, int64_t cell, bool valid, int lane, CellState &st) {
    st.mw = T(0);
    if (AFFINE && valid) {
      const int ji = lane % Ly::N2;
      st.mw = __ldg(oa.mcell + (size_t)cell * MC_COUNT + MC_DET) * Shape<T, N>::W(ji % N) * Shape<T, N>::W(ji / N);
    }
  }
  static __device__ __forceinline__ void store_dofs(const Args &, int64_t, int, bool, const T (&)[NC][N]) {}
  static __device__ __forceinline__ void load_pre(const Args &, int64_t, int, bool, int, const uint32_t (&)[N], T (&)[N]) {}
  static __device__ __forceinline__ void quad_pre(const Args &, const CellState &, bool, const T *, T *, int64_t, int, int,
                                                  const T (&)[NC], const T (&)[NC][3]) {}
  static __device__ __forceinline__ void quad(const Args &oa, const CellState &st, bool active, const T *q, const T *q2, const T *,
                                              int64_t, int iq, int, const T (&v)[NC], const T (&)[NC][3], T (&ov)[NC],
                                              T (&oG)[NC][3]) {
    oG[0][0] = oG[0][1] = oG[0][2] = T(0);
    ov[0] = T(0);
    if (!active) return;
    constexpr int FS = Ly::FS;
    const T w = AFFINE ? st.mw * Shape<T, N>::W(iq) : q[MF_W * FS];
    const T f = oa.stage2 != nullptr ? q2[0] : oa.source_const;
    ov[0] = w * (v[0] + oa.delta_t * f);  // :607-611
  }
};

}  // namespace b200mf
