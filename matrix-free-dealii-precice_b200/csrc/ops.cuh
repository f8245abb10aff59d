// Quadrature-point functors plugged into cell_loop_kernel (cell_kernel.cuh).
// Each one restates a branch of the reference's quadrature loops; citations are
// relative to the reference tree.
#pragma once
#include "cell_kernel.cuh"

namespace b200mf {

// field ids of the deformed-frame "tensor2" table (18 values per q-point, 144 B in FP64):
// K = inverse Jacobian of the *current* configuration, K[e][f] = d xi_f / d x_e
// (MappingQEulerian: J_cur = F J_ref), w = JxW_cur,
// c1 = 2(mu - 2 lambda ln J)/J   (cached_scalar,        mf_nh_operator.h:464-468)
// c2 = 1/J                       (cached_second_scalar, :469)
// T  = tau/J, symmetric          (cached_tensor2,       :470-478; stored as 00,11,22,01,02,12)
enum { QF_K = 0, QF_W = 9, QF_C1 = 10, QF_C2 = 11, QF_T = 12, QF_TENSOR2_COUNT = 18 };
// reference metric table: K_ref (9) + JxW_ref
enum { MF_K = 0, MF_W = 9, MF_COUNT = 10 };

template <typename T>
struct MaterialConsts {  // material.h:78-94, 252-258
  T mu, lambda, rho_alpha, kappa, c_1;
};

// ---------------------------------------------------------------------------
// a4: tensor2 level, formulation 1, deformed frame (mf_nh_operator.h:945-974)
template <typename T, int N>
struct NHTensor2Op {
  static constexpr int NC = 3, NF = QF_TENSOR2_COUNT, NF_IN = QF_TENSOR2_COUNT, NF_IN2 = 0;
  static constexpr bool PLAIN = false, INTEGRATE = true;
  using Ly = Layout<N, NC, sizeof(T) / 4>;
  struct Args {
    const T *stage;   // tensor2 table [group][NF][N][L]
    const T *stage2;  // unused
    const uint8_t *matid;
    int64_t n_cells;
    T lam2[2];       // 2*lambda
    T rho_alpha[2];  // rho*alpha_1
  };
  struct CellState {
    T lam2, rho_alpha;
  };
  static __device__ __forceinline__ void cell_begin(const Args &oa, int64_t cell, bool valid, CellState &st) {
    int m = 0;
    if (oa.matid != nullptr && valid && cell < oa.n_cells) m = oa.matid[cell] != 0;  // :656-658
    st.lam2 = m ? oa.lam2[1] : oa.lam2[0];
    st.rho_alpha = m ? oa.rho_alpha[1] : oa.rho_alpha[0];
  }
  static __device__ __forceinline__ void quad(const Args &, const CellState &st, bool active, const T *qs, const T *,
                                              int64_t, int iq, int lane, const T (&v)[NC], const T (&G)[NC][3],
                                              T (&ov)[NC], T (&oG)[NC][3]) {
    if (!active) {
#pragma unroll
      for (int c = 0; c < NC; ++c) ov[c] = oG[c][0] = oG[c][1] = oG[c][2] = T(0);
      return;
    }
    const T *q = qs + iq * Ly::L + lane;
    constexpr int FS = N * Ly::L;
    T K[9], Ts[6];
#pragma unroll
    for (int i = 0; i < 9; ++i) K[i] = q[(QF_K + i) * FS];
    const T w = q[QF_W * FS], c1 = q[QF_C1 * FS], c2 = q[QF_C2 * FS];
#pragma unroll
    for (int i = 0; i < 6; ++i) Ts[i] = q[(QF_T + i) * FS];
    const T Tm[3][3] = {{Ts[0], Ts[3], Ts[4]}, {Ts[3], Ts[1], Ts[5]}, {Ts[4], Ts[5], Ts[2]}};
    T g[3][3];  // grad_x v, get_gradient :954
#pragma unroll
    for (int c = 0; c < 3; ++c)
#pragma unroll
      for (int e = 0; e < 3; ++e) g[c][e] = K[e * 3 + 0] * G[c][0] + K[e * 3 + 1] * G[c][1] + K[e * 3 + 2] * G[c][2];
    const T tmp = st.lam2 * c2 * (g[0][0] + g[1][1] + g[2][2]);  // 2 lambda c2 tr(sym g), :960-964
    const T hc1 = T(0.5) * c1;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      T sub[3];
#pragma unroll
      for (int e = 0; e < 3; ++e) {
        T s = hc1 * (g[c][e] + g[e][c]) + (c == e ? tmp : T(0));                   // jc
        s += g[c][0] * Tm[0][e] + g[c][1] * Tm[1][e] + g[c][2] * Tm[2][e];         // + grad_v * tau/J  :967-970
        sub[e] = s;
      }
#pragma unroll
      for (int f = 0; f < 3; ++f)  // submit_gradient: JxW * K^T
        oG[c][f] = w * (K[0 * 3 + f] * sub[0] + K[1 * 3 + f] * sub[1] + K[2 * 3 + f] * sub[2]);
      ov[c] = w * st.rho_alpha * c2 * v[c];  // submit_value :971
    }
  }
};

// ---------------------------------------------------------------------------
// a9: cache() for the tensor2 table (mf_nh_operator.h:392-523, formulation 1 branch :440-478)
// + deformed-frame metric (replaces MatrixFree::update_mapping(MappingQEulerian), mf_elasticity.h:955)
template <typename T, int N>
struct NHCacheTensor2Op {
  static constexpr int NC = 3, NF = QF_TENSOR2_COUNT, NF_IN = MF_COUNT, NF_IN2 = 0;
  static constexpr bool PLAIN = true, INTEGRATE = false;
  using Ly = Layout<N, NC, sizeof(T) / 4>;
  struct Args {
    const T *stage;   // reference metric [group][MF_COUNT][N][L]
    const T *stage2;  // unused
    T *qd;            // out: tensor2 table
    const uint8_t *matid;
    int64_t n_cells;
    T mu[2], lambda[2];
    unsigned long long *n_bad;
  };
  struct CellState {
    T mu, lambda;
    bool real;
  };
  static __device__ __forceinline__ void cell_begin(const Args &oa, int64_t cell, bool valid, CellState &st) {
    int m = 0;
    st.real = valid && cell < oa.n_cells;
    if (oa.matid != nullptr && st.real) m = oa.matid[cell] != 0;  // :402-404
    st.mu = m ? oa.mu[1] : oa.mu[0];
    st.lambda = m ? oa.lambda[1] : oa.lambda[0];
  }
  static __device__ __forceinline__ void quad(const Args &oa, const CellState &st, bool active, const T *qs, const T *,
                                              int64_t group, int iq, int lane, const T (&)[NC], const T (&G)[NC][3],
                                              T (&)[NC], T (&)[NC][3]) {
    if (!active) return;
    constexpr size_t FS = (size_t)N * Ly::L;
    const T *m = qs + iq * Ly::L + lane;
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
    Fi[0][0] = F[1][1] * F[2][2] - F[1][2] * F[2][1];
    Fi[0][1] = F[0][2] * F[2][1] - F[0][1] * F[2][2];
    Fi[0][2] = F[0][1] * F[1][2] - F[0][2] * F[1][1];
    Fi[1][0] = F[1][2] * F[2][0] - F[1][0] * F[2][2];
    Fi[1][1] = F[0][0] * F[2][2] - F[0][2] * F[2][0];
    Fi[1][2] = F[0][2] * F[1][0] - F[0][0] * F[1][2];
    Fi[2][0] = F[1][0] * F[2][1] - F[1][1] * F[2][0];
    Fi[2][1] = F[0][1] * F[2][0] - F[0][0] * F[2][1];
    Fi[2][2] = F[0][0] * F[1][1] - F[0][1] * F[1][0];
    const T detF = F[0][0] * Fi[0][0] + F[0][1] * Fi[1][0] + F[0][2] * Fi[2][0];
    if (st.real && !(detF > T(0))) atomicAdd(oa.n_bad, 1ull);  // :430-435
    const T idet = T(1) / detF;
    T *q = oa.qd + (size_t)group * Ly::table_stride(NF) + iq * Ly::L + lane;
    // K_cur[e][f] = sum_m Finv[m][e] K_ref[m][f]   (invjac_cur = F^-T invjac_ref)
#pragma unroll
    for (int e = 0; e < 3; ++e)
#pragma unroll
      for (int f = 0; f < 3; ++f)
        q[(QF_K + e * 3 + f) * FS] = (Fi[0][e] * Kr[0 * 3 + f] + Fi[1][e] * Kr[1 * 3 + f] + Fi[2][e] * Kr[2 * 3 + f]) * idet;
    q[QF_W * FS] = detF * wr;
    const T scalar = st.mu - T(2) * st.lambda * log(detF);  // :455-456
    q[QF_C1 * FS] = scalar * T(2) * idet;                   // :466
    q[QF_C2 * FS] = idet;                                   // :469
    constexpr int I6[6] = {0, 1, 2, 0, 0, 1}, J6[6] = {0, 1, 2, 1, 2, 2};
#pragma unroll
    for (int s = 0; s < 6; ++s) {
      const int i = I6[s], j = J6[s];
      const T bij = F[i][0] * F[j][0] + F[i][1] * F[j][1] + F[i][2] * F[j][2];  // b = F F^T
      q[(QF_T + s) * FS] = (st.mu * bij - (i == j ? scalar : T(0))) * idet;     // tau/J :470-478
    }
  }
};

// ---------------------------------------------------------------------------
// a13: LaplaceOperator::do_operation_on_cell (heat_transfer.h:234-250): (M + dt*c*K) u
template <typename T, int N>
struct LaplaceOp {
  static constexpr int NC = 1, NF_IN = MF_COUNT, NF_IN2 = 1;
  static constexpr bool PLAIN = false, INTEGRATE = true;
  using Ly = Layout<N, NC, sizeof(T) / 4>;
  struct Args {
    const T *stage;   // reference metric [group][MF_COUNT][N][L]
    const T *stage2;  // coefficient [group][1][N][L] or nullptr (Coefficient::value == 1, heat_transfer.h:64-71)
    T delta_t;
  };
  struct CellState {};
  static __device__ __forceinline__ void cell_begin(const Args &, int64_t, bool, CellState &) {}
  static __device__ __forceinline__ void quad(const Args &oa, const CellState &, bool active, const T *qs, const T *qs2,
                                              int64_t, int iq, int lane, const T (&v)[NC], const T (&G)[NC][3],
                                              T (&ov)[NC], T (&oG)[NC][3]) {
    if (!active) {
      ov[0] = oG[0][0] = oG[0][1] = oG[0][2] = T(0);
      return;
    }
    constexpr int FS = N * Ly::L;
    const T *m = qs + iq * Ly::L + lane;
    T K[9];
#pragma unroll
    for (int i = 0; i < 9; ++i) K[i] = m[(MF_K + i) * FS];
    const T w = m[MF_W * FS];
    T k = oa.delta_t;
    if (oa.stage2 != nullptr) k *= qs2[iq * Ly::L + lane];  // :245-247
    T g[3];
#pragma unroll
    for (int e = 0; e < 3; ++e) g[e] = K[e * 3 + 0] * G[0][0] + K[e * 3 + 1] * G[0][1] + K[e * 3 + 2] * G[0][2];
    const T wk = w * k;
#pragma unroll
    for (int f = 0; f < 3; ++f) oG[0][f] = wk * (K[0 * 3 + f] * g[0] + K[1 * 3 + f] * g[1] + K[2 * 3 + f] * g[2]);
    ov[0] = w * v[0];  // submit_value(get_value) :244
  }
};

}  // namespace b200mf
