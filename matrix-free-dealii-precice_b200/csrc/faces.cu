// Interface-face part of Solid::assemble_residual on the device (second half):
//   include/solid_mechanics/mf_elasticity.h:1849-1897
//     for every boundary face with id == interface_id:
//       phi.gather_evaluate(total_displacement, gradients)           (FEFaceEvaluation on the reference MatrixFree)
//       F = I + Grad u;  n_star = det F F^-T N;  t_0 = |n_star| * traction(q)     (pull-back of the preCICE traction)
//       phi.submit_value(t_0, q);  phi.integrate_scatter(values, system_rhs)
// The traction at the face quadrature points arrives from preCICE once per time window (host array); the loop above
// runs in every Newton iteration, so the face data and the traction live on the device and system_rhs never leaves it.
//
// Geometry as in the cells: Q1 (StaticMappingQ1 of the reference configuration) from the 8 vertices of the face's cell.
// One block per face: thread (q1, q2) evaluates its quadrature point (the nodal basis is interpolatory at the end
// points, so phi_a vanishes on the face unless node a lies on it: the value integral only touches the N^2 face nodes,
// the gradient needs the whole cell), then thread (i1, i2) integrates against its face node and scatters.
//
// Value integrals on the same face sets (heat solver):
//   include/adapter/boundary_mass_operator.h:137-161   BoundaryMassOperator::apply_add: dst += int_Gamma u phi dA
//   include/adapter/boundary_mass_operator.h:79-118    compute_diagonal: "diagonal" = M_Gamma * 1 (row sums), inverse only where > 0
//   include/adapter/boundary_projector.h:86-110        project: rhs = residual on the interface DoFs, CG(Jacobi 0.8) from x0 = residual
//   include/heat_transfer/heat_transfer.h:617-652      assemble_rhs: rhs += int_Gamma (-flux(q) dt) phi dA (flux from the coupling partner)
#include <cstring>
#include <memory>
#include <vector>

#include "cell_kernel.cuh"
#include "common.h"

using namespace b200mf;

namespace b200mf {

struct FaceShape {
  double S[MAXN * MAXN];   // l_i(xq)
  double D[MAXN * MAXN];   // l_i'(xq)
  double E[2][MAXN];       // l_i'(0), l_i'(1)
  double xq[MAXN], wq[MAXN];
};

// quadrature-point order on a face normal to direction d: tangential directions t1 < t2, q = q1 + N*q2 (t1 fastest)
template <typename T>
__global__ void face_traction_kernel(int N, int64_t n_faces, const uint32_t *__restrict__ cell_nodes /*[face][N^3] node | flags*/,
                                     const int *__restrict__ face_no, const double *__restrict__ vtx /*[face][8][3]*/,
                                     const double *__restrict__ traction /*[face][N^2][3]*/, int64_t cstride,
                                     const T *__restrict__ u, T *__restrict__ rhs, const __grid_constant__ FaceShape sh) {
  extern __shared__ double fsm[];
  const int N2 = N * N, N3 = N2 * N;
  const int64_t face = blockIdx.x;
  double *uc = fsm;            // [3][N^3] cell-local displacement
  double *tq = fsm + 3 * N3;   // [3][N^2] JxW * t_0 at the face quadrature points
  const uint32_t *nodes = cell_nodes + face * N3;
  for (int a = threadIdx.x; a < N3; a += blockDim.x) {
    const uint32_t id = nodes[a];
    for (int c = 0; c < 3; ++c)  // read_dof_values: constrained entries (homogeneous Dirichlet) read as 0
      uc[c * N3 + a] = (id & node_flag(c)) ? 0.0 : (double)u[(size_t)c * cstride + (id & NODE_MASK)];
  }
  __syncthreads();
  const int fn = face_no[face], d = fn / 2, side = fn % 2;
  const int t1 = d == 0 ? 1 : 0, t2 = d == 2 ? 1 : 2;
  const int stride[3] = {1, N, N2};
  if ((int)threadIdx.x < N2) {
    const int q1 = threadIdx.x % N, q2 = threadIdx.x / N;
    // reference gradient dU_c/dxi_f at the face point: tangential directions by the Gauss tables, normal one by the end-point derivative
    double G[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}};
    const int kf = side ? N - 1 : 0;  // face node layer
    for (int k = 0; k < N; ++k) {
      const double en = sh.E[side][k];
      for (int j = 0; j < N; ++j) {
        const double s2 = sh.S[q2 * N + j], d2 = sh.D[q2 * N + j];
        for (int i = 0; i < N; ++i) {
          const double s1 = sh.S[q1 * N + i], d1 = sh.D[q1 * N + i];
          const int a = k * stride[d] + i * stride[t1] + j * stride[t2];
          const double wn = en * s1 * s2;                        // d/dxi_d
          const double w1 = (k == kf) ? d1 * s2 : 0.0;           // d/dxi_t1 (only face nodes have a value on the face)
          const double w2 = (k == kf) ? s1 * d2 : 0.0;           // d/dxi_t2
          for (int c = 0; c < 3; ++c) {
            const double v = uc[c * N3 + a];
            G[c][d] += wn * v;
            G[c][t1] += w1 * v;
            G[c][t2] += w2 * v;
          }
        }
      }
    }
    // Q1 geometry at the face point
    double xi[3];
    xi[d] = side;
    xi[t1] = sh.xq[q1];
    xi[t2] = sh.xq[q2];
    const double *v = vtx + face * 24;
    double J[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
    for (int m = 0; m < 8; ++m) {
      double l[3], dl[3];
      for (int e = 0; e < 3; ++e) {
        const int bit = (m >> e) & 1;
        l[e] = bit ? xi[e] : 1.0 - xi[e];
        dl[e] = bit ? 1.0 : -1.0;
      }
      const double g[3] = {dl[0] * l[1] * l[2], l[0] * dl[1] * l[2], l[0] * l[1] * dl[2]};
      for (int e = 0; e < 3; ++e)
        for (int f = 0; f < 3; ++f) J[e * 3 + f] += v[m * 3 + e] * g[f];
    }
    double A[3][3];  // adj(J): J^-1 = A / det, A[f][e]
    A[0][0] = J[4] * J[8] - J[5] * J[7];
    A[0][1] = J[2] * J[7] - J[1] * J[8];
    A[0][2] = J[1] * J[5] - J[2] * J[4];
    A[1][0] = J[5] * J[6] - J[3] * J[8];
    A[1][1] = J[0] * J[8] - J[2] * J[6];
    A[1][2] = J[2] * J[3] - J[0] * J[5];
    A[2][0] = J[3] * J[7] - J[4] * J[6];
    A[2][1] = J[1] * J[6] - J[0] * J[7];
    A[2][2] = J[0] * J[4] - J[1] * J[3];
    const double det = J[0] * A[0][0] + J[1] * A[1][0] + J[2] * A[2][0];
    const double idet = 1.0 / det;
    // outward unit normal and surface element: J^-T e_d * det J = row d of adj(J); sign by the side
    double nrm[3] = {A[d][0], A[d][1], A[d][2]};
    const double sgn = (side ? 1.0 : -1.0) * (det > 0 ? 1.0 : -1.0);
    const double len = sqrt(nrm[0] * nrm[0] + nrm[1] * nrm[1] + nrm[2] * nrm[2]);
    for (int e = 0; e < 3; ++e) nrm[e] *= sgn / len;
    const double jxw = len * sh.wq[q1] * sh.wq[q2];  // |det J J^-T e_d| w
    // F = I + Grad u, Grad u[c][e] = sum_f G[c][f] dxi_f/dx_e
    double F[3][3];
    for (int c = 0; c < 3; ++c)
      for (int e = 0; e < 3; ++e)
        F[c][e] = (G[c][0] * A[0][e] + G[c][1] * A[1][e] + G[c][2] * A[2][e]) * idet + (c == e ? 1.0 : 0.0);
    // n_star = det F F^-T N = cof(F) N
    double ns[3];
    ns[0] = (F[1][1] * F[2][2] - F[1][2] * F[2][1]) * nrm[0] + (F[1][2] * F[2][0] - F[1][0] * F[2][2]) * nrm[1] + (F[1][0] * F[2][1] - F[1][1] * F[2][0]) * nrm[2];
    ns[1] = (F[0][2] * F[2][1] - F[0][1] * F[2][2]) * nrm[0] + (F[0][0] * F[2][2] - F[0][2] * F[2][0]) * nrm[1] + (F[0][1] * F[2][0] - F[0][0] * F[2][1]) * nrm[2];
    ns[2] = (F[0][1] * F[1][2] - F[0][2] * F[1][1]) * nrm[0] + (F[0][2] * F[1][0] - F[0][0] * F[1][2]) * nrm[1] + (F[0][0] * F[1][1] - F[0][1] * F[1][0]) * nrm[2];
    const double da = sqrt(ns[0] * ns[0] + ns[1] * ns[1] + ns[2] * ns[2]);  // da/dA
    const double *t = traction + (face * N2 + threadIdx.x) * 3;
    for (int c = 0; c < 3; ++c) tq[c * N2 + threadIdx.x] = t[c] * da * jxw;
  }
  __syncthreads();
  if ((int)threadIdx.x < N2) {
    const int i1 = threadIdx.x % N, i2 = threadIdx.x / N;
    double acc[3] = {0, 0, 0};
    for (int q2 = 0; q2 < N; ++q2)
      for (int q1 = 0; q1 < N; ++q1) {
        const double w = sh.S[q1 * N + i1] * sh.S[q2 * N + i2];
        for (int c = 0; c < 3; ++c) acc[c] += w * tq[c * N2 + q2 * N + q1];
      }
    const int kf = side ? N - 1 : 0;
    const uint32_t id = nodes[kf * stride[d] + i1 * stride[t1] + i2 * stride[t2]];
    for (int c = 0; c < 3; ++c)
      if (!(id & node_flag(c))) atomicAdd(rhs + (size_t)c * cstride + (id & NODE_MASK), (T)acc[c]);
  }
}

// positions of the face quadrature points in the reference configuration (the interface mesh the adapter hands to preCICE)
__global__ void face_points_kernel(int N, int64_t n_faces, const int *__restrict__ face_no, const double *__restrict__ vtx,
                                   double *__restrict__ xyz, const __grid_constant__ FaceShape sh) {
  const int N2 = N * N;
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n_faces * N2) return;
  const int64_t face = idx / N2;
  const int q = (int)(idx % N2), q1 = q % N, q2 = q / N;
  const int fn = face_no[face], d = fn / 2, side = fn % 2;
  const int t1 = d == 0 ? 1 : 0, t2 = d == 2 ? 1 : 2;
  double xi[3];
  xi[d] = side;
  xi[t1] = sh.xq[q1];
  xi[t2] = sh.xq[q2];
  const double *v = vtx + face * 24;
  double x[3] = {0, 0, 0};
  for (int m = 0; m < 8; ++m) {
    double w = 1.0;
    for (int e = 0; e < 3; ++e) w *= ((m >> e) & 1) ? xi[e] : 1.0 - xi[e];
    for (int e = 0; e < 3; ++e) x[e] += w * v[m * 3 + e];
  }
  for (int e = 0; e < 3; ++e) xyz[idx * 3 + e] = x[e];
}

// |det J J^-T e_d| of the Q1 cell map at the reference point xi on the face normal to d: the surface element dA/d(xi_t1 xi_t2)
__device__ inline double face_area_element(const double *v /*[8][3]*/, int d, const double xi[3]) {
  double J[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
  for (int m = 0; m < 8; ++m) {
    double l[3], dl[3];
    for (int e = 0; e < 3; ++e) {
      const int bit = (m >> e) & 1;
      l[e] = bit ? xi[e] : 1.0 - xi[e];
      dl[e] = bit ? 1.0 : -1.0;
    }
    const double g[3] = {dl[0] * l[1] * l[2], l[0] * dl[1] * l[2], l[0] * l[1] * dl[2]};
    for (int e = 0; e < 3; ++e)
      for (int f = 0; f < 3; ++f) J[e * 3 + f] += v[m * 3 + e] * g[f];
  }
  double a[3];  // row d of adj(J)
  if (d == 0) {
    a[0] = J[4] * J[8] - J[5] * J[7];
    a[1] = J[2] * J[7] - J[1] * J[8];
    a[2] = J[1] * J[5] - J[2] * J[4];
  } else if (d == 1) {
    a[0] = J[5] * J[6] - J[3] * J[8];
    a[1] = J[0] * J[8] - J[2] * J[6];
    a[2] = J[2] * J[3] - J[0] * J[5];
  } else {
    a[0] = J[3] * J[7] - J[4] * J[6];
    a[1] = J[1] * J[6] - J[0] * J[7];
    a[2] = J[0] * J[4] - J[1] * J[3];
  }
  return sqrt(a[0] * a[0] + a[1] * a[1] + a[2] * a[2]);
}

// MODE 0: dst += int_face u phi dA with u interpolated from src (boundary mass);  MODE 1: dst += scale * int_face data(q) phi dA.
// One block per face, thread (i1, i2) owns a face node, thread (q1, q2) a face quadrature point; NC components.
template <typename T, int MODE>
__global__ void face_value_kernel(int N, int NC, const uint32_t *__restrict__ cell_nodes, const int *__restrict__ face_no,
                                  const double *__restrict__ vtx, const double *__restrict__ data /*[face][N^2][NC]*/, double scale,
                                  int64_t cstride, const T *__restrict__ src, T *__restrict__ dst, const __grid_constant__ FaceShape sh) {
  extern __shared__ double fsm[];
  const int N2 = N * N, N3 = N2 * N, t = threadIdx.x;
  const int64_t face = blockIdx.x;
  double *uf = fsm;            // [NC][N^2] face-node values of src
  double *vq = fsm + NC * N2;  // [NC][N^2] JxW * value at the quadrature points
  const int fn = face_no[face], d = fn / 2, side = fn % 2;
  const int t1 = d == 0 ? 1 : 0, t2 = d == 2 ? 1 : 2;
  const int stride[3] = {1, N, N2};
  const int kf = side ? N - 1 : 0;
  uint32_t id = 0;
  if (t < N2) {
    id = cell_nodes[face * N3 + kf * stride[d] + (t % N) * stride[t1] + (t / N) * stride[t2]];
    if (MODE == 0)
      for (int c = 0; c < NC; ++c) uf[c * N2 + t] = (id & node_flag(c)) ? 0.0 : (double)src[(size_t)c * cstride + (id & NODE_MASK)];
  }
  __syncthreads();
  if (t < N2) {
    const int q1 = t % N, q2 = t / N;
    double xi[3];
    xi[d] = side;
    xi[t1] = sh.xq[q1];
    xi[t2] = sh.xq[q2];
    const double jxw = face_area_element(vtx + face * 24, d, xi) * sh.wq[q1] * sh.wq[q2];
    for (int c = 0; c < NC; ++c) {
      double val = 0.0;
      if (MODE == 0) {
        for (int i2 = 0; i2 < N; ++i2) {
          double row = 0.0;
          for (int i1 = 0; i1 < N; ++i1) row += sh.S[q1 * N + i1] * uf[c * N2 + i2 * N + i1];
          val += sh.S[q2 * N + i2] * row;
        }
      } else {
        val = scale * data[(face * N2 + t) * NC + c];
      }
      vq[c * N2 + t] = val * jxw;
    }
  }
  __syncthreads();
  if (t < N2) {
    const int i1 = t % N, i2 = t / N;
    for (int c = 0; c < NC; ++c) {
      if (id & node_flag(c)) continue;
      double acc = 0.0;
      for (int q2 = 0; q2 < N; ++q2) {
        double row = 0.0;
        for (int q1 = 0; q1 < N; ++q1) row += sh.S[q1 * N + i1] * vq[c * N2 + q2 * N + q1];
        acc += sh.S[q2 * N + i2] * row;
      }
      atomicAdd(dst + (size_t)c * cstride + (id & NODE_MASK), (T)acc);
    }
  }
}

// mask[c*cstride + node] = 1 for the DoFs on the faces of the set (DoFTools::extract_boundary_dofs on the interface id)
__global__ void face_dof_mask_kernel(int N, int NC, int64_t n_faces, const uint32_t *__restrict__ cell_nodes, const int *__restrict__ face_no,
                                     int64_t cstride, unsigned char *__restrict__ mask) {
  const int N2 = N * N, N3 = N2 * N;
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n_faces * N2) return;
  const int64_t face = idx / N2;
  const int t = (int)(idx % N2);
  const int fn = face_no[face], d = fn / 2, side = fn % 2;
  const int t1 = d == 0 ? 1 : 0, t2 = d == 2 ? 1 : 2;
  const int stride[3] = {1, N, N2};
  const uint32_t id = cell_nodes[face * N3 + (side ? N - 1 : 0) * stride[d] + (t % N) * stride[t1] + (t / N) * stride[t2]];
  for (int c = 0; c < NC; ++c) mask[(size_t)c * cstride + (id & NODE_MASK)] = 1;
}
// dst = src on the owned masked entries, 0 elsewhere (boundary_projector.h:93-95)
template <typename T>
__global__ void masked_copy_kernel(int64_t n_int, int64_t n_blocks_total, int64_t cstride, int64_t n_owned_nodes,
                                   const unsigned char *__restrict__ mask, const T *__restrict__ src, T *__restrict__ dst) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_int) return;
  const bool on = i < n_blocks_total && (i % cstride) < n_owned_nodes && mask[i];
  dst[i] = on ? src[i] : T(0);
}
template <typename T>
__global__ void fill_kernel(int64_t n, T value, T *__restrict__ v) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) v[i] = value;
}
template <typename T>
__global__ void face_identity_kernel(int64_t n, const uint32_t *__restrict__ idx, T *__restrict__ dst, const T *__restrict__ src) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  if (src)
    dst[idx[i]] += src[idx[i]];  // Base::mult_add: constrained rows act as the identity (dst was zeroed by vmult)
  else
    dst[idx[i]] = T(1);          // set_constrained_entries_to_one
}
// inverse only where the entry is positive (boundary_mass_operator.h:108-114): rows without a face keep their 0
template <typename T>
__global__ void invert_positive_kernel(int64_t n, const T *__restrict__ diag, T *__restrict__ inv) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) inv[i] = diag[i] > T(0) ? T(1) / diag[i] : diag[i];
}

}  // namespace b200mf

struct b200mf_face_set {
  b200mf_ctx *ctx;
  int64_t n_faces;
  uint32_t *cell_nodes = nullptr;
  int *face_no = nullptr;
  double *vtx = nullptr, *traction = nullptr;
  bool have_traction = false;
  FaceShape sh;
};

extern "C" int b200mf_face_set_create(b200mf_face_set **out, b200mf_ctx *ctx, int64_t n_faces, const int64_t *cells, const int32_t *face_no,
                                      const double *cell_vertices) {
  if (!out || !ctx || n_faces < 0 || (n_faces && (!cells || !face_no || !cell_vertices))) return set_error(B200MF_ERR_ARG, "null argument");
  std::unique_ptr<b200mf_face_set> fs(new b200mf_face_set());
  fs->ctx = ctx;
  fs->n_faces = n_faces;
  const int N = ctx->N, N3 = N * N * N;
  memset(&fs->sh, 0, sizeof(FaceShape));
  const Shape1D &s = ctx->shape;
  for (int i = 0; i < N * N; ++i) {
    fs->sh.S[i] = s.S[i];
    fs->sh.D[i] = s.D[i];
  }
  for (int q = 0; q < N; ++q) {
    fs->sh.xq[q] = s.xq[q];
    fs->sh.wq[q] = s.w[q];
  }
  for (int side = 0; side < 2; ++side)  // l_i'(x) at the end points of [0, 1]
    for (int i = 0; i < N; ++i) {
      const double x = side;
      double sum = 0.0;
      for (int m = 0; m < N; ++m) {
        if (m == i) continue;
        double prod = 1.0 / (s.nodes[i] - s.nodes[m]);
        for (int l = 0; l < N; ++l)
          if (l != i && l != m) prod *= (x - s.nodes[l]) / (s.nodes[i] - s.nodes[l]);
        sum += prod;
      }
      fs->sh.E[side][i] = sum;
    }
  std::vector<uint32_t> nodes((size_t)std::max<int64_t>(1, n_faces) * N3);
  for (int64_t f = 0; f < n_faces; ++f) {
    if (cells[f] < 0 || cells[f] >= ctx->n_cells) return set_error(B200MF_ERR_ARG, "face set: cell index out of range");
    if (face_no[f] < 0 || face_no[f] > 5) return set_error(B200MF_ERR_ARG, "face set: face number must be 2*direction + side");
    memcpy(&nodes[(size_t)f * N3], &(*ctx->h_cell_nodes)[(size_t)cells[f] * N3], (size_t)N3 * sizeof(uint32_t));
  }
  const size_t nf = (size_t)std::max<int64_t>(1, n_faces);
  B2_CUDA(cudaMalloc(&fs->cell_nodes, nodes.size() * sizeof(uint32_t)));
  B2_CUDA(cudaMalloc(&fs->face_no, nf * sizeof(int)));
  B2_CUDA(cudaMalloc(&fs->vtx, nf * 24 * sizeof(double)));
  B2_CUDA(cudaMalloc(&fs->traction, nf * N * N * ctx->NC * sizeof(double)));
  B2_CUDA(cudaMemset(fs->traction, 0, nf * N * N * ctx->NC * sizeof(double)));
  if (n_faces) {
    B2_CUDA(cudaMemcpy(fs->cell_nodes, nodes.data(), nodes.size() * sizeof(uint32_t), cudaMemcpyHostToDevice));
    B2_CUDA(cudaMemcpy(fs->face_no, face_no, (size_t)n_faces * sizeof(int), cudaMemcpyHostToDevice));
    B2_CUDA(cudaMemcpy(fs->vtx, cell_vertices, (size_t)n_faces * 24 * sizeof(double), cudaMemcpyHostToDevice));
  }
  *out = fs.release();
  return 0;
}

extern "C" int b200mf_face_set_destroy(b200mf_face_set *fs) {
  if (!fs) return 0;
  cudaFree(fs->cell_nodes);
  cudaFree(fs->face_no);
  cudaFree(fs->vtx);
  cudaFree(fs->traction);
  delete fs;
  return 0;
}

extern "C" int64_t b200mf_face_set_n_q_points(const b200mf_face_set *fs) { return fs ? fs->n_faces * fs->ctx->N * fs->ctx->N : -1; }

extern "C" int b200mf_face_set_quadrature_points(b200mf_face_set *fs, double *xyz_host) {
  if (!fs || !xyz_host) return set_error(B200MF_ERR_ARG, "null argument");
  const int64_t n = b200mf_face_set_n_q_points(fs);
  if (n == 0) return 0;
  double *d = nullptr;
  B2_CUDA(cudaMalloc(&d, (size_t)n * 3 * sizeof(double)));
  face_points_kernel<<<(unsigned)((n + 127) / 128), 128, 0, fs->ctx->stream>>>(fs->ctx->N, fs->n_faces, fs->face_no, fs->vtx, d, fs->sh);
  B2_LAUNCHED();
  B2_CUDA(cudaMemcpyAsync(xyz_host, d, (size_t)n * 3 * sizeof(double), cudaMemcpyDeviceToHost, fs->ctx->stream));
  B2_CUDA(cudaStreamSynchronize(fs->ctx->stream));
  cudaFree(d);
  return 0;
}

// traction at the face quadrature points [n_faces][nq^2][3] (precice_adapter->read_on_quadrature_point, :1876-1879),
// or one constant vector (the reference's test partner writes a constant stress, tests/solid/dummy_tester.cpp:193-197)
extern "C" int b200mf_face_set_set_traction(b200mf_face_set *fs, const double *traction_q, const double *traction_const) {
  if (!fs || (!traction_q && !traction_const)) return set_error(B200MF_ERR_ARG, "null argument");
  const int64_t n = b200mf_face_set_n_q_points(fs);
  const int NC = fs->ctx->NC;  // 3: traction vector (solid); 1: heat flux (scalar context)
  if (n) {
    if (traction_q) {
      B2_CUDA(cudaMemcpyAsync(fs->traction, traction_q, (size_t)n * NC * sizeof(double), cudaMemcpyHostToDevice, fs->ctx->stream));
      B2_CUDA(cudaStreamSynchronize(fs->ctx->stream));  // the host array may be reused by the adapter
    } else {
      std::vector<double> h((size_t)n * NC);
      for (int64_t i = 0; i < n; ++i)
        for (int c = 0; c < NC; ++c) h[i * NC + c] = traction_const[c];
      B2_CUDA(cudaMemcpy(fs->traction, h.data(), h.size() * sizeof(double), cudaMemcpyHostToDevice));
    }
  }
  fs->have_traction = true;
  return 0;
}

// rhs += int_faces phi * t_0 dA   (no compress, no set_zero: assemble_residual does both after cells + faces, :1899-1905)
extern "C" int b200mf_face_set_assemble_traction(b200mf_face_set *fs, void *rhs, const void *total_displacement) {
  if (!fs || !rhs || !total_displacement) return set_error(B200MF_ERR_ARG, "null argument");
  if (!fs->have_traction) return set_error(B200MF_ERR_STATE, "assemble_traction before set_traction");
  if (fs->ctx->NC != 3) return set_error(B200MF_ERR_ARG, "face set: the interface traction term needs a vector (3-component) context");
  if (fs->n_faces == 0) return 0;
  b200mf_ctx *c = fs->ctx;
  B2_TRY(update_ghost_values(c, const_cast<void *>(total_displacement)));
  const int N = c->N, N2 = N * N, N3 = N2 * N;
  const size_t smem = (size_t)(3 * N3 + 3 * N2) * sizeof(double);
  const int threads = (N2 + 31) / 32 * 32;
  if (c->dtype == B200MF_F64)
    face_traction_kernel<double><<<(unsigned)fs->n_faces, threads, smem, c->stream>>>(N, fs->n_faces, fs->cell_nodes, fs->face_no, fs->vtx, fs->traction,
                                                                                      c->cstride, (const double *)total_displacement, (double *)rhs, fs->sh);
  else
    face_traction_kernel<float><<<(unsigned)fs->n_faces, threads, smem, c->stream>>>(N, fs->n_faces, fs->cell_nodes, fs->face_no, fs->vtx, fs->traction,
                                                                                     c->cstride, (const float *)total_displacement, (float *)rhs, fs->sh);
  B2_LAUNCHED();
  return 0;
}

// ---------------------------------------------------------------------------
// value integrals: coupling flux and boundary mass operator
template <int MODE>
static int launch_face_value(b200mf_face_set *fs, void *dst, const void *src, double scale) {
  if (fs->n_faces == 0) return 0;
  b200mf_ctx *c = fs->ctx;
  const int N = c->N, N2 = N * N, NC = c->NC;
  const size_t smem = (size_t)2 * NC * N2 * sizeof(double);
  const int threads = (N2 + 31) / 32 * 32;
  if (c->dtype == B200MF_F64)
    face_value_kernel<double, MODE><<<(unsigned)fs->n_faces, threads, smem, c->stream>>>(N, NC, fs->cell_nodes, fs->face_no, fs->vtx, fs->traction, scale,
                                                                                         c->cstride, (const double *)src, (double *)dst, fs->sh);
  else
    face_value_kernel<float, MODE><<<(unsigned)fs->n_faces, threads, smem, c->stream>>>(N, NC, fs->cell_nodes, fs->face_no, fs->vtx, fs->traction, scale,
                                                                                        c->cstride, (const float *)src, (float *)dst, fs->sh);
  B2_LAUNCHED();
  return 0;
}

// rhs += scale * int_faces data(q) phi dA with the data of b200mf_face_set_set_traction (heat: flux at the face quadrature
// points, scale = -delta_t, heat_transfer.h:636-647); no compress (the caller's assemble_rhs does it, :654)
extern "C" int b200mf_face_set_assemble_flux(b200mf_face_set *fs, void *rhs, double scale) {
  if (!fs || !rhs) return set_error(B200MF_ERR_ARG, "null argument");
  if (!fs->have_traction) return set_error(B200MF_ERR_STATE, "assemble_flux before set_traction");
  return launch_face_value<1>(fs, rhs, nullptr, scale);
}

struct b200mf_boundary_mass {
  b200mf_face_set *fs;
  void *diag = nullptr, *inv = nullptr;
  unsigned char *mask = nullptr;
  bool have_diag = false;
};

static size_t bm_esize(const b200mf_ctx *c) { return c->dtype == B200MF_F64 ? 8 : 4; }
static unsigned bm_blocks(int64_t n) { return (unsigned)((n + 255) / 256); }

extern "C" int b200mf_boundary_mass_create(b200mf_boundary_mass **out, b200mf_face_set *fs) {
  if (!out || !fs) return set_error(B200MF_ERR_ARG, "null argument");
  std::unique_ptr<b200mf_boundary_mass> op(new b200mf_boundary_mass());
  op->fs = fs;
  b200mf_ctx *c = fs->ctx;
  B2_CUDA(cudaMalloc(&op->mask, (size_t)c->n_int));
  B2_CUDA(cudaMemsetAsync(op->mask, 0, (size_t)c->n_int, c->stream));
  if (fs->n_faces) {
    const int64_t n = fs->n_faces * c->N * c->N;
    face_dof_mask_kernel<<<bm_blocks(n), 256, 0, c->stream>>>(c->N, c->NC, fs->n_faces, fs->cell_nodes, fs->face_no, c->cstride, op->mask);
    B2_LAUNCHED();
  }
  *out = op.release();
  return 0;
}
extern "C" int b200mf_boundary_mass_destroy(b200mf_boundary_mass *op) {
  if (!op) return 0;
  cudaFree(op->diag);
  cudaFree(op->inv);
  cudaFree(op->mask);
  delete op;
  return 0;
}
extern "C" b200mf_ctx *b200mf_boundary_mass_ctx(b200mf_boundary_mass *op) { return op ? op->fs->ctx : nullptr; }

static int bm_apply(b200mf_boundary_mass *op, void *dst, const void *src, bool zero_dst) {
  if (!op || !dst || !src) return set_error(B200MF_ERR_ARG, "null argument");
  if (dst == src) return set_error(B200MF_ERR_ARG, "vmult: dst aliases src");
  b200mf_ctx *c = op->fs->ctx;
  if (zero_dst)
    B2_CUDA(cudaMemsetAsync(dst, 0, (size_t)c->n_int * bm_esize(c), c->stream));
  else
    B2_TRY(zero_ghosts(c, dst));
  B2_TRY(update_ghost_values(c, const_cast<void *>(src)));  // boundary_mass_operator.h:143
  B2_TRY(launch_face_value<0>(op->fs, dst, src, 1.0));
  B2_TRY(compress_add(c, dst));                             // :159
  if (c->n_constrained) {  // MatrixFreeOperators::Base::vmult: constrained rows act as the identity
    if (c->dtype == B200MF_F64)
      face_identity_kernel<double><<<bm_blocks(c->n_constrained), 256, 0, c->stream>>>(c->n_constrained, c->constrained, (double *)dst, (const double *)src);
    else
      face_identity_kernel<float><<<bm_blocks(c->n_constrained), 256, 0, c->stream>>>(c->n_constrained, c->constrained, (float *)dst, (const float *)src);
    B2_LAUNCHED();
  }
  return 0;
}
extern "C" int b200mf_boundary_mass_vmult(b200mf_boundary_mass *op, void *dst, const void *src) { return bm_apply(op, dst, src, true); }
extern "C" int b200mf_boundary_mass_vmult_add(b200mf_boundary_mass *op, void *dst, const void *src) { return bm_apply(op, dst, src, false); }

extern "C" int b200mf_boundary_mass_compute_diagonal(b200mf_boundary_mass *op) {
  if (!op) return set_error(B200MF_ERR_ARG, "null argument");
  b200mf_ctx *c = op->fs->ctx;
  const size_t bytes = (size_t)c->n_int * bm_esize(c);
  if (!op->diag) B2_CUDA(cudaMalloc(&op->diag, bytes));
  if (!op->inv) B2_CUDA(cudaMalloc(&op->inv, bytes));
  const int64_t n = c->n_int;
  B2_CUDA(cudaMemsetAsync(op->diag, 0, bytes, c->stream));
  // inverse_diagonal_vector = 1; apply_add(diagonal_vector, inverse_diagonal_vector)   (:96-97: row sums, not the diagonal)
  if (c->dtype == B200MF_F64)
    fill_kernel<double><<<bm_blocks(n), 256, 0, c->stream>>>(n, 1.0, (double *)op->inv);
  else
    fill_kernel<float><<<bm_blocks(n), 256, 0, c->stream>>>(n, 1.0f, (float *)op->inv);
  B2_LAUNCHED();
  B2_TRY(launch_face_value<0>(op->fs, op->diag, op->inv, 1.0));
  B2_TRY(compress_add(c, op->diag));
  if (c->n_constrained) {  // set_constrained_entries_to_one (:99)
    if (c->dtype == B200MF_F64)
      face_identity_kernel<double><<<bm_blocks(c->n_constrained), 256, 0, c->stream>>>(c->n_constrained, c->constrained, (double *)op->diag, nullptr);
    else
      face_identity_kernel<float><<<bm_blocks(c->n_constrained), 256, 0, c->stream>>>(c->n_constrained, c->constrained, (float *)op->diag, nullptr);
    B2_LAUNCHED();
  }
  if (c->dtype == B200MF_F64)
    invert_positive_kernel<double><<<bm_blocks(n), 256, 0, c->stream>>>(n, (const double *)op->diag, (double *)op->inv);
  else
    invert_positive_kernel<float><<<bm_blocks(n), 256, 0, c->stream>>>(n, (const float *)op->diag, (float *)op->inv);
  B2_LAUNCHED();
  B2_TRY(update_ghost_values(c, op->inv));   // :116-117
  B2_TRY(update_ghost_values(c, op->diag));
  op->have_diag = true;
  return 0;
}
extern "C" int b200mf_boundary_mass_diagonal(b200mf_boundary_mass *op, const void **diag, const void **inv) {
  if (!op) return set_error(B200MF_ERR_ARG, "null argument");
  if (!op->have_diag) return set_error(B200MF_ERR_STATE, "diagonal requested before compute_diagonal()");
  if (diag) *diag = op->diag;
  if (inv) *inv = op->inv;
  return 0;
}
// dst = src on the locally owned DoFs of the interface faces, 0 elsewhere (boundary_projector.h:93-95)
extern "C" int b200mf_boundary_mass_restrict(b200mf_boundary_mass *op, void *dst, const void *src) {
  if (!op || !dst || !src) return set_error(B200MF_ERR_ARG, "null argument");
  b200mf_ctx *c = op->fs->ctx;
  const int64_t n = c->n_int;
  if (c->dtype == B200MF_F64)
    masked_copy_kernel<double><<<bm_blocks(n), 256, 0, c->stream>>>(n, c->cstride * c->NC, c->cstride, c->n_owned_nodes, op->mask, (const double *)src, (double *)dst);
  else
    masked_copy_kernel<float><<<bm_blocks(n), 256, 0, c->stream>>>(n, c->cstride * c->NC, c->cstride, c->n_owned_nodes, op->mask, (const float *)src, (float *)dst);
  B2_LAUNCHED();
  return 0;
}
