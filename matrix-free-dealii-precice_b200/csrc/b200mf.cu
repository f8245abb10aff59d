// libb200mf: C ABI (include/b200mf.h) over the sm_100a kernels.
// Host logic only in this file: numbering, table layout, launches.
#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <numeric>

#include "common.h"
#include "kernels.h"

namespace b200mf {

thread_local std::string g_last_error;
std::atomic<int64_t> g_launches{0};
std::atomic<uint64_t> g_config_epoch{1};

int set_error(int code, const char *fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  g_last_error = buf;
  return code;
}

// ---------------------------------------------------------------------------
// 1-D basis (host, long double Newton iterations)
static void legendre(int n, long double x, long double &P, long double &dP, long double &ddP) {
  long double p0 = 1.0L, p1 = x;
  if (n == 0) {
    P = 1;
    dP = ddP = 0;
    return;
  }
  for (int k = 2; k <= n; ++k) {
    long double pk = ((2.0L * k - 1.0L) * x * p1 - (k - 1.0L) * p0) / k;
    p0 = p1;
    p1 = pk;
  }
  P = p1;
  dP = n * (x * p1 - p0) / (x * x - 1.0L);
  ddP = (2.0L * x * dP - n * (n + 1.0L) * p1) / (1.0L - x * x);
}

static void lagrange(int n, const long double *nodes, long double x, long double *v, long double *d) {
  for (int i = 0; i < n; ++i) {
    long double den = 1, num = 1, s = 0;
    for (int j = 0; j < n; ++j)
      if (j != i) {
        den *= nodes[i] - nodes[j];
        num *= x - nodes[j];
      }
    for (int m = 0; m < n; ++m)
      if (m != i) {
        long double t = 1;
        for (int j = 0; j < n; ++j)
          if (j != i && j != m) t *= x - nodes[j];
        s += t;
      }
    v[i] = num / den;
    d[i] = s / den;
  }
}

void compute_shape_1d(int p, Shape1D &out) {
  const int n = p + 1;
  const long double pi = 3.141592653589793238462643383279502884L;
  long double nd[8], xq[8], wq[8];
  nd[0] = -1;
  nd[p] = 1;
  for (int i = 1; i < p; ++i) {
    long double z = -cosl(pi * i / p);
    for (int it = 0; it < 100; ++it) {
      long double P, dP, ddP;
      legendre(p, z, P, dP, ddP);
      long double dz = dP / ddP;
      z -= dz;
      if (fabsl(dz) < 1e-19L) break;
    }
    nd[i] = z;
  }
  for (int i = 0; i <= p / 2; ++i) {
    long double a = 0.5L * (nd[p - i] - nd[i]);
    nd[i] = -a;
    nd[p - i] = a;
  }
  for (int i = 0; i < n; ++i) {
    long double z = -cosl(pi * (i + 0.75L) / (n + 0.5L));
    long double P, dP, ddP;
    for (int it = 0; it < 100; ++it) {
      legendre(n, z, P, dP, ddP);
      long double dz = P / dP;
      z -= dz;
      if (fabsl(dz) < 1e-19L) break;
    }
    legendre(n, z, P, dP, ddP);
    xq[i] = z;
    wq[i] = 2.0L / ((1.0L - z * z) * dP * dP);
  }
  for (int i = 0; i < n / 2; ++i) {
    long double a = 0.5L * (xq[n - 1 - i] - xq[i]);
    xq[i] = -a;
    xq[n - 1 - i] = a;
    long double ww = 0.5L * (wq[i] + wq[n - 1 - i]);
    wq[i] = wq[n - 1 - i] = ww;
  }
  if (n % 2) xq[n / 2] = 0;
  for (int i = 0; i < n; ++i) {
    nd[i] = 0.5L * (nd[i] + 1.0L);
    xq[i] = 0.5L * (xq[i] + 1.0L);
    wq[i] *= 0.5L;
  }
  out.n = n;
  long double v[8], d[8];
  for (int q = 0; q < n; ++q) {
    lagrange(n, nd, xq[q], v, d);
    for (int i = 0; i < n; ++i) {
      out.S[q * n + i] = (double)v[i];
      out.D[q * n + i] = (double)d[i];
    }
    lagrange(n, xq, xq[q], v, d);
    for (int r = 0; r < n; ++r) out.C[q * n + r] = (double)d[r];
    out.w[q] = (double)wq[q];
    out.xq[q] = (double)xq[q];
  }
  for (int i = 0; i < n; ++i) out.nodes[i] = (double)nd[i];
}

// host copies of the 1-D matrices every kernel translation unit uploads into its own __constant__ bank (launch.cuh)
void shape_tables(const double **hd_out, const float **hf_out) {
  static double hd[MAXN + 1][SHAPE_ROWS][MAXN * MAXN];
  static float hf[MAXN + 1][SHAPE_ROWS][MAXN * MAXN];
  static bool done = false;
  if (!done) {
    memset(hd, 0, sizeof hd);
    for (int N = 2; N <= MAXN; ++N) {
      Shape1D s;
      compute_shape_1d(N - 1, s);
      for (int i = 0; i < N * N; ++i) {
        hd[N][0][i] = s.S[i];
        hd[N][1][i] = s.C[i];
        hd[N][2][i] = s.D[i];  // derivative of the nodal basis at the Gauss points (= C S)
      }
      for (int q = 0; q < N; ++q) hd[N][3][q] = s.w[q];
    }
    for (int N = 0; N <= MAXN; ++N)
      for (int m = 0; m < SHAPE_ROWS; ++m)
        for (int i = 0; i < MAXN * MAXN; ++i) hf[N][m][i] = (float)hd[N][m][i];
    done = true;
  }
  *hd_out = &hd[0][0][0];
  *hf_out = &hf[0][0][0];
}

// ---------------------------------------------------------------------------
// small kernels
struct QuadPts {
  double xq[MAXN], wq[MAXN];
};

// Q1 metric of StaticMappingQ1 at the Gauss points, written straight into the group layout.
// invjac[e][f] = d xi_f / d x_e (mf_nh_operator.h:872-887), JxW = det J * w.
template <typename T>
__global__ void metric_from_vertices_kernel(int N, int cpg, int L, TableGeom tg, int64_t n_cells, int64_t n_cells_padded,
                                            const double *__restrict__ vtx, QuadPts qp, T *__restrict__ mref) {
  const int N2 = N * N, N3 = N2 * N;
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n_cells_padded * N3) return;
  const int64_t cell = idx / N3;
  const int q = (int)(idx % N3);
  const int iq = q % N, jq = (q / N) % N, kq = q / N2;
  const int64_t group = cell / cpg;
  const int lane = (int)(cell % cpg) * N2 + jq * N + iq;  // column thread (cell, j, i) owns the q-points (i, j, kq)
  double Ki[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1}, jxw = 0.0;
  if (cell < n_cells) {
    const double *v = vtx + cell * 24;
    const double xi[3] = {qp.xq[iq], qp.xq[jq], qp.xq[kq]};
    double J[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
    for (int m = 0; m < 8; ++m) {
      double l[3], dl[3];
      for (int d = 0; d < 3; ++d) {
        const int bit = (m >> d) & 1;
        l[d] = bit ? xi[d] : 1.0 - xi[d];
        dl[d] = bit ? 1.0 : -1.0;
      }
      const double g[3] = {dl[0] * l[1] * l[2], l[0] * dl[1] * l[2], l[0] * l[1] * dl[2]};
      for (int e = 0; e < 3; ++e)
        for (int f = 0; f < 3; ++f) J[e * 3 + f] += v[m * 3 + e] * g[f];
    }
    const double det = J[0] * (J[4] * J[8] - J[5] * J[7]) - J[1] * (J[3] * J[8] - J[5] * J[6]) + J[2] * (J[3] * J[7] - J[4] * J[6]);
    const double id = 1.0 / det;
    double Ji[9];  // Ji[f][e] = d xi_f / d x_e
    Ji[0] = (J[4] * J[8] - J[5] * J[7]) * id;
    Ji[1] = (J[2] * J[7] - J[1] * J[8]) * id;
    Ji[2] = (J[1] * J[5] - J[2] * J[4]) * id;
    Ji[3] = (J[5] * J[6] - J[3] * J[8]) * id;
    Ji[4] = (J[0] * J[8] - J[2] * J[6]) * id;
    Ji[5] = (J[2] * J[3] - J[0] * J[5]) * id;
    Ji[6] = (J[3] * J[7] - J[4] * J[6]) * id;
    Ji[7] = (J[1] * J[6] - J[0] * J[7]) * id;
    Ji[8] = (J[0] * J[4] - J[1] * J[3]) * id;
    for (int e = 0; e < 3; ++e)
      for (int f = 0; f < 3; ++f) Ki[e * 3 + f] = Ji[f * 3 + e];
    jxw = det * qp.wq[iq] * qp.wq[jq] * qp.wq[kq];
  }
  const size_t FS = (size_t)tg.fstride;
  T *o = mref + (size_t)group * tg.gstride + (size_t)kq * tg.kqstride + lane;
  for (int i = 0; i < 9; ++i) o[(MF_K + i) * FS] = (T)Ki[i];
  o[MF_W * FS] = (T)jxw;
}

// natural [cell][q][nf] (double) -> group layout T (plain [group][nf][N][L] or sliced [group][N][nf][L], see table_geom);
// padded cells get `pad`-pattern (identity K)
template <typename T>
__global__ void table_to_group_layout_kernel(int N, int cpg, int L, int nf, TableGeom tg, int64_t n_cells, int64_t n_cells_padded,
                                             const double *__restrict__ in, T *__restrict__ out, int identity_pad,
                                             const int64_t *__restrict__ order = nullptr) {
  const int N2 = N * N, N3 = N2 * N;
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n_cells_padded * N3) return;
  const int64_t cell = idx / N3;
  const int q = (int)(idx % N3);
  const int iq = q % N, jq = (q / N) % N, kq = q / N2;
  const int64_t group = cell / cpg;
  const int lane = (int)(cell % cpg) * N2 + jq * N + iq;  // column thread (cell, j, i) owns the q-points (i, j, kq)
  const size_t FS = (size_t)tg.fstride;
  T *o = out + (size_t)group * tg.gstride + (size_t)kq * tg.kqstride + lane;
  for (int f = 0; f < nf; ++f) {
    double v;
    if (cell < n_cells)
      v = in[((size_t)(order ? order[cell] : cell) * N3 + q) * nf + f];
    else
      v = (identity_pad && (f == 0 || f == 4 || f == 8)) ? 1.0 : 0.0;
    o[f * FS] = (T)v;
  }
}

// group layout -> natural [cell][q] of one field (test hook)
template <typename T>
__global__ void field_from_group_layout_kernel(int N, int cpg, int L, TableGeom tg, int field, int64_t n_cells,
                                               const T *__restrict__ in, double *__restrict__ out, const int64_t *__restrict__ order = nullptr) {
  const int N2 = N * N, N3 = N2 * N;
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n_cells * N3) return;
  const int64_t cell = idx / N3;
  const int q = (int)(idx % N3);
  const int iq = q % N, jq = (q / N) % N, kq = q / N2;
  const int64_t group = cell / cpg;
  const int lane = (int)(cell % cpg) * N2 + jq * N + iq;  // column thread (cell, j, i) owns the q-points (i, j, kq)
  out[(size_t)(order ? order[cell] : cell) * N3 + q] = (double)in[(size_t)group * tg.gstride + (size_t)kq * tg.kqstride + (size_t)field * tg.fstride + lane];
}

template <typename TO, typename TI>
__global__ void permute_in_kernel(int64_t n, const int64_t *__restrict__ int2ext, const TI *__restrict__ ext, TO *__restrict__ in) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) {
    const int64_t e = int2ext ? int2ext[i] : i;
    in[i] = e >= 0 ? (TO)ext[e] : TO(0);  // e < 0: padding slot between component blocks
  }
}
template <typename TO, typename TI>
__global__ void permute_out_kernel(int64_t n, const int64_t *__restrict__ int2ext, TO *__restrict__ ext, const TI *__restrict__ in) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) {
    const int64_t e = int2ext ? int2ext[i] : i;
    if (e >= 0) ext[e] = (TO)in[i];
  }
}

// dst[c] += src[c] on constrained rows (mf_nh_operator.h:623-624; MatrixFreeOperators::Base::vmult)
template <typename T>
__global__ void constrained_identity_kernel(int64_t n, const uint32_t *__restrict__ idx, T *__restrict__ dst, const T *__restrict__ src,
                                            double *dot) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  double s = 0.0;
  if (i < n) {
    const T v = src[idx[i]];
    dst[idx[i]] += v;
    s = (double)v * (double)v;
  }
  if (dot != nullptr) {  // the identity rows' share of src . (A src)
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_down_sync(0xffffffffu, s, o);
    if ((threadIdx.x & 31) == 0 && s != 0.0) atomicAdd(dot, s);
  }
}

// diag[constrained] = 1 (mf_nh_operator.h:1043-1044, heat_transfer.h:220)
template <typename T>
__global__ void set_constrained_one_kernel(int64_t n, const uint32_t *__restrict__ idx, T *__restrict__ diag) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) diag[idx[i]] = T(1);
}
// inv = 1/diag; counts non-positive entries (debug assertion mf_nh_operator.h:1049-1051)
template <typename T>
__global__ void invert_diagonal_kernel(int64_t n, int64_t cstride, int64_t n_nodes, const T *__restrict__ diag, T *__restrict__ inv,
                                       unsigned long long *n_bad) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) {
    if (i % cstride >= n_nodes) {  // padding slot, or a ghost entry (filled by update_ghost_values afterwards, :1056-1057)
      inv[i] = T(0);
      return;
    }
    const T d = diag[i];
    if (!(d > T(0))) atomicAdd(n_bad, 1ull);
    inv[i] = T(1) / d;
  }
}
// precondition_Jacobi: dst = omega * D^-1 src (mf_nh_operator.h:281-291)
template <typename T>
__global__ void jacobi_kernel(int64_t n, T *__restrict__ dst, const T *__restrict__ src, const T *__restrict__ inv, T omega) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dst[i] = omega * inv[i] * src[i];
}

static inline unsigned nblocks(int64_t n, int bs) { return (unsigned)((n + bs - 1) / bs); }

static inline size_t esize(const b200mf_ctx *c) { return c->dtype == B200MF_F64 ? 8 : 4; }
// geometry of a staged table (== Layout::table_stride / kq_stride / FS) or, plain = true, of a table that is only ever
// read or written straight from/to global memory (== Layout::plain_stride / L / FSP)
static inline TableGeom table_geom_rt(const b200mf_ctx *c, int NF, bool plain = false) {
  return table_geom(!plain && table_sliced(c->N, c->NC), c->N, c->L, NF, (int)esize(c));
}
static inline int64_t table_stride_rt(const b200mf_ctx *c, int NF, bool plain = false) { return table_geom_rt(c, NF, plain).gstride; }

}  // namespace b200mf

using namespace b200mf;

// ===========================================================================
// ABI: misc
extern "C" const char *b200mf_last_error(void) { return g_last_error.c_str(); }
extern "C" int64_t b200mf_launch_count(void) { return g_launches.load(); }
extern "C" int b200mf_version(void) { return 100; }

// ===========================================================================
// ABI: context
extern "C" int b200mf_ctx_create(b200mf_ctx **out, const b200mf_ctx_desc *d) {
  if (!out || !d) return set_error(B200MF_ERR_ARG, "null argument");
  *out = nullptr;
  if (d->degree < 1 || d->degree > 6) return set_error(B200MF_ERR_ARG, "degree %d outside 1..6", d->degree);
  if (d->n_q_points_1d != d->degree + 1)
    return set_error(B200MF_ERR_ARG, "n_q_points_1d must be degree+1 (got %d for degree %d)", d->n_q_points_1d, d->degree);
  if (d->n_components != 1 && d->n_components != 3) return set_error(B200MF_ERR_ARG, "n_components must be 1 or 3");
  if (d->dtype != B200MF_F64 && d->dtype != B200MF_F32) return set_error(B200MF_ERR_ARG, "bad dtype");
  if ((d->dof_indices == nullptr) == (d->node_indices == nullptr))
    return set_error(B200MF_ERR_ARG, "exactly one of dof_indices / node_indices must be given");
  if (d->n_cells < 0 || d->n_dofs <= 0 || d->n_owned < 0 || d->n_owned > d->n_dofs) return set_error(B200MF_ERR_ARG, "bad sizes");
  if (!d->vertices && !(d->inverse_jacobian && d->jxw)) return set_error(B200MF_ERR_ARG, "need vertices or inverse_jacobian+jxw");
  if (d->n_constrained > 0 && !d->constrained_dofs) return set_error(B200MF_ERR_ARG, "constrained_dofs is null");

  // every field starts null and b200mf_ctx_destroy frees them all: an early error return below releases what was allocated so far
  std::unique_ptr<b200mf_ctx, int (*)(b200mf_ctx *)> c(new b200mf_ctx(), b200mf_ctx_destroy);
  memset(c.get(), 0, sizeof(b200mf_ctx));
  c->degree = d->degree;
  c->N = d->degree + 1;
  c->NC = d->n_components;
  c->dtype = d->dtype;
  c->n_cells = d->n_cells;
  c->n_dofs = d->n_dofs;
  c->n_owned = d->n_owned;
  // exchange schedule of vmult on a partitioned context: 0 = the reference's serial order update_ghost_values -> cells ->
  // compress(add); 1 = exchanges on a second stream behind the interior cells.  b200mf_ctx_set_partitioner picks the
  // default (1 with the peer-to-peer exchange, 0 with NCCL send/recv, comm.cu); b200mf_ctx_set_overlap overrides it
  c->overlap = 0;
  const int N = c->N, N2 = N * N, N3 = N2 * N, NC = c->NC;
  c->cpg = group_cells(N, NC);
  c->L = c->cpg * N2;
  c->n_groups = (c->n_cells + c->cpg - 1) / c->cpg;
  compute_shape_1d(d->degree, c->shape);
  int dev = 0;
  B2_CUDA(cudaGetDevice(&dev));
  B2_CUDA(cudaDeviceGetAttribute(&c->sm_count, cudaDevAttrMultiProcessorCount, dev));

  // ---- numbering: nodes, internal <-> caller permutation ---------------------
  // internal vector layout: component blocks, DoF (node, c) at c*cstride + node; DoFs no local cell
  // touches ("orphans", e.g. extra ghost slots) are appended after the blocks.
  std::vector<uint32_t> node_of_cell((size_t)c->n_cells * N3);
  std::vector<int64_t> ext_of_nodecomp;  // [node*NC + c] -> caller DoF
  std::vector<int64_t> orphans;
  if (d->node_indices) {
    if (d->n_dofs % NC) return set_error(B200MF_ERR_ARG, "n_dofs not divisible by n_components");
    c->n_nodes = d->n_dofs / NC;
    for (size_t i = 0; i < node_of_cell.size(); ++i) {
      if ((int64_t)d->node_indices[i] >= c->n_nodes) return set_error(B200MF_ERR_ARG, "node index out of range");
      node_of_cell[i] = d->node_indices[i];
    }
    ext_of_nodecomp.resize(d->n_dofs);
    for (int64_t i = 0; i < d->n_dofs; ++i) ext_of_nodecomp[i] = i;
  } else {
    // group the DoFs of the NC components that share a support point into nodes
    const size_t ld = (size_t)NC * N3;
    std::vector<int8_t> comp(d->n_dofs, -1);
    std::vector<int64_t> lead(d->n_dofs, -1);  // DoF -> component-0 DoF of its node
    for (int64_t cell = 0; cell < c->n_cells; ++cell)
      for (int a = 0; a < N3; ++a) {
        const uint32_t d0 = d->dof_indices[cell * ld + a];
        for (int cc = 0; cc < NC; ++cc) {
          const uint32_t dc = d->dof_indices[cell * ld + (size_t)cc * N3 + a];
          if ((int64_t)dc >= d->n_dofs) return set_error(B200MF_ERR_ARG, "dof index %u out of range", dc);
          if (comp[dc] == -1) {
            comp[dc] = (int8_t)cc;
            lead[dc] = d0;
          } else if (comp[dc] != cc || lead[dc] != (int64_t)d0) {
            return set_error(B200MF_ERR_ARG, "dof_indices: components of a support point are not consistent across cells (dof %u)", dc);
          }
        }
      }
    // node ids in ascending order of the component-0 DoF (keeps owned before ghost)
    std::vector<int64_t> node_id(d->n_dofs, -1);
    int64_t nn = 0;
    for (int64_t i = 0; i < d->n_dofs; ++i)
      if (comp[i] == 0) node_id[i] = nn++;
    c->n_nodes = nn;
    ext_of_nodecomp.assign(nn * NC, -1);
    for (int64_t i = 0; i < d->n_dofs; ++i) {
      if (comp[i] == -1)
        orphans.push_back(i);
      else
        ext_of_nodecomp[node_id[lead[i]] * NC + comp[i]] = i;
    }
    for (int64_t cell = 0; cell < c->n_cells; ++cell)
      for (int a = 0; a < N3; ++a) node_of_cell[cell * N3 + a] = (uint32_t)node_id[d->dof_indices[cell * ld + a]];
  }
  if (c->n_nodes > (int64_t)NODE_MASK) return set_error(B200MF_ERR_ARG, "too many nodes per rank (%lld)", (long long)c->n_nodes);
  c->cstride = (int64_t)round_up((size_t)c->n_nodes, 32);
  c->n_int = c->cstride * NC + (int64_t)orphans.size();
  if (c->n_int >= (int64_t)1 << 32) return set_error(B200MF_ERR_ARG, "more than 2^32 local DoFs");
  // scalar contexts with the caller's own numbering need no translation
  bool identity = (NC == 1 && orphans.empty());
  if (identity)
    for (int64_t i = 0; i < c->n_nodes; ++i)
      if (ext_of_nodecomp[i] != i) {
        identity = false;
        break;
      }
  std::vector<int64_t> int2ext, ext2int;
  if (!identity) {
    int2ext.assign(c->n_int, -1);
    ext2int.assign(d->n_dofs, -1);
    for (int64_t node = 0; node < c->n_nodes; ++node)
      for (int cc = 0; cc < NC; ++cc) {
        const int64_t e = ext_of_nodecomp[node * NC + cc], in = cc * c->cstride + node;
        int2ext[in] = e;
        if (e >= 0) ext2int[e] = in;
      }
    for (size_t k = 0; k < orphans.size(); ++k) {
      int2ext[c->cstride * NC + (int64_t)k] = orphans[k];
      ext2int[orphans[k]] = c->cstride * NC + (int64_t)k;
    }
    // owned nodes must precede ghost nodes inside every component block
    for (int64_t node = 0; node + 1 < c->n_nodes; ++node) {
      const bool o0 = ext_of_nodecomp[node * NC] < d->n_owned, o1 = ext_of_nodecomp[(node + 1) * NC] < d->n_owned;
      if (!o0 && o1) return set_error(B200MF_ERR_ARG, "owned DoFs must precede ghost DoFs");
    }
  }
  c->n_owned_nodes = 0;
  for (int64_t node = 0; node < c->n_nodes; ++node)
    if (ext_of_nodecomp[node * NC] < d->n_owned) ++c->n_owned_nodes;

  // ---- cell order: cells that touch a ghost node first ("boundary" cells need the ghost values of src and
  //      produce ghost contributions of dst), so that the two exchanges hide behind the interior cells
  std::vector<int64_t> order;  // position in the library's cell order -> caller's cell
  if (c->n_owned_nodes < c->n_nodes) {
    std::vector<int64_t> bnd, inner;
    for (int64_t cell = 0; cell < c->n_cells; ++cell) {
      bool b = false;
      for (int a = 0; a < N3 && !b; ++a) b = (int64_t)node_of_cell[cell * N3 + a] >= c->n_owned_nodes;
      (b ? bnd : inner).push_back(cell);
    }
    c->n_boundary_groups = ((int64_t)bnd.size() + c->cpg - 1) / c->cpg;
    order = bnd;
    order.insert(order.end(), inner.begin(), inner.end());
  }
  auto caller_cell = [&](int64_t cell) -> int64_t { return order.empty() ? cell : order[cell]; };

  std::vector<uint8_t> nflags(c->n_nodes, 0);
  std::vector<uint32_t> cons_int(d->n_constrained);
  for (int64_t i = 0; i < d->n_constrained; ++i) {
    const uint32_t e = d->constrained_dofs[i];
    if ((int64_t)e >= d->n_dofs) return set_error(B200MF_ERR_ARG, "constrained dof out of range");
    const int64_t in = identity ? (int64_t)e : ext2int[e];
    cons_int[i] = (uint32_t)in;
    if (in < c->cstride * NC) nflags[in % c->cstride] |= (uint8_t)(1u << (in / c->cstride));
  }
  // the index list drives dst[c] += src[c], diag[c] = 1 and constraints.set_zero: locally owned rows only, as
  // MatrixFree::get_constrained_dofs() (ghost rows belong to their owner; the per-node flags above cover them too)
  {
    std::vector<uint32_t> owned_only;
    for (uint32_t in : cons_int)
      if ((int64_t)in < c->cstride * NC ? (int64_t)(in % c->cstride) < c->n_owned_nodes
                                        : orphans[in - c->cstride * NC] < d->n_owned)
        owned_only.push_back(in);
    cons_int.swap(owned_only);
  }
  c->n_constrained = (int64_t)cons_int.size();
  if (!identity) c->h_ext2int = new std::vector<int64_t>(ext2int);

  // ---- node index table in group layout [group][k][lane] ---------------------
  {
    const int L = c->L, cpg = c->cpg;
    std::vector<uint32_t> nidx((size_t)c->n_groups * N * L, NODE_ALL_CONSTRAINED);
    for (int64_t cell = 0; cell < c->n_cells; ++cell) {
      const int64_t g = cell / cpg;
      const int cw = (int)(cell % cpg);
      for (int k = 0; k < N; ++k)
        for (int ji = 0; ji < N2; ++ji) {
          const uint32_t node = node_of_cell[caller_cell(cell) * N3 + k * N2 + ji];
          nidx[((size_t)g * N + k) * L + cw * N2 + ji] = node | ((uint32_t)nflags[node] << 29);
        }
    }
    c->h_cell_nodes = new std::vector<uint32_t>(node_of_cell.size());
    for (size_t i = 0; i < node_of_cell.size(); ++i) (*c->h_cell_nodes)[i] = node_of_cell[i] | ((uint32_t)nflags[node_of_cell[i]] << 29);
    B2_CUDA(cudaMalloc(&c->nidx, nidx.size() * sizeof(uint32_t)));
    B2_CUDA(cudaMemcpy(c->nidx, nidx.data(), nidx.size() * sizeof(uint32_t), cudaMemcpyHostToDevice));
    c->bytes += nidx.size() * sizeof(uint32_t);
  }
  if (!identity) {
    B2_CUDA(cudaMalloc(&c->int2ext, int2ext.size() * sizeof(int64_t)));
    B2_CUDA(cudaMemcpy(c->int2ext, int2ext.data(), int2ext.size() * sizeof(int64_t), cudaMemcpyHostToDevice));
    c->bytes += int2ext.size() * sizeof(int64_t);
  }
  if (c->n_constrained) {
    B2_CUDA(cudaMalloc(&c->constrained, cons_int.size() * sizeof(uint32_t)));
    B2_CUDA(cudaMemcpy(c->constrained, cons_int.data(), cons_int.size() * sizeof(uint32_t), cudaMemcpyHostToDevice));
    c->bytes += cons_int.size() * sizeof(uint32_t);
  }
  if (d->material_id) {
    std::vector<uint8_t> mid(std::max<int64_t>(1, c->n_cells));
    for (int64_t cell = 0; cell < c->n_cells; ++cell) mid[cell] = d->material_id[caller_cell(cell)];
    B2_CUDA(cudaMalloc(&c->matid, mid.size()));
    B2_CUDA(cudaMemcpy(c->matid, mid.data(), mid.size(), cudaMemcpyHostToDevice));
    c->bytes += c->n_cells;
  }
  if (!order.empty()) {
    B2_CUDA(cudaMalloc(&c->cell_order, order.size() * sizeof(int64_t)));
    B2_CUDA(cudaMemcpy(c->cell_order, order.data(), order.size() * sizeof(int64_t), cudaMemcpyHostToDevice));
    c->bytes += order.size() * sizeof(int64_t);
  }
  B2_CUDA(cudaMalloc(&c->stage, std::max<int64_t>(c->n_dofs, c->n_int) * sizeof(double)));
  c->bytes += std::max<int64_t>(c->n_dofs, c->n_int) * sizeof(double);

  // ---- per-cell metric when every cell is affine (deal.II compresses Cartesian/affine cells to one Jacobian
  //      per cell): M = det J (K^T K), det J
  {
    const int64_t padded = c->n_groups * c->cpg;
    std::vector<double> mc((size_t)padded * MC_COUNT, 0.0);
    bool affine = true;
    auto fill = [&](int64_t cell, const double *Ki /* K[e][f] = d xi_f / d x_e */, double det) {
      const int I6[6] = {0, 1, 2, 0, 0, 1}, J6[6] = {0, 1, 2, 1, 2, 2};
      for (int s6 = 0; s6 < 6; ++s6) {
        double a = 0;
        for (int e = 0; e < 3; ++e) a += Ki[e * 3 + I6[s6]] * Ki[e * 3 + J6[s6]];
        mc[cell * MC_COUNT + MC_M + s6] = det * a;
      }
      mc[cell * MC_COUNT + MC_DET] = det;
    };
    if (d->vertices) {
      for (int64_t cell = 0; cell < c->n_cells && affine; ++cell) {
        const double *v = d->vertices + caller_cell(cell) * 24;
        double J[9], h = 0, xmax = 0;  // J[e][f] = d x_e / d xi_f from the three edges at vertex 0
        for (int e = 0; e < 3; ++e)
          for (int f = 0; f < 3; ++f) {
            J[e * 3 + f] = v[(1 << f) * 3 + e] - v[e];
            h = std::max(h, std::fabs(J[e * 3 + f]));
          }
        for (int i = 0; i < 24; ++i) xmax = std::max(xmax, std::fabs(v[i]));
        // a vertex may deviate from the parallelepiped by the rounding of its own coordinates
        const double tol = 1e-13 * h + 16 * 2.3e-16 * xmax;
        for (int m = 0; m < 8 && affine; ++m)
          for (int e = 0; e < 3; ++e) {
            double x = v[e];
            for (int f = 0; f < 3; ++f)
              if ((m >> f) & 1) x += J[e * 3 + f];
            if (std::fabs(x - v[m * 3 + e]) > tol) affine = false;
          }
        if (!affine) break;
        const double det = J[0] * (J[4] * J[8] - J[5] * J[7]) - J[1] * (J[3] * J[8] - J[5] * J[6]) + J[2] * (J[3] * J[7] - J[4] * J[6]);
        const double id = 1.0 / det;
        double Ji[9];  // Ji[f][e] = d xi_f / d x_e
        Ji[0] = (J[4] * J[8] - J[5] * J[7]) * id;
        Ji[1] = (J[2] * J[7] - J[1] * J[8]) * id;
        Ji[2] = (J[1] * J[5] - J[2] * J[4]) * id;
        Ji[3] = (J[5] * J[6] - J[3] * J[8]) * id;
        Ji[4] = (J[0] * J[8] - J[2] * J[6]) * id;
        Ji[5] = (J[2] * J[3] - J[0] * J[5]) * id;
        Ji[6] = (J[3] * J[7] - J[4] * J[6]) * id;
        Ji[7] = (J[1] * J[6] - J[0] * J[7]) * id;
        Ji[8] = (J[0] * J[4] - J[1] * J[3]) * id;
        double Ki[9];
        for (int e = 0; e < 3; ++e)
          for (int f = 0; f < 3; ++f) Ki[e * 3 + f] = Ji[f * 3 + e];
        fill(cell, Ki, det);
      }
    } else {
      for (int64_t cell = 0; cell < c->n_cells && affine; ++cell) {
        const int64_t cc = caller_cell(cell);
        const double *K0 = d->inverse_jacobian + (size_t)cc * N3 * 9;
        double kmax = 0;
        for (int i = 0; i < 9; ++i) kmax = std::max(kmax, std::fabs(K0[i]));
        const double det0 = d->jxw[(size_t)cc * N3] / (c->shape.w[0] * c->shape.w[0] * c->shape.w[0]);
        for (int q = 1; q < N3 && affine; ++q) {
          for (int i = 0; i < 9; ++i)
            if (std::fabs(K0[q * 9 + i] - K0[i]) > 1e-11 * kmax) affine = false;
          const double wq = c->shape.w[q % N] * c->shape.w[(q / N) % N] * c->shape.w[q / N2];
          if (std::fabs(d->jxw[(size_t)cc * N3 + q] / wq - det0) > 1e-11 * std::fabs(det0)) affine = false;
        }
        if (affine) fill(cell, K0, det0);
      }
    }
    if (affine) {
      const size_t es = esize(c.get());
      B2_CUDA(cudaMalloc(&c->mcell, std::max<size_t>(8, mc.size() * es)));
      if (c->dtype == B200MF_F64) {
        B2_CUDA(cudaMemcpy(c->mcell, mc.data(), mc.size() * 8, cudaMemcpyHostToDevice));
      } else {
        std::vector<float> mf32(mc.begin(), mc.end());
        B2_CUDA(cudaMemcpy(c->mcell, mf32.data(), mf32.size() * 4, cudaMemcpyHostToDevice));
      }
      c->bytes += mc.size() * es;
    }
  }

  // ---- reference metric at the quadrature points (not needed by the scalar operator on affine cells) ------
  if (!(c->mcell && NC == 1)) {
    const size_t es = esize(c.get());
    const int64_t gs = table_stride_rt(c.get(), MF_COUNT);
    const size_t nel = (size_t)c->n_groups * gs;
    B2_CUDA(cudaMalloc(&c->mref, nel * es));
    B2_CUDA(cudaMemset(c->mref, 0, nel * es));
    c->bytes += nel * es;
    const int64_t padded = c->n_groups * c->cpg;
    const int64_t nthreads = padded * N3;
    if (d->vertices) {
      double *dv = nullptr;
      B2_CUDA(cudaMalloc(&dv, (size_t)c->n_cells * 24 * sizeof(double)));
      if (order.empty()) {
        B2_CUDA(cudaMemcpy(dv, d->vertices, (size_t)c->n_cells * 24 * sizeof(double), cudaMemcpyHostToDevice));
      } else {
        std::vector<double> pv((size_t)c->n_cells * 24);
        for (int64_t cell = 0; cell < c->n_cells; ++cell) memcpy(&pv[cell * 24], d->vertices + order[cell] * 24, 24 * sizeof(double));
        B2_CUDA(cudaMemcpy(dv, pv.data(), pv.size() * sizeof(double), cudaMemcpyHostToDevice));
      }
      QuadPts qp;
      for (int i = 0; i < N; ++i) {
        qp.xq[i] = c->shape.xq[i];
        qp.wq[i] = c->shape.w[i];
      }
      if (c->dtype == B200MF_F64)
        metric_from_vertices_kernel<double><<<nblocks(nthreads, 256), 256>>>(N, c->cpg, c->L, table_geom_rt(c.get(), MF_COUNT), c->n_cells, padded, dv, qp, (double *)c->mref);
      else
        metric_from_vertices_kernel<float><<<nblocks(nthreads, 256), 256>>>(N, c->cpg, c->L, table_geom_rt(c.get(), MF_COUNT), c->n_cells, padded, dv, qp, (float *)c->mref);
      B2_LAUNCHED();
      B2_CUDA(cudaDeviceSynchronize());
      cudaFree(dv);
    } else {
      // interleave K (9) and JxW (1) into one natural table, then reorder on the device
      std::vector<double> nat((size_t)c->n_cells * N3 * MF_COUNT);
      for (size_t cq = 0; cq < (size_t)c->n_cells * N3; ++cq) {
        for (int i = 0; i < 9; ++i) nat[cq * MF_COUNT + i] = d->inverse_jacobian[cq * 9 + i];
        nat[cq * MF_COUNT + MF_W] = d->jxw[cq];
      }
      double *dn = nullptr;
      B2_CUDA(cudaMalloc(&dn, std::max<size_t>(8, nat.size() * sizeof(double))));
      B2_CUDA(cudaMemcpy(dn, nat.data(), nat.size() * sizeof(double), cudaMemcpyHostToDevice));
      if (c->dtype == B200MF_F64)
        table_to_group_layout_kernel<double><<<nblocks(nthreads, 256), 256>>>(N, c->cpg, c->L, MF_COUNT, table_geom_rt(c.get(), MF_COUNT), c->n_cells, padded, dn, (double *)c->mref, 1, c->cell_order);
      else
        table_to_group_layout_kernel<float><<<nblocks(nthreads, 256), 256>>>(N, c->cpg, c->L, MF_COUNT, table_geom_rt(c.get(), MF_COUNT), c->n_cells, padded, dn, (float *)c->mref, 1, c->cell_order);
      B2_LAUNCHED();
      B2_CUDA(cudaDeviceSynchronize());
      cudaFree(dn);
    }
  }
  *out = c.release();
  return 0;
}

extern "C" int b200mf_ctx_destroy(b200mf_ctx *c) {
  g_config_epoch.fetch_add(1);
  if (!c) return 0;
  cudaFree(c->nidx);
  cudaFree(c->mref);
  cudaFree(c->mcell);
  cudaFree(c->scalars);
  cudaFree(c->cell_order);
  if (c->comm_stream) cudaStreamDestroy(c->comm_stream);
  for (cudaEvent_t e : c->ev) if (e) cudaEventDestroy(e);
  cudaFree(c->matid);
  cudaFree(c->int2ext);
  cudaFree(c->constrained);
  cudaFree(c->stage);
  free_partitioner(c);
  free_hostpipe(c);
  delete c->h_ext2int;
  delete c->h_cell_nodes;
  if (c->kernel_events) {
    for (cudaEvent_t e : *c->kernel_events) cudaEventDestroy(e);
    delete c->kernel_events;
  }
  delete c;
  return 0;
}

extern "C" int b200mf_ctx_set_stream(b200mf_ctx *c, void *s) {
  g_config_epoch.fetch_add(1);
  if (!c) return set_error(B200MF_ERR_ARG, "null ctx");
  c->stream = (cudaStream_t)s;
  return 0;
}
extern "C" int b200mf_ctx_synchronize(b200mf_ctx *c) {
  if (!c) return set_error(B200MF_ERR_ARG, "null ctx");
  B2_CUDA(cudaStreamSynchronize(c->stream));
  return 0;
}
// Measurement aid (bench.py roofline): CUDA events on the launching stream around every cell-loop launch.
extern "C" int b200mf_ctx_time_kernels(b200mf_ctx *c, int on) {
  g_config_epoch.fetch_add(1);
  if (!c) return set_error(B200MF_ERR_ARG, "null ctx");
  if (c->kernel_events) {
    for (cudaEvent_t e : *c->kernel_events) cudaEventDestroy(e);
    delete c->kernel_events;
    c->kernel_events = nullptr;
  }
  if (on) c->kernel_events = new std::vector<cudaEvent_t>();
  return 0;
}
extern "C" int b200mf_ctx_kernel_time(b200mf_ctx *c, double *total_ms, int64_t *n_launches) {
  if (!c || !total_ms || !n_launches) return set_error(B200MF_ERR_ARG, "null argument");
  *total_ms = 0;
  *n_launches = 0;
  if (!c->kernel_events) return 0;
  B2_CUDA(cudaStreamSynchronize(c->stream));
  for (size_t i = 0; i + 1 < c->kernel_events->size(); i += 2) {
    float ms = 0;
    B2_CUDA(cudaEventSynchronize((*c->kernel_events)[i + 1]));
    B2_CUDA(cudaEventElapsedTime(&ms, (*c->kernel_events)[i], (*c->kernel_events)[i + 1]));
    *total_ms += ms;
    ++*n_launches;
  }
  return 0;
}
namespace b200mf {
int kernel_event(b200mf_ctx *c, cudaStream_t s) {
  if (!c->kernel_events) return 0;
  cudaEvent_t e;
  B2_CUDA(cudaEventCreate(&e));
  B2_CUDA(cudaEventRecord(e, s));
  c->kernel_events->push_back(e);
  return 0;
}
}  // namespace b200mf
extern "C" int b200mf_ctx_set_overlap(b200mf_ctx *c, int on) {
  g_config_epoch.fetch_add(1);
  if (!c) return set_error(B200MF_ERR_ARG, "null ctx");
  c->overlap = on < 0 ? -1 : on ? 1 : 0;
  return 0;
}
extern "C" int b200mf_ctx_cell_type(const b200mf_ctx *c) { return c ? (c->mcell ? 1 : 0) : -1; }
extern "C" int64_t b200mf_ctx_n_dofs(const b200mf_ctx *c) { return c ? c->n_dofs : -1; }
extern "C" size_t b200mf_ctx_memory_consumption(const b200mf_ctx *c) { return c ? c->bytes : 0; }

// ===========================================================================
// ABI: vectors
extern "C" int b200mf_vec_create(b200mf_ctx *c, void **v) {
  if (!c || !v) return set_error(B200MF_ERR_ARG, "null argument");
  B2_CUDA(cudaMalloc(v, c->n_int * esize(c)));
  B2_CUDA(cudaMemsetAsync(*v, 0, c->n_int * esize(c), c->stream));
  return 0;
}
extern "C" int b200mf_vec_destroy(b200mf_ctx *, void *v) {
  if (v) B2_CUDA(cudaFree(v));
  return 0;
}

template <typename TD, typename TH>
static int vec_upload_t(b200mf_ctx *c, TD *dvec, const TH *host) {
  if (std::is_same<TD, TH>::value && !c->int2ext) {
    B2_CUDA(cudaMemcpyAsync(dvec, host, c->n_dofs * sizeof(TD), cudaMemcpyHostToDevice, c->stream));
    return 0;
  }
  B2_CUDA(cudaMemcpyAsync(c->stage, host, c->n_dofs * sizeof(TH), cudaMemcpyHostToDevice, c->stream));
  permute_in_kernel<TD, TH><<<nblocks(c->n_int, 256), 256, 0, c->stream>>>(c->n_int, c->int2ext, (const TH *)c->stage, dvec);
  B2_LAUNCHED();
  return 0;
}
template <typename TD, typename TH>
static int vec_download_t(b200mf_ctx *c, TH *host, const TD *dvec) {
  if (std::is_same<TD, TH>::value && !c->int2ext) {
    B2_CUDA(cudaMemcpyAsync(host, dvec, c->n_dofs * sizeof(TD), cudaMemcpyDeviceToHost, c->stream));
    B2_CUDA(cudaStreamSynchronize(c->stream));
    return 0;
  }
  permute_out_kernel<TH, TD><<<nblocks(c->n_int, 256), 256, 0, c->stream>>>(c->n_int, c->int2ext, (TH *)c->stage, dvec);
  B2_LAUNCHED();
  B2_CUDA(cudaMemcpyAsync(host, c->stage, c->n_dofs * sizeof(TH), cudaMemcpyDeviceToHost, c->stream));
  B2_CUDA(cudaStreamSynchronize(c->stream));
  return 0;
}

extern "C" int b200mf_vec_upload(b200mf_ctx *c, void *dvec, const void *host) {
  if (!c || !dvec || !host) return set_error(B200MF_ERR_ARG, "null argument");
  return c->dtype == B200MF_F64 ? vec_upload_t<double, double>(c, (double *)dvec, (const double *)host)
                                : vec_upload_t<float, float>(c, (float *)dvec, (const float *)host);
}
extern "C" int b200mf_vec_download(b200mf_ctx *c, void *host, const void *dvec) {
  if (!c || !dvec || !host) return set_error(B200MF_ERR_ARG, "null argument");
  return c->dtype == B200MF_F64 ? vec_download_t<double, double>(c, (double *)host, (const double *)dvec)
                                : vec_download_t<float, float>(c, (float *)host, (const float *)dvec);
}
extern "C" int b200mf_vec_upload_f64(b200mf_ctx *c, void *dvec, const double *host) {
  if (!c || !dvec || !host) return set_error(B200MF_ERR_ARG, "null argument");
  return c->dtype == B200MF_F64 ? vec_upload_t<double, double>(c, (double *)dvec, host)
                                : vec_upload_t<float, double>(c, (float *)dvec, host);
}
extern "C" int b200mf_vec_download_f64(b200mf_ctx *c, double *host, const void *dvec) {
  if (!c || !dvec || !host) return set_error(B200MF_ERR_ARG, "null argument");
  return c->dtype == B200MF_F64 ? vec_download_t<double, double>(c, host, (const double *)dvec)
                                : vec_download_t<float, double>(c, host, (const float *)dvec);
}

// ===========================================================================
// NeoHookOperator
struct b200mf_nh {
  b200mf_ctx *ctx;
  b200mf_material mat[2];
  int caching;  // requested MF caching level
  int variant;  // NHVariant the kernels run
  bool initialized, cached, have_diag;
  void *qd;   // staged table of the deformed-frame variants, T [group][NF][N][L]
  void *qd2;  // scalar_referential: cached_scalar table
  void *big;  // tensor4 / tensor4_ns tables
  void *pos;  // scalar_referential: cell-local displacement
  size_t table_bytes;
  unsigned long long *n_bad;
  void *diag, *inv_diag;
  void *tmp_src, *tmp_dst;  // for the *_host entry points
  uint64_t diag_gen;        // bumped by compute_diagonal (consumers that keep derived copies of the inverse diagonal)
};

static void derived(const b200mf_material &m, double &kappa, double &c_1, double &lambda, double &rho_alpha) {
  kappa = (2.0 * m.mu * (1.0 + m.nu)) / (3.0 * (1.0 - 2.0 * m.nu));  // material.h:85
  c_1 = m.mu / 2.0;
  lambda = (2.0 * m.mu * m.nu) / (1.0 - 2.0 * m.nu);
  rho_alpha = m.rho * m.alpha_1;
}

extern "C" int b200mf_nh_create(b200mf_nh **out, b200mf_ctx *ctx, const b200mf_material *mat, const b200mf_material *incl) {
  if (!out || !ctx || !mat) return set_error(B200MF_ERR_ARG, "null argument");
  if (ctx->NC != 3) return set_error(B200MF_ERR_ARG, "NeoHookOperator needs a 3-component context");
  const b200mf_material *ms[2] = {mat, incl ? incl : mat};
  for (const b200mf_material *m : ms) {
    if (m->formulation != 0 && m->formulation != 1) return set_error(B200MF_ERR_ARG, "Unknown material formulation");  // :405-406
    if (!(m->nu < 0.5 && m->nu >= -1.0)) return set_error(B200MF_ERR_ARG, "Poisson ratio outside [-1,0.5) (material.h:88-89)");
  }
  if (ms[0]->formulation != ms[1]->formulation)
    return set_error(B200MF_ERR_ARG, "material and material_inclusion must use the same formulation");
  std::unique_ptr<b200mf_nh> op(new b200mf_nh());
  memset(op.get(), 0, sizeof(b200mf_nh));
  op->ctx = ctx;
  op->mat[0] = *ms[0];
  op->mat[1] = *ms[1];
  op->caching = -1;
  B2_CUDA(cudaMalloc(&op->n_bad, sizeof(unsigned long long)));
  *out = op.release();
  return 0;
}

// a device-side failure counter, summed over the ranks of the partitioner before the host reads it: every rank returns
// the same status (a rank that bailed out alone would leave its peers hanging in the next exchange)
static int read_failure_count(b200mf_ctx *c, unsigned long long *dev, unsigned long long *host) {
  B2_TRY(allreduce_sum_u64(c, dev, 1));
  B2_CUDA(cudaMemcpyAsync(host, dev, sizeof *host, cudaMemcpyDeviceToHost, c->stream));
  B2_CUDA(cudaStreamSynchronize(c->stream));
  return 0;
}

static void nh_free_tables(b200mf_nh *op) {
  cudaFree(op->qd);
  cudaFree(op->qd2);
  cudaFree(op->big);
  cudaFree(op->pos);
  op->qd = op->qd2 = op->big = op->pos = nullptr;
  op->table_bytes = 0;
}

extern "C" int b200mf_nh_destroy(b200mf_nh *op) {
  g_config_epoch.fetch_add(1);
  if (!op) return 0;
  nh_free_tables(op);
  cudaFree(op->n_bad);
  cudaFree(op->diag);
  cudaFree(op->inv_diag);
  cudaFree(op->tmp_src);
  cudaFree(op->tmp_dst);
  delete op;
  return 0;
}

extern "C" int b200mf_nh_initialize(b200mf_nh *op, int caching) {
  g_config_epoch.fetch_add(1);
  if (!op) return set_error(B200MF_ERR_ARG, "null op");
  if (caching < B200MF_CACHE_SCALAR_REFERENTIAL || caching > B200MF_CACHE_TENSOR4_NS)
    return set_error(B200MF_ERR_ARG, "Unknown caching");  // mf_nh_operator.h:385
  if (op->mat[0].formulation == 0 && caching != B200MF_CACHE_TENSOR2)
    return set_error(B200MF_ERR_ARG, "formulation 0 requires tensor2 caching (mf_nh_operator.h:421, 706)");
  b200mf_ctx *c = op->ctx;
  op->caching = caching;
  op->variant = op->mat[0].formulation == 0            ? NHV_FORM0
                : caching == B200MF_CACHE_TENSOR2      ? NHV_TENSOR2
                : caching == B200MF_CACHE_TENSOR4      ? NHV_TENSOR4
                : caching == B200MF_CACHE_TENSOR4_NS   ? NHV_TENSOR4_NS
                                                       : NHV_SCALAR_REF;
  nh_free_tables(op);
  auto alloc = [&](void **p, int nf, bool plain = false) -> int {
    const size_t bytes = (size_t)c->n_groups * table_stride_rt(c, nf, plain) * esize(c);
    B2_CUDA(cudaMalloc(p, std::max<size_t>(bytes, 16)));
    B2_CUDA(cudaMemsetAsync(*p, 0, bytes, c->stream));
    op->table_bytes += bytes;
    return 0;
  };
  const int v = op->variant;
  if (nhv_deformed(v)) B2_TRY(alloc(&op->qd, nhv_staged_fields(v)));
  if (nhv_big_fields(v)) B2_TRY(alloc(&op->big, nhv_big_fields(v), true));
  if (v == NHV_SCALAR_REF) {
    B2_TRY(alloc(&op->qd2, 1));
    B2_TRY(alloc(&op->pos, 3, true));
  }
  op->initialized = true;
  op->cached = false;
  op->have_diag = false;
  return 0;
}

template <typename T>
static NHArgs<T> nh_args(const b200mf_nh *op) {
  const b200mf_ctx *c = op->ctx;
  NHArgs<T> oa;
  oa.stage = (const T *)(nhv_deformed(op->variant) ? op->qd : c->mref);
  oa.stage2 = (const T *)op->qd2;
  oa.big = (const T *)op->big;
  oa.pos = (const T *)op->pos;
  oa.matid = c->matid;
  oa.n_cells = c->n_cells;
  for (int m = 0; m < 2; ++m) {
    double kappa, c_1, lambda, ra;
    derived(op->mat[m], kappa, c_1, lambda, ra);
    oa.mu[m] = (T)op->mat[m].mu;
    oa.lambda[m] = (T)lambda;
    oa.rho_alpha[m] = (T)ra;
    oa.kappa[m] = (T)kappa;
    oa.c_1[m] = (T)c_1;
  }
  return oa;
}

template <typename T>
static int nh_cache_t(b200mf_nh *op, const void *disp) {
  b200mf_ctx *c = op->ctx;
  NHCacheArgs<T> oa;
  oa.stage = (const T *)c->mref;
  oa.stage2 = nullptr;
  oa.qd = (T *)op->qd;
  oa.qd2 = (T *)op->qd2;
  oa.big = (T *)op->big;
  oa.pos = (T *)op->pos;
  oa.matid = c->matid;
  oa.n_cells = c->n_cells;
  for (int m = 0; m < 2; ++m) {
    double kappa, c_1, lambda, ra;
    derived(op->mat[m], kappa, c_1, lambda, ra);
    oa.mu[m] = (T)op->mat[m].mu;
    oa.lambda[m] = (T)lambda;
  }
  oa.n_bad = op->n_bad;
  return run_nh_cache<T>(c, op->variant, (const T *)disp, oa);
}

extern "C" int b200mf_nh_cache(b200mf_nh *op, const void *disp, int64_t *n_bad_out) {
  if (!op || !disp) return set_error(B200MF_ERR_ARG, "null argument");
  if (!op->initialized) return set_error(B200MF_ERR_STATE, "cache() before initialize()");
  b200mf_ctx *c = op->ctx;
  B2_CUDA(cudaMemsetAsync(op->n_bad, 0, sizeof(unsigned long long), c->stream));
  B2_TRY(update_ghost_values(c, const_cast<void *>(disp)));  // cache() reads ghosted displacement
  int rc = c->dtype == B200MF_F64 ? nh_cache_t<double>(op, disp) : nh_cache_t<float>(op, disp);
  B2_TRY(rc);
  unsigned long long bad = 0;
  B2_TRY(read_failure_count(c, op->n_bad, &bad));
  if (n_bad_out) *n_bad_out = (int64_t)bad;
  if (op->cached != (bad == 0)) g_config_epoch.fetch_add(1);  // a recorded launch sequence must not outlive the tables' validity
  op->cached = (bad == 0);  // tables built from det F <= 0 must not be applied (the reference asserts, :430-435)
  op->have_diag = false;
  if (bad) return set_error(B200MF_ERR_NUMERIC, "det_F is not positive at %llu quadrature points", bad);  // :430-435
  return 0;
}

static int nh_tmp(b200mf_nh *op) {
  b200mf_ctx *c = op->ctx;
  if (!op->tmp_src) B2_CUDA(cudaMalloc(&op->tmp_src, c->n_int * esize(c)));
  if (!op->tmp_dst) B2_CUDA(cudaMalloc(&op->tmp_dst, c->n_int * esize(c)));
  return 0;
}

extern "C" int b200mf_nh_cache_host(b200mf_nh *op, const double *disp, int64_t *n_bad_out) {
  if (!op || !disp) return set_error(B200MF_ERR_ARG, "null argument");
  B2_TRY(nh_tmp(op));
  B2_TRY(b200mf_vec_upload_f64(op->ctx, op->tmp_src, disp));
  return b200mf_nh_cache(op, op->tmp_src, n_bad_out);
}

// update_ghost_values(src) -> cell loop -> compress(add)(dst) (mf_nh_operator.h:609-620).  The reference leaves
// both exchanges exposed (FIXME at :598-607); here the cells that touch ghost entries come first in the
// library's cell order (ctx_create), so that
//   exchange stream: [ghost update of src]                    [compress(add) of dst]
//   main stream    : [interior cells, 1st half] [boundary cells] [interior cells, 2nd half]
// launch(g0, g1, stream) enqueues the cell loop over the groups [g0, g1).
template <class Launch>
static int apply_cells_with_exchange(b200mf_ctx *c, void *dst, const void *src, Launch &&launch) {
  const int64_t nb = c->n_boundary_groups, ng = c->n_groups;
  if (c->overlap < 0) return launch(0, ng, c->stream);  // measurement only: no exchange at all (rank-local result)
  if (!has_neighbors(c) || !c->overlap || nb >= ng) {
    B2_TRY(update_ghost_values(c, const_cast<void *>(src)));  // :610
    B2_TRY(launch(0, ng, c->stream));
    return compress_add(c, dst);                             // :620
  }
  if (!c->comm_stream) {
    B2_CUDA(cudaStreamCreateWithFlags(&c->comm_stream, cudaStreamNonBlocking));
    for (cudaEvent_t &e : c->ev) B2_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
  }
  const int64_t mid = nb + (ng - nb) / 2;
  B2_CUDA(cudaEventRecord(c->ev[0], c->stream));
  B2_CUDA(cudaStreamWaitEvent(c->comm_stream, c->ev[0], 0));
  B2_TRY(update_ghost_values(c, const_cast<void *>(src), c->comm_stream));
  B2_CUDA(cudaEventRecord(c->ev[1], c->comm_stream));
  B2_TRY(launch(nb, mid, c->stream));
  B2_CUDA(cudaStreamWaitEvent(c->stream, c->ev[1], 0));
  B2_TRY(launch(0, nb, c->stream));
  B2_CUDA(cudaEventRecord(c->ev[2], c->stream));
  B2_CUDA(cudaStreamWaitEvent(c->comm_stream, c->ev[2], 0));
  B2_TRY(compress_add(c, dst, c->comm_stream));
  B2_CUDA(cudaEventRecord(c->ev[3], c->comm_stream));
  B2_TRY(launch(mid, ng, c->stream));
  B2_CUDA(cudaStreamWaitEvent(c->stream, c->ev[3], 0));
  return 0;
}

template <typename T>
static int constrained_identity(b200mf_ctx *c, void *dst, const void *src, double *dot = nullptr) {  // :623-624
  if (c->n_constrained) {
    constrained_identity_kernel<T><<<nblocks(c->n_constrained, 256), 256, 0, c->stream>>>(c->n_constrained, c->constrained, (T *)dst, (const T *)src, dot);
    B2_LAUNCHED();
  }
  return 0;
}

template <typename T>
static int nh_apply_t(b200mf_nh *op, void *dst, const void *src, double *dot, bool identity_rows) {
  b200mf_ctx *c = op->ctx;
  const NHArgs<T> oa = nh_args<T>(op);
  B2_TRY(apply_cells_with_exchange(c, dst, src, [&](int64_t g0, int64_t g1, cudaStream_t s) {
    return run_nh<T>(c, op->variant, (const T *)src, (T *)dst, oa, g0, g1, s, dot);
  }));
  return identity_rows ? constrained_identity<T>(c, dst, src, dot) : 0;
}

// how an application treats dst: vmult zeroes it, vmult_add keeps it, APPLY_CELLS is the smoother hook's operator pass - dst is
// zero on entry by contract (the previous vector update left it so) and the constrained rows are left 0 for the caller's update
enum ApplyMode { APPLY_VMULT = 0, APPLY_ADD = 1, APPLY_CELLS = 2 };

static int nh_apply(b200mf_nh *op, void *dst, const void *src, ApplyMode mode, double *dot = nullptr) {
  if (!op || !dst || !src) return set_error(B200MF_ERR_ARG, "null argument");
  if (dst == src) return set_error(B200MF_ERR_ARG, "vmult: dst aliases src");
  if (!op->cached) return set_error(B200MF_ERR_STATE, "vmult before cache()");
  b200mf_ctx *c = op->ctx;
  if (mode == APPLY_VMULT)
    B2_CUDA(cudaMemsetAsync(dst, 0, c->n_int * esize(c), c->stream));  // vmult: dst = 0, :548
  else if (mode == APPLY_ADD)
    B2_TRY(zero_ghosts(c, dst));  // vmult_add: stale ghost entries of dst must not be summed into their owners by compress(add)
  int rc = c->dtype == B200MF_F64 ? nh_apply_t<double>(op, dst, src, dot, mode != APPLY_CELLS) : nh_apply_t<float>(op, dst, src, dot, mode != APPLY_CELLS);
  B2_TRY(rc);
  B2_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int b200mf_nh_vmult(b200mf_nh *op, void *dst, const void *src) { return nh_apply(op, dst, src, APPLY_VMULT); }
extern "C" int b200mf_nh_vmult_add(b200mf_nh *op, void *dst, const void *src) { return nh_apply(op, dst, src, APPLY_ADD); }
extern "C" int b200mf_nh_Tvmult(b200mf_nh *op, void *dst, const void *src) { return nh_apply(op, dst, src, APPLY_VMULT); }    // :556-561
extern "C" int b200mf_nh_Tvmult_add(b200mf_nh *op, void *dst, const void *src) { return nh_apply(op, dst, src, APPLY_ADD); }  // :566-571
extern "C" int b200mf_nh_vmult_cells(b200mf_nh *op, void *dst, const void *src) { return nh_apply(op, dst, src, APPLY_CELLS); }
extern "C" uint64_t b200mf_nh_diagonal_generation(const b200mf_nh *op) { return op ? op->diag_gen : 0; }

// vmult with the CG dot product fused into the operator pass: *dot_device += (this rank's share of) src . (A src)
extern "C" int b200mf_nh_vmult_dot(b200mf_nh *op, void *dst, const void *src, double *dot_device) {
  if (!dot_device) return set_error(B200MF_ERR_ARG, "null argument");
  return nh_apply(op, dst, src, APPLY_VMULT, dot_device);
}

static int fused_dot_result(b200mf_ctx *c, double *result_host) {
  B2_TRY(allreduce_sum_f64(c, c->scalars, 1));
  B2_CUDA(cudaMemcpyAsync(result_host, c->scalars, sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  B2_CUDA(cudaStreamSynchronize(c->stream));
  return 0;
}
static int fused_dot_slot(b200mf_ctx *c) {
  if (!c->scalars) B2_CUDA(cudaMalloc(&c->scalars, 8 * sizeof(double)));
  B2_CUDA(cudaMemsetAsync(c->scalars, 0, sizeof(double), c->stream));
  return 0;
}
// same, result summed over ranks and returned on the host
extern "C" int b200mf_nh_vmult_dot_host(b200mf_nh *op, void *dst, const void *src, double *result_host) {
  if (!op || !result_host) return set_error(B200MF_ERR_ARG, "null argument");
  B2_TRY(fused_dot_slot(op->ctx));
  B2_TRY(nh_apply(op, dst, src, APPLY_VMULT, op->ctx->scalars));
  return fused_dot_result(op->ctx, result_host);
}

// ---- host-vector vmult as a pipeline -----------------------------------------------------------------------------------
// The host entry point (the "parity mode" of INTEGRATION.md: caller-owned host vectors) used to run upload -> vmult ->
// download back to back: 2 x 8 B/DoF over PCIe, fully exposed.  The cell groups only need the source entries of their own
// cells and finish their destination entries wave by wave, so the three stages overlap:
//   h2d stream    : src[up_0..up_1)      src[up_1..up_2)      ...
//   compute stream:       scatter+cells(wave 0)   scatter+cells(wave 1)   ...   (+ gather of the finished dst range)
//   d2h stream    :                              dst[down_0..down_1)     dst[down_1..down_2) ...
// up_w   = 1 + largest caller index any group of the waves <= w touches (prefix maximum: a contiguous slice of the host vector)
// down_w = smallest caller index any group of a later wave touches (suffix minimum): everything below is final after wave w
// For a locality-preserving numbering both frontiers advance with the waves; for an arbitrary numbering the first slice is
// (nearly) the whole vector and the schedule degenerates to the serial one - still correct.
struct b200mf_hostpipe {
  int W = 0;
  std::vector<int64_t> g, up, down;  // [W+1]
  uint32_t *d_ext2int = nullptr;     // caller index -> internal index (0xffffffff: no internal slot)
  double *stage_out = nullptr;       // finished dst entries in the caller's order
  cudaStream_t h2d = nullptr, d2h = nullptr;
  std::vector<cudaEvent_t> ev_up, ev_out;
  cudaEvent_t ev_start = nullptr;
};
namespace b200mf {
void free_hostpipe(b200mf_ctx *c) {
  b200mf_hostpipe *hp = c->hostpipe;
  if (!hp) return;
  cudaFree(hp->d_ext2int);
  cudaFree(hp->stage_out);
  if (hp->h2d) cudaStreamDestroy(hp->h2d);
  if (hp->d2h) cudaStreamDestroy(hp->d2h);
  for (cudaEvent_t e : hp->ev_up) cudaEventDestroy(e);
  for (cudaEvent_t e : hp->ev_out) cudaEventDestroy(e);
  if (hp->ev_start) cudaEventDestroy(hp->ev_start);
  delete hp;
  c->hostpipe = nullptr;
}
template <typename T>
__global__ void scatter_ext_range_kernel(int64_t a, int64_t b, const uint32_t *__restrict__ ext2int, const double *__restrict__ ext, T *__restrict__ in) {
  const int64_t e = a + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e < b) {
    const uint32_t i = ext2int[e];
    if (i != 0xffffffffu) in[i] = (T)ext[e];
  }
}
template <typename T>
__global__ void gather_ext_range_kernel(int64_t a, int64_t b, const uint32_t *__restrict__ ext2int, double *__restrict__ ext, const T *__restrict__ in) {
  const int64_t e = a + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e < b) {
    const uint32_t i = ext2int[e];
    ext[e] = i != 0xffffffffu ? (double)in[i] : 0.0;
  }
}
// dst[c] += src[c] on the constrained rows whose caller index lies in [a, b) (mf_nh_operator.h:623-624, slice-wise)
template <typename T>
__global__ void constrained_identity_range_kernel(int64_t n, const uint32_t *__restrict__ idx, const int64_t *__restrict__ int2ext, int64_t a,
                                                  int64_t b, T *__restrict__ dst, const T *__restrict__ src) {
  const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (k < n) {
    const uint32_t i = idx[k];
    const int64_t e = int2ext ? int2ext[i] : (int64_t)i;
    if (e >= a && e < b) dst[i] += src[i];
  }
}
}  // namespace b200mf

static int build_hostpipe(b200mf_ctx *c) {
  if (c->hostpipe) return 0;
  std::unique_ptr<b200mf_hostpipe> hp(new b200mf_hostpipe());
  const int N3 = c->N * c->N * c->N, NC = c->NC;
  const int64_t ng = c->n_groups, nd = c->n_dofs;
  hp->W = (int)std::max<int64_t>(1, std::min<int64_t>(16, ng / (4 * (int64_t)c->sm_count)));  // waves of >= 4 groups per SM
  const int W = hp->W;
  // caller index of every internal slot
  std::vector<int64_t> int2ext;
  if (c->h_ext2int) {
    int2ext.assign(c->n_int, -1);
    for (int64_t e = 0; e < nd; ++e)
      if ((*c->h_ext2int)[e] >= 0) int2ext[(*c->h_ext2int)[e]] = e;
  }
  std::vector<int64_t> gmin(ng, nd), gmax(ng, -1);
  const std::vector<uint32_t> &cn = *c->h_cell_nodes;
  for (int64_t cell = 0; cell < c->n_cells; ++cell) {
    const int64_t g = cell / c->cpg;  // no partitioner: library cell order == caller's
    int64_t lo = gmin[g], hi = gmax[g];
    for (int a = 0; a < N3; ++a) {
      const int64_t node = cn[(size_t)cell * N3 + a] & NODE_MASK;
      for (int cc = 0; cc < NC; ++cc) {
        const int64_t in = cc * c->cstride + node, e = int2ext.empty() ? in : int2ext[in];
        if (e < 0) continue;
        lo = std::min(lo, e);
        hi = std::max(hi, e);
      }
    }
    gmin[g] = lo;
    gmax[g] = hi;
  }
  hp->g.resize(W + 1);
  hp->up.assign(W + 1, 0);
  hp->down.assign(W + 1, 0);
  for (int w = 0; w <= W; ++w) hp->g[w] = ng * w / W;
  int64_t run = 0;
  for (int w = 0; w < W; ++w) {
    for (int64_t g = hp->g[w]; g < hp->g[w + 1]; ++g) run = std::max(run, gmax[g] + 1);
    hp->up[w + 1] = run;
  }
  hp->up[W] = nd;
  run = nd;
  hp->down[W] = nd;
  for (int w = W - 1; w >= 1; --w) {
    for (int64_t g = hp->g[w]; g < hp->g[w + 1]; ++g) run = std::min(run, gmin[g]);
    hp->down[w] = std::min(run, hp->up[w]);  // entries no cell touches wait for their source slice
  }
  hp->down[0] = 0;
  // caller -> internal map on the device
  std::vector<uint32_t> e2i(std::max<int64_t>(1, nd), 0xffffffffu);
  for (int64_t e = 0; e < nd; ++e) {
    const int64_t in = c->h_ext2int ? (*c->h_ext2int)[e] : e;
    if (in >= 0) e2i[e] = (uint32_t)in;
  }
  B2_CUDA(cudaMalloc(&hp->d_ext2int, e2i.size() * sizeof(uint32_t)));
  B2_CUDA(cudaMemcpy(hp->d_ext2int, e2i.data(), e2i.size() * sizeof(uint32_t), cudaMemcpyHostToDevice));
  B2_CUDA(cudaMalloc(&hp->stage_out, std::max<int64_t>(1, nd) * sizeof(double)));
  B2_CUDA(cudaStreamCreateWithFlags(&hp->h2d, cudaStreamNonBlocking));
  B2_CUDA(cudaStreamCreateWithFlags(&hp->d2h, cudaStreamNonBlocking));
  hp->ev_up.resize(W);
  hp->ev_out.resize(W);
  for (int w = 0; w < W; ++w) {
    B2_CUDA(cudaEventCreateWithFlags(&hp->ev_up[w], cudaEventDisableTiming));
    B2_CUDA(cudaEventCreateWithFlags(&hp->ev_out[w], cudaEventDisableTiming));
  }
  B2_CUDA(cudaEventCreateWithFlags(&hp->ev_start, cudaEventDisableTiming));
  c->bytes += e2i.size() * sizeof(uint32_t) + (size_t)nd * sizeof(double);
  c->hostpipe = hp.release();
  return 0;
}

// launch(g0, g1) enqueues the cell loop over the groups [g0, g1) on the context's stream (dst = tmp_dst, src = tmp_src)
template <typename T, class Launch>
static int vmult_host_pipelined(b200mf_ctx *c, double *dst, const double *src, T *tmp_dst, T *tmp_src, Launch &&launch) {
  B2_TRY(build_hostpipe(c));
  b200mf_hostpipe *hp = c->hostpipe;
  const int W = hp->W;
  cudaStream_t cs = c->stream;
  B2_CUDA(cudaMemsetAsync(tmp_dst, 0, c->n_int * sizeof(T), cs));
  B2_CUDA(cudaEventRecord(hp->ev_start, cs));  // earlier work on the context's stream may still read the staging buffers
  B2_CUDA(cudaStreamWaitEvent(hp->h2d, hp->ev_start, 0));
  double *stage_in = (double *)c->stage;
  for (int w = 0; w < W; ++w) {
    const int64_t a = hp->up[w], b = hp->up[w + 1];
    if (b > a) B2_CUDA(cudaMemcpyAsync(stage_in + a, src + a, (size_t)(b - a) * sizeof(double), cudaMemcpyHostToDevice, hp->h2d));
    B2_CUDA(cudaEventRecord(hp->ev_up[w], hp->h2d));
  }
  for (int w = 0; w < W; ++w) {
    const int64_t a = hp->up[w], b = hp->up[w + 1];
    B2_CUDA(cudaStreamWaitEvent(cs, hp->ev_up[w], 0));
    if (b > a) {
      scatter_ext_range_kernel<T><<<nblocks(b - a, 256), 256, 0, cs>>>(a, b, hp->d_ext2int, stage_in, tmp_src);
      B2_LAUNCHED();
    }
    B2_TRY(launch(hp->g[w], hp->g[w + 1]));
    const int64_t da = hp->down[w], db = hp->down[w + 1];
    if (db > da) {
      if (c->n_constrained) {
        constrained_identity_range_kernel<T><<<nblocks(c->n_constrained, 256), 256, 0, cs>>>(c->n_constrained, c->constrained, c->int2ext, da, db, tmp_dst, tmp_src);
        B2_LAUNCHED();
      }
      gather_ext_range_kernel<T><<<nblocks(db - da, 256), 256, 0, cs>>>(da, db, hp->d_ext2int, hp->stage_out, tmp_dst);
      B2_LAUNCHED();
    }
    B2_CUDA(cudaEventRecord(hp->ev_out[w], cs));
    B2_CUDA(cudaStreamWaitEvent(hp->d2h, hp->ev_out[w], 0));
    if (db > da) B2_CUDA(cudaMemcpyAsync(dst + da, hp->stage_out + da, (size_t)(db - da) * sizeof(double), cudaMemcpyDeviceToHost, hp->d2h));
  }
  B2_CUDA(cudaStreamSynchronize(hp->d2h));
  B2_CUDA(cudaStreamSynchronize(cs));
  return 0;
}

static bool host_pipeline_applies(const b200mf_ctx *c) {
  static const bool off = getenv("B200MF_HOST_PIPELINE") && atoi(getenv("B200MF_HOST_PIPELINE")) == 0;
  return !off && !c->part && !c->cell_order && c->n_groups >= 8 * (int64_t)c->sm_count;
}

extern "C" int b200mf_nh_vmult_host(b200mf_nh *op, double *dst, const double *src) {
  if (!op || !dst || !src) return set_error(B200MF_ERR_ARG, "null argument");
  if (!op->cached) return set_error(B200MF_ERR_STATE, "vmult before cache()");
  B2_TRY(nh_tmp(op));
  b200mf_ctx *c = op->ctx;
  if (host_pipeline_applies(c)) {
    if (c->dtype == B200MF_F64) {
      const NHArgs<double> oa = nh_args<double>(op);
      return vmult_host_pipelined<double>(c, dst, src, (double *)op->tmp_dst, (double *)op->tmp_src, [&](int64_t g0, int64_t g1) {
        return run_nh<double>(c, op->variant, (const double *)op->tmp_src, (double *)op->tmp_dst, oa, g0, g1, c->stream, nullptr);
      });
    }
    const NHArgs<float> oa = nh_args<float>(op);
    return vmult_host_pipelined<float>(c, dst, src, (float *)op->tmp_dst, (float *)op->tmp_src, [&](int64_t g0, int64_t g1) {
      return run_nh<float>(c, op->variant, (const float *)op->tmp_src, (float *)op->tmp_dst, oa, g0, g1, c->stream, nullptr);
    });
  }
  B2_TRY(b200mf_vec_upload_f64(c, op->tmp_src, src));
  B2_TRY(nh_apply(op, op->tmp_dst, op->tmp_src, APPLY_VMULT));
  return b200mf_vec_download_f64(c, dst, op->tmp_dst);
}

// ---- cell part of Solid::assemble_residual (mf_elasticity.h:1764-1846) ------------------
template <typename T>
static int nh_residual_t(b200mf_nh *op, void *rhs, const void *disp, const void *acc, const double *bf) {
  b200mf_ctx *c = op->ctx;
  NHResidualArgs<T> oa;
  oa.stage = (const T *)c->mref;
  oa.stage2 = nullptr;
  oa.acc = (const T *)acc;
  oa.cstride = c->cstride;
  oa.matid = c->matid;
  oa.n_cells = c->n_cells;
  for (int m = 0; m < 2; ++m) {
    double kappa, c_1, lambda, rho_alpha;
    derived(op->mat[m], kappa, c_1, lambda, rho_alpha);
    oa.mu[m] = (T)op->mat[m].mu;
    oa.lambda[m] = (T)lambda;
    oa.kappa[m] = (T)kappa;
    oa.c_1[m] = (T)c_1;
    oa.rho[m] = (T)op->mat[m].rho;
  }
  oa.formulation = op->mat[0].formulation;
  for (int d = 0; d < 3; ++d) oa.body_force[d] = (T)(bf ? bf[d] : 0.0);
  oa.n_bad = op->n_bad;
  return run_nh_residual<T>(c, (const T *)disp, (T *)rhs, oa);
}

extern "C" int b200mf_nh_assemble_residual(b200mf_nh *op, void *rhs, const void *total_displacement, const void *acceleration,
                                           const double *body_force, int64_t *n_bad_out) {
  if (!op || !rhs || !total_displacement || !acceleration) return set_error(B200MF_ERR_ARG, "null argument");
  b200mf_ctx *c = op->ctx;
  if (!c->mref) return set_error(B200MF_ERR_STATE, "context holds no reference metric");
  B2_CUDA(cudaMemsetAsync(op->n_bad, 0, sizeof(unsigned long long), c->stream));
  B2_TRY(update_ghost_values(c, const_cast<void *>(total_displacement)));  // both are read with read_dof_values_plain
  B2_TRY(update_ghost_values(c, const_cast<void *>(acceleration)));
  B2_TRY(c->dtype == B200MF_F64 ? nh_residual_t<double>(op, rhs, total_displacement, acceleration, body_force)
                                : nh_residual_t<float>(op, rhs, total_displacement, acceleration, body_force));
  unsigned long long bad = 0;
  B2_TRY(read_failure_count(c, op->n_bad, &bad));
  if (n_bad_out) *n_bad_out = (int64_t)bad;
  if (bad) return set_error(B200MF_ERR_NUMERIC, "assemble_residual: det F <= 0 at %llu quadrature points", bad);
  return 0;
}

template <typename T>
__global__ void zero_constrained_kernel(int64_t n, const uint32_t *__restrict__ idx, T *__restrict__ v) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) v[idx[i]] = T(0);
}
// AffineConstraints::set_zero (mf_elasticity.h:1905, 2059)
extern "C" int b200mf_vec_set_zero_constrained(b200mf_ctx *c, void *v) {
  if (!c || !v) return set_error(B200MF_ERR_ARG, "null argument");
  if (!c->n_constrained) return 0;
  if (c->dtype == B200MF_F64)
    zero_constrained_kernel<double><<<nblocks(c->n_constrained, 256), 256, 0, c->stream>>>(c->n_constrained, c->constrained, (double *)v);
  else
    zero_constrained_kernel<float><<<nblocks(c->n_constrained, 256), 256, 0, c->stream>>>(c->n_constrained, c->constrained, (float *)v);
  B2_LAUNCHED();
  return 0;
}

extern "C" b200mf_ctx *b200mf_nh_ctx(b200mf_nh *op) { return op ? op->ctx : nullptr; }
extern "C" int64_t b200mf_nh_m(const b200mf_nh *op) { return op ? op->ctx->n_dofs : -1; }
extern "C" size_t b200mf_nh_memory_consumption(const b200mf_nh *op) { return op ? op->table_bytes : 0; }
// number of fields b200mf_nh_qdata_host can return: staged table (deformed variants) or c (scalar_referential), then the big table
extern "C" int b200mf_nh_n_qfields(const b200mf_nh *op) {
  if (!op || !op->initialized) return 0;
  const int v = op->variant;
  return (nhv_deformed(v) ? nhv_staged_fields(v) : v == NHV_SCALAR_REF ? 1 : 0) + nhv_big_fields(v);
}

extern "C" int b200mf_nh_qdata_host(b200mf_nh *op, int field, double *out) {
  if (!op || !out) return set_error(B200MF_ERR_ARG, "null argument");
  if (!op->cached) return set_error(B200MF_ERR_STATE, "qdata before cache()");
  if (field < 0 || field >= b200mf_nh_n_qfields(op)) return set_error(B200MF_ERR_ARG, "bad field");
  b200mf_ctx *c = op->ctx;
  const int v = op->variant;
  const int ns = nhv_deformed(v) ? nhv_staged_fields(v) : v == NHV_SCALAR_REF ? 1 : 0;
  const void *tab = nhv_deformed(v) ? op->qd : op->qd2;
  int nf = ns;
  bool plain = false;
  if (field >= ns) {
    tab = op->big;
    nf = nhv_big_fields(v);
    field -= ns;
    plain = true;  // never staged
  }
  const int64_t n = c->n_cells * c->N * c->N * c->N;
  double *tmp = nullptr;
  B2_CUDA(cudaMalloc(&tmp, n * sizeof(double)));
  if (c->dtype == B200MF_F64)
    field_from_group_layout_kernel<double><<<nblocks(n, 256), 256, 0, c->stream>>>(c->N, c->cpg, c->L, table_geom_rt(c, nf, plain), field, c->n_cells, (const double *)tab, tmp, c->cell_order);
  else
    field_from_group_layout_kernel<float><<<nblocks(n, 256), 256, 0, c->stream>>>(c->N, c->cpg, c->L, table_geom_rt(c, nf, plain), field, c->n_cells, (const float *)tab, tmp, c->cell_order);
  B2_LAUNCHED();
  B2_CUDA(cudaMemcpyAsync(out, tmp, n * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  B2_CUDA(cudaStreamSynchronize(c->stream));
  cudaFree(tmp);
  return 0;
}

template <class MakeArgsAndLaunch>
static int finish_diagonal(b200mf_ctx *c, void *diag, void *inv, unsigned long long *n_bad, MakeArgsAndLaunch &&launch) {
  B2_CUDA(cudaMemsetAsync(diag, 0, c->n_int * esize(c), c->stream));
  B2_CUDA(cudaMemsetAsync(n_bad, 0, sizeof(unsigned long long), c->stream));
  B2_TRY(launch());
  B2_TRY(compress_add(c, diag));
  if (c->dtype == B200MF_F64) {
    if (c->n_constrained) {
      set_constrained_one_kernel<double><<<nblocks(c->n_constrained, 256), 256, 0, c->stream>>>(c->n_constrained, c->constrained, (double *)diag);
      B2_LAUNCHED();
    }
    // orphan entries (ghost slots no local cell touches) keep diag = 0 -> make them 1 as well
    invert_diagonal_kernel<double><<<nblocks(c->cstride * c->NC, 256), 256, 0, c->stream>>>(c->cstride * c->NC, c->cstride, c->n_owned_nodes, (const double *)diag, (double *)inv, n_bad);
  } else {
    if (c->n_constrained) {
      set_constrained_one_kernel<float><<<nblocks(c->n_constrained, 256), 256, 0, c->stream>>>(c->n_constrained, c->constrained, (float *)diag);
      B2_LAUNCHED();
    }
    invert_diagonal_kernel<float><<<nblocks(c->cstride * c->NC, 256), 256, 0, c->stream>>>(c->cstride * c->NC, c->cstride, c->n_owned_nodes, (const float *)diag, (float *)inv, n_bad);
  }
  B2_LAUNCHED();
  B2_TRY(update_ghost_values(c, diag));  // mf_nh_operator.h:1056-1057
  B2_TRY(update_ghost_values(c, inv));
  unsigned long long bad = 0;
  B2_TRY(read_failure_count(c, n_bad, &bad));
  if (bad)
    return set_error(B200MF_ERR_NUMERIC, "No diagonal entry in a positive definite operator should be zero or negative (%llu entries)", bad);
  return 0;
}

extern "C" int b200mf_nh_compute_diagonal(b200mf_nh *op) {
  if (!op) return set_error(B200MF_ERR_ARG, "null op");
  if (!op->cached) return set_error(B200MF_ERR_STATE, "compute_diagonal before cache()");
  b200mf_ctx *c = op->ctx;
  if (!op->diag) {
    B2_CUDA(cudaMalloc(&op->diag, c->n_int * esize(c)));
    B2_CUDA(cudaMalloc(&op->inv_diag, c->n_int * esize(c)));
    B2_CUDA(cudaMemsetAsync(op->inv_diag, 0, c->n_int * esize(c), c->stream));
  }
  int rc0 = finish_diagonal(c, op->diag, op->inv_diag, op->n_bad, [&]() {
    int rc = c->dtype == B200MF_F64 ? run_nh_diag<double>(c, op->variant, (double *)op->diag, nh_args<double>(op))
                                    : run_nh_diag<float>(c, op->variant, (float *)op->diag, nh_args<float>(op));
    return rc;
  });
  B2_TRY(rc0);
  op->have_diag = true;
  ++op->diag_gen;
  return 0;
}

extern "C" int b200mf_nh_diagonal(b200mf_nh *op, const void **diag, const void **inv) {
  if (!op) return set_error(B200MF_ERR_ARG, "null op");
  if (!op->have_diag) return set_error(B200MF_ERR_STATE, "diagonal not computed");
  if (diag) *diag = op->diag;
  if (inv) *inv = op->inv_diag;
  return 0;
}

extern "C" int b200mf_nh_diagonal_host(b200mf_nh *op, double *diag, double *inv) {
  if (!op) return set_error(B200MF_ERR_ARG, "null op");
  if (!op->have_diag) return set_error(B200MF_ERR_STATE, "diagonal not computed");
  if (diag) B2_TRY(b200mf_vec_download_f64(op->ctx, diag, op->diag));
  if (inv) B2_TRY(b200mf_vec_download_f64(op->ctx, inv, op->inv_diag));
  return 0;
}

extern "C" int b200mf_nh_precondition_jacobi(b200mf_nh *op, void *dst, const void *src, double omega) {
  if (!op || !dst || !src) return set_error(B200MF_ERR_ARG, "null argument");
  if (!op->have_diag) return set_error(B200MF_ERR_STATE, "precondition_Jacobi: inverse diagonal entries not available (mf_nh_operator.h:286-288)");
  b200mf_ctx *c = op->ctx;
  if (c->dtype == B200MF_F64)
    jacobi_kernel<double><<<nblocks(c->n_int, 256), 256, 0, c->stream>>>(c->n_int, (double *)dst, (const double *)src, (const double *)op->inv_diag, omega);
  else
    jacobi_kernel<float><<<nblocks(c->n_int, 256), 256, 0, c->stream>>>(c->n_int, (float *)dst, (const float *)src, (const float *)op->inv_diag, (float)omega);
  B2_LAUNCHED();
  return 0;
}

// one entry of a device vector addressed by the caller's row number (el(): a diagonal entry)
static int read_entry(b200mf_ctx *c, const void *vec, int64_t row, double *value) {
  if (row < 0 || row >= c->n_dofs) return set_error(B200MF_ERR_ARG, "row out of range");
  const int64_t in = c->h_ext2int ? (*c->h_ext2int)[row] : row;
  if (c->dtype == B200MF_F64) {
    B2_CUDA(cudaMemcpyAsync(value, (const double *)vec + in, sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    B2_CUDA(cudaStreamSynchronize(c->stream));
  } else {
    float f;
    B2_CUDA(cudaMemcpyAsync(&f, (const float *)vec + in, sizeof(float), cudaMemcpyDeviceToHost, c->stream));
    B2_CUDA(cudaStreamSynchronize(c->stream));
    *value = f;
  }
  return 0;
}

extern "C" int b200mf_nh_el(b200mf_nh *op, int64_t row, double *value) {
  if (!op || !value) return set_error(B200MF_ERR_ARG, "null argument");
  if (!op->have_diag) return set_error(B200MF_ERR_STATE, "el(): diagonal not available (mf_nh_operator.h:1069-1071)");
  return read_entry(op->ctx, op->diag, row, value);
}

// ===========================================================================
// LaplaceOperator
struct b200mf_laplace {
  b200mf_ctx *ctx;
  double delta_t;
  void *coef;  // T [group][1][N][L] or nullptr
  void *diag, *inv_diag;
  void *tmp_src, *tmp_dst;
  unsigned long long *n_bad;
  bool have_diag;
  uint64_t diag_gen;
};

extern "C" int b200mf_laplace_create(b200mf_laplace **out, b200mf_ctx *ctx, double delta_t, const double *coef) {
  if (!out || !ctx) return set_error(B200MF_ERR_ARG, "null argument");
  if (ctx->NC != 1) return set_error(B200MF_ERR_ARG, "LaplaceOperator needs a scalar context");
  std::unique_ptr<b200mf_laplace> op(new b200mf_laplace());
  memset(op.get(), 0, sizeof(b200mf_laplace));
  op->ctx = ctx;
  op->delta_t = delta_t;
  if (coef) {
    b200mf_ctx *c = ctx;
    const int N = c->N, N3 = N * N * N;
    const int64_t gs = table_stride_rt(c, 1);
    const size_t nel = (size_t)c->n_groups * gs;
    double *dn = nullptr;
    B2_CUDA(cudaMalloc(&dn, (size_t)c->n_cells * N3 * sizeof(double)));
    B2_CUDA(cudaMemcpy(dn, coef, (size_t)c->n_cells * N3 * sizeof(double), cudaMemcpyHostToDevice));
    B2_CUDA(cudaMalloc(&op->coef, nel * esize(c)));
    const int64_t padded = c->n_groups * c->cpg;
    if (c->dtype == B200MF_F64)
      table_to_group_layout_kernel<double><<<nblocks(padded * N3, 256), 256>>>(N, c->cpg, c->L, 1, table_geom_rt(c, 1), c->n_cells, padded, dn, (double *)op->coef, 0, c->cell_order);
    else
      table_to_group_layout_kernel<float><<<nblocks(padded * N3, 256), 256>>>(N, c->cpg, c->L, 1, table_geom_rt(c, 1), c->n_cells, padded, dn, (float *)op->coef, 0, c->cell_order);
    B2_LAUNCHED();
    B2_CUDA(cudaDeviceSynchronize());
    cudaFree(dn);
  }
  *out = op.release();
  return 0;
}

extern "C" int b200mf_laplace_destroy(b200mf_laplace *op) {
  g_config_epoch.fetch_add(1);
  if (!op) return 0;
  cudaFree(op->coef);
  cudaFree(op->n_bad);
  cudaFree(op->diag);
  cudaFree(op->inv_diag);
  cudaFree(op->tmp_src);
  cudaFree(op->tmp_dst);
  delete op;
  return 0;
}

extern "C" b200mf_ctx *b200mf_laplace_ctx(b200mf_laplace *op) { return op ? op->ctx : nullptr; }
extern "C" int b200mf_laplace_set_delta_t(b200mf_laplace *op, double dt) {
  g_config_epoch.fetch_add(1);
  if (!op) return set_error(B200MF_ERR_ARG, "null op");
  op->delta_t = dt;
  op->have_diag = false;
  return 0;
}

template <typename T>
static LaplaceArgs<T> laplace_args(const b200mf_laplace *op) {
  LaplaceArgs<T> oa;
  oa.stage = (const T *)op->ctx->mref;
  oa.stage2 = (const T *)op->coef;
  oa.mcell = (const T *)op->ctx->mcell;
  oa.delta_t = (T)op->delta_t;
  return oa;
}
template <typename T>
static int laplace_apply_t(b200mf_laplace *op, void *dst, const void *src, double *dot, bool identity_rows) {
  b200mf_ctx *c = op->ctx;
  const LaplaceArgs<T> oa = laplace_args<T>(op);
  B2_TRY(apply_cells_with_exchange(c, dst, src, [&](int64_t g0, int64_t g1, cudaStream_t s) {
    return run_laplace<T>(c, false, (const T *)src, (T *)dst, oa, g0, g1, s, dot);
  }));
  return identity_rows ? constrained_identity<T>(c, dst, src, dot) : 0;
}

static int laplace_apply(b200mf_laplace *op, void *dst, const void *src, ApplyMode mode, double *dot = nullptr) {
  if (!op || !dst || !src) return set_error(B200MF_ERR_ARG, "null argument");
  if (dst == src) return set_error(B200MF_ERR_ARG, "vmult: dst aliases src");
  b200mf_ctx *c = op->ctx;
  if (mode == APPLY_VMULT)
    B2_CUDA(cudaMemsetAsync(dst, 0, c->n_int * esize(c), c->stream));
  else if (mode == APPLY_ADD)
    B2_TRY(zero_ghosts(c, dst));  // see nh_apply
  const bool id = mode != APPLY_CELLS;
  int rc = c->dtype == B200MF_F64 ? laplace_apply_t<double>(op, dst, src, dot, id) : laplace_apply_t<float>(op, dst, src, dot, id);
  B2_TRY(rc);
  B2_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int b200mf_laplace_vmult(b200mf_laplace *op, void *dst, const void *src) { return laplace_apply(op, dst, src, APPLY_VMULT); }
extern "C" int b200mf_laplace_vmult_add(b200mf_laplace *op, void *dst, const void *src) { return laplace_apply(op, dst, src, APPLY_ADD); }
extern "C" int b200mf_laplace_Tvmult(b200mf_laplace *op, void *dst, const void *src) { return laplace_apply(op, dst, src, APPLY_VMULT); }
extern "C" int b200mf_laplace_Tvmult_add(b200mf_laplace *op, void *dst, const void *src) { return laplace_apply(op, dst, src, APPLY_ADD); }
extern "C" int b200mf_laplace_vmult_cells(b200mf_laplace *op, void *dst, const void *src) { return laplace_apply(op, dst, src, APPLY_CELLS); }
extern "C" uint64_t b200mf_laplace_diagonal_generation(const b200mf_laplace *op) { return op ? op->diag_gen : 0; }
// MatrixFreeOperators::Base::vmult_interface_down / _up: dst = 0, then the refinement-edge rows.  The reference only refines
// globally (mf_elasticity.h:638, heat_transfer.h:421-425: no hanging nodes, no refinement edges on any level), so there are no
// edge rows and both are "dst = 0".
static int laplace_interface(b200mf_laplace *op, void *dst) {
  if (!op || !dst) return set_error(B200MF_ERR_ARG, "null argument");
  B2_CUDA(cudaMemsetAsync(dst, 0, op->ctx->n_int * esize(op->ctx), op->ctx->stream));
  return 0;
}
extern "C" int b200mf_laplace_vmult_interface_down(b200mf_laplace *op, void *dst, const void *) { return laplace_interface(op, dst); }
extern "C" int b200mf_laplace_vmult_interface_up(b200mf_laplace *op, void *dst, const void *) { return laplace_interface(op, dst); }
extern "C" int64_t b200mf_laplace_m(const b200mf_laplace *op) { return op ? op->ctx->n_dofs : -1; }
// Base::el(row, row): the diagonal entry (asserts row == col and an initialised diagonal like deal.II)
extern "C" int b200mf_laplace_el(b200mf_laplace *op, int64_t row, double *value) {
  if (!op || !value) return set_error(B200MF_ERR_ARG, "null argument");
  if (!op->have_diag) return set_error(B200MF_ERR_STATE, "el(): call compute_diagonal() first");
  return read_entry(op->ctx, op->diag, row, value);
}

extern "C" int b200mf_laplace_vmult_dot(b200mf_laplace *op, void *dst, const void *src, double *dot_device) {
  if (!dot_device) return set_error(B200MF_ERR_ARG, "null argument");
  return laplace_apply(op, dst, src, APPLY_VMULT, dot_device);
}

extern "C" int b200mf_laplace_vmult_dot_host(b200mf_laplace *op, void *dst, const void *src, double *result_host) {
  if (!op || !result_host) return set_error(B200MF_ERR_ARG, "null argument");
  B2_TRY(fused_dot_slot(op->ctx));
  B2_TRY(laplace_apply(op, dst, src, APPLY_VMULT, op->ctx->scalars));
  return fused_dot_result(op->ctx, result_host);
}

extern "C" int b200mf_laplace_vmult_host(b200mf_laplace *op, double *dst, const double *src) {
  if (!op || !dst || !src) return set_error(B200MF_ERR_ARG, "null argument");
  b200mf_ctx *c = op->ctx;
  if (!op->tmp_src) B2_CUDA(cudaMalloc(&op->tmp_src, c->n_int * esize(c)));
  if (!op->tmp_dst) B2_CUDA(cudaMalloc(&op->tmp_dst, c->n_int * esize(c)));
  B2_TRY(b200mf_vec_upload_f64(c, op->tmp_src, src));
  B2_TRY(laplace_apply(op, op->tmp_dst, op->tmp_src, APPLY_VMULT));
  return b200mf_vec_download_f64(c, dst, op->tmp_dst);
}

// ---- cell loop of LaplaceProblem::assemble_rhs (heat_transfer.h:599-615) ---------------------
extern "C" int b200mf_laplace_assemble_rhs(b200mf_laplace *op, void *rhs, const void *solution_old, const double *source_q,
                                           double source_const) {
  if (!op || !rhs || !solution_old) return set_error(B200MF_ERR_ARG, "null argument");
  b200mf_ctx *c = op->ctx;
  const int N = c->N, N3 = N * N * N;
  DeviceScratch tab, nat;  // freed on every return path (cudaFree waits for the work that still uses them)
  void *&src_tab = tab.p;
  if (source_q) {  // f at the quadrature points, natural [cell][q] -> group layout
    const int64_t gs = table_stride_rt(c, 1);
    const size_t nel = (size_t)c->n_groups * gs;
    B2_CUDA(cudaMalloc(&src_tab, nel * esize(c)));
    B2_CUDA(cudaMemsetAsync(src_tab, 0, nel * esize(c), c->stream));
    B2_CUDA(cudaMalloc(&nat.p, std::max<size_t>(8, (size_t)c->n_cells * N3 * sizeof(double))));
    double *dn = (double *)nat.p;
    B2_CUDA(cudaMemcpyAsync(dn, source_q, (size_t)c->n_cells * N3 * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    const int64_t padded = c->n_groups * c->cpg;
    if (c->dtype == B200MF_F64)
      table_to_group_layout_kernel<double><<<nblocks(padded * N3, 256), 256, 0, c->stream>>>(N, c->cpg, c->L, 1, table_geom_rt(c, 1), c->n_cells, padded, dn, (double *)src_tab, 0, c->cell_order);
    else
      table_to_group_layout_kernel<float><<<nblocks(padded * N3, 256), 256, 0, c->stream>>>(N, c->cpg, c->L, 1, table_geom_rt(c, 1), c->n_cells, padded, dn, (float *)src_tab, 0, c->cell_order);
    B2_LAUNCHED();
    B2_CUDA(cudaStreamSynchronize(c->stream));  // the host array may be reused by the caller
  }
  B2_TRY(update_ghost_values(c, const_cast<void *>(solution_old)));  // :599
  int rc;
  if (c->dtype == B200MF_F64) {
    LaplaceRhsArgs<double> oa{(const double *)c->mref, (const double *)src_tab, (const double *)c->mcell, op->delta_t, source_const};
    rc = run_laplace_rhs<double>(c, (const double *)solution_old, (double *)rhs, oa);
  } else {
    LaplaceRhsArgs<float> oa{(const float *)c->mref, (const float *)src_tab, (const float *)c->mcell, (float)op->delta_t, (float)source_const};
    rc = run_laplace_rhs<float>(c, (const float *)solution_old, (float *)rhs, oa);
  }
  if (src_tab) cudaStreamSynchronize(c->stream);
  return rc;
}

extern "C" int b200mf_laplace_compute_diagonal(b200mf_laplace *op) {
  if (!op) return set_error(B200MF_ERR_ARG, "null op");
  b200mf_ctx *c = op->ctx;
  if (!op->diag) {
    B2_CUDA(cudaMalloc(&op->diag, c->n_int * esize(c)));
    B2_CUDA(cudaMalloc(&op->inv_diag, c->n_int * esize(c)));
    B2_CUDA(cudaMemsetAsync(op->inv_diag, 0, c->n_int * esize(c), c->stream));
    B2_CUDA(cudaMalloc(&op->n_bad, sizeof(unsigned long long)));
  }
  int rc0 = finish_diagonal(c, op->diag, op->inv_diag, op->n_bad, [&]() {
    int rc = c->dtype == B200MF_F64 ? run_laplace<double>(c, true, nullptr, (double *)op->diag, laplace_args<double>(op), 0, c->n_groups, c->stream)
                                    : run_laplace<float>(c, true, nullptr, (float *)op->diag, laplace_args<float>(op), 0, c->n_groups, c->stream);
    return rc;
  });
  B2_TRY(rc0);
  op->have_diag = true;
  ++op->diag_gen;
  return 0;
}
extern "C" int b200mf_laplace_diagonal(b200mf_laplace *op, const void **diag, const void **inv) {
  if (!op) return set_error(B200MF_ERR_ARG, "null op");
  if (!op->have_diag) return set_error(B200MF_ERR_STATE, "diagonal not computed");
  if (diag) *diag = op->diag;
  if (inv) *inv = op->inv_diag;
  return 0;
}
extern "C" int b200mf_laplace_diagonal_host(b200mf_laplace *op, double *diag, double *inv) {
  if (!op) return set_error(B200MF_ERR_ARG, "null op");
  if (!op->have_diag) return set_error(B200MF_ERR_STATE, "diagonal not computed");
  if (diag) B2_TRY(b200mf_vec_download_f64(op->ctx, diag, op->diag));
  if (inv) B2_TRY(b200mf_vec_download_f64(op->ctx, inv, op->inv_diag));
  return 0;
}
