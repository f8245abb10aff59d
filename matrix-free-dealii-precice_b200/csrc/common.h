// Host-side internals of libb200mf (not part of the ABI).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>
#include <cstdio>
#include <string>
#include <vector>

#include "../../include/b200mf.h"

struct b200mf_ctx;

namespace b200mf {

extern thread_local std::string g_last_error;
extern std::atomic<int64_t> g_launches;
// bumped by every call that changes what a recorded launch sequence would look like (streams, schedules, operator tables,
// operator constants): captured CUDA graphs (multigrid coarse levels, solver.cu) are re-recorded when it moves
extern std::atomic<uint64_t> g_config_epoch;

int set_error(int code, const char *fmt, ...);

#define B2_CUDA(call)                                                                                  \
  do {                                                                                                 \
    cudaError_t e_ = (call);                                                                           \
    if (e_ != cudaSuccess)                                                                             \
      return b200mf::set_error(B200MF_ERR_CUDA, "%s:%d: %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e_)); \
  } while (0)

#define B2_TRY(call)            \
  do {                          \
    int rc_ = (call);           \
    if (rc_ != 0) return rc_;   \
  } while (0)

#define B2_LAUNCHED()                      \
  do {                                     \
    b200mf::g_launches.fetch_add(1);       \
    B2_CUDA(cudaGetLastError());           \
  } while (0)

// 1-D basis on [0,1]: FE_Q(p) on Gauss-Lobatto points, QGauss(n)
struct Shape1D {
  int n;
  double S[49], D[49], C[49], w[7], xq[7], nodes[7];
};
void compute_shape_1d(int p, Shape1D &out);

// device scratch that is released on every return path (early B2_TRY / B2_CUDA returns included)
struct DeviceScratch {
  void *p = nullptr;
  DeviceScratch() = default;
  DeviceScratch(const DeviceScratch &) = delete;
  DeviceScratch &operator=(const DeviceScratch &) = delete;
  ~DeviceScratch() {
    if (p) cudaFree(p);
  }
};

int kernel_event(b200mf_ctx *c, cudaStream_t s);  // b200mf.cu

// comm.cu
void free_partitioner(b200mf_ctx *c);
void free_hostpipe(b200mf_ctx *c);  // b200mf.cu
int update_ghost_values(b200mf_ctx *c, void *vec, cudaStream_t stream = nullptr);  // nullptr: the context's stream
int compress_add(b200mf_ctx *c, void *vec, cudaStream_t stream = nullptr);
int zero_ghosts(b200mf_ctx *c, void *vec);
bool has_neighbors(const b200mf_ctx *c);
int allreduce_sum_f64(b200mf_ctx *c, double *dptr, int n);
int allreduce_sum_u64(b200mf_ctx *c, unsigned long long *dptr, int n);
int allreduce_vector(b200mf_ctx *c, void *dptr, int64_t n, int dtype, int op, cudaStream_t s);

}  // namespace b200mf

struct b200mf_partitioner;

struct b200mf_ctx {
  int degree, N, NC, dtype;
  int64_t n_cells, n_groups, n_dofs, n_owned;  // caller's local vector size
  int64_t n_nodes, n_owned_nodes, cstride, n_int;  // internal vector: NC blocks of cstride (>= n_nodes) + orphans
  int cpg, L;                                  // cells per group, dense lanes per group
  cudaStream_t stream;
  int sm_count;
  // device data
  uint32_t *nidx;        // [group][N][L]
  void *mref;            // T [group][MF_COUNT][N][L] reference-configuration metric (null: scalar context on affine cells)
  int64_t *cell_order;   // library cell -> caller's cell (boundary cells first) or nullptr when identity
  int64_t n_boundary_groups;  // groups [0, n_boundary_groups) touch ghost nodes
  cudaStream_t comm_stream;   // exchanges that overlap the interior cells run here
  cudaEvent_t ev[4];
  int overlap;                // 1: hide the exchanges behind interior cells (opt-in, default 0); -1: no exchange (measurement)
  double *scalars;            // a few device doubles for fused reductions (lazily allocated)
  void *mcell;           // T [cell][MC_COUNT] per-cell metric when every cell is affine, else null
  uint8_t *matid;        // [n_cells] or nullptr
  int64_t *int2ext;      // [n_int] or nullptr when identity
  uint32_t *constrained; // internal indices of constrained DoFs
  int64_t n_constrained;
  void *stage;           // staging buffer, n_dofs elements of double
  size_t bytes;
  b200mf::Shape1D shape;
  b200mf_partitioner *part;         // nullptr on a single rank
  std::vector<int64_t> *h_ext2int;  // caller index -> internal index, nullptr when identity
  std::vector<cudaEvent_t> *kernel_events;  // non-null: event pairs recorded around every cell-loop launch (b200mf_ctx_time_kernels)
  std::vector<uint32_t> *h_cell_nodes;  // [n_cells][N^3] node | constrained flags (bits 29..31), caller's cell order (MG transfer set-up)
  struct b200mf_hostpipe *hostpipe;     // lazily built schedule of the pipelined host-vector vmult (b200mf_*_vmult_host)
};
