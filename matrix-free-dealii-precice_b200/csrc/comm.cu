// Multi-GPU plumbing of libb200mf: NCCL communicator (dlopen'ed, so the library has no
// link-time NCCL dependency and shares the copy the host process already loaded), the
// vector partitioner (Utilities::MPI::Partitioner) and update_ghost_values / compress(add)
// (reference call sites mf_nh_operator.h:610, 620).
#include <dlfcn.h>
#include <nccl.h>

#include <cstring>
#include <memory>

#include "common.h"

namespace b200mf {

struct NcclApi {
  void *lib = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  ncclResult_t (*Send)(const void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Recv)(void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
  const char *(*GetErrorString)(ncclResult_t) = nullptr;
};

static NcclApi g_nccl;

static int load_nccl() {
  if (g_nccl.lib) return 0;
  const char *names[] = {"libnccl.so.2", "libnccl.so"};
  void *h = nullptr;
  for (const char *n : names) {
    h = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
    if (h) break;
  }
  if (!h) return set_error(B200MF_ERR_NCCL, "cannot dlopen libnccl.so.2: %s", dlerror());
#define B2_SYM(field, name)                                                                    \
  g_nccl.field = (decltype(g_nccl.field))dlsym(h, name);                                       \
  if (!g_nccl.field) return set_error(B200MF_ERR_NCCL, "libnccl: missing symbol %s", name);
  B2_SYM(GetUniqueId, "ncclGetUniqueId")
  B2_SYM(CommInitRank, "ncclCommInitRank")
  B2_SYM(CommDestroy, "ncclCommDestroy")
  B2_SYM(GroupStart, "ncclGroupStart")
  B2_SYM(GroupEnd, "ncclGroupEnd")
  B2_SYM(Send, "ncclSend")
  B2_SYM(Recv, "ncclRecv")
  B2_SYM(AllReduce, "ncclAllReduce")
  B2_SYM(GetErrorString, "ncclGetErrorString")
#undef B2_SYM
  g_nccl.lib = h;
  return 0;
}

#define B2_NCCL(call)                                                                                         \
  do {                                                                                                        \
    ncclResult_t r_ = (call);                                                                                 \
    if (r_ != ncclSuccess)                                                                                    \
      return b200mf::set_error(B200MF_ERR_NCCL, "%s:%d: %s -> %s", __FILE__, __LINE__, #call, g_nccl.GetErrorString(r_)); \
  } while (0)

template <typename T>
__global__ void pack_kernel(int64_t n, const uint32_t *__restrict__ idx, const T *__restrict__ vec, T *__restrict__ buf) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) buf[i] = vec[idx[i]];
}
template <typename T>
__global__ void unpack_set_kernel(int64_t n, const uint32_t *__restrict__ idx, T *__restrict__ vec, const T *__restrict__ buf) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) vec[idx[i]] = buf[i];
}
// an owned entry can be ghosted by several neighbours -> atomics
template <typename T>
__global__ void unpack_add_kernel(int64_t n, const uint32_t *__restrict__ idx, T *__restrict__ vec, const T *__restrict__ buf) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) atomicAdd(vec + idx[i], buf[i]);
}
template <typename T>
__global__ void zero_indexed_kernel(int64_t n, const uint32_t *__restrict__ idx, T *__restrict__ vec) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) vec[idx[i]] = T(0);
}

static inline unsigned nblk(int64_t n) { return (unsigned)((n + 255) / 256); }

}  // namespace b200mf

using namespace b200mf;

struct b200mf_comm {
  ncclComm_t comm;
  int rank, world;
};

struct b200mf_partitioner {
  b200mf_comm *comm;
  int nn;
  std::vector<int> ranks;
  std::vector<int64_t> send_off, recv_off;  // [nn+1] element offsets in the exchange buffers
  uint32_t *d_send_idx, *d_recv_idx;        // internal indices
  void *buf_send, *buf_recv;
  int64_t total_send, total_recv;
};

extern "C" int b200mf_comm_unique_id(void *id128) {
  if (!id128) return set_error(B200MF_ERR_ARG, "null argument");
  B2_TRY(load_nccl());
  ncclUniqueId id;
  B2_NCCL(g_nccl.GetUniqueId(&id));
  memcpy(id128, &id, sizeof id);
  return 0;
}

extern "C" int b200mf_comm_create(b200mf_comm **out, int rank, int world, const void *id128) {
  if (!out || !id128 || world < 1 || rank < 0 || rank >= world) return set_error(B200MF_ERR_ARG, "bad argument");
  B2_TRY(load_nccl());
  ncclUniqueId id;
  memcpy(&id, id128, sizeof id);
  std::unique_ptr<b200mf_comm> c(new b200mf_comm());
  c->rank = rank;
  c->world = world;
  B2_NCCL(g_nccl.CommInitRank(&c->comm, world, id, rank));
  *out = c.release();
  return 0;
}

extern "C" int b200mf_comm_destroy(b200mf_comm *c) {
  if (!c) return 0;
  if (g_nccl.lib) g_nccl.CommDestroy(c->comm);
  delete c;
  return 0;
}

extern "C" int b200mf_ctx_set_partitioner(b200mf_ctx *c, b200mf_comm *comm, const b200mf_partitioner_desc *d) {
  if (!c || !comm || !d) return set_error(B200MF_ERR_ARG, "null argument");
  if (d->n_neighbors < 0) return set_error(B200MF_ERR_ARG, "bad n_neighbors");
  std::unique_ptr<b200mf_partitioner> p(new b200mf_partitioner());
  p->comm = comm;
  p->nn = d->n_neighbors;
  p->d_send_idx = p->d_recv_idx = nullptr;
  p->buf_send = p->buf_recv = nullptr;
  p->send_off.assign(1, 0);
  p->recv_off.assign(1, 0);
  std::vector<uint32_t> sidx, ridx;
  auto to_int = [&](int64_t ext) -> int64_t { return c->h_ext2int ? (*c->h_ext2int)[ext] : ext; };
  for (int i = 0; i < p->nn; ++i) {
    const int r = d->neighbor_ranks[i];
    if (r < 0 || r >= comm->world || r == comm->rank) return set_error(B200MF_ERR_ARG, "bad neighbour rank %d", r);
    p->ranks.push_back(r);
    for (int64_t k = d->send_offsets[i]; k < d->send_offsets[i + 1]; ++k) {
      const int64_t e = d->send_indices[k];
      if (e >= c->n_owned) return set_error(B200MF_ERR_ARG, "import index %lld is not locally owned", (long long)e);
      sidx.push_back((uint32_t)to_int(e));
    }
    for (int64_t k = 0; k < d->ghost_counts[i]; ++k) {
      const int64_t e = d->ghost_offsets[i] + k;
      if (e < c->n_owned || e >= c->n_dofs) return set_error(B200MF_ERR_ARG, "ghost index %lld outside the ghost range", (long long)e);
      ridx.push_back((uint32_t)to_int(e));
    }
    p->send_off.push_back((int64_t)sidx.size());
    p->recv_off.push_back((int64_t)ridx.size());
  }
  p->total_send = (int64_t)sidx.size();
  p->total_recv = (int64_t)ridx.size();
  const size_t es = c->dtype == B200MF_F64 ? 8 : 4;
  B2_CUDA(cudaMalloc(&p->d_send_idx, std::max<size_t>(4, sidx.size() * 4)));
  B2_CUDA(cudaMalloc(&p->d_recv_idx, std::max<size_t>(4, ridx.size() * 4)));
  B2_CUDA(cudaMemcpy(p->d_send_idx, sidx.data(), sidx.size() * 4, cudaMemcpyHostToDevice));
  B2_CUDA(cudaMemcpy(p->d_recv_idx, ridx.data(), ridx.size() * 4, cudaMemcpyHostToDevice));
  B2_CUDA(cudaMalloc(&p->buf_send, std::max<size_t>(8, sidx.size() * es)));
  B2_CUDA(cudaMalloc(&p->buf_recv, std::max<size_t>(8, ridx.size() * es)));
  if (c->part) {
    cudaFree(c->part->d_send_idx);
    cudaFree(c->part->d_recv_idx);
    cudaFree(c->part->buf_send);
    cudaFree(c->part->buf_recv);
    delete c->part;
  }
  c->part = p.release();
  return 0;
}

namespace b200mf {
void free_partitioner(b200mf_ctx *c) {
  if (c && c->part) {
    cudaFree(c->part->d_send_idx);
    cudaFree(c->part->d_recv_idx);
    cudaFree(c->part->buf_send);
    cudaFree(c->part->buf_recv);
    delete c->part;
    c->part = nullptr;
  }
}

// forward == true : owner -> ghost (update_ghost_values); false: ghost -> owner accumulate (compress(add))
template <typename T>
static int exchange(b200mf_ctx *c, T *vec, bool forward, cudaStream_t s) {
  b200mf_partitioner *p = c->part;
  if (!p || p->nn == 0) return 0;
  const ncclDataType_t dt = sizeof(T) == 8 ? ncclFloat64 : ncclFloat32;
  if (!s) s = c->stream;
  T *bs = (T *)p->buf_send, *br = (T *)p->buf_recv;
  if (forward) {
    if (p->total_send) {
      pack_kernel<T><<<nblk(p->total_send), 256, 0, s>>>(p->total_send, p->d_send_idx, vec, bs);
      B2_LAUNCHED();
    }
  } else if (p->total_recv) {
    pack_kernel<T><<<nblk(p->total_recv), 256, 0, s>>>(p->total_recv, p->d_recv_idx, vec, br);
    B2_LAUNCHED();
  }
  B2_NCCL(g_nccl.GroupStart());
  for (int i = 0; i < p->nn; ++i) {
    const int64_t ns = p->send_off[i + 1] - p->send_off[i], nr = p->recv_off[i + 1] - p->recv_off[i];
    if (forward) {
      if (ns) B2_NCCL(g_nccl.Send(bs + p->send_off[i], (size_t)ns, dt, p->ranks[i], p->comm->comm, s));
      if (nr) B2_NCCL(g_nccl.Recv(br + p->recv_off[i], (size_t)nr, dt, p->ranks[i], p->comm->comm, s));
    } else {
      if (nr) B2_NCCL(g_nccl.Send(br + p->recv_off[i], (size_t)nr, dt, p->ranks[i], p->comm->comm, s));
      if (ns) B2_NCCL(g_nccl.Recv(bs + p->send_off[i], (size_t)ns, dt, p->ranks[i], p->comm->comm, s));
    }
  }
  B2_NCCL(g_nccl.GroupEnd());
  if (forward) {
    if (p->total_recv) {
      unpack_set_kernel<T><<<nblk(p->total_recv), 256, 0, s>>>(p->total_recv, p->d_recv_idx, vec, br);
      B2_LAUNCHED();
    }
  } else {
    if (p->total_send) {
      unpack_add_kernel<T><<<nblk(p->total_send), 256, 0, s>>>(p->total_send, p->d_send_idx, vec, bs);
      B2_LAUNCHED();
    }
    if (p->total_recv) {  // compress zeroes the ghost entries
      zero_indexed_kernel<T><<<nblk(p->total_recv), 256, 0, s>>>(p->total_recv, p->d_recv_idx, vec);
      B2_LAUNCHED();
    }
  }
  return 0;
}

bool has_neighbors(const b200mf_ctx *c) { return c->part && c->part->nn > 0; }
int update_ghost_values(b200mf_ctx *c, void *vec, cudaStream_t s) {
  return c->dtype == B200MF_F64 ? exchange<double>(c, (double *)vec, true, s) : exchange<float>(c, (float *)vec, true, s);
}
int compress_add(b200mf_ctx *c, void *vec, cudaStream_t s) {
  return c->dtype == B200MF_F64 ? exchange<double>(c, (double *)vec, false, s) : exchange<float>(c, (float *)vec, false, s);
}
int zero_ghosts(b200mf_ctx *c, void *vec) {
  b200mf_partitioner *p = c->part;
  if (!p || !p->total_recv) return 0;
  if (c->dtype == B200MF_F64)
    zero_indexed_kernel<double><<<nblk(p->total_recv), 256, 0, c->stream>>>(p->total_recv, p->d_recv_idx, (double *)vec);
  else
    zero_indexed_kernel<float><<<nblk(p->total_recv), 256, 0, c->stream>>>(p->total_recv, p->d_recv_idx, (float *)vec);
  B2_LAUNCHED();
  return 0;
}
// sum over ranks of n doubles living on the device (in place), on the context's stream
int allreduce_sum_f64(b200mf_ctx *c, double *dptr, int n) {
  b200mf_partitioner *p = c->part;
  if (!p || p->comm->world == 1) return 0;
  B2_NCCL(g_nccl.AllReduce(dptr, dptr, (size_t)n, ncclFloat64, ncclSum, p->comm->comm, c->stream));
  return 0;
}
// sum over ranks of a device counter (failure counts: every rank must take the same error path, or the peers of a
// failing rank would walk into the next exchange alone)
int allreduce_sum_u64(b200mf_ctx *c, unsigned long long *dptr, int n) {
  b200mf_partitioner *p = c->part;
  if (!p || p->comm->world == 1) return 0;
  B2_NCCL(g_nccl.AllReduce(dptr, dptr, (size_t)n, ncclUint64, ncclSum, p->comm->comm, c->stream));
  return 0;
}
}  // namespace b200mf

extern "C" int b200mf_vec_update_ghost_values(b200mf_ctx *c, void *dvec) {
  if (!c || !dvec) return set_error(B200MF_ERR_ARG, "null argument");
  return update_ghost_values(c, dvec);
}
extern "C" int b200mf_vec_compress_add(b200mf_ctx *c, void *dvec) {
  if (!c || !dvec) return set_error(B200MF_ERR_ARG, "null argument");
  return compress_add(c, dvec);
}
extern "C" int b200mf_vec_zero_out_ghost_values(b200mf_ctx *c, void *dvec) {
  if (!c || !dvec) return set_error(B200MF_ERR_ARG, "null argument");
  return zero_ghosts(c, dvec);
}
