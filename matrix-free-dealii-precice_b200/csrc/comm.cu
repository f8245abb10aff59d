// Multi-GPU plumbing of libb200mf: NCCL communicator (dlopen'ed, so the library has no
// link-time NCCL dependency and shares the copy the host process already loaded), the
// vector partitioner (Utilities::MPI::Partitioner) and update_ghost_values / compress(add)
// (reference call sites mf_nh_operator.h:610, 620).
#include <dlfcn.h>
#include <nccl.h>

#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <vector>

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

// ---------------------------------------------------------------------------
// Peer-to-peer exchange over NVLink (all ranks on one box): every rank exports one "exchange block" through CUDA IPC
//   [ state: 2 epochs + 2 block counters ][ per neighbour: flag_fwd, flag_cmp, ack_fwd, ack_cmp ][ fwd slots 0/1 ][ cmp slots 0/1 ]
// and the sender stores its contribution straight into the receiver's slot (no pack -> ncclSend/ncclRecv -> unpack, no
// NCCL proxy on the critical path; 2 launches per exchange instead of 3 + the NCCL kernel).
//   push:  data -> peer slot (epoch parity picks the slot); the last block publishes the epoch in the peers' flag words
//   pull:  wait for the flags of all contributing neighbours, copy / accumulate the slot into the vector, acknowledge
// A slot is reused every second exchange; before writing it the sender checks the receiver's acknowledgement of the
// exchange before last.  Epochs live in device memory (bumped by the pull kernel), so the launches are CUDA-graph safe.
constexpr int MAX_NB = 26;
struct PeerTable {
  int nn;
  int64_t send_off[MAX_NB + 1], recv_off[MAX_NB + 1];
  // remote addresses (in the neighbours' exchange blocks), per neighbour
  unsigned long long *r_flag[2][MAX_NB];  // [dir] flag word I raise at the neighbour when my data has landed there
  unsigned long long *r_ack[2][MAX_NB];   // [dir] acknowledgement word I raise at the neighbour after consuming its data
  char *r_data[2][2][MAX_NB];             // [dir][slot] where my segment for this neighbour starts
  // local block
  unsigned long long *state;              // [0] fwd epoch, [1] cmp epoch, [2] push counter, [3] pull counter
  unsigned long long *flag[2], *ack[2];   // [dir][neighbour]
  char *data[2][2];                       // [dir][slot] local receive slots (fwd: total_recv, cmp: total_send elements)
};

__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long *p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_sys(unsigned long long *p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

// dir 0 (update_ghost_values): element k of my import list goes to the owner-side segment of neighbour i
// dir 1 (compress(add))      : element k of my ghost range goes back to its owner and is zeroed here
template <typename T, int DIR>
__global__ void p2p_push_kernel(const __grid_constant__ PeerTable pt, const uint32_t *__restrict__ idx, T *__restrict__ vec) {
  const unsigned long long epoch = pt.state[DIR] + 1;  // this exchange
  const int slot = (int)(epoch & 1ull);
  const int64_t *off = DIR == 0 ? pt.send_off : pt.recv_off;
  const int64_t n = off[pt.nn];
  // the slot was last used by exchange epoch-2: wait until every receiver has acknowledged it (normally long done)
  if (epoch > 2) {
    for (int i = threadIdx.x; i < pt.nn; i += blockDim.x)
      if (off[i + 1] > off[i])
        while (ld_acquire_sys(pt.ack[DIR] + i) < epoch - 2) {
        }
    __syncthreads();
  }
  const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (k < n) {
    int i = 0;
    while (k >= off[i + 1]) ++i;
    const uint32_t e = idx[k];
    T *dst = reinterpret_cast<T *>(pt.r_data[DIR][slot][i]);
    dst[k - off[i]] = vec[e];
    if (DIR == 1) vec[e] = T(0);
  }
  // publish: all blocks' stores must be visible at the peers before the flags are
  __threadfence_system();
  __syncthreads();
  __shared__ bool last;
  if (threadIdx.x == 0) last = atomicAdd(pt.state + 2, 1ull) == gridDim.x - 1;
  __syncthreads();
  if (last) {
    __threadfence_system();
    for (int i = threadIdx.x; i < pt.nn; i += blockDim.x)
      if (off[i + 1] > off[i]) st_release_sys(pt.r_flag[DIR][i], epoch);
    if (threadIdx.x == 0) pt.state[2] = 0;
  }
}

template <typename T, int DIR>
__global__ void p2p_pull_kernel(const __grid_constant__ PeerTable pt, const uint32_t *__restrict__ idx, T *__restrict__ vec) {
  const unsigned long long epoch = pt.state[DIR] + 1;
  const int slot = (int)(epoch & 1ull);
  const int64_t *off = DIR == 0 ? pt.recv_off : pt.send_off;  // what I receive: my ghost range / contributions to my import list
  const int64_t n = off[pt.nn];
  for (int i = threadIdx.x; i < pt.nn; i += blockDim.x)
    if (off[i + 1] > off[i])
      while (ld_acquire_sys(pt.flag[DIR] + i) < epoch) {
      }
  __syncthreads();
  const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (k < n) {
    const T v = __ldcg(reinterpret_cast<const T *>(pt.data[DIR][slot]) + k);  // written by a peer: L2 is the point of coherence
    if (DIR == 0)
      vec[idx[k]] = v;
    else
      atomicAdd(vec + idx[k], v);  // an owned entry can be ghosted by several neighbours
  }
  __syncthreads();
  __shared__ bool last;
  if (threadIdx.x == 0) last = atomicAdd(pt.state + 3, 1ull) == gridDim.x - 1;
  __syncthreads();
  if (last) {  // everything of this exchange has been consumed: acknowledge, then open the next epoch
    for (int i = threadIdx.x; i < pt.nn; i += blockDim.x)
      if (off[i + 1] > off[i]) st_release_sys(pt.r_ack[DIR][i], epoch);
    __syncthreads();
    if (threadIdx.x == 0) {
      pt.state[3] = 0;
      __threadfence();
      pt.state[DIR] = epoch;
    }
  }
}

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
  // peer-to-peer path (null/false: NCCL send/recv)
  bool p2p = false;
  void *xblock = nullptr;            // my exchange block (IPC-exported)
  std::vector<void *> peer_blocks;   // IPC-mapped blocks of the neighbours
  PeerTable table;
};

struct PeerHello {  // what a rank tells each neighbour about its exchange block
  cudaIpcMemHandle_t handle;
  int64_t my_index_of_you, nn, total_recv, total_send, fwd_off, cmp_off, elem;
};
static inline size_t xblock_state_bytes(int nn) { return (size_t)(4 + 4 * nn) * sizeof(unsigned long long); }
static inline size_t xblock_bytes(int nn, int64_t total_recv, int64_t total_send, size_t es) {
  return ((xblock_state_bytes(nn) + 255) / 256) * 256 + 2 * ((size_t)total_recv + (size_t)total_send) * es + 1024;
}
// addresses inside an exchange block laid out as described above
static void xblock_layout(char *base, int nn, int64_t total_recv, int64_t total_send, size_t es, unsigned long long **state,
                          unsigned long long *flag[2], unsigned long long *ack[2], char *data[2][2]) {
  unsigned long long *w = reinterpret_cast<unsigned long long *>(base);
  *state = w;
  flag[0] = w + 4;
  flag[1] = w + 4 + nn;
  ack[0] = w + 4 + 2 * nn;
  ack[1] = w + 4 + 3 * nn;
  char *d = base + ((xblock_state_bytes(nn) + 255) / 256) * 256;
  data[0][0] = d;
  data[0][1] = d + (size_t)total_recv * es;
  data[1][0] = d + 2 * (size_t)total_recv * es;
  data[1][1] = d + 2 * (size_t)total_recv * es + (size_t)total_send * es;
}

// exchange block + handshake with the neighbours (handles travel through NCCL); any failure leaves the NCCL path in place
static int setup_p2p(b200mf_ctx *c, b200mf_partitioner *p) {
  const char *env = getenv("B200MF_P2P");
  if (env && atoi(env) == 0) return 0;
  if (p->nn == 0 || p->nn > MAX_NB) return 0;
  const size_t es = c->dtype == B200MF_F64 ? 8 : 4;
  const size_t bytes = xblock_bytes(p->nn, p->total_recv, p->total_send, es);
  if (cudaMalloc(&p->xblock, bytes) != cudaSuccess) {
    cudaGetLastError();
    p->xblock = nullptr;
    return 0;
  }
  B2_CUDA(cudaMemset(p->xblock, 0, bytes));
  PeerTable &t = p->table;
  memset(&t, 0, sizeof t);
  t.nn = p->nn;
  for (int i = 0; i <= p->nn; ++i) {
    t.send_off[i] = p->send_off[i];
    t.recv_off[i] = p->recv_off[i];
  }
  xblock_layout((char *)p->xblock, p->nn, p->total_recv, p->total_send, es, &t.state, t.flag, t.ack, t.data);
  // hello messages: one per neighbour, through device buffers
  std::vector<PeerHello> out(p->nn), in(p->nn);
  cudaIpcMemHandle_t mine;
  int ok = cudaIpcGetMemHandle(&mine, p->xblock) == cudaSuccess;
  if (!ok) cudaGetLastError();
  for (int i = 0; i < p->nn; ++i) {
    memset(&out[i], 0, sizeof(PeerHello));
    out[i].handle = mine;
    out[i].my_index_of_you = ok ? i : -1;  // -1: "I cannot do peer-to-peer" - both sides fall back
    out[i].nn = p->nn;
    out[i].total_recv = p->total_recv;
    out[i].total_send = p->total_send;
    out[i].fwd_off = p->recv_off[i];  // where neighbour i's forward data starts inside my fwd slots
    out[i].cmp_off = p->send_off[i];  // where neighbour i's compress data starts inside my cmp slots
    out[i].elem = (int64_t)es;
  }
  PeerHello *dout = nullptr, *din = nullptr;
  B2_CUDA(cudaMalloc(&dout, sizeof(PeerHello) * p->nn));
  B2_CUDA(cudaMalloc(&din, sizeof(PeerHello) * p->nn));
  B2_CUDA(cudaMemcpy(dout, out.data(), sizeof(PeerHello) * p->nn, cudaMemcpyHostToDevice));
  B2_NCCL(g_nccl.GroupStart());
  for (int i = 0; i < p->nn; ++i) {
    B2_NCCL(g_nccl.Send(dout + i, sizeof(PeerHello), ncclUint8, p->ranks[i], p->comm->comm, c->stream));
    B2_NCCL(g_nccl.Recv(din + i, sizeof(PeerHello), ncclUint8, p->ranks[i], p->comm->comm, c->stream));
  }
  B2_NCCL(g_nccl.GroupEnd());
  B2_CUDA(cudaStreamSynchronize(c->stream));
  B2_CUDA(cudaMemcpy(in.data(), din, sizeof(PeerHello) * p->nn, cudaMemcpyDeviceToHost));
  cudaFree(dout);
  cudaFree(din);
  p->peer_blocks.assign(p->nn, nullptr);
  for (int i = 0; i < p->nn && ok; ++i) {
    if (in[i].my_index_of_you < 0 || in[i].elem != (int64_t)es) {
      ok = 0;
      break;
    }
    if (cudaIpcOpenMemHandle(&p->peer_blocks[i], in[i].handle, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
      cudaGetLastError();
      p->peer_blocks[i] = nullptr;
      ok = 0;
      break;
    }
    unsigned long long *rstate, *rflag[2], *rack[2];
    char *rdata[2][2];
    xblock_layout((char *)p->peer_blocks[i], (int)in[i].nn, in[i].total_recv, in[i].total_send, es, &rstate, rflag, rack, rdata);
    const int j = (int)in[i].my_index_of_you;  // my slot in the neighbour's tables
    for (int dir = 0; dir < 2; ++dir) {
      t.r_flag[dir][i] = rflag[dir] + j;
      t.r_ack[dir][i] = rack[dir] + j;
      for (int slot = 0; slot < 2; ++slot)
        t.r_data[dir][slot][i] = rdata[dir][slot] + (size_t)(dir == 0 ? in[i].fwd_off : in[i].cmp_off) * es;
    }
  }
  // all ranks must agree (a rank pushing into a peer that waits inside NCCL would hang): min over the communicator
  double *agree = nullptr;
  B2_CUDA(cudaMalloc(&agree, sizeof(double)));
  const double mine_ok = ok ? 0.0 : 1.0;
  B2_CUDA(cudaMemcpy(agree, &mine_ok, sizeof(double), cudaMemcpyHostToDevice));
  B2_NCCL(g_nccl.AllReduce(agree, agree, 1, ncclFloat64, ncclSum, p->comm->comm, c->stream));
  double failures = 0.0;
  B2_CUDA(cudaMemcpyAsync(&failures, agree, sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  B2_CUDA(cudaStreamSynchronize(c->stream));
  cudaFree(agree);
  p->p2p = failures == 0.0;
  return 0;
}

static void free_p2p(b200mf_partitioner *p) {
  for (void *b : p->peer_blocks)
    if (b) cudaIpcCloseMemHandle(b);
  p->peer_blocks.clear();
  cudaFree(p->xblock);
  p->xblock = nullptr;
  p->p2p = false;
}

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
  g_config_epoch.fetch_add(1);
  if (!c) return 0;
  if (g_nccl.lib) g_nccl.CommDestroy(c->comm);
  delete c;
  return 0;
}

extern "C" int b200mf_ctx_set_partitioner(b200mf_ctx *c, b200mf_comm *comm, const b200mf_partitioner_desc *d) {
  g_config_epoch.fetch_add(1);
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
  B2_TRY(setup_p2p(c, p.get()));
  if (c->part) free_partitioner(c);
  // with the peer-to-peer exchange the overlapped schedule (interior cells behind the exchanges) pays: measured on B200,
  // Q3, 5e7 DoF per GPU: 1.425 vs 1.480 ms at 2 ranks, 1.437 vs 1.499 ms at 4 (profiles/r02_scale*.json); with NCCL
  // send/recv it does not (round 1), so the serial order stays the default there
  c->overlap = getenv("B200MF_OVERLAP") ? atoi(getenv("B200MF_OVERLAP")) : (p->p2p ? 1 : 0);
  c->part = p.release();
  return 0;
}

// 1 if the ghost exchanges of this context go peer to peer over NVLink (CUDA IPC), 0 if they use ncclSend/ncclRecv
extern "C" int b200mf_ctx_exchange_is_p2p(const b200mf_ctx *c) { return c && c->part && c->part->p2p ? 1 : 0; }

namespace b200mf {
void free_partitioner(b200mf_ctx *c) {
  if (c && c->part) {
    free_p2p(c->part);
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
  if (p->p2p) {
    // every rank launches both kernels (possibly on zero elements): the epoch advances in lockstep everywhere
    const int64_t n_push = forward ? p->total_send : p->total_recv, n_pull = forward ? p->total_recv : p->total_send;
    const uint32_t *i_push = forward ? p->d_send_idx : p->d_recv_idx, *i_pull = forward ? p->d_recv_idx : p->d_send_idx;
    if (forward) {
      p2p_push_kernel<T, 0><<<std::max(1u, nblk(n_push)), 256, 0, s>>>(p->table, i_push, vec);
      B2_LAUNCHED();
      p2p_pull_kernel<T, 0><<<std::max(1u, nblk(n_pull)), 256, 0, s>>>(p->table, i_pull, vec);
      B2_LAUNCHED();
    } else {
      p2p_push_kernel<T, 1><<<std::max(1u, nblk(n_push)), 256, 0, s>>>(p->table, i_push, vec);
      B2_LAUNCHED();
      p2p_pull_kernel<T, 1><<<std::max(1u, nblk(n_pull)), 256, 0, s>>>(p->table, i_pull, vec);
      B2_LAUNCHED();
    }
    return 0;
  }
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
// element-wise reduction of a whole device vector over the ranks of c's partitioner (op 0 = sum, 1 = max, 2 = min): the
// replicated coarse levels of the multigrid transfer
int allreduce_vector(b200mf_ctx *c, void *dptr, int64_t n, int dtype, int op, cudaStream_t s) {
  b200mf_partitioner *p = c->part;
  if (!p || p->comm->world == 1) return 0;
  const ncclRedOp_t ops[3] = {ncclSum, ncclMax, ncclMin};
  B2_NCCL(g_nccl.AllReduce(dptr, dptr, (size_t)n, dtype == B200MF_F64 ? ncclFloat64 : ncclFloat32, ops[op], p->comm->comm, s));
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
