// Solver-side work of the hot path on the device: the deal.II callees the
// reference instantiates on its operators — PreconditionChebyshev (mf_elasticity.h:371-372, 1149-1180;
// heat_transfer.h:764-788), MGTransferMatrixFree<dim,float> (mf_elasticity.h:872-874, 920-922),
// Multigrid V-cycle + PreconditionMG (:1209-1224, 1272-1276), SolverCG (:2024-2051; heat_transfer.h:810-811) —
// restated from their documented algorithms on top of the library's operators.
// Everything is ordered on the context's stream; CG reads two scalars per iteration back to the host,
// the Chebyshev smoother and the whole V-cycle run without any host synchronisation.
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <memory>

#include "kernels.h"

namespace b200mf {

static inline size_t esz(const b200mf_ctx *c) { return c->dtype == B200MF_F64 ? 8 : 4; }
static inline unsigned vgrid(int64_t n, const b200mf_ctx *c) {
  return (unsigned)std::max<int64_t>(1, std::min<int64_t>((n + 255) / 256, (int64_t)c->sm_count * 8));
}

// ---------------------------------------------------------------------------
// vector kernels (grid-stride).  Reductions run over the OWNED entries only (LinearAlgebra::distributed::Vector
// semantics): internal index i = c*cstride + node is owned iff node < n_owned_nodes.
__device__ __forceinline__ double block_sum(double v) {
  __shared__ double red[32];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  if (l == 0) red[w] = v;
  __syncthreads();
  v = (threadIdx.x < (blockDim.x >> 5)) ? red[threadIdx.x] : 0.0;
  if (w == 0) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
  }
  __syncthreads();
  return v;  // valid in thread 0
}
struct Own {
  int64_t n_main, cstride, n_owned_nodes;
  int nc;
  // i = c * cstride + node is owned iff node < n_owned_nodes; the component blocks are peeled off by comparison (nc <= 3) - a
  // 64-bit modulo per vector entry costs more instructions than the update it guards
  __device__ __forceinline__ bool operator()(int64_t i) const {
    if (i >= n_main) return false;
    for (int c = 1; c < nc; ++c) i -= i >= cstride ? cstride : 0;
    return i < n_owned_nodes;
  }
};
static Own own_of(const b200mf_ctx *c) { return Own{c->cstride * c->NC, c->cstride, c->n_owned_nodes, c->NC}; }

// Streaming loops over a vector: per trip a thread takes U block-strided entries and issues the loads of all of them
// (load(u, i)) before the first store (store(u, i)), so that U independent cache lines per array are in flight per thread - a
// plain grid-stride loop that writes what it reads is serialised by the compiler behind each store.
constexpr int VU = 4;
template <class Load, class Store>
__device__ __forceinline__ void stream_loop(int64_t n, Load load, Store store) {
  const int64_t tile = (int64_t)blockDim.x * VU;
  for (int64_t base = (int64_t)blockIdx.x * tile + threadIdx.x; base < n; base += (int64_t)gridDim.x * tile) {
#pragma unroll
    for (int u = 0; u < VU; ++u) {
      const int64_t i = base + (int64_t)u * blockDim.x;
      if (i < n) load(u, i);
    }
#pragma unroll
    for (int u = 0; u < VU; ++u) {
      const int64_t i = base + (int64_t)u * blockDim.x;
      if (i < n) store(u, i);
    }
  }
}

template <typename T>
__global__ void dot_kernel(Own own, const T *__restrict__ a, const T *__restrict__ b, double *out) {
  double s = 0.0;
  T av[VU], bv[VU];
  stream_loop(
      own.n_main,
      [&](int u, int64_t i) {
        av[u] = a[i];
        bv[u] = b[i];
      },
      [&](int u, int64_t i) {
        if (own(i)) s += (double)av[u] * (double)bv[u];
      });
  s = block_sum(s);
  if (threadIdx.x == 0) atomicAdd(out, s);
}
// y = a*x + b*y
template <typename T>
__global__ void axpby_kernel(int64_t n, T a, const T *__restrict__ x, T b, T *__restrict__ y) {
  T xv[VU], yv[VU];
  stream_loop(
      n,
      [&](int u, int64_t i) {
        xv[u] = x[i];
        yv[u] = b == T(0) ? T(0) : y[i];
      },
      [&](int u, int64_t i) { y[i] = a * xv[u] + (b == T(0) ? T(0) : b * yv[u]); });
}
// SolverCG: x += alpha p ; r -= alpha v ; out += r.r  (add_and_dot)
template <typename T>
__device__ __forceinline__ void cg_update_body(int64_t n, const Own &own, T alpha, const T *__restrict__ p, const T *__restrict__ v,
                                               T *__restrict__ x, T *__restrict__ r, double *out) {
  double s = 0.0;
  T pv[VU], vv[VU], xv[VU], rv[VU];
  stream_loop(
      n,
      [&](int u, int64_t i) {
        pv[u] = p[i];
        vv[u] = v[i];
        xv[u] = x[i];
        rv[u] = r[i];
      },
      [&](int u, int64_t i) {
        x[i] = xv[u] + alpha * pv[u];
        const T rn = rv[u] - alpha * vv[u];
        r[i] = rn;
        if (own(i)) s += (double)rn * (double)rn;
      });
  s = block_sum(s);
  if (threadIdx.x == 0) atomicAdd(out, s);
}
template <typename T>
__global__ void cg_update_kernel(int64_t n, Own own, T alpha, const T *__restrict__ p, const T *__restrict__ v, T *__restrict__ x,
                                 T *__restrict__ r, double *out) {
  cg_update_body<T>(n, own, alpha, p, v, x, r, out);
}
// z = omega * inv .* r ; out += r.z   (inv == nullptr: z = r)
template <typename T>
__global__ void diag_precond_dot_kernel(int64_t n, Own own, T omega, const T *__restrict__ inv, const T *__restrict__ r, T *__restrict__ z,
                                        double *out) {
  double s = 0.0;
  T rv[VU], iv[VU];
  stream_loop(
      n,
      [&](int u, int64_t i) {
        rv[u] = r[i];
        iv[u] = inv ? inv[i] : T(1);
      },
      [&](int u, int64_t i) {
        const T zi = inv ? omega * iv[u] * rv[u] : rv[u];
        z[i] = zi;
        if (own(i)) s += (double)rv[u] * (double)zi;
      });
  s = block_sum(s);
  if (threadIdx.x == 0) atomicAdd(out, s);
}
// PreconditionChebyshev update:
//   first (zero start):  dst = f2 * inv .* src ; old = 0
//   general:             new = dst + f1*(dst - old) + f2 * inv .* (src - t) ; old = dst ; dst = new
template <typename T>
__global__ void cheb_first_kernel(int64_t n, T f2, const T *__restrict__ inv, const T *__restrict__ src, T *__restrict__ dst,
                                  T *__restrict__ old) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    dst[i] = f2 * inv[i] * src[i];
    old[i] = T(0);
  }
}
template <typename T>
__global__ void cheb_update_kernel(int64_t n, T f1, T f2, const T *__restrict__ inv, const T *__restrict__ src, const T *__restrict__ t,
                                   T *__restrict__ dst, T *__restrict__ old) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const T d = dst[i], o = old[i];
    old[i] = d;
    dst[i] = d + f1 * (d - o) + f2 * inv[i] * (src[i] - t[i]);
  }
}
// The same update as the operator's "after" hook: t = A_cells dst comes without identity rows and is consumed here, so
//   - constrained rows (marked by a negative entry of invm = -1/diag = -1) use t := dst (A acts as the identity there),
//   - t is zeroed on the way out: the next operator pass accumulates into it without a memset of its own.
template <typename T>
__global__ void cheb_update_hook_kernel(int64_t n, T f1, T f2, const T *__restrict__ invm, const T *__restrict__ src, T *__restrict__ t,
                                        T *__restrict__ dst, T *__restrict__ old) {
  // t is read and written: all loads of a batch are issued before its stores so that they stay in flight together
  constexpr int U = 4;
  const int64_t tile = (int64_t)blockDim.x * U;
  for (int64_t base = (int64_t)blockIdx.x * tile + threadIdx.x; base < n; base += (int64_t)gridDim.x * tile) {
    T d[U], o[U], iv[U], ti[U], sr[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int64_t i = base + (int64_t)u * blockDim.x;
      if (i < n) {
        d[u] = dst[i];
        o[u] = old[i];
        iv[u] = invm[i];
        ti[u] = t[i];
        sr[u] = src[i];
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int64_t i = base + (int64_t)u * blockDim.x;
      if (i < n) {
        T a = iv[u], tv = ti[u];
        if (a < T(0)) {
          a = -a;
          tv = d[u];
        }
        t[i] = T(0);
        old[i] = d[u];
        dst[i] = d[u] + f1 * (d[u] - o[u]) + f2 * a * (sr[u] - tv);
      }
    }
  }
}
template <typename T>
__global__ void mark_inverse_kernel(int64_t n, const T *__restrict__ inv, T *__restrict__ invm) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) invm[i] = inv[i];
}
template <typename T>
__global__ void negate_indexed_kernel(int64_t n, const uint32_t *__restrict__ idx, T *__restrict__ v) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) v[idx[i]] = -fabs(v[idx[i]]);
}
template <typename TO, typename TI>
__global__ void convert_kernel(int64_t n, const TI *__restrict__ in, TO *__restrict__ out) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) out[i] = (TO)in[i];
}
// deal.II's deterministic Chebyshev start vector: (global index % 11), then the mean is removed (A.5);
// index = caller numbering (int2ext), constrained entries are zeroed afterwards
template <typename T>
__global__ void cheb_guess_kernel(int64_t n, Own own, const int64_t *__restrict__ int2ext, int64_t first_global, T *__restrict__ v,
                                  double *sum) {
  double s = 0.0;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    T val = T(0);
    if (own(i)) {
      const int64_t e = int2ext ? int2ext[i] : i;
      val = (T)((e + first_global) % 11);
      s += (double)val;
    }
    v[i] = val;
  }
  s = block_sum(s);
  if (threadIdx.x == 0) atomicAdd(sum, s);
}
template <typename T>
__global__ void shift_owned_kernel(int64_t n, Own own, T shift, T *__restrict__ v) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    if (own(i)) v[i] += shift;
}
template <typename T>
__global__ void zero_indexed_kernel2(int64_t n, const uint32_t *__restrict__ idx, T *__restrict__ v) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) v[idx[i]] = T(0);
}

// ---------------------------------------------------------------------------
// a linear operator as the solvers see it (duck-typed operator surface of mf_nh_operator.h:102-173)
struct LinOp {
  b200mf_ctx *ctx = nullptr;
  std::function<int(void *, const void *)> vmult;
  std::function<int(void *, const void *, double *)> vmult_dot;  // dst = A src and *slot += src . (A src) in one operator pass
  std::function<int(const void **)> inv_diag;
  // the smoother hook (SURVEY.md §8b "Optional fusion hook"): operator pass without the zeroing of dst before and the
  // identity rows after it - both move into the smoother's vector update (cheb_update_hook_kernel)
  std::function<int(void *, const void *)> vmult_cells;
  std::function<uint64_t()> diag_gen;  // changes whenever the operator's diagonal was recomputed
};

struct Scratch {  // a few device doubles + pinned host mirror for the reductions
  double *d = nullptr, *h = nullptr;
  int init() {
    if (d) return 0;
    B2_CUDA(cudaMalloc(&d, 8 * sizeof(double)));
    B2_CUDA(cudaMallocHost(&h, 8 * sizeof(double)));
    return 0;
  }
  ~Scratch() {
    cudaFree(d);
    cudaFreeHost(h);
  }
};

static int vec_alloc(b200mf_ctx *c, void **p) {
  B2_CUDA(cudaMalloc(p, c->n_int * esz(c)));
  B2_CUDA(cudaMemsetAsync(*p, 0, c->n_int * esz(c), c->stream));
  return 0;
}

template <typename T>
static int dot_t(b200mf_ctx *c, const void *a, const void *b, double *dslot) {
  dot_kernel<T><<<vgrid(c->n_int, c), 256, 0, c->stream>>>(own_of(c), (const T *)a, (const T *)b, dslot);
  B2_LAUNCHED();
  return 0;
}
// host value of `n` device slots after the cross-rank sum
static int fetch(b200mf_ctx *c, Scratch &s, int n) {
  B2_TRY(allreduce_sum_f64(c, s.d, n));
  B2_CUDA(cudaMemcpyAsync(s.h, s.d, n * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  B2_CUDA(cudaStreamSynchronize(c->stream));
  return 0;
}

// extreme eigenvalues of the symmetric tridiagonal Lanczos matrix (Sturm bisection)
static void tridiag_extreme(const std::vector<double> &d, const std::vector<double> &e, double &lmin, double &lmax) {
  const int n = (int)d.size();
  if (n == 0) {
    lmin = lmax = 1.0;
    return;
  }
  double lo = d[0], hi = d[0];
  for (int i = 0; i < n; ++i) {
    const double r = (i > 0 ? std::fabs(e[i - 1]) : 0.0) + (i + 1 < n ? std::fabs(e[i]) : 0.0);
    lo = std::min(lo, d[i] - r);
    hi = std::max(hi, d[i] + r);
  }
  auto count_below = [&](double x) {  // number of eigenvalues < x
    int cnt = 0;
    double q = 1.0;
    for (int i = 0; i < n; ++i) {
      const double e2 = i > 0 ? e[i - 1] * e[i - 1] : 0.0;
      q = d[i] - x - (i > 0 ? e2 / q : 0.0);
      if (q == 0.0) q = 1e-300;
      if (q < 0.0) ++cnt;
    }
    return cnt;
  };
  auto kth = [&](int k) {  // k-th smallest (0-based)
    double a = lo, b = hi;
    for (int it = 0; it < 200 && (b - a) > 1e-15 * std::max(std::fabs(a), std::fabs(b)); ++it) {
      const double m = 0.5 * (a + b);
      if (count_below(m) > k)
        b = m;
      else
        a = m;
    }
    return 0.5 * (a + b);
  };
  lmin = kth(0);
  lmax = kth(n - 1);
}

// ---------------------------------------------------------------------------
// SolverCG.  precond: 0 none, 1 diagonal (omega * inv_diag), 2 callback
struct CGPrecond {
  int kind = 0;
  double omega = 1.0;
  const void *inv = nullptr;
  std::function<int(void *, const void *)> apply;
};
struct CGResult {
  int iterations = 0;
  double residual = 0.0, initial_residual = 0.0;
  bool converged = false;
  std::vector<double> diag, offdiag;  // Lanczos tridiagonal (connect_eigenvalues_slot)
};

// reduction > 0: ReductionControl - also converged once the residual has dropped to reduction * (initial residual)
template <typename T>
static int cg_solve_t(LinOp &A, const CGPrecond &P, T *x, const T *b, double tol, int max_it, bool iteration_number_control,
                      bool x_is_zero, void *const work[3], Scratch &sc, CGResult &res, double reduction = 0.0) {
  b200mf_ctx *c = A.ctx;
  const int64_t n = c->n_int;
  T *r = (T *)work[0], *p = (T *)work[1], *v = (T *)work[2];
  cudaStream_t st = c->stream;
  const Own own = own_of(c);
  const unsigned g = vgrid(n, c);
  B2_TRY(sc.init());
  // start-up: r = b - A x
  if (x_is_zero) {
    B2_CUDA(cudaMemcpyAsync(r, b, n * sizeof(T), cudaMemcpyDeviceToDevice, st));
  } else {
    B2_TRY(A.vmult(r, x));
    axpby_kernel<T><<<g, 256, 0, st>>>(n, T(1), b, T(-1), r);
    B2_LAUNCHED();
  }
  B2_CUDA(cudaMemsetAsync(sc.d, 0, 8 * sizeof(double), st));
  B2_TRY(dot_t<T>(c, r, r, sc.d));
  B2_TRY(fetch(c, sc, 1));
  double res_norm = std::sqrt(std::fabs(sc.h[0]));
  res.initial_residual = res_norm;
  if (reduction > 0.0) tol = std::max(tol, reduction * res_norm);
  double rPr = 0.0, alpha = 0.0, beta = 0.0;
  int it = 0;
  auto done = [&](int step, double val) {
    if (val <= tol) return 1;
    if (step >= max_it) return iteration_number_control ? 1 : -1;
    return 0;
  };
  int state = done(0, res_norm);
  while (state == 0) {
    ++it;
    const double old_rPr = rPr;
    const T *direction;
    B2_CUDA(cudaMemsetAsync(sc.d, 0, 8 * sizeof(double), st));
    if (P.kind == 0) {
      rPr = res_norm * res_norm;
      direction = r;
    } else {
      if (P.kind == 1) {
        diag_precond_dot_kernel<T><<<g, 256, 0, st>>>(n, own, (T)P.omega, (const T *)P.inv, r, v, sc.d);
        B2_LAUNCHED();
      } else {
        B2_TRY(P.apply(v, r));
        B2_TRY(dot_t<T>(c, r, v, sc.d));
      }
      B2_TRY(fetch(c, sc, 1));
      rPr = sc.h[0];
      direction = v;
    }
    const double previous_alpha = alpha;
    if (it > 1) {
      beta = rPr / old_rPr;
      axpby_kernel<T><<<g, 256, 0, st>>>(n, T(1), direction, (T)beta, p);  // p = direction + beta p
    } else {
      axpby_kernel<T><<<g, 256, 0, st>>>(n, T(1), direction, T(0), p);
    }
    B2_LAUNCHED();
    B2_CUDA(cudaMemsetAsync(sc.d, 0, 8 * sizeof(double), st));
    if (A.vmult_dot) {
      B2_TRY(A.vmult_dot(v, p, sc.d));  // p . Ap accumulated inside the cell loop
    } else {
      B2_TRY(A.vmult(v, p));
      B2_TRY(dot_t<T>(c, p, v, sc.d));
    }
    B2_TRY(fetch(c, sc, 1));
    const double pAp = sc.h[0];
    alpha = rPr / pAp;
    B2_CUDA(cudaMemsetAsync(sc.d, 0, 8 * sizeof(double), st));
    cg_update_kernel<T><<<g, 256, 0, st>>>(n, own, (T)alpha, p, v, x, r, sc.d);
    B2_LAUNCHED();
    B2_TRY(fetch(c, sc, 1));
    res_norm = std::sqrt(std::fabs(sc.h[0]));
    if (it > 1) {
      res.diag.push_back(1.0 / alpha + beta / previous_alpha);
      res.offdiag.push_back(std::sqrt(beta) / previous_alpha);
    } else {
      res.diag.push_back(1.0 / alpha);
    }
    state = done(it, res_norm);
    if (!(res_norm == res_norm)) state = -1;  // NaN
  }
  res.iterations = it;
  res.residual = res_norm;
  res.converged = state == 1;
  return 0;
}

// ---------------------------------------------------------------------------
// Sync-free CG for the eigenvalue estimate of PreconditionChebyshev (fixed iteration count, diagonal preconditioner):
// alpha and beta never leave the device; every iteration enqueues its kernels without a host round trip (the
// synchronising loop above costs 3 stream synchronisations per iteration, which dominated the coarse levels).
// Device scalars S[]: 0 r.z (new)  1 r.z (old)  2 p.Ap  3 r.r  4 alpha  5 beta;  history H[3][max_it + 1] = alpha, beta, r.r
template <typename T>
__global__ void axpby_dev_kernel(int64_t n, const T *__restrict__ x, const double *__restrict__ b, T *__restrict__ y) {
  const T beta = (T)*b;
  T xv[VU], yv[VU];
  stream_loop(
      n,
      [&](int u, int64_t i) {
        xv[u] = x[i];
        yv[u] = beta == T(0) ? T(0) : y[i];
      },
      [&](int u, int64_t i) { y[i] = xv[u] + (beta == T(0) ? T(0) : beta * yv[u]); });
}
template <typename T>
__global__ void cg_update_dev_kernel(int64_t n, Own own, const double *__restrict__ a, const T *__restrict__ p, const T *__restrict__ v,
                                     T *__restrict__ x, T *__restrict__ r, double *out) {
  cg_update_body<T>(n, own, (T)*a, p, v, x, r, out);
}
// after r.z: beta = r.z / (r.z)_old, keep the previous r.r in the history, clear the accumulators of this iteration
__global__ void cg_scalars_beta_kernel(double *S, double *H, int it, int stride) {
  if (it > 1) H[2 * stride + (it - 1)] = S[3];
  const double beta = it > 1 ? S[0] / S[1] : 0.0;
  S[5] = beta;
  H[stride + it] = beta;
  S[1] = S[0];
  S[0] = 0.0;
  S[2] = 0.0;
  S[3] = 0.0;
}
// after p.Ap: alpha = r.z / p.Ap
__global__ void cg_scalars_alpha_kernel(double *S, double *H, int it) {
  const double alpha = S[1] / S[2];
  S[4] = alpha;
  H[it] = alpha;
}
__global__ void cg_scalars_final_kernel(double *S, double *H, int it, int stride) { H[2 * stride + it] = S[3]; }

template <typename T>
static int cg_fixed_iterations_t(LinOp &A, double omega, const void *inv, T *x, const T *b, double tol, int max_it, void *const work[3],
                                 Scratch &sc, CGResult &res) {
  b200mf_ctx *c = A.ctx;
  const int64_t n = c->n_int;
  T *r = (T *)work[0], *p = (T *)work[1], *v = (T *)work[2];
  cudaStream_t st = c->stream;
  const Own own = own_of(c);
  const unsigned g = vgrid(n, c);
  B2_TRY(sc.init());
  const int stride = max_it + 1;
  DeviceScratch history;  // released on every return path
  B2_CUDA(cudaMalloc(&history.p, 3 * (size_t)stride * sizeof(double)));
  double *H = (double *)history.p;
  B2_CUDA(cudaMemsetAsync(H, 0, 3 * (size_t)stride * sizeof(double), st));
  B2_CUDA(cudaMemcpyAsync(r, b, n * sizeof(T), cudaMemcpyDeviceToDevice, st));  // x = 0
  B2_CUDA(cudaMemsetAsync(sc.d, 0, 8 * sizeof(double), st));
  B2_TRY(dot_t<T>(c, r, r, sc.d));
  B2_TRY(fetch(c, sc, 1));
  res.initial_residual = std::sqrt(std::fabs(sc.h[0]));
  res.diag.clear();
  res.offdiag.clear();
  if (res.initial_residual <= tol || max_it <= 0) {
    res.iterations = 0;
    res.residual = res.initial_residual;
    res.converged = true;
    return 0;
  }
  B2_CUDA(cudaMemsetAsync(sc.d, 0, 8 * sizeof(double), st));
  double *S = sc.d;
  // the residual history is inspected every CHUNK iterations (one host read), so a solve that reaches the tolerance
  // early (level 0 asks for up to m() iterations, heat_transfer.h:780-782) stops within CHUNK iterations of it
  constexpr int CHUNK = 32;
  std::vector<double> h(3 * (size_t)stride, 0.0);
  bool rr_reduced = false;  // r.r of the previous iteration already summed over the ranks
  int done_it = 0, last_it = 0;
  for (int it = 1; it <= max_it && !done_it; ++it) {
    diag_precond_dot_kernel<T><<<g, 256, 0, st>>>(n, own, (T)omega, (const T *)inv, r, v, S + 0);
    B2_LAUNCHED();
    B2_TRY(allreduce_sum_f64(c, S + 0, 1));
    if (it > 1 && !rr_reduced) B2_TRY(allreduce_sum_f64(c, S + 3, 1));
    rr_reduced = false;
    cg_scalars_beta_kernel<<<1, 1, 0, st>>>(S, H, it, stride);
    B2_LAUNCHED();
    axpby_dev_kernel<T><<<g, 256, 0, st>>>(n, v, S + 5, p);  // p = z + beta p
    B2_LAUNCHED();
    if (A.vmult_dot) {
      B2_TRY(A.vmult_dot(v, p, S + 2));
    } else {
      B2_TRY(A.vmult(v, p));
      B2_TRY(dot_t<T>(c, p, v, S + 2));
    }
    B2_TRY(allreduce_sum_f64(c, S + 2, 1));
    cg_scalars_alpha_kernel<<<1, 1, 0, st>>>(S, H, it);
    B2_LAUNCHED();
    cg_update_dev_kernel<T><<<g, 256, 0, st>>>(n, own, S + 4, p, v, x, r, S + 3);
    B2_LAUNCHED();
    last_it = it;
    if (it % CHUNK == 0 || it == max_it) {
      B2_TRY(allreduce_sum_f64(c, S + 3, 1));
      rr_reduced = true;
      cg_scalars_final_kernel<<<1, 1, 0, st>>>(S, H, it, stride);
      B2_LAUNCHED();
      const int first = it - ((it - 1) % CHUNK);  // first iteration of this chunk
      B2_CUDA(cudaMemcpyAsync(h.data() + 2 * stride + first, H + 2 * stride + first, (size_t)(it - first + 1) * sizeof(double),
                              cudaMemcpyDeviceToHost, st));
      B2_CUDA(cudaStreamSynchronize(st));
      for (int k = first; k <= it; ++k) {
        const double rn = std::sqrt(std::fabs(h[2 * stride + k]));
        if (rn <= tol || !(rn == rn)) {
          done_it = k;
          break;
        }
      }
    }
  }
  B2_CUDA(cudaMemcpyAsync(h.data(), H, 2 * (size_t)stride * sizeof(double), cudaMemcpyDeviceToHost, st));  // alpha, beta
  B2_CUDA(cudaStreamSynchronize(st));
  const int n_run = done_it ? done_it : last_it;
  // the stopping rule of the synchronising loop, replayed on the recorded history
  const int n_it = n_run;
  for (int it = 1; it <= n_it; ++it) {
    const double alpha = h[it], beta = h[stride + it];
    if (it > 1) {
      res.diag.push_back(1.0 / alpha + beta / h[it - 1]);
      res.offdiag.push_back(std::sqrt(beta) / h[it - 1]);
    } else {
      res.diag.push_back(1.0 / alpha);
    }
  }
  res.iterations = n_it;
  res.residual = std::sqrt(std::fabs(h[2 * stride + n_it]));
  res.converged = true;
  return 0;
}

}  // namespace b200mf

using namespace b200mf;

// ===========================================================================
// PreconditionChebyshev
struct b200mf_chebyshev {
  LinOp A;
  b200mf_chebyshev_data data;
  bool eig_ready = false;
  double min_eig = 1.0, max_eig = 1.0, theta = 1.0, delta = 0.0;
  int degree = 1, cg_iterations = 0;
  void *old = nullptr, *tmp = nullptr, *w[3] = {nullptr, nullptr, nullptr};
  Scratch sc;
  // smoother hook: inverse diagonal with the constrained rows marked by their sign, rebuilt when the operator's diagonal moves
  void *invm = nullptr;
  const void *invm_of = nullptr;
  uint64_t invm_gen = ~0ull;
};

static LinOp linop_nh(b200mf_nh *op, b200mf_ctx *ctx) {
  LinOp A;
  A.ctx = ctx;
  A.vmult = [op](void *d, const void *s) { return b200mf_nh_vmult(op, d, s); };
  A.vmult_dot = [op](void *d, const void *s, double *slot) { return b200mf_nh_vmult_dot(op, d, s, slot); };
  A.inv_diag = [op](const void **inv) { return b200mf_nh_diagonal(op, nullptr, inv); };
  A.vmult_cells = [op](void *d, const void *s) { return b200mf_nh_vmult_cells(op, d, s); };
  A.diag_gen = [op]() { return b200mf_nh_diagonal_generation(op); };
  return A;
}
static LinOp linop_laplace(b200mf_laplace *op, b200mf_ctx *ctx) {
  LinOp A;
  A.ctx = ctx;
  A.vmult = [op](void *d, const void *s) { return b200mf_laplace_vmult(op, d, s); };
  A.vmult_dot = [op](void *d, const void *s, double *slot) { return b200mf_laplace_vmult_dot(op, d, s, slot); };
  A.inv_diag = [op](const void **inv) { return b200mf_laplace_diagonal(op, nullptr, inv); };
  A.vmult_cells = [op](void *d, const void *s) { return b200mf_laplace_vmult_cells(op, d, s); };
  A.diag_gen = [op]() { return b200mf_laplace_diagonal_generation(op); };
  return A;
}

static LinOp linop_boundary_mass(b200mf_boundary_mass *op) {
  LinOp A;
  A.ctx = b200mf_boundary_mass_ctx(op);
  A.vmult = [op](void *d, const void *s) { return b200mf_boundary_mass_vmult(op, d, s); };
  A.inv_diag = [op](const void **inv) { return b200mf_boundary_mass_diagonal(op, nullptr, inv); };
  return A;
}

static int chebyshev_create(b200mf_chebyshev **out, LinOp A, const b200mf_chebyshev_data *d) {
  if (!out || !d) return set_error(B200MF_ERR_ARG, "null argument");
  if (d->degree == 0 || d->degree < -1) return set_error(B200MF_ERR_ARG, "Chebyshev degree must be >= 1 or -1 (solver mode)");
  std::unique_ptr<b200mf_chebyshev> ch(new b200mf_chebyshev());
  ch->A = A;
  ch->data = *d;
  B2_TRY(vec_alloc(A.ctx, &ch->old));
  B2_TRY(vec_alloc(A.ctx, &ch->tmp));
  *out = ch.release();
  return 0;
}
extern "C" int b200mf_chebyshev_create_nh(b200mf_chebyshev **out, b200mf_nh *op, const b200mf_chebyshev_data *d) {
  if (!op) return set_error(B200MF_ERR_ARG, "null op");
  return chebyshev_create(out, linop_nh(op, b200mf_nh_ctx(op)), d);
}
extern "C" int b200mf_chebyshev_create_laplace(b200mf_chebyshev **out, b200mf_laplace *op, const b200mf_chebyshev_data *d) {
  if (!op) return set_error(B200MF_ERR_ARG, "null op");
  return chebyshev_create(out, linop_laplace(op, b200mf_laplace_ctx(op)), d);
}
extern "C" int b200mf_chebyshev_destroy(b200mf_chebyshev *ch) {
  if (!ch) return 0;
  cudaFree(ch->old);
  cudaFree(ch->tmp);
  cudaFree(ch->invm);
  for (void *p : ch->w) cudaFree(p);
  delete ch;
  return 0;
}

// smoother hook: (re)build the sign-marked inverse diagonal when the operator's diagonal has moved
template <typename T>
static int cheb_refresh_marked_inverse(b200mf_chebyshev *ch, const void *invp) {
  b200mf_ctx *c = ch->A.ctx;
  const uint64_t gen = ch->A.diag_gen();
  if (ch->invm && ch->invm_of == invp && ch->invm_gen == gen) return 0;
  if (!ch->invm) B2_TRY(vec_alloc(c, &ch->invm));
  mark_inverse_kernel<T><<<vgrid(c->n_int, c), 256, 0, c->stream>>>(c->n_int, (const T *)invp, (T *)ch->invm);
  B2_LAUNCHED();
  if (c->n_constrained) {
    negate_indexed_kernel<T><<<(unsigned)((c->n_constrained + 255) / 256), 256, 0, c->stream>>>(c->n_constrained, c->constrained, (T *)ch->invm);
    B2_LAUNCHED();
  }
  ch->invm_of = invp;
  ch->invm_gen = gen;
  return 0;
}
static bool smoother_hook_enabled() {
  static const bool on = !(getenv("B200MF_SMOOTHER_HOOK") && atoi(getenv("B200MF_SMOOTHER_HOOK")) == 0);
  return on;
}

template <typename T>
static int cheb_estimate_t(b200mf_chebyshev *ch) {
  b200mf_ctx *c = ch->A.ctx;
  const void *inv = nullptr;
  B2_TRY(ch->A.inv_diag(&inv));
  double mn, mx;
  if (ch->data.eig_cg_n_iterations > 0) {
    for (void *&p : ch->w)
      if (!p) B2_TRY(vec_alloc(c, &p));
    B2_TRY(ch->sc.init());
    // rhs = deterministic high-frequency guess, x0 = 0 (A.5)
    T *rhs = (T *)ch->tmp, *x = (T *)ch->old;
    B2_CUDA(cudaMemsetAsync(ch->sc.d, 0, 8 * sizeof(double), c->stream));
    B2_CUDA(cudaMemsetAsync(x, 0, c->n_int * sizeof(T), c->stream));
    cheb_guess_kernel<T><<<vgrid(c->n_int, c), 256, 0, c->stream>>>(c->n_int, own_of(c), c->int2ext, ch->data.first_global_index, rhs, ch->sc.d);
    B2_LAUNCHED();
    B2_TRY(fetch(c, ch->sc, 1));
    const double n_global = ch->data.n_global_dofs > 0 ? (double)ch->data.n_global_dofs : (double)c->n_owned;
    shift_owned_kernel<T><<<vgrid(c->n_int, c), 256, 0, c->stream>>>(c->n_int, own_of(c), (T)(-ch->sc.h[0] / n_global), rhs);
    B2_LAUNCHED();
    if (c->n_constrained) {  // constraints.set_zero
      zero_indexed_kernel2<T><<<(unsigned)((c->n_constrained + 255) / 256), 256, 0, c->stream>>>(c->n_constrained, c->constrained, rhs);
      B2_LAUNCHED();
    }
    CGPrecond P;
    P.kind = 1;
    P.omega = 1.0;
    P.inv = inv;
    CGResult r;
    if (getenv("B200MF_SYNC_EIG_CG"))  // the synchronising loop (3 host round trips per iteration), kept for A/B runs
      B2_TRY(cg_solve_t<T>(ch->A, P, x, rhs, 1e-10, ch->data.eig_cg_n_iterations, true, true, ch->w, ch->sc, r));
    else
      B2_TRY(cg_fixed_iterations_t<T>(ch->A, P.omega, P.inv, x, rhs, 1e-10, ch->data.eig_cg_n_iterations, ch->w, ch->sc, r));
    ch->cg_iterations = r.iterations;
    if (r.diag.empty()) {
      mn = mx = 1.0;
    } else {
      tridiag_extreme(r.diag, r.offdiag, mn, mx);
      mx *= 1.2;  // safety factor: the CG estimate is in general not converged
    }
  } else {
    mx = ch->data.max_eigenvalue;
    mn = mx / ch->data.smoothing_range;
  }
  const double alpha = ch->data.smoothing_range > 1.0 ? mx / ch->data.smoothing_range : std::min(0.9 * mx, mn);
  ch->degree = ch->data.degree;
  if (ch->data.degree < 0) {  // solver mode: smoothing_range is the target reduction (Varga, sec. 5.1)
    const double actual_range = mx / alpha;
    const double sigma = (1.0 - std::sqrt(1.0 / actual_range)) / (1.0 + std::sqrt(1.0 / actual_range));
    const double eps = ch->data.smoothing_range;
    ch->degree = 1 + (int)(std::log(1.0 / eps + std::sqrt(1.0 / eps / eps - 1.0)) / std::log(1.0 / sigma));
  }
  ch->min_eig = mn;
  ch->max_eig = mx;
  ch->delta = (mx - alpha) * 0.5;
  ch->theta = (mx + alpha) * 0.5;
  ch->eig_ready = true;
  if (smoother_hook_enabled() && ch->A.vmult_cells && ch->A.diag_gen) {
    B2_TRY(cheb_refresh_marked_inverse<T>(ch, inv));  // here, not inside a recorded V-cycle
    // invariant of the hooked smoother: tmp is zero whenever no application is running (vec_alloc zero-fills, every hooked update
    // re-zeroes it) - the eigenvalue CG above used it as its right-hand side.  A recorded V-cycle relies on it as well.
    B2_CUDA(cudaMemsetAsync(ch->tmp, 0, c->n_int * sizeof(T), c->stream));
  }
  return 0;
}

extern "C" int b200mf_chebyshev_estimate_eigenvalues(b200mf_chebyshev *ch, double *min_eig, double *max_eig, int *cg_its, int *degree) {
  if (!ch) return set_error(B200MF_ERR_ARG, "null argument");
  B2_TRY(ch->A.ctx->dtype == B200MF_F64 ? cheb_estimate_t<double>(ch) : cheb_estimate_t<float>(ch));
  if (min_eig) *min_eig = ch->min_eig;
  if (max_eig) *max_eig = ch->max_eig;
  if (cg_its) *cg_its = ch->cg_iterations;
  if (degree) *degree = ch->degree;
  return 0;
}

// vmult (zero start: degree-1 matvecs) and step (degree matvecs), A.5
template <typename T>
static int cheb_apply_t(b200mf_chebyshev *ch, T *dst, const T *src, bool zero_start) {
  b200mf_ctx *c = ch->A.ctx;
  if (!ch->eig_ready) B2_TRY(cheb_estimate_t<T>(ch));
  const void *invp = nullptr;
  B2_TRY(ch->A.inv_diag(&invp));
  const T *inv = (const T *)invp;
  const int64_t n = c->n_int;
  const unsigned g = vgrid(n, c);
  T *old = (T *)ch->old, *t = (T *)ch->tmp;
  const bool hook = smoother_hook_enabled() && ch->A.vmult_cells && ch->A.diag_gen;
  const T *invm = nullptr;
  if (hook) {
    B2_TRY(cheb_refresh_marked_inverse<T>(ch, invp));
    invm = (const T *)ch->invm;
  }
  // one matrix-vector product + update: t = A dst, then the three-term recurrence
  auto step = [&](double f1, double f2) -> int {
    if (hook) {
      B2_TRY(ch->A.vmult_cells(t, dst));
      cheb_update_hook_kernel<T><<<g, 256, 0, c->stream>>>(n, (T)f1, (T)f2, invm, src, t, dst, old);
    } else {
      B2_TRY(ch->A.vmult(t, dst));
      cheb_update_kernel<T><<<g, 256, 0, c->stream>>>(n, (T)f1, (T)f2, inv, src, t, dst, old);
    }
    B2_LAUNCHED();
    return 0;
  };
  if (zero_start) {
    cheb_first_kernel<T><<<g, 256, 0, c->stream>>>(n, (T)(1.0 / ch->theta), inv, src, dst, old);
    B2_LAUNCHED();
  } else {
    B2_TRY(step(0.0, 1.0 / ch->theta));
  }
  if (ch->degree < 2 || std::fabs(ch->delta) < 1e-40) return 0;
  double rhok = ch->delta / ch->theta;
  const double sigma = ch->theta / ch->delta;
  for (int k = 0; k < ch->degree - 1; ++k) {
    const double rhokp = 1.0 / (2.0 * sigma - rhok);
    const double f1 = rhokp * rhok, f2 = 2.0 * rhokp / ch->delta;
    rhok = rhokp;
    B2_TRY(step(f1, f2));
  }
  return 0;
}
static int cheb_apply(b200mf_chebyshev *ch, void *dst, const void *src, bool zero_start) {
  if (!ch || !dst || !src) return set_error(B200MF_ERR_ARG, "null argument");
  if (dst == src) return set_error(B200MF_ERR_ARG, "Chebyshev: dst aliases src");
  return ch->A.ctx->dtype == B200MF_F64 ? cheb_apply_t<double>(ch, (double *)dst, (const double *)src, zero_start)
                                        : cheb_apply_t<float>(ch, (float *)dst, (const float *)src, zero_start);
}
extern "C" int b200mf_chebyshev_vmult(b200mf_chebyshev *ch, void *dst, const void *src) { return cheb_apply(ch, dst, src, true); }
extern "C" int b200mf_chebyshev_step(b200mf_chebyshev *ch, void *dst, const void *src) { return cheb_apply(ch, dst, src, false); }

// ===========================================================================
// MGTransferMatrixFree (A.8): uniformly refined parent/children, tensor-product 1-D embedding
namespace b200mf {
constexpr int MAXM = 2 * MAXN - 1;
struct Embed {  // in the number type of the level vectors: an FP64 hierarchy / interpolate_to_mg is not truncated to FP32
  double Pd[MAXM * MAXN];  // P[m][j] = l_j(x_m): fine patch node m (2p+1 per direction) <- coarse node j
  double Rd[MAXN * MAXM];  // R[i][m]: coarse node i <- fine patch nodes (evaluation of the fine function at x_i)
  float Pf[MAXM * MAXN];
  float Rf[MAXN * MAXM];
  template <typename S> __device__ __forceinline__ const S *P() const;
  template <typename S> __device__ __forceinline__ const S *R() const;
};
template <> __device__ __forceinline__ const double *Embed::P<double>() const { return Pd; }
template <> __device__ __forceinline__ const double *Embed::R<double>() const { return Rd; }
template <> __device__ __forceinline__ const float *Embed::P<float>() const { return Pf; }
template <> __device__ __forceinline__ const float *Embed::R<float>() const { return Rf; }

// one block per parent cell; NC components one after the other through shared memory
// MODE 0: prolongate (dst_fine = P src_coarse), 1: prolongate_and_add, 2: restrict_and_add (dst_coarse += P^T w.*src_fine),
// 3: interpolate (dst_coarse = R src_fine, plain values: interpolate_to_mg of the displacement)
template <typename T, int MODE>
__global__ void transfer_kernel(int N, int NC, const uint32_t *__restrict__ cidx /*[cell][N^3]*/, const uint32_t *__restrict__ pidx /*[cell][M^3]*/,
                                const float *__restrict__ wfine /*[fine node] 1/multiplicity*/, int64_t cstride_c, int64_t cstride_f,
                                const T *__restrict__ src, T *__restrict__ dst, const __grid_constant__ Embed E) {
  extern __shared__ __align__(16) unsigned char sm_raw[];
  using S = T;  // scratch and 1-D embedding in the number type of the level
  S *sm = reinterpret_cast<S *>(sm_raw);
  const S *EP = E.P<S>(), *ER = E.R<S>();
  const int M = 2 * N - 1;
  const int N3 = N * N * N, M3 = M * M * M;
  S *a = sm;           // up to M^3
  S *b = sm + M3;      // up to M^3
  const int64_t cell = blockIdx.x;
  const uint32_t *ci = cidx + cell * N3, *pi = pidx + cell * M3;
  const bool up = (MODE == 0 || MODE == 1);
  for (int c = 0; c < NC; ++c) {
    if (up) {
      // gather coarse (constrained coarse DoFs hold 0 in level vectors; read plainly)
      for (int i = threadIdx.x; i < N3; i += blockDim.x) a[i] = (S)src[(size_t)c * cstride_c + (ci[i] & NODE_MASK)];
      __syncthreads();
      // x: [k][j][i] (N,N,N) -> (N,N,M)
      for (int o = threadIdx.x; o < N * N * M; o += blockDim.x) {
        const int m = o % M, r = o / M;
        S s = S(0);
        for (int i = 0; i < N; ++i) s += EP[m * N + i] * a[r * N + i];
        b[o] = s;
      }
      __syncthreads();
      // y: (N,N,M) -> (N,M,M)
      for (int o = threadIdx.x; o < N * M * M; o += blockDim.x) {
        const int mx = o % M, my = (o / M) % M, k = o / (M * M);
        S s = S(0);
        for (int j = 0; j < N; ++j) s += EP[my * N + j] * b[(k * N + j) * M + mx];
        a[o] = s;
      }
      __syncthreads();
      // z: (N,M,M) -> (M,M,M), write
      for (int o = threadIdx.x; o < M3; o += blockDim.x) {
        const int mxy = o % (M * M), mz = o / (M * M);
        S s = S(0);
        for (int k = 0; k < N; ++k) s += EP[mz * N + k] * a[k * M * M + mxy];
        const uint32_t id = pi[o];
        if (!(id & node_flag(c))) {  // level Dirichlet DoFs stay zero
          T *p = dst + (size_t)c * cstride_f + (id & NODE_MASK);
          if (MODE == 0)
            *p = (T)s;  // every parent of a shared node writes the same value
          else
            atomicAdd(p, (T)(s * (S)wfine[id & NODE_MASK]));
        }
      }
      __syncthreads();
    } else {
      for (int o = threadIdx.x; o < M3; o += blockDim.x) {
        const uint32_t id = pi[o];
        S v = (S)src[(size_t)c * cstride_f + (id & NODE_MASK)];
        if (MODE == 2) v = (id & node_flag(c)) ? S(0) : v * (S)wfine[id & NODE_MASK];
        a[o] = v;
      }
      __syncthreads();
      const S *Mx = (MODE == 2) ? EP : ER;  // restriction: P^T ; interpolation: R
      // z: (M,M,M) -> (N,M,M)
      for (int o = threadIdx.x; o < N * M * M; o += blockDim.x) {
        const int mxy = o % (M * M), k = o / (M * M);
        S s = S(0);
        for (int mz = 0; mz < M; ++mz) s += (MODE == 2 ? Mx[mz * N + k] : Mx[k * M + mz]) * a[mz * M * M + mxy];
        b[o] = s;
      }
      __syncthreads();
      // y: (N,M,M) -> (N,N,M)
      for (int o = threadIdx.x; o < N * N * M; o += blockDim.x) {
        const int mx = o % M, j = (o / M) % N, k = o / (M * N);
        S s = S(0);
        for (int my = 0; my < M; ++my) s += (MODE == 2 ? Mx[my * N + j] : Mx[j * M + my]) * b[(k * M + my) * M + mx];
        a[o] = s;
      }
      __syncthreads();
      // x: (N,N,M) -> (N,N,N)
      for (int o = threadIdx.x; o < N3; o += blockDim.x) {
        const int i = o % N, r = o / N;
        S s = S(0);
        for (int mx = 0; mx < M; ++mx) s += (MODE == 2 ? Mx[mx * N + i] : Mx[i * M + mx]) * a[r * M + mx];
        const uint32_t id = ci[o];
        T *p = dst + (size_t)c * cstride_c + (id & NODE_MASK);
        if (MODE == 2) {
          if (!(id & node_flag(c))) atomicAdd(p, (T)s);
        } else {
          *p = (T)s;
        }
      }
      __syncthreads();
    }
  }
}

static double lagrange_1d(const double *nodes, int n, int j, double x) {
  double v = 1.0;
  for (int m = 0; m < n; ++m)
    if (m != j) v *= (x - nodes[m]) / (nodes[j] - nodes[m]);
  return v;
}
}  // namespace b200mf

struct b200mf_mg_transfer {
  b200mf_ctx *coarse, *fine;
  uint32_t *cidx = nullptr, *pidx = nullptr;
  float *wfine = nullptr;
  int64_t n_coarse_cells = 0;  // coarse cells with children on this rank (all of them unless the coarse level is replicated)
  // replicated coarse level: the coarse context holds the WHOLE coarse mesh on every rank (no partitioner), the fine one is
  // partitioned.  Each rank transfers between its fine cells and their parents; restriction sums the partial coarse vectors
  // over the ranks, interpolation merges them - every rank then continues with an identical copy of the coarse vector.
  bool replicated = false;
  void *ctmp[2] = {nullptr, nullptr};
  Embed E;
};

extern "C" int b200mf_mg_transfer_create(b200mf_mg_transfer **out, b200mf_ctx *coarse, b200mf_ctx *fine, const int64_t *children) {
  if (!out || !coarse || !fine || !children) return set_error(B200MF_ERR_ARG, "null argument");
  if (coarse->N != fine->N || coarse->NC != fine->NC || coarse->dtype != fine->dtype)
    return set_error(B200MF_ERR_ARG, "MG transfer: levels must share degree, components and number type");
  const bool replicated = !has_neighbors(coarse) && has_neighbors(fine);
  // the coarse cells that have their children on this rank (replicated coarse level: a subset; children = -1 elsewhere)
  std::vector<int64_t> present;
  for (int64_t cell = 0; cell < coarse->n_cells; ++cell) {
    int n_here = 0;
    for (int ch = 0; ch < 8; ++ch) n_here += children[cell * 8 + ch] >= 0;
    if (n_here == 8)
      present.push_back(cell);
    else if (n_here != 0 || !replicated)
      return set_error(B200MF_ERR_ARG, "MG transfer: coarse cell %lld has %d of its 8 children on this rank", (long long)cell, n_here);
  }
  if (fine->n_cells != 8 * (int64_t)present.size())
    return set_error(B200MF_ERR_ARG, "MG transfer: fine level must have 8 children per coarse cell");
  std::unique_ptr<b200mf_mg_transfer, int (*)(b200mf_mg_transfer *)> t(new b200mf_mg_transfer(), b200mf_mg_transfer_destroy);
  t->coarse = coarse;
  t->fine = fine;
  t->replicated = replicated;
  const int N = coarse->N, p = N - 1, M = 2 * N - 1, N3 = N * N * N, M3 = M * M * M;
  t->n_coarse_cells = (int64_t)present.size();
  // 1-D embedding: fine patch node m sits at x_m = nodes[m]/2 (child 0), 1/2 + nodes[m-p]/2 (child 1)
  const double *nd = coarse->shape.nodes;
  memset(&t->E, 0, sizeof(Embed));
  std::vector<double> xm(M);
  for (int m = 0; m < M; ++m) xm[m] = m <= p ? 0.5 * nd[m] : 0.5 + 0.5 * nd[m - p];
  for (int m = 0; m < M; ++m)
    for (int j = 0; j < N; ++j) t->E.Pf[m * N + j] = (float)(t->E.Pd[m * N + j] = lagrange_1d(nd, N, j, xm[m]));
  for (int i = 0; i < N; ++i) {  // coarse node x_i lies in child ch; fine basis functions of that child evaluated there
    const int ch = (nd[i] <= 0.5) ? 0 : 1;
    const double xi = 2.0 * nd[i] - ch;  // child-local coordinate
    for (int l = 0; l < N; ++l) t->E.Rf[i * M + ch * p + l] = (float)(t->E.Rd[i * M + ch * p + l] = lagrange_1d(nd, N, l, xi));
  }
  // patch index table: fine node of patch position (mx,my,mz) through the child that contains it
  const int64_t n_here = (int64_t)present.size();
  std::vector<uint32_t> pidx((size_t)std::max<int64_t>(1, n_here) * M3), cidx((size_t)std::max<int64_t>(1, n_here) * N3);
  const std::vector<uint32_t> &fn = *fine->h_cell_nodes, &cn = *coarse->h_cell_nodes;
  for (int64_t k = 0; k < n_here; ++k) {
    const int64_t cell = present[k];
    memcpy(&cidx[(size_t)k * N3], &cn[(size_t)cell * N3], (size_t)N3 * sizeof(uint32_t));
    for (int mz = 0; mz < M; ++mz)
      for (int my = 0; my < M; ++my)
        for (int mx = 0; mx < M; ++mx) {
          const int cx = mx > p, cy = my > p, cz = mz > p;
          const int64_t child = children[cell * 8 + cx + 2 * cy + 4 * cz];
          if (child >= fine->n_cells) return set_error(B200MF_ERR_ARG, "MG transfer: child index out of range");
          const int lx = mx - cx * p, ly = my - cy * p, lz = mz - cz * p;
          pidx[(size_t)k * M3 + (mz * M + my) * M + mx] = fn[(size_t)child * N3 + (lz * N + ly) * N + lx];
        }
  }
  // consistency: shared patch positions must map to the same fine node through either child
  for (int64_t k = 0; k < std::min<int64_t>(n_here, 4); ++k) {
    const int64_t c0 = children[present[k] * 8], c1 = children[present[k] * 8 + 1];
    if ((fn[(size_t)c0 * N3 + p] & NODE_MASK) != (fn[(size_t)c1 * N3] & NODE_MASK))
      return set_error(B200MF_ERR_ARG, "MG transfer: children are not in lexicographic (x fastest) order");
  }
  // multiplicity weights of the fine nodes (number of parent patches containing the node; ghosts via compress/update)
  std::vector<float> cnt(fine->n_int, 0.f);
  for (size_t i = 0; i < pidx.size(); ++i) cnt[pidx[i] & NODE_MASK] += 1.f;
  B2_CUDA(cudaMalloc(&t->cidx, cidx.size() * 4));
  B2_CUDA(cudaMemcpy(t->cidx, cidx.data(), cidx.size() * 4, cudaMemcpyHostToDevice));
  if (replicated)
    for (void *&q : t->ctmp) B2_TRY(vec_alloc(coarse, &q));
  B2_CUDA(cudaMalloc(&t->pidx, pidx.size() * 4));
  B2_CUDA(cudaMemcpy(t->pidx, pidx.data(), pidx.size() * 4, cudaMemcpyHostToDevice));
  if (fine->part) {
    // sum the counts of all ranks through the vector partitioner (component-0 block carries the node counts)
    void *tmp = nullptr;
    B2_TRY(vec_alloc(fine, &tmp));
    if (fine->dtype == B200MF_F64) {
      std::vector<double> h(fine->n_int, 0.0);
      for (int64_t i = 0; i < fine->n_nodes; ++i) h[i] = cnt[i];
      B2_CUDA(cudaMemcpy(tmp, h.data(), h.size() * 8, cudaMemcpyHostToDevice));
      B2_TRY(compress_add(fine, tmp));
      B2_TRY(update_ghost_values(fine, tmp));
      B2_CUDA(cudaStreamSynchronize(fine->stream));
      B2_CUDA(cudaMemcpy(h.data(), tmp, h.size() * 8, cudaMemcpyDeviceToHost));
      for (int64_t i = 0; i < fine->n_nodes; ++i) cnt[i] = (float)h[i];
    } else {
      std::vector<float> h(fine->n_int, 0.f);
      for (int64_t i = 0; i < fine->n_nodes; ++i) h[i] = cnt[i];
      B2_CUDA(cudaMemcpy(tmp, h.data(), h.size() * 4, cudaMemcpyHostToDevice));
      B2_TRY(compress_add(fine, tmp));
      B2_TRY(update_ghost_values(fine, tmp));
      B2_CUDA(cudaStreamSynchronize(fine->stream));
      B2_CUDA(cudaMemcpy(h.data(), tmp, h.size() * 4, cudaMemcpyDeviceToHost));
      for (int64_t i = 0; i < fine->n_nodes; ++i) cnt[i] = h[i];
    }
    cudaFree(tmp);
  }
  std::vector<float> w(fine->n_nodes);
  for (int64_t i = 0; i < fine->n_nodes; ++i) w[i] = cnt[i] > 0.f ? 1.f / cnt[i] : 0.f;
  B2_CUDA(cudaMalloc(&t->wfine, std::max<size_t>(4, w.size() * 4)));
  B2_CUDA(cudaMemcpy(t->wfine, w.data(), w.size() * 4, cudaMemcpyHostToDevice));
  *out = t.release();
  return 0;
}
extern "C" int b200mf_mg_transfer_destroy(b200mf_mg_transfer *t) {
  if (!t) return 0;
  cudaFree(t->cidx);
  cudaFree(t->pidx);
  cudaFree(t->wfine);
  for (void *q : t->ctmp) cudaFree(q);
  delete t;
  return 0;
}

template <typename T, int MODE>
static int transfer_launch(b200mf_mg_transfer *t, void *dst, const void *src) {
  const int N = t->coarse->N, M = 2 * N - 1;
  const size_t smem = 2 * (size_t)M * M * M * sizeof(T);
  auto kern = transfer_kernel<T, MODE>;
  static bool attr = false;
  if (!attr) {
    B2_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * MAXM * MAXM * MAXM * (int)sizeof(T)));
    attr = true;
  }
  if (t->n_coarse_cells == 0) return 0;
  kern<<<(unsigned)t->n_coarse_cells, 128, smem, t->fine->stream>>>(N, t->coarse->NC, t->cidx, t->pidx, t->wfine, t->coarse->cstride,
                                                                     t->fine->cstride, (const T *)src, (T *)dst, t->E);
  B2_LAUNCHED();
  return 0;
}
template <int MODE>
static int transfer_dispatch(b200mf_mg_transfer *t, void *dst, const void *src) {
  // the exchanges of the coarse vector run on the coarse context's stream, the kernel on the fine one's: nothing orders two
  // different streams here, so they must be one and the same (b200mf.h: b200mf_ctx_set_stream)
  if (t->coarse->stream != t->fine->stream)
    return set_error(B200MF_ERR_ARG, "MG transfer: the coarse and the fine context must use the same CUDA stream");
  return t->coarse->dtype == B200MF_F64 ? transfer_launch<double, MODE>(t, dst, src) : transfer_launch<float, MODE>(t, dst, src);
}

// prolongate: fine_dst = P coarse_src (coarse ghosts are updated first)
extern "C" int b200mf_mg_transfer_prolongate(b200mf_mg_transfer *t, void *fine_dst, const void *coarse_src) {
  if (!t || !fine_dst || !coarse_src) return set_error(B200MF_ERR_ARG, "null argument");
  B2_CUDA(cudaMemsetAsync(fine_dst, 0, t->fine->n_int * esz(t->fine), t->fine->stream));
  B2_TRY(update_ghost_values(t->coarse, const_cast<void *>(coarse_src)));
  B2_TRY(transfer_dispatch<0>(t, fine_dst, coarse_src));
  return zero_ghosts(t->fine, fine_dst);
}
extern "C" int b200mf_mg_transfer_prolongate_and_add(b200mf_mg_transfer *t, void *fine_dst, const void *coarse_src) {
  if (!t || !fine_dst || !coarse_src) return set_error(B200MF_ERR_ARG, "null argument");
  B2_TRY(update_ghost_values(t->coarse, const_cast<void *>(coarse_src)));
  B2_TRY(zero_ghosts(t->fine, fine_dst));
  B2_TRY(transfer_dispatch<1>(t, fine_dst, coarse_src));
  return compress_add(t->fine, fine_dst);
}
// restrict_and_add: coarse_dst += P^T fine_src
extern "C" int b200mf_mg_transfer_restrict_and_add(b200mf_mg_transfer *t, void *coarse_dst, const void *fine_src) {
  if (!t || !coarse_dst || !fine_src) return set_error(B200MF_ERR_ARG, "null argument");
  B2_TRY(update_ghost_values(t->fine, const_cast<void *>(fine_src)));
  if (t->replicated) {  // this rank's share of P^T fine_src, summed over the ranks, then added to every rank's copy
    b200mf_ctx *cc = t->coarse;
    B2_CUDA(cudaMemsetAsync(t->ctmp[0], 0, cc->n_int * esz(cc), cc->stream));
    B2_TRY(transfer_dispatch<2>(t, t->ctmp[0], fine_src));
    B2_TRY(allreduce_vector(t->fine, t->ctmp[0], cc->n_int, cc->dtype, 0, cc->stream));
    if (cc->dtype == B200MF_F64)
      axpby_kernel<double><<<vgrid(cc->n_int, cc), 256, 0, cc->stream>>>(cc->n_int, 1.0, (const double *)t->ctmp[0], 1.0, (double *)coarse_dst);
    else
      axpby_kernel<float><<<vgrid(cc->n_int, cc), 256, 0, cc->stream>>>(cc->n_int, 1.f, (const float *)t->ctmp[0], 1.f, (float *)coarse_dst);
    B2_LAUNCHED();
    return 0;
  }
  B2_TRY(zero_ghosts(t->coarse, coarse_dst));
  B2_TRY(transfer_dispatch<2>(t, coarse_dst, fine_src));
  return compress_add(t->coarse, coarse_dst);
}
// interpolate_to_mg step (mf_elasticity.h:920-922): coarse_dst = values of the fine function at the coarse support points
extern "C" int b200mf_mg_transfer_interpolate(b200mf_mg_transfer *t, void *coarse_dst, const void *fine_src) {
  if (!t || !coarse_dst || !fine_src) return set_error(B200MF_ERR_ARG, "null argument");
  B2_TRY(update_ghost_values(t->fine, const_cast<void *>(fine_src)));
  if (t->replicated) {
    // every rank evaluates the coarse support points inside its own parents; points shared by parents of two ranks get the
    // same value from both, all others are 0 elsewhere: max + min over the ranks merges the pieces without double counting
    b200mf_ctx *cc = t->coarse;
    const size_t bytes = cc->n_int * esz(cc);
    B2_CUDA(cudaMemsetAsync(t->ctmp[0], 0, bytes, cc->stream));
    B2_TRY(transfer_dispatch<3>(t, t->ctmp[0], fine_src));
    B2_CUDA(cudaMemcpyAsync(t->ctmp[1], t->ctmp[0], bytes, cudaMemcpyDeviceToDevice, cc->stream));
    B2_TRY(allreduce_vector(t->fine, t->ctmp[0], cc->n_int, cc->dtype, 1, cc->stream));
    B2_TRY(allreduce_vector(t->fine, t->ctmp[1], cc->n_int, cc->dtype, 2, cc->stream));
    B2_CUDA(cudaMemcpyAsync(coarse_dst, t->ctmp[0], bytes, cudaMemcpyDeviceToDevice, cc->stream));
    if (cc->dtype == B200MF_F64)
      axpby_kernel<double><<<vgrid(cc->n_int, cc), 256, 0, cc->stream>>>(cc->n_int, 1.0, (const double *)t->ctmp[1], 1.0, (double *)coarse_dst);
    else
      axpby_kernel<float><<<vgrid(cc->n_int, cc), 256, 0, cc->stream>>>(cc->n_int, 1.f, (const float *)t->ctmp[1], 1.f, (float *)coarse_dst);
    B2_LAUNCHED();
    return 0;
  }
  B2_TRY(transfer_dispatch<3>(t, coarse_dst, fine_src));
  return zero_ghosts(t->coarse, coarse_dst);
}

// ===========================================================================
// Multigrid (V-cycle) + PreconditionMG (A.7)
struct b200mf_mg {
  int n_levels;
  std::vector<b200mf_chebyshev *> smoother;   // [level]; [0] is the coarse solver (MGCoarseGridApplySmoother)
  std::vector<b200mf_mg_transfer *> transfer; // [level-1]: level-1 <-> level
  std::vector<void *> defect, solution, t;
  b200mf_ctx *outer;  // context of the (FP64) system vectors handed to vmult
  // coarse grid: 0 = the level-0 smoother as solver (MGCoarseGridApplySmoother), 1 = SolverCG under a ReductionControl,
  // preconditioned with the level-0 smoother (MGCoarseGridIterativeSolver, mf_elasticity.h:1183-1198)
  int coarse_kind = 0, coarse_max_it = 100, coarse_last_iterations = 0;
  double coarse_abs_tol = 1e-10, coarse_reduction = 1e-3;
  void *coarse_work[3] = {nullptr, nullptr, nullptr};
  Scratch coarse_sc;
  // Levels 0..graph_top of the V-cycle as one CUDA graph: they are launch-bound (hundreds of small kernels per cycle), the
  // vectors they touch belong to this object and the host-side numbers that enter their launches (Chebyshev theta/delta/
  // degree, operator constants) change only when the smoothers are re-estimated - the sequence is recorded once per such
  // change (stamp) on a private stream and replayed on the contexts' stream by every further V-cycle.
  int graph_top = -1;       // -1: off
  int graph_state = 0;      // 0 = nothing recorded yet, 1 = replaying, -1 = recording failed (eager launches from then on)
  bool graph_warm = false;  // the first cycle runs eagerly (lazy allocations of the operators and exchanges)
  uint64_t graph_stamp = 0;
  int64_t graph_launches = 0, graph_records = 0, graph_replays = 0;
  cudaGraphExec_t graph_exec = nullptr;
  cudaStream_t graph_stream = nullptr;
  ~b200mf_mg() {  // also runs on the error paths of b200mf_mg_create
    for (void *p : defect) cudaFree(p);
    for (void *p : solution) cudaFree(p);
    for (void *p : t) cudaFree(p);
    for (void *p : coarse_work) cudaFree(p);
    if (graph_exec) cudaGraphExecDestroy(graph_exec);
    if (graph_stream) cudaStreamDestroy(graph_stream);
  }
};

extern "C" int b200mf_mg_create(b200mf_mg **out, int n_levels, b200mf_chebyshev *const *smoothers, b200mf_mg_transfer *const *transfers,
                                b200mf_ctx *outer) {
  if (!out || n_levels < 1 || !smoothers || (n_levels > 1 && !transfers) || !outer) return set_error(B200MF_ERR_ARG, "bad argument");
  std::unique_ptr<b200mf_mg> mg(new b200mf_mg());
  mg->n_levels = n_levels;
  mg->outer = outer;
  for (int l = 0; l < n_levels; ++l) {
    if (!smoothers[l]) return set_error(B200MF_ERR_ARG, "null smoother");
    mg->smoother.push_back(smoothers[l]);
    b200mf_ctx *c = smoothers[l]->A.ctx;
    if (l > 0) {
      if (!transfers[l - 1] || transfers[l - 1]->fine != c || transfers[l - 1]->coarse != smoothers[l - 1]->A.ctx)
        return set_error(B200MF_ERR_ARG, "MG: transfer %d does not connect levels %d and %d", l - 1, l - 1, l);
      mg->transfer.push_back(transfers[l - 1]);
    }
    for (std::vector<void *> *vs : {&mg->defect, &mg->solution, &mg->t}) {
      vs->push_back(nullptr);  // owned by mg from here on
      B2_TRY(vec_alloc(c, &vs->back()));
    }
  }
  if (smoothers[n_levels - 1]->A.ctx->n_int != outer->n_int)
    return set_error(B200MF_ERR_ARG, "MG: finest level and system context have different vector layouts");
  const char *ge = getenv("B200MF_MG_GRAPH");  // 0: eager launches on every level
  mg->graph_top = (ge && atoi(ge) == 0) ? -1 : n_levels - 2;
  *out = mg.release();
  return 0;
}
extern "C" int b200mf_mg_destroy(b200mf_mg *mg) {
  delete mg;
  return 0;
}

extern "C" int b200mf_mg_set_coarse_solver(b200mf_mg *mg, int kind, int max_iterations, double abs_tolerance, double reduction) {
  if (!mg || kind < 0 || kind > 1) return set_error(B200MF_ERR_ARG, "bad argument");
  mg->coarse_kind = kind;
  if (kind == 1) {
    mg->coarse_max_it = max_iterations;
    mg->coarse_abs_tol = abs_tolerance;
    mg->coarse_reduction = reduction;
    for (void *&p : mg->coarse_work)
      if (!p) B2_TRY(vec_alloc(mg->smoother[0]->A.ctx, &p));
  }
  return 0;
}
extern "C" int b200mf_mg_coarse_iterations(const b200mf_mg *mg) { return mg ? mg->coarse_last_iterations : -1; }

extern "C" int b200mf_mg_set_graph_levels(b200mf_mg *mg, int top_level) {
  if (!mg || top_level >= mg->n_levels) return set_error(B200MF_ERR_ARG, "bad argument");
  mg->graph_top = top_level < 0 ? -1 : top_level;
  mg->graph_stamp = 0;
  if (mg->graph_state > 0) mg->graph_state = 0;
  return 0;
}
extern "C" int b200mf_mg_graph_info(const b200mf_mg *mg, int *top_level, int *state, int64_t *records, int64_t *replays, int64_t *launches) {
  if (!mg) return set_error(B200MF_ERR_ARG, "null argument");
  if (top_level) *top_level = mg->graph_top;
  if (state) *state = mg->graph_state;
  if (records) *records = mg->graph_records;
  if (replays) *replays = mg->graph_replays;
  if (launches) *launches = mg->graph_launches;
  return 0;
}

template <typename T>
static int mg_level_v_step(b200mf_mg *mg, int level);

static uint64_t stamp_mix(uint64_t h, uint64_t v) { return h ^ (v + 0x9e3779b97f4a7c15ull + (h << 6) + (h >> 2)); }
static uint64_t stamp_bits(double x) {
  uint64_t u;
  memcpy(&u, &x, sizeof u);
  return u;
}

// levels 0..level of the V-cycle (defect[level] -> solution[level]), replayed from a CUDA graph where possible
template <typename T>
static int mg_lower_levels(b200mf_mg *mg, int level) {
  bool usable = level == mg->graph_top && mg->graph_state >= 0 && mg->coarse_kind == 0;
  for (int l = 0; usable && l <= level; ++l) {
    b200mf_chebyshev *sm = mg->smoother[l];
    usable = sm->eig_ready && !sm->A.ctx->kernel_events;  // estimates synchronise; timed launches record events the graph would own
  }
  if (!usable || !mg->graph_warm) {
    if (level == mg->graph_top) mg->graph_warm = true;
    return mg_level_v_step<T>(mg, level);
  }
  uint64_t stamp = stamp_mix(g_config_epoch.load(), (uint64_t)level);
  for (int l = 0; l <= level; ++l) {
    b200mf_chebyshev *sm = mg->smoother[l];
    const void *inv = nullptr;
    B2_TRY(sm->A.inv_diag(&inv));
    stamp = stamp_mix(stamp, stamp_bits(sm->theta));
    stamp = stamp_mix(stamp, stamp_bits(sm->delta));
    stamp = stamp_mix(stamp, (uint64_t)sm->degree);
    stamp = stamp_mix(stamp, (uint64_t)(uintptr_t)inv);
  }
  stamp |= 1;
  cudaStream_t run = mg->smoother[level]->A.ctx->stream;
  if (mg->graph_exec && mg->graph_stamp == stamp) {
    B2_CUDA(cudaGraphLaunch(mg->graph_exec, run));
    g_launches.fetch_add(mg->graph_launches);
    ++mg->graph_replays;
    return 0;
  }
  // record: the levels' launches go to a private stream (the contexts' stream may be the legacy default stream, which cannot
  // be captured); the graph is then launched on the contexts' stream, which orders it with the finer levels
  if (mg->graph_exec) {
    cudaGraphExecDestroy(mg->graph_exec);
    mg->graph_exec = nullptr;
  }
  if (!mg->graph_stream) B2_CUDA(cudaStreamCreateWithFlags(&mg->graph_stream, cudaStreamNonBlocking));
  const int64_t l0 = g_launches.load();
  for (int l = 0; l <= level; ++l) mg->smoother[l]->A.ctx->stream = mg->graph_stream;
  cudaError_t e = cudaStreamBeginCapture(mg->graph_stream, cudaStreamCaptureModeThreadLocal);
  int rc = 0;
  cudaGraph_t graph = nullptr;
  if (e == cudaSuccess) {
    rc = mg_level_v_step<T>(mg, level);
    e = cudaStreamEndCapture(mg->graph_stream, &graph);
  }
  for (int l = 0; l <= level; ++l) mg->smoother[l]->A.ctx->stream = run;
  const int64_t recorded = g_launches.load() - l0;
  if (rc == 0 && e == cudaSuccess) e = cudaGraphInstantiate(&mg->graph_exec, graph, 0);
  if (graph) cudaGraphDestroy(graph);
  if (rc != 0 || e != cudaSuccess) {  // not fatal: these levels run as plain launches from now on (b200mf_mg_graph_info: state -1)
    fprintf(stderr, "b200mf: multigrid levels 0..%d could not be recorded as a CUDA graph (%s); launching them one by one\n", level,
            rc != 0 ? b200mf_last_error() : cudaGetErrorString(e));
    cudaGetLastError();
    mg->graph_exec = nullptr;
    mg->graph_state = -1;
    g_launches.fetch_add(-recorded);
    return mg_level_v_step<T>(mg, level);
  }
  mg->graph_stamp = stamp;
  mg->graph_launches = recorded;
  mg->graph_state = 1;
  ++mg->graph_records;
  B2_CUDA(cudaGraphLaunch(mg->graph_exec, run));  // recording did not execute anything
  return 0;
}

template <typename T>
static int mg_level_v_step(b200mf_mg *mg, int level) {
  b200mf_chebyshev *sm = mg->smoother[level];
  b200mf_ctx *c = sm->A.ctx;
  T *sol = (T *)mg->solution[level], *def = (T *)mg->defect[level], *t = (T *)mg->t[level];
  if (level == 0) {
    if (mg->coarse_kind == 1) {  // SolverCG(ReductionControl) preconditioned with the level-0 smoother, start vector 0
      CGPrecond P;
      P.kind = 2;
      P.apply = [sm](void *d, const void *s) { return cheb_apply_t<T>(sm, (T *)d, (const T *)s, true); };
      CGResult r;
      B2_CUDA(cudaMemsetAsync(sol, 0, c->n_int * sizeof(T), c->stream));
      B2_TRY(cg_solve_t<T>(sm->A, P, sol, def, mg->coarse_abs_tol, mg->coarse_max_it, false, true, mg->coarse_work, mg->coarse_sc, r,
                           mg->coarse_reduction));
      mg->coarse_last_iterations = r.iterations;
      if (!r.converged) return set_error(B200MF_ERR_NUMERIC, "MG coarse solver did not converge in %d iterations", r.iterations);
      return 0;
    }
    return cheb_apply_t<T>(sm, sol, def, true);  // coarse: Chebyshev as solver
  }
  B2_TRY(cheb_apply_t<T>(sm, sol, def, true));                  // pre-smoothing (zero start)
  B2_TRY(sm->A.vmult(t, sol));
  axpby_kernel<T><<<vgrid(c->n_int, c), 256, 0, c->stream>>>(c->n_int, T(1), def, T(-1), t);  // t = defect - A sol
  B2_LAUNCHED();
  b200mf_ctx *cc = mg->smoother[level - 1]->A.ctx;
  B2_CUDA(cudaMemsetAsync(mg->defect[level - 1], 0, cc->n_int * sizeof(T), cc->stream));
  B2_TRY(b200mf_mg_transfer_restrict_and_add(mg->transfer[level - 1], mg->defect[level - 1], t));
  B2_TRY(mg_lower_levels<T>(mg, level - 1));
  B2_TRY(b200mf_mg_transfer_prolongate_and_add(mg->transfer[level - 1], sol, mg->solution[level - 1]));
  return cheb_apply_t<T>(sm, sol, def, false);  // post-smoothing (step)
}

// PreconditionMG::vmult: copy_to_mg, V-cycle, copy_from_mg
extern "C" int b200mf_mg_vmult(b200mf_mg *mg, void *dst, const void *src) {
  if (!mg || !dst || !src) return set_error(B200MF_ERR_ARG, "null argument");
  const int top = mg->n_levels - 1;
  b200mf_ctx *c = mg->smoother[top]->A.ctx, *o = mg->outer;
  const int64_t n = c->n_int;
  const unsigned g = vgrid(n, c);
  const bool lf = c->dtype == B200MF_F32, of = o->dtype == B200MF_F32;
  // the V-cycle hands vectors from level to level and to the outer context without events: one stream for all of them
  for (int l = 0; l <= top; ++l)
    if (mg->smoother[l]->A.ctx->stream != o->stream)
      return set_error(B200MF_ERR_ARG, "MG: level %d and the system context use different CUDA streams (b200mf_ctx_set_stream)", l);
  if (lf && !of)
    convert_kernel<float, double><<<g, 256, 0, c->stream>>>(n, (const double *)src, (float *)mg->defect[top]);
  else if (!lf && !of)
    convert_kernel<double, double><<<g, 256, 0, c->stream>>>(n, (const double *)src, (double *)mg->defect[top]);
  else if (lf && of)
    convert_kernel<float, float><<<g, 256, 0, c->stream>>>(n, (const float *)src, (float *)mg->defect[top]);
  else
    return set_error(B200MF_ERR_ARG, "MG: FP64 levels under an FP32 system are not supported");
  B2_LAUNCHED();
  B2_TRY(lf ? mg_level_v_step<float>(mg, top) : mg_level_v_step<double>(mg, top));
  if (lf && !of)
    convert_kernel<double, float><<<g, 256, 0, c->stream>>>(n, (const float *)mg->solution[top], (double *)dst);
  else if (!lf && !of)
    convert_kernel<double, double><<<g, 256, 0, c->stream>>>(n, (const double *)mg->solution[top], (double *)dst);
  else
    convert_kernel<float, float><<<g, 256, 0, c->stream>>>(n, (const float *)mg->solution[top], (float *)dst);
  B2_LAUNCHED();
  return 0;
}

// ===========================================================================
// SolverCG on the library's operators
struct b200mf_cg {
  LinOp A;
  void *w[3] = {nullptr, nullptr, nullptr};
  Scratch sc;
  CGResult last;
};
static int cg_create(b200mf_cg **out, LinOp A) {
  if (!out) return set_error(B200MF_ERR_ARG, "null argument");
  std::unique_ptr<b200mf_cg> s(new b200mf_cg());
  s->A = A;
  for (void *&p : s->w) B2_TRY(vec_alloc(A.ctx, &p));
  *out = s.release();
  return 0;
}
extern "C" int b200mf_cg_create_nh(b200mf_cg **out, b200mf_nh *op) {
  if (!op) return set_error(B200MF_ERR_ARG, "null op");
  return cg_create(out, linop_nh(op, b200mf_nh_ctx(op)));
}
extern "C" int b200mf_cg_create_laplace(b200mf_cg **out, b200mf_laplace *op) {
  if (!op) return set_error(B200MF_ERR_ARG, "null op");
  return cg_create(out, linop_laplace(op, b200mf_laplace_ctx(op)));
}
extern "C" int b200mf_cg_create_boundary_mass(b200mf_cg **out, b200mf_boundary_mass *op) {
  if (!op) return set_error(B200MF_ERR_ARG, "null op");
  return cg_create(out, linop_boundary_mass(op));
}
extern "C" int b200mf_cg_destroy(b200mf_cg *s) {
  if (!s) return 0;
  for (void *p : s->w) cudaFree(p);
  delete s;
  return 0;
}
extern "C" int b200mf_cg_solve(b200mf_cg *s, void *x, const void *b, const b200mf_cg_control *ctl, b200mf_cg_result *result) {
  if (!s || !x || !b || !ctl) return set_error(B200MF_ERR_ARG, "null argument");
  CGPrecond P;
  P.kind = 0;
  if (ctl->preconditioner == B200MF_PRECOND_JACOBI) {
    P.kind = 1;
    P.omega = ctl->jacobi_omega;
    B2_TRY(s->A.inv_diag(&P.inv));
  } else if (ctl->preconditioner == B200MF_PRECOND_MG) {
    if (!ctl->mg) return set_error(B200MF_ERR_ARG, "CG: multigrid preconditioner is null");
    P.kind = 2;
    b200mf_mg *mg = ctl->mg;
    P.apply = [mg](void *d, const void *r) { return b200mf_mg_vmult(mg, d, r); };
  } else if (ctl->preconditioner != B200MF_PRECOND_NONE) {
    return set_error(B200MF_ERR_ARG, "CG: unknown preconditioner");
  }
  s->last = CGResult();
  b200mf_ctx *c = s->A.ctx;
  int rc = c->dtype == B200MF_F64
               ? cg_solve_t<double>(s->A, P, (double *)x, (const double *)b, ctl->tolerance, ctl->max_iterations, false, ctl->x_is_zero != 0, s->w, s->sc, s->last)
               : cg_solve_t<float>(s->A, P, (float *)x, (const float *)b, ctl->tolerance, ctl->max_iterations, false, ctl->x_is_zero != 0, s->w, s->sc, s->last);
  B2_TRY(rc);
  if (result) {
    result->iterations = s->last.iterations;
    result->residual = s->last.residual;
    result->initial_residual = s->last.initial_residual;
    result->converged = s->last.converged ? 1 : 0;
    tridiag_extreme(s->last.diag, s->last.offdiag, result->min_eigenvalue, result->max_eigenvalue);
  }
  if (!s->last.converged) return set_error(B200MF_ERR_NUMERIC, "SolverCG: no convergence after %d iterations (residual %g)", s->last.iterations, s->last.residual);
  return 0;
}

// BoundaryProjector::project (include/adapter/boundary_projector.h:86-110): L2 projection of a nodal residual onto the
// interface - M_Gamma x = residual|Gamma by CG under ReductionControl(500, 0, 1e-12) with PreconditionJacobi(0.8), started
// from x = residual (entries off the interface have a zero row, a zero right-hand side and a zero preconditioner: they keep
// their value); the operator's compute_diagonal() runs inside, as in the reference; ghost values updated at the end.
extern "C" int b200mf_boundary_mass_project(b200mf_boundary_mass *op, void *residual, int *iterations, double *final_residual) {
  if (!op || !residual) return set_error(B200MF_ERR_ARG, "null argument");
  LinOp A = linop_boundary_mass(op);
  b200mf_ctx *c = A.ctx;
  B2_TRY(b200mf_boundary_mass_compute_diagonal(op));
  void *w[4] = {nullptr, nullptr, nullptr, nullptr};  // rhs + the three CG vectors
  struct Free {
    void **w;
    ~Free() {
      for (int i = 0; i < 4; ++i) cudaFree(w[i]);
    }
  } guard{w};
  for (void *&p : w) B2_TRY(vec_alloc(c, &p));
  B2_TRY(b200mf_boundary_mass_restrict(op, w[0], residual));
  CGPrecond P;
  P.kind = 1;
  P.omega = 0.8;
  B2_TRY(A.inv_diag(&P.inv));
  Scratch sc;
  CGResult r;
  int rc = c->dtype == B200MF_F64 ? cg_solve_t<double>(A, P, (double *)residual, (const double *)w[0], 0.0, 500, false, false, w + 1, sc, r, 1e-12)
                                  : cg_solve_t<float>(A, P, (float *)residual, (const float *)w[0], 0.0, 500, false, false, w + 1, sc, r, 1e-12);
  B2_TRY(rc);
  if (iterations) *iterations = r.iterations;
  if (final_residual) *final_residual = r.residual;
  if (!r.converged) return set_error(B200MF_ERR_NUMERIC, "BoundaryProjector: no convergence after %d iterations (residual %g)", r.iterations, r.residual);
  return update_ghost_values(c, residual);
}

// ===========================================================================
// vector algebra on device vectors (LinearAlgebra::distributed::Vector: l2_norm, operator*, sadd, =)
extern "C" int b200mf_vec_dot(b200mf_ctx *c, const void *a, const void *b, double *out) {
  if (!c || !a || !b || !out) return set_error(B200MF_ERR_ARG, "null argument");
  static thread_local Scratch sc;
  B2_TRY(sc.init());
  B2_CUDA(cudaMemsetAsync(sc.d, 0, sizeof(double), c->stream));
  B2_TRY(c->dtype == B200MF_F64 ? dot_t<double>(c, a, b, sc.d) : dot_t<float>(c, a, b, sc.d));
  B2_TRY(fetch(c, sc, 1));
  *out = sc.h[0];
  return 0;
}
extern "C" int b200mf_vec_l2_norm(b200mf_ctx *c, const void *a, double *out) {
  B2_TRY(b200mf_vec_dot(c, a, a, out));
  *out = std::sqrt(*out);
  return 0;
}
// y = s*y + a*x (Vector::sadd)
extern "C" int b200mf_vec_sadd(b200mf_ctx *c, void *y, double b, double a, const void *x) {
  if (!c || !y || !x) return set_error(B200MF_ERR_ARG, "null argument");
  if (c->dtype == B200MF_F64)
    axpby_kernel<double><<<vgrid(c->n_int, c), 256, 0, c->stream>>>(c->n_int, a, (const double *)x, b, (double *)y);
  else
    axpby_kernel<float><<<vgrid(c->n_int, c), 256, 0, c->stream>>>(c->n_int, (float)a, (const float *)x, (float)b, (float *)y);
  B2_LAUNCHED();
  return 0;
}
extern "C" int b200mf_vec_set_zero(b200mf_ctx *c, void *y) {
  if (!c || !y) return set_error(B200MF_ERR_ARG, "null argument");
  B2_CUDA(cudaMemsetAsync(y, 0, c->n_int * esz(c), c->stream));
  return 0;
}
// copy between two contexts over the same mesh/numbering (FP64 system <-> FP32 level: copy_to_mg / copy_from_mg / level displacement)
extern "C" int b200mf_vec_convert(b200mf_ctx *dst_ctx, void *dst, b200mf_ctx *src_ctx, const void *src) {
  if (!dst_ctx || !src_ctx || !dst || !src) return set_error(B200MF_ERR_ARG, "null argument");
  if (dst_ctx->n_int != src_ctx->n_int) return set_error(B200MF_ERR_ARG, "convert: contexts have different vector layouts");
  const int64_t n = dst_ctx->n_int;
  const unsigned g = vgrid(n, dst_ctx);
  cudaStream_t st = dst_ctx->stream;
  const bool df = dst_ctx->dtype == B200MF_F32, sf = src_ctx->dtype == B200MF_F32;
  if (df && sf) convert_kernel<float, float><<<g, 256, 0, st>>>(n, (const float *)src, (float *)dst);
  else if (df) convert_kernel<float, double><<<g, 256, 0, st>>>(n, (const double *)src, (float *)dst);
  else if (sf) convert_kernel<double, float><<<g, 256, 0, st>>>(n, (const float *)src, (double *)dst);
  else convert_kernel<double, double><<<g, 256, 0, st>>>(n, (const double *)src, (double *)dst);
  B2_LAUNCHED();
  return 0;
}
