/* b200mf — B200-native (sm_100a) matrix-free operator application.
 *
 * C ABI of the drop-in boundary.  Every entry point names the
 * reference interface it replaces; paths are relative to the reference tree
 * (davidscn/matrix-free-dealii-precice).  The reference-side binding (a C++
 * shim class with the duck-typed deal.II operator surface) is shown in
 * INTEGRATION.md.
 *
 * Conventions
 *  - plain pointers and sizes only; no C++/torch types cross this boundary;
 *  - every function returns 0 on success, a negative b200mf_status otherwise;
 *    b200mf_last_error() returns a thread-local message (the shim turns a
 *    non-zero status into deal.II's AssertThrow, src/solid.cc:64-89);
 *  - "dvec" arguments are DEVICE pointers to vectors in the library's internal
 *    numbering (see b200mf_vec_*); "host" arguments are caller-owned host arrays
 *    in the caller's (deal.II) numbering, local size n_dofs = owned + ghosts;
 *  - one caller thread per context (the reference runs one thread per rank,
 *    src/solid.cc:30-31); all work of a context is ordered on its stream.
 */
#ifndef B200MF_H
#define B200MF_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum {
  B200MF_OK = 0,
  B200MF_ERR_ARG = -1,     /* invalid argument / unsupported configuration */
  B200MF_ERR_CUDA = -2,    /* CUDA runtime error */
  B200MF_ERR_STATE = -3,   /* call order violated (e.g. vmult before cache) */
  B200MF_ERR_NUMERIC = -4, /* det F <= 0, non-positive diagonal */
  B200MF_ERR_NCCL = -5
} b200mf_status;

typedef enum { B200MF_F64 = 0, B200MF_F32 = 1 } b200mf_dtype;

/* MF caching levels, include/solid_mechanics/mf_nh_operator.h:215-223 and the
 * "MF caching" .prm key (include/parameter/parameter_handling.h:177-254). */
typedef enum {
  B200MF_CACHE_SCALAR_REFERENTIAL = 0,
  B200MF_CACHE_TENSOR2 = 1,
  B200MF_CACHE_TENSOR4 = 2,
  B200MF_CACHE_TENSOR4_NS = 3
} b200mf_caching;

typedef struct b200mf_ctx b200mf_ctx;         /* replaces MatrixFree<3,Number> (+DoFInfo, MappingInfo) */
typedef struct b200mf_nh b200mf_nh;           /* replaces NeoHookOperator<3,Number>, mf_nh_operator.h:99-237 */
typedef struct b200mf_laplace b200mf_laplace; /* replaces LaplaceOperator<3,number>, heat_transfer.h:81-127 */

const char *b200mf_last_error(void);
/* number of CUDA kernels this library has launched in this process (bench.py "gpu_launches") */
int64_t b200mf_launch_count(void);
int b200mf_version(void);

/* ------------------------------------------------------------------------
 * Cell-batch data ("MatrixFree").  Replaces MatrixFree::reinit as called by
 * Solid::reinit_matrix_free / reinit_multi_grid_matrix_free
 * (include/solid_mechanics/mf_elasticity.h:933-1055) and
 * LaplaceProblem::setup_system (include/heat_transfer/heat_transfer.h:430-509).
 * All host arrays are copied; they may be freed after the call.
 * ------------------------------------------------------------------------ */
typedef struct {
  int degree;        /* FE_Q degree p, 1..6 (README.md:60; include/base/fe_integrator.h:10-22) */
  int n_q_points_1d; /* must be degree+1 (every shipped .prm uses Quadrature order = p+1) */
  int n_components;  /* 3 = FESystem(FE_Q(p),3) solid; 1 = scalar heat */
  int dtype;         /* b200mf_dtype: FP64 system operator or FP32 level operator (mf_elasticity.h:363-367) */
  int64_t n_cells;
  int64_t n_dofs;  /* local vector size: owned + ghost entries (LA::distributed::Vector) */
  int64_t n_owned; /* owned entries come first; == n_dofs on one rank */

  /* Exactly one of dof_indices / node_indices.
   * dof_indices: [n_cells][n_components*(p+1)^3] local DoF indices in FEEvaluation order
   *   (component-major, lexicographic within a component, x fastest) — what
   *   FEEvaluation::read_dof_values walks (mf_nh_operator.h:640-646).
   * node_indices: [n_cells][(p+1)^3] scalar node indices; DoF = n_components*node + c. */
  const uint32_t *dof_indices;
  const uint32_t *node_indices;

  /* MatrixFree::get_constrained_dofs(): local indices (mf_nh_operator.h:623-624, 1043-1044) */
  const uint32_t *constrained_dofs;
  int64_t n_constrained;

  /* Reference-configuration mapping (StaticMappingQ1, mf_elasticity.h:739-746).
   * Either the 8 vertices per cell (lexicographic, x fastest): [n_cells][8][3] — the
   * library evaluates the Q1 metric on the device — or explicit MappingInfo arrays
   * inverse_jacobian [n_cells][nq^3][3][3] with [e][f] = d xi_f / d x_e
   * (mf_nh_operator.h:872-887) and jxw [n_cells][nq^3]. */
  const double *vertices;
  const double *inverse_jacobian;
  const double *jxw;

  const uint8_t *material_id; /* [n_cells] or NULL (all 0); mf_nh_operator.h:402-404, 656-658 */
} b200mf_ctx_desc;

int b200mf_ctx_create(b200mf_ctx **ctx, const b200mf_ctx_desc *desc);
int b200mf_ctx_destroy(b200mf_ctx *ctx);
/* all work of this context (and of operators built on it) is enqueued on `cuda_stream` (cudaStream_t; NULL = default) */
int b200mf_ctx_set_stream(b200mf_ctx *ctx, void *cuda_stream);
int b200mf_ctx_synchronize(b200mf_ctx *ctx);
/* measurement aid: with on != 0 every cell-loop launch is bracketed by CUDA events on the launching stream;
 * b200mf_ctx_kernel_time synchronises and returns their summed duration and the number of launches since `on` */
int b200mf_ctx_time_kernels(b200mf_ctx *ctx, int on);
int b200mf_ctx_kernel_time(b200mf_ctx *ctx, double *total_ms, int64_t *n_launches);
/* multi-rank contexts: 1 hides update_ghost_values / compress(add) of vmult behind the interior cells on a second stream,
 * 0 (default; env B200MF_OVERLAP) runs exchange -> cells -> exchange on one stream as the reference does (mf_nh_operator.h:598-620),
 * -1 skips the exchanges (rank-local result; only for measuring the exposed exchange time) */
int b200mf_ctx_set_overlap(b200mf_ctx *ctx, int on);
/* 1: every cell is affine (one Jacobian per cell, as deal.II's MappingInfo compresses Cartesian/affine cells), 0: general */
int b200mf_ctx_cell_type(const b200mf_ctx *ctx);
int64_t b200mf_ctx_n_dofs(const b200mf_ctx *ctx);
size_t b200mf_ctx_memory_consumption(const b200mf_ctx *ctx); /* MatrixFree::memory_consumption, mf_elasticity.h:982-990 */

/* ------------------------------------------------------------------------
 * Device vectors (LinearAlgebra::distributed::Vector<Number> local part).
 * MatrixFree::initialize_dof_vector (mf_elasticity.h:693-706).  Element type is
 * the context's dtype; numbering is the library's internal one (nodes
 * interleaved n_components*node+c), upload/download translate.
 * ------------------------------------------------------------------------ */
int b200mf_vec_create(b200mf_ctx *ctx, void **dvec); /* zero-initialised */
int b200mf_vec_destroy(b200mf_ctx *ctx, void *dvec);
int b200mf_vec_upload(b200mf_ctx *ctx, void *dvec, const void *host);         /* host dtype = ctx dtype */
int b200mf_vec_download(b200mf_ctx *ctx, void *host, const void *dvec);
int b200mf_vec_upload_f64(b200mf_ctx *ctx, void *dvec, const double *host);   /* converts when ctx is FP32 */
int b200mf_vec_download_f64(b200mf_ctx *ctx, double *host, const void *dvec);

/* ------------------------------------------------------------------------
 * Multi-GPU: one rank per GPU.  Replaces the MPI side of
 * LinearAlgebra::distributed::Vector / Utilities::MPI::Partitioner:
 * update_ghost_values (mf_nh_operator.h:610), compress(add) (:620),
 * zero_out_ghost_values, and the allreduce inside l2_norm / operator*
 * (mf_elasticity.h:1907, 1993).  Transport: NCCL over NVLink (libnccl is
 * dlopen'ed at first use, so single-GPU users need no NCCL).
 * ------------------------------------------------------------------------ */
typedef struct b200mf_comm b200mf_comm;
/* rank 0 fills a 128-byte id (ncclUniqueId) and broadcasts it to the other ranks by any host means (MPI_Bcast, a file,
 * torch.distributed) before everyone calls b200mf_comm_create — the role of MPI_COMM_WORLD in src/solid.cc:30-31 */
int b200mf_comm_unique_id(void *id128);
int b200mf_comm_create(b200mf_comm **comm, int rank, int world, const void *id128);
int b200mf_comm_destroy(b200mf_comm *comm);

typedef struct { /* Utilities::MPI::Partitioner of MatrixFree::get_vector_partitioner(); caller numbering */
  int n_neighbors;
  const int *neighbor_ranks;     /* [n_neighbors] */
  const int64_t *send_offsets;   /* [n_neighbors+1] into send_indices: import_indices per neighbour */
  const uint32_t *send_indices;  /* locally owned entries the neighbour ghosts (import_indices) */
  const int64_t *ghost_offsets;  /* [n_neighbors] first local index of the contiguous ghost range owned by the neighbour */
  const int64_t *ghost_counts;   /* [n_neighbors] */
} b200mf_partitioner_desc;
int b200mf_ctx_set_partitioner(b200mf_ctx *ctx, b200mf_comm *comm, const b200mf_partitioner_desc *desc);
/* 1: the ghost exchanges of this context go peer to peer over NVLink (CUDA IPC exchange blocks, direct stores into the
 * receiver's slot); 0: ncclSend/ncclRecv (peers not IPC-reachable, or B200MF_P2P=0 in the environment) */
int b200mf_ctx_exchange_is_p2p(const b200mf_ctx *ctx);
int b200mf_vec_update_ghost_values(b200mf_ctx *ctx, void *dvec);   /* Vector::update_ghost_values */
int b200mf_vec_compress_add(b200mf_ctx *ctx, void *dvec);          /* Vector::compress(VectorOperation::add) */
int b200mf_vec_zero_out_ghost_values(b200mf_ctx *ctx, void *dvec); /* Vector::zero_out_ghost_values */

/* ------------------------------------------------------------------------
 * NeoHookOperator (include/solid_mechanics/mf_nh_operator.h)
 * ------------------------------------------------------------------------ */
typedef struct { /* Material_Compressible_Neo_Hook_One_Field ctor, include/solid_mechanics/material.h:78-94 */
  double mu, nu, rho, alpha_1;
  int formulation; /* 0: vol/iso split, 1: coupled (material.h:116-161) */
} b200mf_material;

/* NeoHookOperator() + set_material (mf_nh_operator.h:527-539): material for material_id==0, inclusion otherwise */
int b200mf_nh_create(b200mf_nh **op, b200mf_ctx *ctx, const b200mf_material *material,
                     const b200mf_material *material_inclusion);
int b200mf_nh_destroy(b200mf_nh *op); /* clear(), mf_nh_operator.h:313-327 */
/* initialize(data_current, data_reference, displacement, caching), mf_nh_operator.h:331-388.  The
 * deformed-frame MatrixFree is not needed: the library derives J_cur = F J_ref (MappingQEulerian). */
int b200mf_nh_initialize(b200mf_nh *op, int caching);
/* cache(), mf_nh_operator.h:392-523.  *n_nonpositive_detF receives the number of quadrature points
 * violating the reference's debug assertion det F > 0 (:430-435); returns B200MF_ERR_NUMERIC if > 0. */
int b200mf_nh_cache(b200mf_nh *op, const void *displacement_dvec, int64_t *n_nonpositive_detF);
int b200mf_nh_cache_host(b200mf_nh *op, const double *displacement_host, int64_t *n_nonpositive_detF);
/* vmult / Tvmult / vmult_add / Tvmult_add, mf_nh_operator.h:543-625 (operator is symmetric) */
int b200mf_nh_vmult(b200mf_nh *op, void *dst_dvec, const void *src_dvec);
int b200mf_nh_vmult_add(b200mf_nh *op, void *dst_dvec, const void *src_dvec);
int b200mf_nh_Tvmult(b200mf_nh *op, void *dst_dvec, const void *src_dvec);
int b200mf_nh_Tvmult_add(b200mf_nh *op, void *dst_dvec, const void *src_dvec);
/* "Fused solver vector work": vmult that also accumulates *dot_device += src . (A src) (this rank's cells and identity
 * rows; sum over ranks = the global value) inside the operator pass - the p . Ap of SolverCG without a second sweep
 * over the vectors.  dot_device is a device double the caller zeroes. */
int b200mf_nh_vmult_dot(b200mf_nh *op, void *dst_dvec, const void *src_dvec, double *dot_device);
int b200mf_nh_vmult_dot_host(b200mf_nh *op, void *dst_dvec, const void *src_dvec, double *result_host); /* summed over ranks */
/* parity / drop-in mode: caller-owned host vectors, H2D + vmult + D2H */
int b200mf_nh_vmult_host(b200mf_nh *op, double *dst_host, const double *src_host);
/* compute_diagonal(), mf_nh_operator.h:1023-1060: fills diagonal and inverse (constrained -> 1) */
int b200mf_nh_compute_diagonal(b200mf_nh *op);
int b200mf_nh_diagonal(b200mf_nh *op, const void **diag_dvec, const void **inv_diag_dvec); /* get_matrix_diagonal_inverse :157-161 */
/* The smoother hook (deal.II's vmult(dst, src, before, after) overload, which the reference does not provide - SURVEY.md §8b):
 * the operator pass alone.  dst must be ZERO on entry (all local entries, ghosts included) and the constrained rows of the result
 * are left 0: the smoother's vector update that consumes dst treats them as identity rows and zeroes dst again, so a Chebyshev
 * iteration is two launches (cell loop + update) instead of four (memset, cell loop, identity rows, update).  Ghost exchange and
 * compress(add) inside as in vmult.  diagonal_generation changes whenever compute_diagonal ran (for cached copies of the diagonal). */
int b200mf_nh_vmult_cells(b200mf_nh *op, void *dst_zero_dvec, const void *src_dvec);
uint64_t b200mf_nh_diagonal_generation(const b200mf_nh *op);
int b200mf_nh_diagonal_host(b200mf_nh *op, double *diag_host, double *inv_diag_host);
/* precondition_Jacobi(dst, src, omega), mf_nh_operator.h:281-291 */
int b200mf_nh_precondition_jacobi(b200mf_nh *op, void *dst_dvec, const void *src_dvec, double omega);
/* el(row,row), mf_nh_operator.h:1064-1073 (row in caller numbering) */
int b200mf_nh_el(b200mf_nh *op, int64_t row, double *value);
int64_t b200mf_nh_m(const b200mf_nh *op); /* m() == n(), mf_nh_operator.h:295-309 */
size_t b200mf_nh_memory_consumption(const b200mf_nh *op); /* mf_nh_operator.h:241-260 */
/* Cell part of Solid::assemble_residual, fast FEEvaluation path (mf_elasticity.h:1764-1846):
 *   rhs += int [ -tau(F) F^-T : Grad v + rho (body_force - acceleration) . v ] dV,   F = I + Grad total_displacement,
 * both vectors read without constraints (read_dof_values_plain), constrained rows of rhs not written.  The caller
 * zeroes rhs before (:1655), adds the interface-traction faces (:1849-1897, preCICE data) and then calls
 * b200mf_vec_compress_add (:1899) and b200mf_vec_set_zero_constrained (:1905).  Usable right after
 * b200mf_nh_create (needs no cache()).  *n_nonpositive_detF as in b200mf_nh_cache. */
int b200mf_nh_assemble_residual(b200mf_nh *op, void *rhs_dvec, const void *total_displacement_dvec, const void *acceleration_dvec,
                                const double body_force[3], int64_t *n_nonpositive_detF);
/* test hook: copy one cached quadrature table to the host in natural order [n_cells][nq^3].
 * field ids: see B200MF_QF_* */
int b200mf_nh_qdata_host(b200mf_nh *op, int field, double *out);
int b200mf_nh_n_qfields(const b200mf_nh *op);

/* ------------------------------------------------------------------------
 * LaplaceOperator: (M + dt*c*K) u, include/heat_transfer/heat_transfer.h:81-250
 * ------------------------------------------------------------------------ */
/* coefficient: [n_cells][nq^3] table (evaluate_coefficient :150-166) or NULL for c == 1 (:64-71) */
int b200mf_laplace_create(b200mf_laplace **op, b200mf_ctx *ctx, double delta_t, const double *coefficient);
int b200mf_laplace_destroy(b200mf_laplace *op);
int b200mf_laplace_set_delta_t(b200mf_laplace *op, double delta_t); /* set_delta_t :141-146 */
int b200mf_laplace_vmult(b200mf_laplace *op, void *dst_dvec, const void *src_dvec);     /* Base::vmult -> apply_add :194-201 */
int b200mf_laplace_vmult_add(b200mf_laplace *op, void *dst_dvec, const void *src_dvec);
int b200mf_laplace_Tvmult(b200mf_laplace *op, void *dst_dvec, const void *src_dvec);
int b200mf_laplace_Tvmult_add(b200mf_laplace *op, void *dst_dvec, const void *src_dvec);
/* MatrixFreeOperators::Base::vmult_interface_down/_up (used by mg::Matrix on adaptively refined levels): dst = 0 + the
 * refinement-edge rows; the reference refines globally only (heat_transfer.h:421-425), so both reduce to dst = 0 */
int b200mf_laplace_vmult_interface_down(b200mf_laplace *op, void *dst_dvec, const void *src_dvec);
int b200mf_laplace_vmult_interface_up(b200mf_laplace *op, void *dst_dvec, const void *src_dvec);
int b200mf_laplace_vmult_cells(b200mf_laplace *op, void *dst_zero_dvec, const void *src_dvec); /* smoother hook, see b200mf_nh_vmult_cells */
uint64_t b200mf_laplace_diagonal_generation(const b200mf_laplace *op);
int64_t b200mf_laplace_m(const b200mf_laplace *op);                        /* Base::m() == n(), heat_transfer.h:81 (inherited) */
int b200mf_laplace_el(b200mf_laplace *op, int64_t row, double *value);     /* Base::el(row, row): diagonal entry, needs compute_diagonal */
int b200mf_laplace_vmult_dot(b200mf_laplace *op, void *dst_dvec, const void *src_dvec, double *dot_device);
int b200mf_laplace_vmult_dot_host(b200mf_laplace *op, void *dst_dvec, const void *src_dvec, double *result_host);
int b200mf_laplace_vmult_host(b200mf_laplace *op, double *dst_host, const double *src_host);
int b200mf_laplace_compute_diagonal(b200mf_laplace *op); /* :205-230 */
/* Cell loop of LaplaceProblem::assemble_rhs (heat_transfer.h:599-615): rhs += int (solution_old + delta_t f(x_q)) phi on
 * this operator's MatrixFree (the reference uses the unconstrained inhomogeneous_operator's).  source_q: f at the
 * quadrature points [n_cells][nq^3] (evaluate_function on phi.quadrature_point(q)) or NULL for the constant source_const.
 * The caller adds the lifting -A_inhom * solution_bc (:586-592, a vmult) before, the interface-flux faces (:617-652) and
 * b200mf_vec_compress_add (:654) after. */
int b200mf_laplace_assemble_rhs(b200mf_laplace *op, void *rhs_dvec, const void *solution_old_dvec, const double *source_q,
                                double source_const);
int b200mf_laplace_diagonal(b200mf_laplace *op, const void **diag_dvec, const void **inv_diag_dvec);
int b200mf_laplace_diagonal_host(b200mf_laplace *op, double *diag_host, double *inv_diag_host);

/* ------------------------------------------------------------------------
 * Interface faces of Solid::assemble_residual (include/solid_mechanics/mf_elasticity.h:1849-1897): the traction read
 * from preCICE at the face quadrature points is pulled back (t_0 = |det F F^-T N| t) and integrated against the test
 * functions.  A face set is built once from the boundary faces with id == interface_id (cases/case_base.h:25):
 *   cells[f]          the caller's cell index of face f (cell order of b200mf_ctx_create)
 *   face_no[f]        2*direction + side (deal.II's GeometryInfo face numbering)
 *   cell_vertices     [n_faces][8][3] vertices of those cells (Q1 geometry of the reference configuration, vertex m at
 *                     (m&1, m>>1&1, m>>2&1)), as b200mf_ctx_desc::vertices
 * Quadrature points of a face normal to direction d are ordered q = q_t1 + nq*q_t2 with the tangential directions
 * t1 < t2 (deal.II orders the y-faces (z, x): the shim permutes the preCICE data once per window).
 * ------------------------------------------------------------------------ */
typedef struct b200mf_face_set b200mf_face_set;
int b200mf_face_set_create(b200mf_face_set **fs, b200mf_ctx *ctx, int64_t n_faces, const int64_t *cells, const int32_t *face_no,
                           const double *cell_vertices);
int b200mf_face_set_destroy(b200mf_face_set *fs);
int64_t b200mf_face_set_n_q_points(const b200mf_face_set *fs); /* n_faces * nq^2 */
/* [n_faces][nq^2][3] positions in the reference configuration: the interface mesh of Adapter::initialize
 * (include/adapter/adapter.h: the read data lives on the face quadrature points) */
int b200mf_face_set_quadrature_points(b200mf_face_set *fs, double *xyz_host);
/* once per coupling window: traction_q [n_faces][nq^2][3] (precice_adapter->read_on_quadrature_point, :1876-1879) or,
 * traction_q == NULL, one constant vector traction_const[3] */
int b200mf_face_set_set_traction(b200mf_face_set *fs, const double *traction_q, const double *traction_const);
/* every Newton iteration: rhs += face integrals (no compress(add), no set_zero - the caller does both after cells + faces,
 * :1899-1905, with b200mf_vec_compress_add and b200mf_vec_set_zero_constrained) */
int b200mf_face_set_assemble_traction(b200mf_face_set *fs, void *rhs_dvec, const void *total_displacement_dvec);

/* Heat solver on the same face sets (scalar context: the data of b200mf_face_set_set_traction has ONE component per face
 * quadrature point, the heat flux read from the coupling partner).
 * assemble_flux: rhs += scale * int_Gamma flux(q) phi dA - the interface faces of LaplaceProblem::assemble_rhs
 * (include/heat_transfer/heat_transfer.h:617-652, scale = -delta_t); no compress(add) (the caller's :654). */
int b200mf_face_set_assemble_flux(b200mf_face_set *fs, void *rhs_dvec, double scale);

/* MatrixFreeOperators::BoundaryMassOperator (include/adapter/boundary_mass_operator.h): dst = int_Gamma u phi dA over the
 * faces of the set (apply_add :137-161, with the ghost update and compress(add) of the reference inside; constrained rows
 * act as the identity like every MatrixFreeOperators::Base operator).  compute_diagonal (:79-118) stores M_Gamma * 1 (the
 * row sums - not the diagonal) and its inverse where the entry is positive; rows off the interface keep 0.
 * restrict: dst = src on the locally owned DoFs of the faces, 0 elsewhere (include/adapter/boundary_projector.h:93-95).
 * project (boundary_projector.h:86-110, called from LaplaceProblem::evaluate_boundary_flux, heat_transfer.h:935-939):
 * residual <- solution of M_Gamma x = residual|Gamma by CG, ReductionControl(500, 0, 1e-12), PreconditionJacobi(0.8),
 * start vector = residual; ghost values of the result updated.  B200MF_ERR_NUMERIC if CG does not converge. */
typedef struct b200mf_boundary_mass b200mf_boundary_mass;
int b200mf_boundary_mass_create(b200mf_boundary_mass **op, b200mf_face_set *fs); /* the face set must outlive the operator */
int b200mf_boundary_mass_destroy(b200mf_boundary_mass *op);
b200mf_ctx *b200mf_boundary_mass_ctx(b200mf_boundary_mass *op);
int b200mf_boundary_mass_vmult(b200mf_boundary_mass *op, void *dst_dvec, const void *src_dvec);
int b200mf_boundary_mass_vmult_add(b200mf_boundary_mass *op, void *dst_dvec, const void *src_dvec);
int b200mf_boundary_mass_compute_diagonal(b200mf_boundary_mass *op);
int b200mf_boundary_mass_diagonal(b200mf_boundary_mass *op, const void **diag_dvec, const void **inv_diag_dvec);
int b200mf_boundary_mass_restrict(b200mf_boundary_mass *op, void *dst_dvec, const void *src_dvec);
int b200mf_boundary_mass_project(b200mf_boundary_mass *op, void *residual_dvec, int *iterations, double *final_residual);

b200mf_ctx *b200mf_nh_ctx(b200mf_nh *op);           /* the MatrixFree the operator was built on (data_in_use, :213) */
b200mf_ctx *b200mf_laplace_ctx(b200mf_laplace *op); /* MatrixFreeOperators::Base::get_matrix_free */

/* ------------------------------------------------------------------------
 * Vector algebra on device vectors: the LinearAlgebra::distributed::Vector members the solvers and the
 * Newton loop call (l2_norm mf_elasticity.h:1907, 1993, 2061; operator* / sadd / = inside SolverCG).
 * Reductions run over owned entries and are summed over ranks.
 * ------------------------------------------------------------------------ */
int b200mf_vec_dot(b200mf_ctx *ctx, const void *a_dvec, const void *b_dvec, double *out);
int b200mf_vec_l2_norm(b200mf_ctx *ctx, const void *a_dvec, double *out);
int b200mf_vec_sadd(b200mf_ctx *ctx, void *y_dvec, double s, double a, const void *x_dvec); /* y = s*y + a*x */
int b200mf_vec_set_zero(b200mf_ctx *ctx, void *y_dvec);
int b200mf_vec_set_zero_constrained(b200mf_ctx *ctx, void *y_dvec); /* AffineConstraints::set_zero, mf_elasticity.h:1905, 2059 */
/* copy between two contexts built on the same mesh and numbering but different number types
 * (double system vector <-> float level vector: MGTransferMatrixFree::copy_to_mg / copy_from_mg,
 * level displacement mf_elasticity.h:314, 920-922) */
int b200mf_vec_convert(b200mf_ctx *dst_ctx, void *dst_dvec, b200mf_ctx *src_ctx, const void *src_dvec);

/* ------------------------------------------------------------------------
 * PreconditionChebyshev<LevelMatrixType, LevelVectorType> with the operator's inverse diagonal as inner
 * preconditioner (mf_elasticity.h:371-372, 1149-1180; heat_transfer.h:764-788).  Algorithm: SURVEY.md A.5.
 * ------------------------------------------------------------------------ */
typedef struct b200mf_chebyshev b200mf_chebyshev;
typedef struct { /* PreconditionChebyshev::AdditionalData */
  int degree;              /* >= 1, or -1 = numbers::invalid_unsigned_int: solver mode (mf_elasticity.h:1160-1161) */
  double smoothing_range;  /* > 1: [lambda_max/range, lambda_max]; <= 1 in solver mode: target reduction (:1158-1159) */
  int eig_cg_n_iterations; /* CG iterations of the eigenvalue estimate; 0: use max_eigenvalue (:1162-1165, 1173-1174) */
  double eig_cg_residual;  /* kept for interface parity (deal.II >= 9.6 runs a fixed iteration count to 1e-10) */
  double max_eigenvalue;
  int64_t first_global_index; /* global index of this rank's first owned DoF (deterministic start vector) */
  int64_t n_global_dofs;      /* global problem size for its mean value; 0 = local owned size */
} b200mf_chebyshev_data;
int b200mf_chebyshev_create_nh(b200mf_chebyshev **cheb, b200mf_nh *op, const b200mf_chebyshev_data *data);
int b200mf_chebyshev_create_laplace(b200mf_chebyshev **cheb, b200mf_laplace *op, const b200mf_chebyshev_data *data);
int b200mf_chebyshev_destroy(b200mf_chebyshev *cheb);
/* estimate_eigenvalues(): (re)runs the CG/Lanczos estimate — the reference does so for every level at every Newton
 * step (setup_gmg(false), mf_elasticity.h:1352); out pointers may be NULL */
int b200mf_chebyshev_estimate_eigenvalues(b200mf_chebyshev *cheb, double *min_eigenvalue, double *max_eigenvalue, int *cg_iterations,
                                          int *degree);
int b200mf_chebyshev_vmult(b200mf_chebyshev *cheb, void *dst_dvec, const void *src_dvec); /* zero initial guess: degree-1 matvecs */
int b200mf_chebyshev_step(b200mf_chebyshev *cheb, void *dst_dvec, const void *src_dvec);  /* dst is the initial guess: degree matvecs */

/* ------------------------------------------------------------------------
 * MGTransferMatrixFree<3,float> between two uniformly refined levels (mf_elasticity.h:863-877, 881-929;
 * heat_transfer.h:761-762).  children: [n_coarse_cells][8] cell indices on the fine context, child (cx,cy,cz)
 * at cx + 2 cy + 4 cz (the order of the caller's cell arrays).
 * Replicated coarse level: if `coarse` has no partitioner but `fine` has one, the coarse context is the WHOLE coarse mesh, held by
 * every rank (the levels a distributed triangulation leaves most ranks without cells on).  children rows are then all -1 for the
 * parents whose children live on other ranks (a parent has all 8 or none of them here).  restrict_and_add sums the ranks' shares
 * with one all-reduce and adds the sum to every rank's copy, prolongate* need no communication, interpolate merges the ranks'
 * pieces; the coarse vector is identical on all ranks afterwards and everything below runs redundantly without exchanges.
 * ------------------------------------------------------------------------ */
typedef struct b200mf_mg_transfer b200mf_mg_transfer;
int b200mf_mg_transfer_create(b200mf_mg_transfer **t, b200mf_ctx *coarse, b200mf_ctx *fine, const int64_t *children);
int b200mf_mg_transfer_destroy(b200mf_mg_transfer *t);
int b200mf_mg_transfer_prolongate(b200mf_mg_transfer *t, void *fine_dst, const void *coarse_src);
int b200mf_mg_transfer_prolongate_and_add(b200mf_mg_transfer *t, void *fine_dst, const void *coarse_src);
int b200mf_mg_transfer_restrict_and_add(b200mf_mg_transfer *t, void *coarse_dst, const void *fine_src);
int b200mf_mg_transfer_interpolate(b200mf_mg_transfer *t, void *coarse_dst, const void *fine_src); /* one step of interpolate_to_mg */

/* ------------------------------------------------------------------------
 * Multigrid<LevelVectorType> V-cycle wrapped as PreconditionMG (mf_elasticity.h:1209-1224, 1272-1276):
 * level l > 0: Chebyshev pre-smoothing (vmult), residual, restrict_and_add, recursion, prolongate_and_add,
 * Chebyshev post-smoothing (step); level 0: smoothers[0] applied as solver (MGCoarseGridApplySmoother, :1180).
 * ------------------------------------------------------------------------ */
typedef struct b200mf_mg b200mf_mg;
int b200mf_mg_create(b200mf_mg **mg, int n_levels, b200mf_chebyshev *const *smoothers /*[n_levels]*/,
                     b200mf_mg_transfer *const *transfers /*[n_levels-1]*/, b200mf_ctx *system_ctx);
int b200mf_mg_destroy(b200mf_mg *mg);
/* coarse-grid solver of the V-cycle.  kind 0 (default): the level-0 smoother applied as a solver (MGCoarseGridApplySmoother,
 * 'MF Chebyshev coarse = true'); kind 1: SolverCG under ReductionControl(max_iterations, abs_tolerance, reduction) preconditioned
 * with the level-0 smoother (MGCoarseGridIterativeSolver, include/solid_mechanics/mf_elasticity.h:1183-1198: max(m(), 100),
 * 1e-10, 1e-3) */
int b200mf_mg_set_coarse_solver(b200mf_mg *mg, int kind, int max_iterations, double abs_tolerance, double reduction);
int b200mf_mg_coarse_iterations(const b200mf_mg *mg); /* CG iterations of the last coarse solve (kind 1) */
/* Levels 0..top_level of the V-cycle are recorded once as a CUDA graph and replayed by every further b200mf_mg_vmult until a
 * smoother is re-estimated or an operator/context of those levels is reconfigured (then recorded again).  Default: n_levels-2
 * (everything below the finest level; those levels are launch-bound), -1 = plain launches (also: environment B200MF_MG_GRAPH=0).
 * Not used with coarse solver kind 1 (its convergence check synchronises).  graph_info: state 0 = nothing recorded yet,
 * 1 = replaying, -1 = recording failed and the levels run as plain launches; launches = kernels per replay. */
int b200mf_mg_set_graph_levels(b200mf_mg *mg, int top_level);
int b200mf_mg_graph_info(const b200mf_mg *mg, int *top_level, int *state, int64_t *records, int64_t *replays, int64_t *launches);
int b200mf_mg_vmult(b200mf_mg *mg, void *dst_dvec, const void *src_dvec); /* system-context vectors (copy_to_mg / copy_from_mg inside) */

/* ------------------------------------------------------------------------
 * SolverCG<VectorType> (mf_elasticity.h:2024-2051; heat_transfer.h:724-737, 810-811), SURVEY.md A.6
 * ------------------------------------------------------------------------ */
typedef enum { B200MF_PRECOND_NONE = 0, B200MF_PRECOND_JACOBI = 1, B200MF_PRECOND_MG = 2 } b200mf_precond;
typedef struct b200mf_cg b200mf_cg;
typedef struct {
  double tolerance;    /* absolute: tol_lin * ||rhs|| (mf_elasticity.h:1993, 2018-2021) */
  int max_iterations;  /* SolverControl max steps (:2016) */
  int preconditioner;  /* b200mf_precond */
  double jacobi_omega; /* PreconditionJacobi relaxation (:2028-2030) */
  b200mf_mg *mg;       /* PreconditionMG (:2048) */
  int x_is_zero;       /* skip the start-up matvec (the solid solver zeroes newton_update, :1990) */
} b200mf_cg_control;
typedef struct {
  int iterations; /* SolverControl::last_step() */
  double residual, initial_residual;
  int converged;
  double min_eigenvalue, max_eigenvalue; /* Lanczos estimates (condition-number slot, :2000-2010) */
} b200mf_cg_result;
int b200mf_cg_create_nh(b200mf_cg **solver, b200mf_nh *op);
int b200mf_cg_create_laplace(b200mf_cg **solver, b200mf_laplace *op);
int b200mf_cg_create_boundary_mass(b200mf_cg **solver, b200mf_boundary_mass *op); /* boundary_projector.h:97-104 as separate calls */
int b200mf_cg_destroy(b200mf_cg *solver);
int b200mf_cg_solve(b200mf_cg *solver, void *x_dvec, const void *b_dvec, const b200mf_cg_control *control, b200mf_cg_result *result);

#ifdef __cplusplus
}
#endif
#endif /* B200MF_H */
