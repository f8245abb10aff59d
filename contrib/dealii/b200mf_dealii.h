// b200mf_dealii.h — the deal.II-side shim over the C ABI of libb200mf (include/b200mf.h).
//
// STATUS: deal.II (>= 9.6), p4est, MPI and preCICE are not installed in the image this library is built and tested in, so
// this header has never been compiled against deal.II.  tests/test_shim_syntax.py compiles it (g++ -fsyntax-only, every
// template instantiated) against tests/mock_dealii/, a minimal stand-in that declares the deal.II classes and members used
// here with deal.II's documented signatures - that catches syntax and type errors, not semantic drift.  It lives under
// contrib/ (not include/) until a deal.II build compiles it in CI.  It is the binding a maintainer of
// davidscn/matrix-free-dealii-precice adds; the identical call sequence is what tests/ exercise through the ctypes mirror
// (matrix-free-dealii-precice_b200/b200mf/operators.py).  Public deal.II API only.
//
// It provides, with the member set the reference's solvers call:
//   B200MF::make_context(MatrixFree, comm)  -> b200mf_ctx*      arrays extracted once per (re)initialisation
//   B200MF::DeviceVector                                         RAII handle of a library-owned device vector (performance mode)
//   B200MF::NeoHookOperator<dim, Number>    drop-in for ::NeoHookOperator<dim, Number> (include/solid_mechanics/mf_nh_operator.h:102-173)
//   B200MF::LaplaceOperator<dim, Number>    drop-in for Heat_Transfer::LaplaceOperator (include/heat_transfer/heat_transfer.h:81-127,
//                                           with the MatrixFreeOperators::Base surface the heat solver and mg::Matrix call)
// Parity mode: deal.II's SolverCG / PreconditionChebyshev / Multigrid keep running on the host and call vmult with
// LinearAlgebra::distributed::Vector; every operator keeps two pinned staging buffers, so a vmult allocates nothing.
// Performance mode: the same operators applied to B200MF::DeviceVector, and b200mf_cg_solve / b200mf_mg_vmult on the
// handles (handle()) replace the deal.II solver objects wholesale.
#pragma once
#include <deal.II/base/mpi.h>
#include <deal.II/lac/diagonal_matrix.h>
#include <deal.II/lac/la_parallel_vector.h>
#include <deal.II/matrix_free/fe_evaluation.h>
#include <deal.II/matrix_free/matrix_free.h>

#include <b200mf.h>
#include <cuda_runtime_api.h> // cudaMallocHost / cudaFreeHost for the pinned staging buffers

#include <cstdint>
#include <map>
#include <memory>
#include <string>
#include <type_traits>
#include <vector>

namespace B200MF
{
  inline void
  check(const int rc)
  {
    // deal.II error convention: exceptions caught in main (src/solid.cc:64-89)
    AssertThrow(rc == 0, dealii::ExcMessage(b200mf_last_error()));
  }

  // One NCCL communicator per MPI communicator: rank 0 creates the id, everybody receives it (the role MPI_COMM_WORLD
  // plays in src/solid.cc:30-31).
  inline b200mf_comm *
  make_communicator(const MPI_Comm mpi_comm)
  {
    unsigned char id[128] = {};
    const int     rank = dealii::Utilities::MPI::this_mpi_process(mpi_comm), world = dealii::Utilities::MPI::n_mpi_processes(mpi_comm);
    if (rank == 0)
      check(b200mf_comm_unique_id(id));
    MPI_Bcast(id, 128, MPI_BYTE, 0, mpi_comm);
    b200mf_comm *comm = nullptr;
    check(b200mf_comm_create(&comm, rank, world, id));
    return comm;
  }

  // Extract from a MatrixFree (the *reference*-configuration one, mf_elasticity.h:739-746) everything the library needs.
  // Cell-local DoF indices are obtained through FEEvaluation itself: a vector whose entries are their own local index is
  // read with read_dof_values_plain, so the ordering is exactly FEEvaluation's (component-major, lexicographic)
  // without touching DoFInfo internals.  Indices are split into 20-bit halves so that a float MatrixFree stays exact.
  template <int dim, int n_components, typename Number>
  b200mf_ctx *
  make_context(const dealii::MatrixFree<dim, Number> &mf, b200mf_comm *comm /* nullptr on one rank */)
  {
    static_assert(dim == 3, "the device kernels are 3-D");
    using VectorType         = dealii::LinearAlgebra::distributed::Vector<Number>;
    const unsigned int degree = mf.get_dof_handler().get_fe().degree, n = degree + 1, n3 = n * n * n;
    const unsigned int n_lanes = dealii::VectorizedArray<Number>::size();

    VectorType lo, hi;
    mf.initialize_dof_vector(lo);
    mf.initialize_dof_vector(hi);
    const auto        &part     = *mf.get_vector_partitioner();
    const unsigned int n_owned = part.locally_owned_size(), n_local = n_owned + part.n_ghost_indices();
    for (unsigned int i = 0; i < n_local; ++i)
      {
        lo.local_element(i) = Number(i & 0xFFFFFu);
        hi.local_element(i) = Number(i >> 20);
      }

    std::vector<std::uint32_t> dof;    // [cell][n_components * n3]
    std::vector<double>        invjac; // [cell][n3][9],  [e][f] = d xi_f / d x_e (mf_nh_operator.h:872-887)
    std::vector<double>        jxw;    // [cell][n3]
    std::vector<std::uint8_t>  matid;
    dealii::FEEvaluation<dim, -1, 0, n_components, Number> phi(mf); // include/base/fe_integrator.h:10-22
    for (unsigned int b = 0; b < mf.n_cell_batches(); ++b)
      {
        phi.reinit(b);
        std::vector<std::uint32_t> idx(n_components * n3 * n_lanes);
        phi.read_dof_values_plain(lo);
        for (unsigned int i = 0; i < n_components * n3; ++i)
          for (unsigned int l = 0; l < n_lanes; ++l)
            idx[l * n_components * n3 + i] = static_cast<std::uint32_t>(phi.begin_dof_values()[i][l]);
        phi.read_dof_values_plain(hi);
        for (unsigned int i = 0; i < n_components * n3; ++i)
          for (unsigned int l = 0; l < n_lanes; ++l)
            idx[l * n_components * n3 + i] |= static_cast<std::uint32_t>(phi.begin_dof_values()[i][l]) << 20;
        for (unsigned int l = 0; l < mf.n_active_entries_per_cell_batch(b); ++l)
          {
            dof.insert(dof.end(), idx.begin() + l * n_components * n3, idx.begin() + (l + 1) * n_components * n3);
            for (unsigned int q = 0; q < n3; ++q)
              {
                const auto J = phi.inverse_jacobian(q);
                for (unsigned int e = 0; e < 3; ++e)
                  for (unsigned int f = 0; f < 3; ++f)
                    invjac.push_back(J[e][f][l]);
                jxw.push_back(phi.JxW(q)[l]);
              }
            matid.push_back(static_cast<std::uint8_t>(mf.get_cell_iterator(b, l)->material_id())); // mf_nh_operator.h:402-404
          }
      }

    const auto                &con = mf.get_constrained_dofs(); // local indices (mf_nh_operator.h:623-624)
    std::vector<std::uint32_t> con32(con.begin(), con.end());

    b200mf_ctx_desc d{};
    d.degree           = degree;
    d.n_q_points_1d    = n; // the reference always runs Quadrature order = degree + 1 (elasticity.prm:5-8)
    d.n_components     = n_components;
    d.dtype            = std::is_same<Number, double>::value ? B200MF_F64 : B200MF_F32;
    d.n_cells          = matid.size();
    d.n_owned          = n_owned;
    d.n_dofs           = n_local;
    d.dof_indices      = dof.data();
    d.constrained_dofs = con32.data();
    d.n_constrained    = con32.size();
    d.inverse_jacobian = invjac.data();
    d.jxw              = jxw.data();
    d.material_id      = matid.data();
    b200mf_ctx *ctx    = nullptr;
    check(b200mf_ctx_create(&ctx, &d));

    if (comm != nullptr)
      {
        // Utilities::MPI::Partitioner -> b200mf_partitioner_desc: import indices are stored as compressed
        // [begin, end) ranges, grouped by target in the order of import_targets(); ghosts are grouped by owner in the
        // order of ghost_targets().
        std::vector<int>           ranks;
        std::vector<std::int64_t>  send_off(1, 0), ghost_off, ghost_cnt;
        std::vector<std::uint32_t> send_idx;
        for (const auto &r : part.import_indices())
          for (unsigned int i = r.first; i < r.second; ++i)
            send_idx.push_back(i);
        // a neighbour may appear in only one of the two lists: take the union, ordered by rank
        std::map<unsigned int, std::pair<unsigned int, unsigned int>> nb; // rank -> (n_import, n_ghost)
        for (const auto &t : part.import_targets())
          nb[t.first].first = t.second;
        for (const auto &t : part.ghost_targets())
          nb[t.first].second = t.second;
        std::int64_t                         ghost_begin = n_owned;
        std::map<unsigned int, std::int64_t> import_begin, ghost_first;
        {
          std::int64_t o = 0;
          for (const auto &t : part.import_targets())
            {
              import_begin[t.first] = o;
              o += t.second;
            }
          for (const auto &t : part.ghost_targets())
            {
              ghost_first[t.first] = ghost_begin;
              ghost_begin += t.second;
            }
        }
        std::vector<std::uint32_t> send_sorted;
        for (const auto &kv : nb)
          {
            ranks.push_back(kv.first);
            if (kv.second.first)
              send_sorted.insert(send_sorted.end(), send_idx.begin() + import_begin[kv.first],
                                 send_idx.begin() + import_begin[kv.first] + kv.second.first);
            send_off.push_back(send_sorted.size());
            ghost_off.push_back(kv.second.second ? ghost_first[kv.first] : n_owned);
            ghost_cnt.push_back(kv.second.second);
          }
        b200mf_partitioner_desc pd{};
        pd.n_neighbors    = ranks.size();
        pd.neighbor_ranks = ranks.data();
        pd.send_offsets   = send_off.data();
        pd.send_indices   = send_sorted.data();
        pd.ghost_offsets  = ghost_off.data();
        pd.ghost_counts   = ghost_cnt.data();
        check(b200mf_ctx_set_partitioner(ctx, comm, &pd));
      }
    return ctx;
  }


  // Library-owned device vector in the context's internal numbering (performance mode: vectors never leave the GPU
  // between solver calls).  upload()/download() translate from/to the MatrixFree partitioner's local numbering
  // [owned | ghosts] in double precision.
  class DeviceVector
  {
  public:
    explicit DeviceVector(b200mf_ctx *ctx_)
      : ctx(ctx_)
    {
      check(b200mf_vec_create(ctx, &ptr));
    }
    DeviceVector(const DeviceVector &) = delete;
    DeviceVector &
    operator=(const DeviceVector &) = delete;
    ~DeviceVector()
    {
      b200mf_vec_destroy(ctx, ptr);
    }
    template <typename VectorType>
    void
    upload(const VectorType &v, std::vector<double> &scratch)
    {
      scratch.assign(b200mf_ctx_n_dofs(ctx), 0.0);
      for (unsigned int i = 0; i < v.locally_owned_size(); ++i)
        scratch[i] = v.local_element(i);
      check(b200mf_vec_upload_f64(ctx, ptr, scratch.data()));
    }
    template <typename VectorType>
    void
    download(VectorType &v, std::vector<double> &scratch) const
    {
      scratch.resize(b200mf_ctx_n_dofs(ctx));
      check(b200mf_vec_download_f64(ctx, scratch.data(), ptr));
      for (unsigned int i = 0; i < v.locally_owned_size(); ++i)
        v.local_element(i) = typename VectorType::value_type(scratch[i]);
    }
    void *
    data() const
    {
      return ptr;
    }
    double
    l2_norm() const // LinearAlgebra::distributed::Vector::l2_norm over the owned entries of all ranks
    {
      double v = 0;
      check(b200mf_vec_l2_norm(ctx, ptr, &v));
      return v;
    }
    void
    sadd(const double s, const double a, const DeviceVector &x) // *this = s * *this + a * x
    {
      check(b200mf_vec_sadd(ctx, ptr, s, a, x.ptr));
    }

  private:
    b200mf_ctx *ctx;
    void       *ptr = nullptr;
  };

  // Two pinned host buffers of the context's local vector size: the staging area of the host-vector entry points
  // (b200mf_*_vmult_host pipelines H2D, the cell waves and D2H out of / into them).
  struct PinnedStaging
  {
    double     *src = nullptr, *dst = nullptr;
    std::size_t n   = 0;
    void
    reinit(const std::size_t n_)
    {
      release();
      n = n_;
      AssertThrow(cudaMallocHost(reinterpret_cast<void **>(&src), n * sizeof(double)) == cudaSuccess, dealii::ExcMessage("cudaMallocHost"));
      AssertThrow(cudaMallocHost(reinterpret_cast<void **>(&dst), n * sizeof(double)) == cudaSuccess, dealii::ExcMessage("cudaMallocHost"));
      for (std::size_t i = 0; i < n; ++i)
        src[i] = dst[i] = 0.0;
    }
    void
    release()
    {
      if (src)
        cudaFreeHost(src);
      if (dst)
        cudaFreeHost(dst);
      src = dst = nullptr;
      n         = 0;
    }
    ~PinnedStaging()
    {
      release();
    }
  };

  // Drop-in for ::NeoHookOperator<dim, Number> (mf_nh_operator.h:99-237).
  template <int dim, typename Number>
  class NeoHookOperator : public dealii::Subscriptor
  {
  public:
    using value_type = Number;
    using size_type  = dealii::types::global_dof_index;
    using VectorType = dealii::LinearAlgebra::distributed::Vector<Number>;

    ~NeoHookOperator()
    {
      clear();
    }

    void
    clear() // :313-327
    {
      b200mf_nh_destroy(op);
      b200mf_ctx_destroy(ctx);
      op  = nullptr;
      ctx = nullptr;
      staging.release();
      diagonal_is_available = false;
    }

    template <typename Material> // Material_Compressible_Neo_Hook_One_Field (material.h:78-94, 252-258)
    void
    set_material(const std::shared_ptr<Material> &material, const std::shared_ptr<Material> &material_inclusion) // :527-539
    {
      mat  = b200mf_material{material->mu, material->nu, material->rho, material->rho_alpha / material->rho, int(material->formulation)};
      incl = b200mf_material{material_inclusion->mu, material_inclusion->nu, material_inclusion->rho,
                             material_inclusion->rho_alpha / material_inclusion->rho, int(material_inclusion->formulation)};
    }

    void
    initialize(std::shared_ptr<const dealii::MatrixFree<dim, Number>> /*data_current: not needed, J_cur = F J_ref*/,
               std::shared_ptr<const dealii::MatrixFree<dim, Number>> data_reference,
               const VectorType                                      &displacement_,
               const std::string                                      caching,
               b200mf_comm                                           *comm = nullptr) // :331-388
    {
      clear();
      data         = data_reference;
      ctx          = make_context<dim, dim, Number>(*data_reference, comm);
      displacement = &displacement_; // raw pointer, caller guarantees the lifetime (:339-341)
      check(b200mf_nh_create(&op, ctx, &mat, &incl));
      const int level = caching == "scalar_referential" ? B200MF_CACHE_SCALAR_REFERENTIAL :
                        caching == "tensor2"            ? B200MF_CACHE_TENSOR2 :
                        caching == "tensor4"            ? B200MF_CACHE_TENSOR4 :
                        caching == "tensor4_ns"         ? B200MF_CACHE_TENSOR4_NS :
                                                          -1;
      AssertThrow(level >= 0, dealii::ExcMessage("Unknown caching")); // :385
      check(b200mf_nh_initialize(op, level));
      staging.reinit(b200mf_ctx_n_dofs(ctx));
    }

    void
    cache() // :392-523; the debug assertion det F > 0 (:430-435) comes back as an error
    {
      std::int64_t n_bad = 0;
      for (unsigned int i = 0; i < displacement->locally_owned_size(); ++i) // the library updates the ghosts itself
        staging.src[i] = displacement->local_element(i);
      check(b200mf_nh_cache_host(op, staging.src, &n_bad));
    }
    void
    cache(const DeviceVector &displacement_on_device) // performance mode
    {
      std::int64_t n_bad = 0;
      check(b200mf_nh_cache(op, displacement_on_device.data(), &n_bad));
    }

    // performance mode: the same operator on device vectors (no host copy, asynchronous on the context's stream)
    void
    vmult(DeviceVector &dst, const DeviceVector &src) const
    {
      check(b200mf_nh_vmult(op, dst.data(), src.data()));
    }
    void
    vmult_add(DeviceVector &dst, const DeviceVector &src) const
    {
      check(b200mf_nh_vmult_add(op, dst.data(), src.data()));
    }
    void
    Tvmult(DeviceVector &dst, const DeviceVector &src) const
    {
      check(b200mf_nh_Tvmult(op, dst.data(), src.data()));
    }
    void
    precondition_Jacobi(DeviceVector &dst, const DeviceVector &src, const Number omega) const
    {
      check(b200mf_nh_precondition_jacobi(op, dst.data(), src.data(), omega));
    }

    void
    vmult(VectorType &dst, const VectorType &src) const // :543-551
    {
      host_apply(dst, src, false);
    }
    void
    Tvmult(VectorType &dst, const VectorType &src) const // :556-561, the operator is symmetric
    {
      host_apply(dst, src, false);
    }
    void
    vmult_add(VectorType &dst, const VectorType &src) const // :575-625
    {
      host_apply(dst, src, true);
    }
    void
    Tvmult_add(VectorType &dst, const VectorType &src) const
    {
      host_apply(dst, src, true);
    }

    void
    compute_diagonal() // :1023-1060
    {
      check(b200mf_nh_compute_diagonal(op));
      inverse_diagonal_entries = std::make_shared<dealii::DiagonalMatrix<VectorType>>();
      diagonal_entries         = std::make_shared<dealii::DiagonalMatrix<VectorType>>();
      data->initialize_dof_vector(inverse_diagonal_entries->get_vector());
      data->initialize_dof_vector(diagonal_entries->get_vector());
      check(b200mf_nh_diagonal_host(op, staging.src, staging.dst));
      for (unsigned int i = 0; i < data->get_vector_partitioner()->locally_owned_size(); ++i)
        {
          diagonal_entries->get_vector().local_element(i)         = Number(staging.src[i]);
          inverse_diagonal_entries->get_vector().local_element(i) = Number(staging.dst[i]);
        }
      diagonal_entries->get_vector().update_ghost_values(); // :1056-1057
      inverse_diagonal_entries->get_vector().update_ghost_values();
      diagonal_is_available = true;
    }

    void
    precondition_Jacobi(VectorType &dst, const VectorType &src, const Number omega) const // :281-291
    {
      Assert(inverse_diagonal_entries.get() && inverse_diagonal_entries->m() > 0, dealii::ExcNotInitialized());
      inverse_diagonal_entries->vmult(dst, src);
      dst *= omega;
    }

    Number
    el(const unsigned int row, const unsigned int col) const // :1064-1073
    {
      Assert(row == col, dealii::ExcNotImplemented());
      (void)col;
      double v = 0;
      check(b200mf_nh_el(op, row, &v));
      return Number(v);
    }

    size_type
    m() const // :295-302
    {
      return data->get_vector_partitioner()->size();
    }
    size_type
    n() const
    {
      return m();
    }

    const std::shared_ptr<dealii::DiagonalMatrix<VectorType>>
    get_matrix_diagonal_inverse() const // :157-161
    {
      return inverse_diagonal_entries;
    }

    std::size_t
    memory_consumption() const // :241-260
    {
      return b200mf_nh_memory_consumption(op) + b200mf_ctx_memory_consumption(ctx);
    }

    b200mf_nh *
    handle() const // performance mode: b200mf_cg_create_nh(handle()), b200mf_chebyshev_create_nh(handle(), ...)
    {
      return op;
    }

  private:
    void
    host_apply(VectorType &dst, const VectorType &src, const bool add) const
    {
      // vectors may carry another ghost set than the MatrixFree partitioner (adjust_ghost_range_if_necessary, :76-90):
      // the library only needs the owned entries and does its own ghost exchange
      const unsigned int n_owned = data->get_vector_partitioner()->locally_owned_size();
      for (unsigned int i = 0; i < n_owned; ++i)
        staging.src[i] = src.local_element(i);
      check(b200mf_nh_vmult_host(op, staging.dst, staging.src)); // pinned buffers: H2D || cells || D2H pipeline inside
      for (unsigned int i = 0; i < n_owned; ++i)
        dst.local_element(i) = (add ? dst.local_element(i) : Number(0)) + Number(staging.dst[i]);
    }

    std::shared_ptr<const dealii::MatrixFree<dim, Number>> data;
    const VectorType                                       *displacement = nullptr;
    b200mf_material                                         mat{}, incl{};
    b200mf_ctx                                             *ctx = nullptr;
    b200mf_nh                                              *op  = nullptr;
    mutable PinnedStaging                                   staging;
    std::shared_ptr<dealii::DiagonalMatrix<VectorType>>     inverse_diagonal_entries, diagonal_entries;
    bool                                                    diagonal_is_available = false;
  };


  // Drop-in for Heat_Transfer::LaplaceOperator<dim, fe_degree, number> (heat_transfer.h:81-127): (M + dt c K) u with the
  // MatrixFreeOperators::Base surface the heat solver, PreconditionChebyshev and mg::Matrix use.
  template <int dim, typename Number>
  class LaplaceOperator : public dealii::Subscriptor
  {
  public:
    using value_type = Number;
    using size_type  = dealii::types::global_dof_index;
    using VectorType = dealii::LinearAlgebra::distributed::Vector<Number>;

    ~LaplaceOperator()
    {
      clear();
    }

    void
    clear() // Base::clear + :129-137
    {
      b200mf_laplace_destroy(op);
      b200mf_ctx_destroy(ctx);
      op  = nullptr;
      ctx = nullptr;
      staging.release();
      inverse_diagonal_entries.reset();
      diagonal_entries.reset();
    }

    // Base::initialize(data) (+ level variant: the caller hands the level MatrixFree) and evaluate_coefficient (:150-166):
    // coefficient = [cell][q] values in FEEvaluation's cell/q order, or empty for Coefficient::value == 1 (:64-71)
    void
    initialize(std::shared_ptr<const dealii::MatrixFree<dim, Number>> data_,
               const double                                           delta_t_,
               const std::vector<double>                             &coefficient = {},
               b200mf_comm                                           *comm        = nullptr)
    {
      clear();
      data    = data_;
      delta_t = delta_t_;
      ctx     = make_context<dim, 1, Number>(*data, comm);
      check(b200mf_laplace_create(&op, ctx, delta_t, coefficient.empty() ? nullptr : coefficient.data()));
      staging.reinit(b200mf_ctx_n_dofs(ctx));
    }

    void
    set_delta_t(const double delta_t_) // :141-146
    {
      delta_t = delta_t_;
      check(b200mf_laplace_set_delta_t(op, delta_t));
    }

    void
    vmult(VectorType &dst, const VectorType &src) const // Base::vmult -> apply_add (:194-201)
    {
      host_apply(dst, src, false);
    }
    void
    Tvmult(VectorType &dst, const VectorType &src) const
    {
      host_apply(dst, src, false);
    }
    void
    vmult_add(VectorType &dst, const VectorType &src) const
    {
      host_apply(dst, src, true);
    }
    void
    Tvmult_add(VectorType &dst, const VectorType &src) const
    {
      host_apply(dst, src, true);
    }
    // mg::Matrix / MGInterfaceOperator: no refinement edges on globally refined meshes (heat_transfer.h:421-425) -> dst = 0
    void
    vmult_interface_down(VectorType &dst, const VectorType &) const
    {
      dst = Number(0);
    }
    void
    vmult_interface_up(VectorType &dst, const VectorType &) const
    {
      dst = Number(0);
    }

    // performance mode
    void
    vmult(DeviceVector &dst, const DeviceVector &src) const
    {
      check(b200mf_laplace_vmult(op, dst.data(), src.data()));
    }
    void
    vmult_add(DeviceVector &dst, const DeviceVector &src) const
    {
      check(b200mf_laplace_vmult_add(op, dst.data(), src.data()));
    }

    void
    compute_diagonal() // :205-230
    {
      check(b200mf_laplace_compute_diagonal(op));
      inverse_diagonal_entries = std::make_shared<dealii::DiagonalMatrix<VectorType>>();
      diagonal_entries         = std::make_shared<dealii::DiagonalMatrix<VectorType>>();
      data->initialize_dof_vector(inverse_diagonal_entries->get_vector());
      data->initialize_dof_vector(diagonal_entries->get_vector());
      check(b200mf_laplace_diagonal_host(op, staging.src, staging.dst));
      for (unsigned int i = 0; i < data->get_vector_partitioner()->locally_owned_size(); ++i)
        {
          diagonal_entries->get_vector().local_element(i)         = Number(staging.src[i]);
          inverse_diagonal_entries->get_vector().local_element(i) = Number(staging.dst[i]);
        }
    }

    const std::shared_ptr<dealii::DiagonalMatrix<VectorType>> &
    get_matrix_diagonal_inverse() const
    {
      return inverse_diagonal_entries;
    }
    const std::shared_ptr<dealii::DiagonalMatrix<VectorType>> &
    get_matrix_diagonal() const
    {
      return diagonal_entries;
    }

    void
    initialize_dof_vector(VectorType &vec) const // Base::initialize_dof_vector
    {
      data->initialize_dof_vector(vec);
    }
    std::shared_ptr<const dealii::MatrixFree<dim, Number>>
    get_matrix_free() const
    {
      return data;
    }

    Number
    el(const unsigned int row, const unsigned int col) const // Base::el: diagonal only
    {
      Assert(row == col, dealii::ExcNotImplemented());
      (void)col;
      double v = 0;
      check(b200mf_laplace_el(op, row, &v));
      return Number(v);
    }
    size_type
    m() const
    {
      return data->get_vector_partitioner()->size();
    }
    size_type
    n() const
    {
      return m();
    }

    b200mf_laplace *
    handle() const // b200mf_cg_create_laplace(handle()), b200mf_chebyshev_create_laplace(handle(), ...)
    {
      return op;
    }

  private:
    void
    host_apply(VectorType &dst, const VectorType &src, const bool add) const
    {
      const unsigned int n_owned = data->get_vector_partitioner()->locally_owned_size();
      for (unsigned int i = 0; i < n_owned; ++i)
        staging.src[i] = src.local_element(i);
      check(b200mf_laplace_vmult_host(op, staging.dst, staging.src));
      for (unsigned int i = 0; i < n_owned; ++i)
        dst.local_element(i) = (add ? dst.local_element(i) : Number(0)) + Number(staging.dst[i]);
    }

    std::shared_ptr<const dealii::MatrixFree<dim, Number>> data;
    double                                                  delta_t = 0;
    b200mf_ctx                                             *ctx     = nullptr;
    b200mf_laplace                                         *op      = nullptr;
    mutable PinnedStaging                                   staging;
    std::shared_ptr<dealii::DiagonalMatrix<VectorType>>     inverse_diagonal_entries, diagonal_entries;
  };

  // The boundary faces with a given id as a library face set (b200mf_face_set_create): cell = position of (batch, lane) in the
  // cell order of make_context, face number = deal.II's 2 * direction + side, Q1 vertices from the cell iterator (deal.II numbers
  // hexahedron vertices lexicographically, vertex m at (m & 1, m >> 1 & 1, m >> 2 & 1)).  Face order = (face batch, lane): the
  // order in which the reference walks them (mf_elasticity.h:1849-1897, heat_transfer.h:617-652, boundary_mass_operator.h:146-158).
  template <int dim, typename Number>
  b200mf_face_set *
  make_face_set(const dealii::MatrixFree<dim, Number> &mf, b200mf_ctx *ctx, const dealii::types::boundary_id interface_id)
  {
    static_assert(dim == 3, "the device kernels are 3-D");
    const unsigned int        n_lanes = dealii::VectorizedArray<Number>::size();
    std::vector<std::int64_t> compact(mf.n_cell_batches() * n_lanes, -1); // batch * n_lanes + lane -> cell of the context
    std::int64_t              running = 0;
    for (unsigned int b = 0; b < mf.n_cell_batches(); ++b)
      for (unsigned int l = 0; l < mf.n_active_entries_per_cell_batch(b); ++l)
        compact[b * n_lanes + l] = running++;
    std::vector<std::int64_t> cells;
    std::vector<std::int32_t> face_no;
    std::vector<double>       vertices;
    for (unsigned int face = mf.n_inner_face_batches(); face < mf.n_inner_face_batches() + mf.n_boundary_face_batches(); ++face)
      {
        if (mf.get_boundary_id(face) != interface_id)
          continue;
        const auto &info = mf.get_face_info(face);
        for (unsigned int l = 0; l < mf.n_active_entries_per_face_batch(face); ++l)
          {
            const unsigned int cell = info.cells_interior[l];
            cells.push_back(compact[cell]);
            face_no.push_back(info.interior_face_no);
            const auto it = mf.get_cell_iterator(cell / n_lanes, cell % n_lanes);
            for (unsigned int m = 0; m < 8; ++m)
              for (unsigned int e = 0; e < 3; ++e)
                vertices.push_back(it->vertex(m)[e]);
          }
      }
    b200mf_face_set *fs = nullptr;
    check(b200mf_face_set_create(&fs, ctx, cells.size(), cells.data(), face_no.data(), vertices.data()));
    return fs;
  }

  // Drop-in for ::BoundaryProjector<dim> (include/adapter/boundary_projector.h:16-110), used by
  // LaplaceProblem::evaluate_boundary_flux (heat_transfer.h:935-939): the boundary mass operator, its row-sum diagonal and the
  // Jacobi-preconditioned CG all run inside b200mf_boundary_mass_project.
  template <int dim>
  class BoundaryProjector
  {
  public:
    using VectorType = dealii::LinearAlgebra::distributed::Vector<double>;

    BoundaryProjector(const dealii::types::boundary_id                       interface_id,
                      std::shared_ptr<const dealii::MatrixFree<dim, double>> data_,
                      b200mf_comm                                           *comm = nullptr)
      : data(data_)
    {
      ctx = make_context<dim, 1, double>(*data, comm);
      fs  = make_face_set<dim, double>(*data, ctx, interface_id);
      check(b200mf_boundary_mass_create(&op, fs));
    }
    BoundaryProjector(const BoundaryProjector &) = delete;
    BoundaryProjector &
    operator=(const BoundaryProjector &) = delete;
    ~BoundaryProjector()
    {
      b200mf_boundary_mass_destroy(op);
      b200mf_face_set_destroy(fs);
      b200mf_ctx_destroy(ctx);
    }

    void
    project(VectorType &residual) // :86-110
    {
      DeviceVector        r(ctx);
      std::vector<double> scratch;
      r.upload(residual, scratch);
      project(r);
      r.download(residual, scratch);
      residual.update_ghost_values(); // :109
    }
    void
    project(DeviceVector &residual) // performance mode: the residual was assembled on the device
    {
      check(b200mf_boundary_mass_project(op, residual.data(), &last_iterations, nullptr));
    }
    int
    last_step() const
    {
      return last_iterations;
    }
    b200mf_ctx *
    context() const
    {
      return ctx;
    }

  private:
    std::shared_ptr<const dealii::MatrixFree<dim, double>> data;
    b200mf_ctx                                            *ctx = nullptr;
    b200mf_face_set                                       *fs  = nullptr;
    b200mf_boundary_mass                                  *op  = nullptr;
    int                                                    last_iterations = 0;
  };
} // namespace B200MF
