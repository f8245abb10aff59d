// b200mf_dealii.h — the deal.II-side shim over the C ABI of libb200mf (include/b200mf.h).
//
// STATUS: deal.II (>= 9.6), p4est, MPI and preCICE are not installed in the image this library is built and tested in,
// so THIS HEADER HAS NEVER BEEN COMPILED.  It is the binding a maintainer of davidscn/matrix-free-dealii-precice adds on a
// machine that has deal.II; the identical call sequence is what tests/ exercise through the ctypes mirror
// (matrix-free-dealii-precice_b200/b200mf/operators.py).  It uses public deal.II API only.
//
// It provides, with the member set the reference's solvers call (include/solid_mechanics/mf_nh_operator.h:102-173):
//   B200MF::make_context(MatrixFree, comm)  -> b200mf_ctx*      arrays extracted once per (re)initialisation
//   B200MF::NeoHookOperator<dim, Number>                         drop-in for ::NeoHookOperator<dim, Number>
// LaplaceOperator (include/heat_transfer/heat_transfer.h:81-127) follows the same pattern on b200mf_laplace_*.
#pragma once
#include <deal.II/base/mpi.h>
#include <deal.II/lac/diagonal_matrix.h>
#include <deal.II/lac/la_parallel_vector.h>
#include <deal.II/matrix_free/fe_evaluation.h>
#include <deal.II/matrix_free/matrix_free.h>

#include <b200mf.h>

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


  // Drop-in for ::NeoHookOperator<dim, Number> (mf_nh_operator.h:99-237), parity mode: deal.II's SolverCG / Chebyshev /
  // Multigrid keep running on the host and call vmult with host vectors.  Performance mode replaces those solver objects by
  // b200mf_cg_solve / b200mf_mg_vmult on device vectors.
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
    }

    void
    cache() // :392-523; the debug assertion det F > 0 (:430-435) comes back as an error
    {
      std::int64_t     n_bad = 0;
      std::vector<double> u(displacement->locally_owned_size() + displacement->get_partitioner()->n_ghost_indices());
      for (unsigned int i = 0; i < u.size(); ++i)
        u[i] = displacement->local_element(i);
      check(b200mf_nh_cache_host(op, u.data(), &n_bad));
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
      const unsigned int  n_local = b200mf_ctx_n_dofs(ctx);
      std::vector<double> diag(n_local), inv(n_local);
      check(b200mf_nh_diagonal_host(op, diag.data(), inv.data()));
      for (unsigned int i = 0; i < data->get_vector_partitioner()->locally_owned_size(); ++i)
        {
          diagonal_entries->get_vector().local_element(i)         = diag[i];
          inverse_diagonal_entries->get_vector().local_element(i) = inv[i];
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
      const unsigned int  n_owned = data->get_vector_partitioner()->locally_owned_size(), n_local = b200mf_ctx_n_dofs(ctx);
      std::vector<double> s(n_local, 0.0), d(n_local, 0.0);
      for (unsigned int i = 0; i < n_owned; ++i)
        s[i] = src.local_element(i);
      check(b200mf_nh_vmult_host(op, d.data(), s.data()));
      for (unsigned int i = 0; i < n_owned; ++i)
        dst.local_element(i) = (add ? dst.local_element(i) : Number(0)) + Number(d[i]);
    }

    std::shared_ptr<const dealii::MatrixFree<dim, Number>> data;
    const VectorType                                       *displacement = nullptr;
    b200mf_material                                         mat{}, incl{};
    b200mf_ctx                                             *ctx = nullptr;
    b200mf_nh                                              *op  = nullptr;
    std::shared_ptr<dealii::DiagonalMatrix<VectorType>>     inverse_diagonal_entries, diagonal_entries;
    bool                                                    diagonal_is_available = false;
  };
} // namespace B200MF
