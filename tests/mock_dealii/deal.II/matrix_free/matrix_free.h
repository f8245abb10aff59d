#pragma once
#include <array>
#include <deal.II/base/mpi.h>
#include <deal.II/lac/la_parallel_vector.h>
namespace dealii
{
  template <typename Number>
  class VectorizedArray
  {
  public:
    static constexpr std::size_t size() { return 8; }
    Number operator[](const unsigned int) const { return Number(); }
  };
  template <int rank, int dim, typename Number>
  struct Tensor
  {
    struct Row
    {
      Number operator[](const unsigned int) const { return Number(); }
    };
    Row operator[](const unsigned int) const { return Row(); }
  };
  struct FiniteElementData
  {
    unsigned int degree = 1;
  };
  struct DoFHandlerMock
  {
    const FiniteElementData &get_fe() const { return fe; }
    FiniteElementData        fe;
  };
  namespace types
  {
    using boundary_id = unsigned int;
  }
  struct PointMock
  {
    double operator[](const unsigned int) const { return 0.0; }
  };
  struct CellMock
  {
    unsigned int material_id() const { return 0; }
    PointMock    vertex(const unsigned int) const { return PointMock(); }
    const CellMock *operator->() const { return this; }
  };
  struct FaceToCellTopologyMock // internal::MatrixFreeFunctions::FaceToCellTopology
  {
    std::array<unsigned int, 8> cells_interior{};
    unsigned char               interior_face_no = 0;
  };
  template <int dim, typename Number>
  class MatrixFree
  {
  public:
    const DoFHandlerMock &get_dof_handler() const { return dh; }
    template <typename VectorType>
    void initialize_dof_vector(VectorType &) const {}
    const std::shared_ptr<const Utilities::MPI::Partitioner> &get_vector_partitioner() const { return part; }
    unsigned int n_cell_batches() const { return 0; }
    unsigned int n_active_entries_per_cell_batch(const unsigned int) const { return 0; }
    CellMock     get_cell_iterator(const unsigned int, const unsigned int) const { return CellMock(); }
    const std::vector<unsigned int> &get_constrained_dofs() const { return con; }
    unsigned int n_inner_face_batches() const { return 0; }
    unsigned int n_boundary_face_batches() const { return 0; }
    unsigned int n_active_entries_per_face_batch(const unsigned int) const { return 0; }
    types::boundary_id get_boundary_id(const unsigned int) const { return 0; }
    const FaceToCellTopologyMock &get_face_info(const unsigned int) const { return fi; }

  private:
    DoFHandlerMock                                     dh;
    std::shared_ptr<const Utilities::MPI::Partitioner> part;
    std::vector<unsigned int>                          con;
    FaceToCellTopologyMock                             fi;
  };
} // namespace dealii
