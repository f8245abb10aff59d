#pragma once
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
  struct CellMock
  {
    unsigned int material_id() const { return 0; }
    const CellMock *operator->() const { return this; }
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

  private:
    DoFHandlerMock                                     dh;
    std::shared_ptr<const Utilities::MPI::Partitioner> part;
    std::vector<unsigned int>                          con;
  };
} // namespace dealii
