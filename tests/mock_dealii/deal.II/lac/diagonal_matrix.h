#pragma once
#include <deal.II/base/mpi.h>
namespace dealii
{
  template <typename VectorType>
  class DiagonalMatrix
  {
  public:
    VectorType       &get_vector() { return vec; }
    const VectorType &get_vector() const { return vec; }
    void              vmult(VectorType &, const VectorType &) const {}
    unsigned int      m() const { return 0; }

  private:
    VectorType vec;
  };
} // namespace dealii
