// tests/mock_dealii: a minimal stand-in for the deal.II (>= 9.6) declarations contrib/dealii/b200mf_dealii.h uses, with
// deal.II's documented signatures, so that the shim can be syntax- and type-checked where deal.II is not installed
// (tests/test_shim_syntax.py).  Nothing here is functional.
#pragma once
#include <cstddef>
#include <cstdint>
#include <memory>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

using MPI_Comm = int;
#define MPI_BYTE 1
inline int MPI_Bcast(void *, int, int, int, MPI_Comm) { return 0; }

namespace dealii
{
  namespace types
  {
    using global_dof_index = unsigned int;
  }
  class Subscriptor
  {};
  struct ExcMessage
  {
    explicit ExcMessage(const std::string &s)
      : msg(s)
    {}
    std::string msg;
  };
  struct ExcNotImplemented
  {};
  struct ExcNotInitialized
  {};
  namespace Utilities
  {
    namespace MPI
    {
      inline unsigned int this_mpi_process(const MPI_Comm) { return 0; }
      inline unsigned int n_mpi_processes(const MPI_Comm) { return 1; }
      class Partitioner
      {
      public:
        unsigned int locally_owned_size() const { return 0; }
        unsigned int n_ghost_indices() const { return 0; }
        unsigned int size() const { return 0; }
        const std::vector<std::pair<unsigned int, unsigned int>> &import_indices() const { return pairs; }
        const std::vector<std::pair<unsigned int, unsigned int>> &import_targets() const { return pairs; }
        const std::vector<std::pair<unsigned int, unsigned int>> &ghost_targets() const { return pairs; }

      private:
        std::vector<std::pair<unsigned int, unsigned int>> pairs;
      };
    } // namespace MPI
  }   // namespace Utilities
} // namespace dealii
#define AssertThrow(cond, exc)                     \
  do                                               \
    {                                              \
      if (!(cond))                                 \
        {                                          \
          auto e_ = exc;                           \
          (void)e_;                                \
          throw std::runtime_error("AssertThrow"); \
        }                                          \
    }                                              \
  while (0)
#define Assert(cond, exc) AssertThrow(cond, exc)
