"""b200mf — B200-native matrix-free operator application (host-side mirror of
the reference's operator interface over libb200mf.so).  No CPU fallback."""
from .lib import B200mfError, load  # noqa: F401
from .operators import (BoundaryMassOperator, Communicator, DeviceVector, InterfaceTraction, LaplaceOperator, MatrixFree,  # noqa: F401
                        NeoHookOperator)
from .solvers import MGTransferMatrixFree, PreconditionChebyshev, PreconditionMG, SolverCG  # noqa: F401,E402
