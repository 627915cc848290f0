"""lsa_fw_b200 -- B200-native shift-and-invert eigensolve backend with the API of LSA-FW's Solver/eigen.py."""

from .carriers import iComplexPETScVector, iPETScBlockMatrix, iPETScMatrix, iPETScNullSpace, iPETScVector
from .eigen import EigenSolver, EigensolverConfig
from .linear import KSPType, LinearSolver, iKSP
from .sensitivity import direct_and_adjoint_modes, eigenvalue_sensitivity, normalize_adjoint, select_mode
from .utils import (LsaError, PreconditionerType, clear_symbolic_cache, iEpsProblemType, iEpsSolver, iEpsWhich,
                    iSTType)

__all__ = [
    "EigenSolver", "EigensolverConfig", "iEpsSolver", "iEpsProblemType", "iEpsWhich", "iSTType",
    "PreconditionerType", "iPETScMatrix", "iPETScVector", "iComplexPETScVector", "iPETScNullSpace", "iPETScBlockMatrix", "LsaError",
    "clear_symbolic_cache", "select_mode", "normalize_adjoint", "direct_and_adjoint_modes", "eigenvalue_sensitivity",
    "KSPType", "iKSP", "LinearSolver",
]
