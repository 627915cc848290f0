"""LSA-FW Eigensolver, B200 backend.

Drop-in for the reference's `Solver/eigen.py` (`:48-61` EigensolverConfig, `:67-155` EigenSolver):
same dataclass fields and defaults, same constructor validation and warnings, same `solve()` return
value `list[(eigenvalue, iComplexPETScVector)]` of length `min(nconv, num_eig)`, same three log
lines.  The numerical work behind `self._solver.solve()` runs as hand-written sm_100a CUDA
(multifrontal LU of A - sigma M, supernodal triangular solves, SpMV, CGS2 Krylov-Schur) through the
C ABI of `include/lsa_b200.h`.

Example (shift-and-invert, as `.examples/eigenvalues.py:95-102` of the reference):

    cfg = EigensolverConfig(num_eig=10, atol=1e-8)
    es = EigenSolver(A, M, cfg, check_hermitian=False)
    es.solver.set_st_type(iSTType.SINVERT)
    es.solver.set_target(0.05 + 0.74j)
    es.solver.set_st_pc_type(PreconditionerType.LU)
    pairs = es.solve()
"""

from __future__ import annotations

import logging
import time
from dataclasses import dataclass

from .carriers import iComplexPETScVector, iPETScMatrix
from .utils import iEpsProblemType, iEpsSolver

logger = logging.getLogger(__name__)


def log_global(lg: logging.Logger, level: int, msg: str, *args) -> None:
    """Rank-0 logging (reference `lib/loggingutils.py:81-84`); one process per GPU here."""
    import os

    if int(os.environ.get("RANK", "0")) == 0:
        lg.log(level, msg, *args)


_HERMITIAN_TYPES: set[iEpsProblemType] = {
    iEpsProblemType.HEP,
    iEpsProblemType.GHEP,
    iEpsProblemType.GHIEP,
}


@dataclass(frozen=True)
class EigensolverConfig:
    """Eigensolver configuration (reference `Solver/eigen.py:48-61`)."""

    num_eig: int = 5
    """Number of computed eigenpairs."""
    problem_type: iEpsProblemType = iEpsProblemType.GNHEP
    """Problem type."""
    atol: float = 1e-6
    """Tolerance (handed to the solver as its RELATIVE tolerance, as the reference does)."""
    max_it: int = 500
    """Maximum number of iterations (restarts)."""
    ncv: int = 80
    """Subspace dimension."""


class EigenSolver:
    """Solver for the generalized eigenvalue problem Ax = λMx on one B200."""

    def __init__(self, *args, check_hermitian: bool = True, **kwargs) -> None:
        """Initialize eigensolver: `EigenSolver(A, M=None, cfg=None, *, check_hermitian=True)`.

        Both argument orders found in the reference are accepted: the current
        `EigenSolver(A, M, cfg)` (`Solver/eigen.py:67-74`) and the older `EigenSolver(cfg, A=..., M=...)`
        still used by its tests, CLI and docs (`tests/unit/Solver/test_eigen.py:91`, `Solver/cli.py:168`).
        """
        A, M, cfg = self._parse_arguments(args, kwargs)
        self._cfg = cfg or EigensolverConfig()

        nrows, ncols = A.shape
        if nrows != ncols:
            raise ValueError(f"Operator A must be square, got shape ({nrows}, {ncols})")

        if M is not None:
            mrows, mcols = M.shape
            if (mrows, mcols) != (nrows, ncols):
                raise ValueError(f"Operator M shape {M.shape} does not match A's shape {A.shape}")
        if self._cfg.problem_type in _HERMITIAN_TYPES and check_hermitian:
            if not A.is_numerically_hermitian():
                log_global(
                    logger,
                    logging.WARNING,
                    f"Problem type '{self._cfg.problem_type.name}' assumes Hermitian A,"
                    " but A is not (numerically) Hermitian.",
                )
            if (
                M is not None
                and self._cfg.problem_type in {iEpsProblemType.GHEP, iEpsProblemType.GHIEP}
                and not M.is_numerically_hermitian()
            ):
                log_global(
                    logger,
                    logging.WARNING,
                    f"Problem type '{self._cfg.problem_type.name}' assumes Hermitian M,"
                    " but M is not (numerically) Hermitian.",
                )

        self._solver = iEpsSolver(A, M)
        self._solver.set_problem_type(self._cfg.problem_type)
        self._solver.set_tolerances(self._cfg.atol, self._cfg.max_it)
        self._solver.set_dimensions(self._cfg.num_eig, self._cfg.ncv)

    @staticmethod
    def _parse_arguments(args: tuple, kwargs: dict):
        unknown = set(kwargs) - {"A", "M", "cfg"}
        if unknown:
            raise TypeError(f"EigenSolver() got unexpected keyword argument(s) {sorted(unknown)}")
        if len(args) > 3:
            raise TypeError("EigenSolver() takes at most 3 positional arguments")
        if args and isinstance(args[0], EigensolverConfig):
            names = ("cfg", "A", "M")  # legacy order
        else:
            names = ("A", "M", "cfg")
        bound = dict(kwargs)
        for name, value in zip(names, args):
            if name in bound:
                raise TypeError(f"EigenSolver() got multiple values for argument '{name}'")
            bound[name] = value
        A, M, cfg = bound.get("A"), bound.get("M"), bound.get("cfg")
        if A is None:
            raise TypeError("EigenSolver() missing the operator A")
        if cfg is not None and not isinstance(cfg, EigensolverConfig):
            raise TypeError("cfg must be an EigensolverConfig")
        return A, M, cfg

    @property
    def solver(self) -> iEpsSolver:
        """Get the solver object."""
        return self._solver

    @property
    def config(self) -> EigensolverConfig:
        """Get the solver configuration."""
        return self._cfg

    def solve(self) -> list[tuple[float | complex, iComplexPETScVector]]:
        """Run the solver and return eigenpairs."""
        log_global(
            logger,
            logging.INFO,
            f"Started eigenvalue solve: type={self._cfg.problem_type.name}, "
            f"nev={self._cfg.num_eig}, "
            f"tol={self._cfg.atol}, max_it={self._cfg.max_it}",
        )

        t0 = time.time()
        self._solver.solve()
        elapsed = time.time() - t0

        nconv = self._solver.get_num_converged()
        try:
            its = self._solver.raw.getST().getKSP().getIterationNumber()
        except Exception:
            its = None

        log_global(
            logger,
            logging.INFO,
            f"Solve completed in {elapsed:.2f} s; converged {nconv} eigenpairs"
            + (f"; iterations={its}" if its is not None else ""),
        )

        pairs = list(self._solver.get_all_eigenpairs_up_to(self._cfg.num_eig))
        log_global(logger, logging.INFO, f"Retrieved {len(pairs)} eigenpairs")
        return pairs
