"""Direct linear solves on the LU kernels of the eigen path.

The reference's base-flow Newton iteration solves its Jacobian systems with `KSP gmres + PC lu (MUMPS)`
(`Solver/nonlinear2.py:61-70`), i.e. one sparse LU per Newton step on a matrix with the sparsity of A, real FP64 --
~95 % of its published pipeline time.  This module mirrors the linear-solver seam those callers use
(`Solver/utils.py:96-129` KSPType, `:331-420` iKSP; `Solver/linear.py:38-87` LinearSolver.solve) on top of the
SAME native handle as the eigensolver: one host symbolic analysis per sparsity pattern (reused by every Newton
step), numeric multifrontal LU + triangular sweeps on the GPU, optional iterative refinement.

Built: PREONLY + LU / CHOLESKY (one factor solve), and GMRES / RICHARDSON / BICGSTAB / CG *preconditioned by LU*
(what the reference's Newton solver configures): with an exact factorisation as preconditioner these converge in
one or two iterations, realised here as the factor solve followed by iterative refinement down to `rtol`.
Unpreconditioned Krylov methods and the other preconditioners are not built (NotImplementedError) -- never a
silent CPU fallback.
"""

from __future__ import annotations

import logging
import time
from enum import StrEnum, auto

import numpy as np
import scipy.sparse as sp

from . import _lib
from .carriers import iPETScMatrix, iPETScVector
from .utils import PreconditionerType, _SYM_CACHE, _as_csr, _raw_csr

logger = logging.getLogger(__name__)

__all__ = ["KSPType", "iKSP", "LinearSolver"]


class KSPType(StrEnum):
    """KSP solver types (reference `Solver/utils.py:96-129`, same member names)."""

    CG = auto()
    GMRES = auto()
    BICG = auto()
    BICGSTAB = auto()
    RICHARDSON = auto()
    CHEBYSHEV = auto()
    PREONLY = auto()
    QCG = auto()
    CGS = auto()
    GCR = auto()
    LSQR = auto()
    LGMRES = auto()
    FGMRES = auto()

    def to_petsc(self) -> str:
        return self.value


class _RawKSPShim:
    def __init__(self, owner: "iKSP") -> None:
        self._o = owner

    def getType(self) -> str:  # noqa: N802
        return self._o.get_type()

    def getIterationNumber(self) -> int:  # noqa: N802
        return self._o.get_iteration_number()

    def getResidualNorm(self) -> float:  # noqa: N802
        return self._o.get_residual_norm()

    def setGMRESRestart(self, *_):  # noqa: N802  (Solver/linear.py:71 calls it; nothing to restart here)
        return None


class iKSP:  # noqa: N801
    """Linear solver object with the interface of the reference's KSP wrapper (`Solver/utils.py:331-420`),
    executing on one B200 through liblsa_b200.so."""

    def __init__(self, A: iPETScMatrix | None = None, comm=None) -> None:
        self._A: iPETScMatrix | None = None
        self._type = KSPType.GMRES
        self._pc = PreconditionerType.LU
        self._atol, self._rtol, self._max_it = 1e-12, 1e-8, 1000
        self._handle: _lib.Handle | None = None
        self._factor_gen = -1
        self._values_ref = None
        self._sym_failed = False
        self._its = 0
        self._rnorm = float("nan")
        self._sol: iPETScVector | None = None
        self._opts = dict(device=0, leaf_size=64, tiny_pivot=0.0, nthreads=0, coords=None)
        self.stats: dict = {}
        if A is not None:
            self.set_operators(A)

    @property
    def raw(self) -> _RawKSPShim:
        return _RawKSPShim(self)

    def set_operators(self, A: iPETScMatrix, P: iPETScMatrix | None = None) -> None:
        """Set the system matrix (a separate preconditioning matrix is not supported: the LU is of A itself)."""
        if P is not None and P is not A:
            raise NotImplementedError("a preconditioning matrix different from A is not built on the B200 backend")
        self._A = A
        self._factor_gen = -1
        self._sym_failed = False

    def set_type(self, ksp_type: KSPType) -> None:
        self._type = KSPType(ksp_type)

    def get_type(self) -> str:
        return self._type.to_petsc()

    def set_tolerances(self, tol: float = 1e-12, max_it: int = 1000, rtol: float = 1e-8) -> None:
        self._atol, self._rtol, self._max_it = float(tol), float(rtol), int(max_it)

    def set_preconditioner(self, pc_type: PreconditionerType) -> None:
        self._pc = PreconditionerType(pc_type)

    def set_initial_guess_nonzero(self, flag: bool) -> None:  # a direct solve ignores the guess
        self._nonzero_guess = bool(flag)

    def set_from_options(self, prefix: str | None = None) -> None:
        return None

    def set_backend_options(self, **kw) -> None:
        unknown = set(kw) - set(self._opts)
        if unknown:
            raise TypeError(f"unknown backend option(s): {sorted(unknown)}")
        self._opts.update(kw)

    # ------------------------------------------------------------------ numeric path
    def _ensure_factor(self) -> tuple["_lib.Handle", sp.csr_matrix, bool]:
        if self._A is None:
            raise ValueError("Operators must be set before solve().")
        if self._pc not in (PreconditionerType.LU, PreconditionerType.CHOLESKY):
            raise NotImplementedError(f"preconditioner '{self._pc}' is not built on the B200 backend (direct LU only)")
        if self._type in (KSPType.CHEBYSHEV, KSPType.QCG, KSPType.LSQR):
            raise NotImplementedError(f"KSP type '{self._type}' is not built on the B200 backend")
        t0 = time.perf_counter()
        coords = self._opts["coords"]
        A = _raw_csr(self._A)
        # PCCHOLESKY on real data: symmetric L D L^T at half the factor storage (falls back to LU on a vanishing pivot)
        use_sym = (self._pc is PreconditionerType.CHOLESKY and not self._sym_failed
                   and not np.iscomplexobj((A if A is not None else _as_csr(self._A)).data))
        extra = ("linear", use_sym, self._opts["leaf_size"], self._opts["device"],
                 None if coords is None else np.ascontiguousarray(coords).tobytes()[:64])
        h = _SYM_CACHE.lookup(A, None, extra) if A is not None and len(_SYM_CACHE) else None
        if h is None:
            A = _as_csr(self._A)
            h = _SYM_CACHE.lookup(A, None, extra)
        n = A.shape[0]
        if A.shape[0] != A.shape[1]:
            raise ValueError("Matrix must be square.")
        cplx = np.iscomplexobj(A.data)
        if h is None:
            h = _lib.Handle(n, self._opts["device"])
            if use_sym:
                h.set_option("symmetric", 1)
            h.analyze(A.indptr, A.indices, None, None, leaf_size=self._opts["leaf_size"], coords=coords,
                      order_last=(A.diagonal() == 0).astype(np.uint8), nthreads=self._opts["nthreads"])
            _SYM_CACHE.store(A, None, extra, h)
            self.stats["symbolic_cached"] = False
        else:
            self.stats["symbolic_cached"] = True
        self.stats["symbolic_seconds"] = time.perf_counter() - t0
        if h is not self._handle or h.gen_factor != self._factor_gen or self._values_ref is not A.data:
            h.set_values(A.data, None)
            fs = h.factor(1.0, 0.0, _lib.LSA_C128 if cplx else _lib.LSA_F64, self._opts["tiny_pivot"])
            if use_sym and (fs.n_perturbed > 0 or fs.max_multiplier > 1e8):
                self._sym_failed = True
                return self._ensure_factor()
            self._factor_gen = h.gen_factor
            self._values_ref = A.data
            self.stats.update(symmetric_factorisation=use_sym, factor_seconds=fs.seconds, factor_flops=fs.flops, n_perturbed=int(fs.n_perturbed),
                              max_multiplier=fs.max_multiplier, scalar="c128" if cplx else "f64")
        self._handle = h
        return h, A, cplx

    def solve(self, b: iPETScVector, x: iPETScVector | None = None) -> iPETScVector:
        """Solve A x = b.  PREONLY: one factor solve.  Krylov types (preconditioned by the LU): factor solve, then
        refinement steps until `||b - A x|| <= max(rtol ||b||, atol)` or `max_it`."""
        h, A, cplx = self._ensure_factor()
        rhs = np.asarray(b.raw.getArray() if hasattr(b, "raw") else b)
        bc = np.ascontiguousarray(rhs, dtype=np.complex128)
        t0 = time.perf_counter()
        xs = h.solve(bc)
        its, bnorm = 1, float(np.linalg.norm(bc))
        rn = float(np.linalg.norm(bc - h.spmv(_lib.LSA_MAT_A, xs)))
        if self._type is not KSPType.PREONLY:
            target = max(self._rtol * bnorm, self._atol)
            while rn > target and its < max(1, self._max_it):
                r = bc - h.spmv(_lib.LSA_MAT_A, xs)
                dx = h.solve(r)
                xs = xs + dx
                rn_new = float(np.linalg.norm(bc - h.spmv(_lib.LSA_MAT_A, xs)))
                its += 1
                if not rn_new < rn:      # stagnation at the level of the factorisation's accuracy
                    rn = min(rn, rn_new)
                    break
                rn = rn_new
        self.stats["solve_seconds"] = time.perf_counter() - t0
        self._its, self._rnorm = its, rn
        out = xs if (cplx or np.iscomplexobj(rhs)) else np.ascontiguousarray(xs.real)
        if x is None:
            x = iPETScVector.from_array(out)
        else:
            arr = x.raw.getArray()
            if np.iscomplexobj(out) and not np.iscomplexobj(arr):
                raise TypeError("complex solution cannot be written into a real vector")
            arr[...] = out
        self._sol = x
        return x

    def get_solution(self) -> iPETScVector:
        if self._sol is None:
            raise RuntimeError("solve() has not been called")
        return self._sol

    def get_iteration_number(self) -> int:
        return int(self._its)

    def reset(self) -> None:
        """Forget the solver state for fresh reuse (`Solver/utils.py:417-419`, `KSPReset`): the next solve factors
        again; the symbolic analysis of the pattern stays in the cache."""
        self._factor_gen = -1
        self._values_ref = None
        self._sym_failed = False
        self._its, self._rnorm, self._sol = 0, float("nan"), None

    def get_residual_norm(self) -> float:
        return float(self._rnorm)


class LinearSolver:
    """Assembler-free part of the reference's `LinearSolver` (`Solver/linear.py:38-87`)."""

    @staticmethod
    def solve(A: iPETScMatrix, b: iPETScVector, *, ksp_type: KSPType, tol: float = 1e-12, rtol: float = 1e-8,
              max_it: int = 1_000, backend_options: dict | None = None) -> iPETScVector:
        """Solve A x = b with `KSPType.PREONLY` (direct LU) or `KSPType.GMRES`.  The reference runs its static GMRES
        without a preconditioner; here GMRES is preconditioned by the LU of A (as its Newton solver configures it,
        `Solver/nonlinear2.py:61-70`), which reaches the same tolerances in a couple of iterations."""
        if ksp_type not in (KSPType.PREONLY, KSPType.GMRES):
            raise ValueError("KSP type not supported.")
        solver = iKSP(A)
        solver.set_type(ksp_type)
        solver.set_preconditioner(PreconditionerType.LU)
        if ksp_type is not KSPType.PREONLY:
            solver.set_tolerances(tol=tol, rtol=rtol, max_it=max_it)
        if backend_options:
            solver.set_backend_options(**backend_options)
        t0 = time.perf_counter()
        sol = solver.solve(b)
        logger.info("%s solve time: %.3f s", ksp_type.name, time.perf_counter() - t0)
        return sol
