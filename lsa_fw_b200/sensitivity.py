"""Post-processing of direct/adjoint eigenpairs the way the reference's sensitivity analysis consumes them
(`Sensitivity/__init__.py:171-203` direct mode, `:230-311` adjoint mode).  Host-side helpers on top of the
carriers; the heavy part (both eigensolves, the second one on the factors of the first) is the CUDA path.

SURVEY 8f row 2 ("next"): only the mode selection and the bi-orthonormal scaling `a^H M v = 1` live here; the
derivative contraction `a^H (dA/dRe) v` needs the reference's UFL forms and stays out of scope.
"""

from __future__ import annotations

from typing import Iterable, Sequence

import numpy as np

from .carriers import iComplexPETScVector, iPETScMatrix, iPETScVector

__all__ = ["select_mode", "normalize_adjoint", "direct_and_adjoint_modes"]


def select_mode(pairs: Iterable[tuple[complex, iComplexPETScVector]], target: complex):
    """The pair whose eigenvalue is closest to `target` (`Sensitivity/__init__.py:190-194, 277-278`:
    `min(pairs, key=|lambda - target|)`; the adjoint solve passes `conj(sigma)`)."""
    pairs = list(pairs)
    if not pairs:
        raise RuntimeError("No eigenpairs returned by the eigensolver.")
    return min(pairs, key=lambda p: abs(p[0] - target))


def _as_complex_vector(v) -> iComplexPETScVector:
    if isinstance(v, iComplexPETScVector):
        return v
    if isinstance(v, iPETScVector):
        return iComplexPETScVector(v)
    return iComplexPETScVector.from_array(np.asarray(v))


def normalize_adjoint(a_vec: iComplexPETScVector, M: iPETScMatrix, v) -> complex:
    """Scale the adjoint eigenvector in place so that `a^H M v = 1` (`Sensitivity/__init__.py:280-287`):
    `prod = a.dot(M v)` with the carriers' conjugating dot, `a.scale(1 / prod)`.  Returns `prod`.

    `v` may be a carrier or a plain array (the reference passes the array of a dolfinx Function)."""
    vc = _as_complex_vector(v)
    mv = M.as_scipy_array() @ vc.as_array()
    complex_build = a_vec.imag is None and np.iscomplexobj(a_vec.real.raw.getArray())
    if complex_build:
        # one complex vector, VecDot conjugates the argument: prod = (M v)^H a, so that a / prod has a^H M v = 1
        Mv = iComplexPETScVector(iPETScVector.from_array(np.asarray(mv, dtype=np.complex128)))
    else:
        # real build: (real, imag) pair, the dot conjugates `self` (`FEM/utils.py:1208-1212`); as in the reference
        # the result then satisfies |a^H M v| = 1
        Mv = iComplexPETScVector.from_array(mv)
    prod = a_vec.dot(Mv)
    if prod == 0:
        raise RuntimeError("Bi-orthonormal normalization failed (a^H B v = 0).")
    a_vec.scale(1.0 / prod)
    return prod


def direct_and_adjoint_modes(A: iPETScMatrix, M: iPETScMatrix, sigma: complex, cfg=None, *, adjoint_cfg=None,
                             backend_options: dict | None = None):
    """Direct mode near `sigma` and the matching adjoint mode, bi-orthonormalised: the sequence of
    `Sensitivity.solve_direct_mode` + `solve_adjoint_mode`, with the adjoint eigensolve running conjugate-
    transposed sweeps on the factorisation of the direct one (one LU instead of two).

    Returns `((lambda, v), (lambda_adj, a))`."""
    from .eigen import EigenSolver, EigensolverConfig
    from .utils import PreconditionerType, iEpsWhich, iSTType

    cfg = cfg or EigensolverConfig()
    adjoint_cfg = adjoint_cfg or cfg

    def run(mat_a, mat_m, shift, conf, which):
        es = EigenSolver(mat_a, mat_m, conf, check_hermitian=False)
        es.solver.set_st_type(iSTType.SINVERT)
        es.solver.set_st_pc_type(PreconditionerType.LU)
        es.solver.set_target(shift)
        es.solver.set_which_eigenpairs(which)
        if backend_options:
            es.solver.set_backend_options(**backend_options)
        return es, es.solve()

    es_d, pairs = run(A, M, sigma, cfg, iEpsWhich.TARGET_MAGNITUDE)
    lam, v = select_mode(pairs, sigma)
    es_a, pairs_adj = run(A.H, M.H, np.conj(sigma), adjoint_cfg, iEpsWhich.TARGET_REAL)
    lam_adj, a = select_mode(pairs_adj, np.conj(sigma))
    normalize_adjoint(a, M, v)
    return (lam, v), (lam_adj, a)
