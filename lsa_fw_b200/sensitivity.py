"""Post-processing of direct/adjoint eigenpairs the way the reference's sensitivity analysis consumes them
(`Sensitivity/__init__.py:171-203` direct mode, `:230-311` adjoint mode).  Host-side helpers on top of the
carriers; the heavy part (both eigensolves, the second one on the factors of the first) is the CUDA path.

SURVEY 8f row 2: mode selection, the bi-orthonormal scaling `a^H M v = 1`, and the eigenvalue sensitivity
`d lambda / d p = a^H (dA/dp - lambda dM/dp) v / (a^H M v)` as contractions with PRE-ASSEMBLED sparse derivative
operators on the device (`lsa_bilinear`), standing in for the UFL integrals of `Sensitivity/__init__.py:354-385`
(the explicit Reynolds term is `dA/dRe = -(1/Re^2) dA/d(1/Re)`, the viscous operator the assembler splits off).
"""

from __future__ import annotations

from typing import Iterable, Sequence

import numpy as np

from .carriers import iComplexPETScVector, iPETScMatrix, iPETScVector

__all__ = ["select_mode", "normalize_adjoint", "direct_and_adjoint_modes", "eigenvalue_sensitivity"]


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


def eigenvalue_sensitivity(solver, lam: complex, v, a, dA_values, *, dM_values=None, m_values=None) -> complex:
    """First-order change of the eigenvalue `lam` of `A v = lam M v` with left eigenvector `a` under a change of
    the operators given as VALUE arrays on the patterns of A and M (CSR entry order of the matrices the solver
    analysed):

        d lambda = a^H (dA - lam dM) v / (a^H M v)

    evaluated on the GPU that holds the pencil (`solver`: the `iEpsSolver` / `EigenSolver` that computed `v`).
    With `dA_values = -(1/Re^2) * pencil.a_visc` this is the explicit part of d sigma / d Re
    (`Sensitivity/__init__.py:372-374`); the implicit base-flow part is the same contraction with the operator
    `dA/dU . u_mu` assembled by the caller (`:376-381`).  `m_values`: values of M (default: the solver's own M)."""
    from . import _lib
    from .utils import _as_csr

    eps = getattr(solver, "solver", solver)
    h = eps.handle
    if h is None or h.closed:
        raise RuntimeError("the solver holds no device state; call solve() first")
    vc = _as_complex_vector(v).as_array().astype(np.complex128)
    ac = _as_complex_vector(a).as_array().astype(np.complex128)
    num = h.bilinear(_lib.LSA_MAT_A, np.asarray(dA_values), ac, vc)
    if dM_values is not None:
        num -= lam * h.bilinear(_lib.LSA_MAT_M, np.asarray(dM_values), ac, vc)
    if eps._M is not None:
        mv = _as_csr(eps._M).data if m_values is None else np.asarray(m_values)
        den = h.bilinear(_lib.LSA_MAT_M, mv, ac, vc)
    else:
        den = complex(np.vdot(ac, vc))
    if den == 0:
        raise RuntimeError("a^H M v = 0: the modes are not a direct / adjoint pair")
    return num / den
