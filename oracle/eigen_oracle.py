"""CPU oracle for the shift-and-invert generalized eigensolve  A x = lambda M x.

TEST INFRASTRUCTURE ONLY.  Nothing under `lsa_fw_b200/` may import this module; it is used
by `tests/`, by `__graft_entry__.smoke()` and by `bench.py`'s cpu_baseline / `--impl
reference` legs as the checker and the timed CPU baseline, never as the product path.

What it restates (reference = ferdean/lsa-fw, mounted at /root/reference when authoring):

* `Solver/eigen2.py:109-111`   C = A - sigma M
* `Solver/eigen2.py:120-151`   direct LU of C  (PETSc LU/MUMPS there, SciPy SuperLU here; SciPy's
                               `splu` is itself a backend the reference uses, `Solver/linear.py:148-155`)
* `Solver/eigen2.py:164-201`   OP(x) = C^-1 (M x), optional zeroing of the pressure DOFs
* `Solver/eigen2.py:224-242`   ARPACK `eigs(OP, which="LM")`, lambda = sigma + 1/mu, sort
* `Solver/eigen2.py:48-56`     relative residual definition
* `Sensitivity/__init__.py:246-287`  adjoint modes: the same solve on (A^H, M^H) at conj(sigma);
                               here realised with `lu.solve(., trans="H")` on the SAME factors.
* `Solver/utils.py:152-187`    `which` selection semantics of `iEpsWhich`

PARITY STATUS: the arithmetic of the reference path lives in SLEPc 3.22 / PETSc 3.22 (+MUMPS),
which are not vendored in the reference and not installable here, so this oracle is pinned
against (a) every known-answer case in `tests/unit/Solver/test_eigen.py:107-304`
(tests/golden/kat_eigen.json), (b) the analytic membrane spectrum of
`tests/benchmark/vibrating_membrane.py:130-141`, and (c) dense `scipy.linalg.eig` on small
pencils.  For the actual linearised-NS pencils no golden eigenvalues exist in the reference
("matrices are not included in the repo", `.examples/eigenvalues.py:6-8`):  PARITY UNPINNED
there; agreement is oracle-vs-device plus residuals.
"""

from __future__ import annotations

import time
from dataclasses import dataclass, field

import numpy as np
import scipy.linalg as sla
import scipy.sparse as sp
import scipy.sparse.linalg as spla

__all__ = [
    "OracleResult",
    "shift_invert_arpack",
    "krylov_schur",
    "shift_invert_krylov_schur",
    "dense_pencil_eigs",
    "eigen2_residuals",
    "north_star_residuals",
    "which_key",
]


@dataclass
class OracleResult:
    eigenvalues: np.ndarray
    eigenvectors: np.ndarray
    residuals: np.ndarray
    n_op_applies: int = 0
    n_restarts: int = 0
    seconds: dict = field(default_factory=dict)
    nnz_lu: int = 0


# ----------------------------------------------------------------------------- residuals


def eigen2_residuals(A, M, lam: np.ndarray, V: np.ndarray) -> np.ndarray:
    """||A v - lam M v|| / (||A v|| + |lam| ||M v|| + 1e-16)   (`Solver/eigen2.py:48-56`)."""
    Av = A @ V
    Mv = M @ V
    R = Av - Mv * lam[np.newaxis, :]
    num = np.linalg.norm(R, axis=0)
    den = np.linalg.norm(Av, axis=0) + np.abs(lam) * np.linalg.norm(Mv, axis=0) + 1e-16
    return num / den


def north_star_residuals(A, M, lam: np.ndarray, V: np.ndarray) -> np.ndarray:
    """||A x - lam M x||_2 / (||A||_F ||x||_2)   (BASELINE.json north_star acceptance bar)."""
    R = A @ V - (M @ V) * lam[np.newaxis, :]
    return np.linalg.norm(R, axis=0) / (spla.norm(A) * np.linalg.norm(V, axis=0))


# ----------------------------------------------------------------------------- selection


def which_key(which: str, lam: np.ndarray, target: complex = 0.0) -> np.ndarray:
    """Sort key (ascending = preferred first) for SLEPc's `EPSWhich` applied to lambda.

    Names follow `Solver/utils.py:152-187`.  Note the reference's alias quirk: its
    `SMALLEST_MAGNITUDE` member has the value of `LARGEST_REAL` (`Solver/utils.py:157-158`).
    """
    lam = np.asarray(lam, dtype=complex)
    w = which.upper()
    if w == "LARGEST_MAGNITUDE":
        return -np.abs(lam)
    if w == "SMALLEST_MAGNITUDE_TRUE":
        return np.abs(lam)
    if w in ("LARGEST_REAL", "SMALLEST_MAGNITUDE"):
        return -lam.real
    if w == "SMALLEST_REAL":
        return lam.real
    if w == "LARGEST_IMAGINARY":
        return -lam.imag
    if w == "SMALLEST_IMAGINARY":
        return lam.imag
    if w == "TARGET_MAGNITUDE":
        return np.abs(lam - target)
    if w == "TARGET_REAL":
        return np.abs(lam.real - np.real(target))
    if w == "TARGET_IMAGINARY":
        return np.abs(lam.imag - np.imag(target))
    raise ValueError(f"unsupported which = {which!r}")


# ----------------------------------------------------------------------------- LU helpers


def _shifted(A, M, sigma, force_complex: bool):
    is_complex = force_complex or np.iscomplexobj(sigma) and complex(sigma).imag != 0.0
    is_complex = is_complex or np.iscomplexobj(A.data) or (M is not None and np.iscomplexobj(M.data))
    dt = np.complex128 if is_complex else np.float64
    A = sp.csc_matrix(A, dtype=dt)
    if M is None:
        M = sp.identity(A.shape[0], dtype=dt, format="csc")
    else:
        M = sp.csc_matrix(M, dtype=dt)
    s = complex(sigma) if is_complex else float(np.real(sigma))
    C = (A - s * M).tocsc()  # Solver/eigen2.py:109-111
    return C, M.tocsr(), dt


def factorize(A, M, sigma, *, force_complex=False, permc_spec="COLAMD", diag_pivot_thresh=None,
              perm: np.ndarray | None = None):
    """SuperLU of A - sigma M.  With `perm`, factor the symmetrically permuted matrix with
    permc_spec='NATURAL' (the GPU path's own ordering: the fair CPU comparison of BASELINE.md 4.3a)."""
    C, Mc, dt = _shifted(A, M, sigma, force_complex)
    opts = {}
    if diag_pivot_thresh is not None:
        opts["diag_pivot_thresh"] = diag_pivot_thresh
    if perm is not None:
        Cp = C[perm][:, perm].tocsc()
        lu = spla.splu(Cp, permc_spec="NATURAL", **opts)
        iperm = np.empty_like(perm)
        iperm[perm] = np.arange(len(perm))

        class _Permuted:
            L, U = lu.L, lu.U
            shape = lu.shape

            @staticmethod
            def solve(b, trans="N"):
                return lu.solve(np.ascontiguousarray(b[perm]), trans=trans)[iperm]

        return _Permuted, Mc, dt
    lu = spla.splu(C, permc_spec=permc_spec, **opts)
    return lu, Mc, dt


# ----------------------------------------------------------------------------- ARPACK path


def shift_invert_arpack(A, M, sigma, nev, *, ncv=None, tol=1e-10, maxiter=500, which_sort="TARGET_MAGNITUDE",
                        adjoint=False, dofs_p=None, seed=0, permc_spec="COLAMD",
                        diag_pivot_thresh=None, perm=None, force_complex=True) -> OracleResult:
    """`ArpackEigenSolver` restated (`Solver/eigen2.py:71-265`).

    adjoint=True solves the left problem  (A^H - conj(sigma) M^H) y = ...  i.e. eigenpairs of
    (A^H, M^H) nearest conj(sigma) (`Sensitivity/__init__.py:246-262`) on the same LU.
    """
    n = A.shape[0]
    t0 = time.perf_counter()
    lu, Mc, dt = factorize(A, M, sigma, force_complex=force_complex, permc_spec=permc_spec,
                           diag_pivot_thresh=diag_pivot_thresh, perm=perm)
    t_factor = time.perf_counter() - t0
    MH = Mc.conj().T.tocsr() if adjoint else Mc
    trans = "H" if adjoint else "N"
    count = [0]

    def op(x):
        count[0] += 1
        x = np.asarray(x, dtype=dt).ravel()
        if dofs_p is not None:
            x = x.copy()
            x[dofs_p] = 0.0  # Solver/eigen2.py:166-167
        y = lu.solve(MH @ x, trans=trans)  # Solver/eigen2.py:174-178
        if not np.isfinite(y).all():
            raise RuntimeError("non-finite values in shift-invert apply")  # eigen2.py:186-189
        if dofs_p is not None:
            y[dofs_p] = 0.0
        return y

    ncv = ncv if ncv is not None else max(4 * nev, 40)  # Solver/eigen2.py:222
    ncv = min(ncv, n - 1)
    rng = np.random.default_rng(seed)
    v0 = op(rng.standard_normal(n).astype(dt))
    lop = spla.LinearOperator((n, n), matvec=op, dtype=dt)
    t0 = time.perf_counter()
    mu, W = spla.eigs(lop, k=nev, which="LM", tol=tol, maxiter=maxiter, ncv=ncv, v0=v0)
    t_eigs = time.perf_counter() - t0
    sig = np.conj(sigma) if adjoint else sigma
    lam = sig + 1.0 / mu  # Solver/eigen2.py:206-207
    idx = np.argsort(which_key(which_sort, lam, sig), kind="stable")
    lam, W = lam[idx], W[:, idx]
    W = W / np.linalg.norm(W, axis=0)
    if adjoint:
        AH, MHs = sp.csr_matrix(A).conj().T, (sp.csr_matrix(M).conj().T if M is not None else sp.identity(n))
        res = eigen2_residuals(AH, MHs, lam, W)
    else:
        res = eigen2_residuals(sp.csr_matrix(A), sp.csr_matrix(M) if M is not None else sp.identity(n), lam, W)
    nnz_lu = int(lu.L.nnz + lu.U.nnz)
    return OracleResult(lam, W, res, count[0], 0, {"factor": t_factor, "eigs": t_eigs}, nnz_lu)


# ----------------------------------------------------------------------------- Krylov-Schur


def _givens(f: complex, g: complex):
    """(c real, s complex) with [c s; -conj(s) c] [f; g] = [r; 0]   (LAPACK zlartg)."""
    if g == 0:
        return 1.0, 0.0j
    if f == 0:
        return 0.0, np.conj(g) / abs(g)
    d = np.hypot(abs(f), abs(g))
    c = abs(f) / d
    s = (f / abs(f)) * np.conj(g) / d
    return c, s


def _swap_schur(T: np.ndarray, Q: np.ndarray, i: int) -> None:
    """Swap diagonal entries i, i+1 of the complex upper-triangular T; T = Q^H S Q kept (ztrexc)."""
    t11, t22 = T[i, i], T[i + 1, i + 1]
    c, s = _givens(T[i, i + 1], t22 - t11)
    n = T.shape[0]
    if i + 2 < n:
        a, b = T[i, i + 2:].copy(), T[i + 1, i + 2:].copy()
        T[i, i + 2:] = c * a + s * b
        T[i + 1, i + 2:] = c * b - np.conj(s) * a
    if i > 0:
        a, b = T[:i, i].copy(), T[:i, i + 1].copy()
        T[:i, i] = c * a + np.conj(s) * b
        T[:i, i + 1] = c * b - s * a
    T[i, i], T[i + 1, i + 1] = t22, t11
    a, b = Q[:, i].copy(), Q[:, i + 1].copy()
    Q[:, i] = c * a + np.conj(s) * b
    Q[:, i + 1] = c * b - s * a


def sorted_schur(S: np.ndarray, key_fn):
    """Complex Schur form S = Q T Q^H with the diagonal of T[lock:, lock:] ordered by key_fn."""
    m = S.shape[0]
    T, Q = sla.schur(S.astype(complex), output="complex")
    for pos in range(m - 1):
        keys = key_fn(np.diag(T)[pos:])
        best = pos + int(np.argmin(keys))
        for j in range(best - 1, pos - 1, -1):
            _swap_schur(T, Q, j)
    return T, Q


def krylov_schur(op, n: int, nev: int, ncv: int, tol: float, max_it: int, *, key_fn, back=lambda th: th,
                 v0: np.ndarray, conv_rel_to_back: bool = False):
    """Krylov-Schur (Stewart 2001) with SLEPc's default policy [from memory of SLEPc 3.22 krylovschur.c]:
    locking, keep = 0.5, mpd = ncv, relative convergence `beta |s_i| <= tol |theta_i|`, CGS2.

    `op`  : callable applying OP to a complex vector;   `key_fn(theta_array) -> sort key`.
    Returns (theta (nconv,), X (n, nconv) unit 2-norm, n_applies, n_restarts).
    """
    ncv = max(1, min(ncv, n))
    nev = min(nev, n)
    V = np.zeros((n, ncv + 1), dtype=complex)
    S = np.zeros((ncv + 1, ncv), dtype=complex)
    v = np.asarray(v0, dtype=complex)
    V[:, 0] = v / np.linalg.norm(v)
    nconv, l, applies, its = 0, 0, 0, 0
    breakdown = False
    m = ncv
    T = Q = None
    while True:
        its += 1
        m = ncv
        for j in range(nconv + l, ncv):
            w = op(V[:, j])
            applies += 1
            h = V[:, : j + 1].conj().T @ w
            w = w - V[:, : j + 1] @ h
            h2 = V[:, : j + 1].conj().T @ w
            w = w - V[:, : j + 1] @ h2
            S[: j + 1, j] = h + h2
            beta = np.linalg.norm(w)
            S[j + 1, j] = beta
            if beta <= 1e-14 * max(1.0, np.abs(S[: j + 1, j]).max()):
                m = j + 1
                breakdown = True
                S[j + 1, j] = 0.0
                break
            V[:, j + 1] = w / beta
        b_last = S[m, m - 1]
        # Schur form of the active block only (locked block S[:nconv,:nconv] is already triangular)
        T = S[:m, :m].copy()
        T2, Q2 = sorted_schur(T[nconv:, nconv:], key_fn)
        T[nconv:, nconv:] = T2
        T[:nconv, nconv:] = T[:nconv, nconv:] @ Q2
        Q = np.eye(m, dtype=complex)
        Q[nconv:, nconv:] = Q2
        brow = b_last * Q[m - 1, :]
        # convergence of the leading run
        k = nconv
        while k < m:
            y = np.zeros(k + 1, dtype=complex)
            y[k] = 1.0
            if k > 0:
                Tk = T[:k, :k] - T[k, k] * np.eye(k)
                d = np.diag(Tk).copy()
                smin = max(np.abs(T).max(), 1e-300) * 2.2e-16
                Tk[np.diag_indices(k)] = np.where(np.abs(d) < smin, smin, d)
                y[:k] = sla.solve_triangular(Tk, -T[:k, k])
            y /= np.linalg.norm(y)
            resid = abs(brow[: k + 1] @ y)
            theta = T[k, k]
            ref = abs(back(theta)) if conv_rel_to_back else abs(theta)
            if resid <= tol * ref or breakdown:
                k += 1
            else:
                break
        done = k >= nev or its >= max_it or breakdown
        l_new = 0 if done else max(1, int((m - k) * 0.5))
        keep = k + l_new
        V[:, nconv:keep] = V[:, nconv:m] @ Q2[:, : keep - nconv]
        if not done:
            V[:, keep] = V[:, m]
        S[:, :] = 0.0
        S[:keep, :keep] = T[:keep, :keep]
        if not done:
            S[keep, :keep] = brow[:keep]
            S[keep, :k] = 0.0  # locked
        nconv, l = k, l_new
        if done:
            break
    # eigenvectors of the converged block
    nc = nconv
    Tc = T[:nc, :nc]
    Y = np.zeros((nc, nc), dtype=complex)
    for i in range(nc):
        Y[i, i] = 1.0
        if i > 0:
            Tk = Tc[:i, :i] - Tc[i, i] * np.eye(i)
            d = np.diag(Tk).copy()
            smin = max(np.abs(Tc).max(), 1e-300) * 2.2e-16
            Tk[np.diag_indices(i)] = np.where(np.abs(d) < smin, smin, d)
            Y[:i, i] = sla.solve_triangular(Tk, -Tc[:i, i])
    X = V[:, :nc] @ Y
    X /= np.linalg.norm(X, axis=0)
    return np.diag(Tc).copy(), X, applies, its


def shift_invert_krylov_schur(A, M, sigma, nev, *, ncv=80, tol=1e-10, max_it=500, which="TARGET_MAGNITUDE",
                              adjoint=False, seed=0, permc_spec="COLAMD", diag_pivot_thresh=None,
                              perm=None, force_complex=True, purify=True) -> OracleResult:
    """SLEPc-default path restated: STSINVERT + Krylov-Schur + LU  (`Solver/utils.py:244-270`)."""
    n = A.shape[0]
    t0 = time.perf_counter()
    lu, Mc, dt = factorize(A, M, sigma, force_complex=force_complex, permc_spec=permc_spec,
                           diag_pivot_thresh=diag_pivot_thresh, perm=perm)
    t_factor = time.perf_counter() - t0
    MH = Mc.conj().T.tocsr() if adjoint else Mc
    trans = "H" if adjoint else "N"
    sig = np.conj(sigma) if adjoint else sigma

    def op(x):
        return lu.solve(np.asarray(MH @ x, dtype=dt) if dt == np.complex128 else MH @ x, trans=trans) \
            if dt == np.complex128 else (lu.solve(np.ascontiguousarray((MH @ x).real), trans=trans)
                                         + 1j * lu.solve(np.ascontiguousarray((MH @ x).imag), trans=trans))

    back = lambda th: sig + 1.0 / th  # noqa: E731
    key_fn = lambda th: which_key(which, back(np.where(th == 0, 1e-300, th)), sig)  # noqa: E731
    rng = np.random.default_rng(seed)
    v0 = op(rng.standard_normal(n).astype(complex))
    t0 = time.perf_counter()
    theta, X, applies, its = krylov_schur(op, n, nev, ncv, tol, max_it, key_fn=key_fn, back=back, v0=v0)
    if purify and len(theta):
        X = np.stack([op(X[:, i]) for i in range(X.shape[1])], axis=1)
        applies += X.shape[1]
        X /= np.linalg.norm(X, axis=0)
    t_eigs = time.perf_counter() - t0
    lam = back(theta)
    idx = np.argsort(which_key(which, lam, sig), kind="stable")
    lam, X = lam[idx], X[:, idx]
    Ah = sp.csr_matrix(A).conj().T if adjoint else sp.csr_matrix(A)
    Mh = (sp.csr_matrix(M).conj().T if adjoint else sp.csr_matrix(M)) if M is not None else sp.identity(n)
    res = eigen2_residuals(Ah, Mh, lam, X) if len(lam) else np.zeros(0)
    return OracleResult(lam, X, res, applies + 1, its, {"factor": t_factor, "eigs": t_eigs},
                        int(lu.L.nnz + lu.U.nnz))


# ----------------------------------------------------------------------------- dense


def dense_pencil_eigs(A, M=None):
    """All finite eigenpairs of a small pencil with LAPACK (ground truth for n <= ~3000)."""
    Ad = A.toarray() if sp.issparse(A) else np.asarray(A)
    if M is None:
        w, V = sla.eig(Ad)
        return w, V
    Md = M.toarray() if sp.issparse(M) else np.asarray(M)
    w, V = sla.eig(Ad, Md)
    ok = np.isfinite(w)
    return w[ok], V[:, ok]
