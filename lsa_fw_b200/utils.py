"""LSA-FW Solver utilities, B200 backend: enums and the `iEpsSolver` wrapper.

Mirror of the EPS half of the reference's `Solver/utils.py` (`:27-63` iEpsProblemType, `:66-93`
PreconditionerType, `:131-149` iSTType, `:152-187` iEpsWhich, `:190-328` iEpsSolver) with the same
member names, setter/getter names, argument meaning and error behaviour, so that code written
against the SLEPc-backed class runs unchanged.  Where the reference forwards to `SLEPc.EPS`, this
class records the setting and `solve()` drives the CUDA library through the C ABI of
`include/lsa_b200.h` (ctypes, `_lib.py`).  There is no CPU path: without the CUDA extension or
without a GPU `solve()` raises.
"""

from __future__ import annotations

import hashlib
import logging
import weakref
from enum import Enum, StrEnum, auto
from typing import Iterator

import numpy as np
import scipy.sparse as sp

from . import _lib
from ._lib import LsaError
from .carriers import iComplexPETScVector, iPETScMatrix, iPETScNullSpace, iPETScVector

logger = logging.getLogger(__name__)

__all__ = [
    "iEpsProblemType", "PreconditionerType", "iSTType", "iEpsWhich", "iEpsSolver", "LsaError",
    "clear_symbolic_cache",
]


class iEpsProblemType(Enum):  # noqa: N801
    """EPS problem types (reference `Solver/utils.py:27-63`)."""

    HEP = 1
    """Standard Hermitian eigenvalue problem. Ax = λx, with A Hermitian."""
    NHEP = 2
    """Standard non-Hermitian eigenvalue problem: Ax = λx, with A arbitrary."""
    GHEP = 3
    """Generalized Hermitian eigenvalue problem: Ax = λBx, with A, B Hermitian."""
    GNHEP = 4
    """Generalized non-Hermitian eigenvalue problem: Ax = λBx, with arbitrary A, B."""
    PGNHEP = 5
    """Generalized non-Hermitian eigenvalue problem with positive (semi-)definite B."""
    GHIEP = 6
    """Generalized Hermitian indefinite eigenvalue problem."""

    def to_slepc(self) -> int:
        """Backend code of this problem type (SLEPc's enum value in the reference)."""
        return self.value

    @classmethod
    def from_slepc(cls, problem_type) -> "iEpsProblemType":
        try:
            return cls(int(problem_type))
        except ValueError:
            raise ValueError(f"Unsupported SLEPc EPS ProblemType: {problem_type}")

    @classmethod
    def from_string(cls, name: str) -> "iEpsProblemType":
        try:
            return cls[name.upper()]
        except KeyError:
            raise ValueError(f"Invalid problem type: {name}. Choose from {list(cls.__members__.keys())}.")


class PreconditionerType(StrEnum):
    """Preconditioner names accepted by `set_st_pc_type` (reference `Solver/utils.py:66-93`).

    Only the direct factorisations (LU, CHOLESKY) are executed by the B200 backend; the other names
    are recorded (and round-trip through `.raw.getST().getKSP().getPC().getType()`), and `solve()`
    raises NotImplementedError for them.
    """

    NONE = auto()
    JACOBI = auto()
    SOR = auto()
    ASM = auto()
    ILU = auto()
    ICC = auto()
    LU = auto()
    CHOLESKY = auto()
    GAMG = auto()
    HYPRE = auto()
    REDUNDANT = auto()
    SHELL = auto()


class iSTType(Enum):  # noqa: N801
    """Spectral transformations (reference `Solver/utils.py:131-149`); SHIFT and SINVERT are built."""

    SHELL = "shell"
    SHIFT = "shift"
    SINVERT = "sinvert"
    CAYLEY = "cayley"
    PRECOND = "precond"
    FILTER = "filter"

    def to_slepc(self) -> str:
        return self.value


class iEpsWhich(Enum):  # noqa: N801
    """Which eigenpairs (reference `Solver/utils.py:152-187`).

    The reference maps BOTH `SMALLEST_MAGNITUDE` and `LARGEST_REAL` to `SLEPc.EPS.Which.LARGEST_REAL`
    (`Solver/utils.py:157-158`), which makes them the same enum member; that alias is kept, so callers
    observe identical behaviour.
    """

    ALL = "ALL"
    LARGEST_MAGNITUDE = "LARGEST_MAGNITUDE"
    SMALLEST_MAGNITUDE = "LARGEST_REAL"
    LARGEST_REAL = "LARGEST_REAL"
    SMALLEST_REAL = "SMALLEST_REAL"
    LARGEST_IMAGINARY = "LARGEST_IMAGINARY"
    SMALLEST_IMAGINARY = "SMALLEST_IMAGINARY"
    TARGET_MAGNITUDE = "TARGET_MAGNITUDE"
    TARGET_REAL = "TARGET_REAL"
    TARGET_IMAGINARY = "TARGET_IMAGINARY"
    USER = "USER"

    def to_slepc(self) -> str:
        return self.value

    def to_arpack(self) -> str:
        """ARPACK-style code (reference `Solver/utils.py:170-187`)."""
        match self:
            case iEpsWhich.LARGEST_REAL:
                return "LR"
            case iEpsWhich.LARGEST_IMAGINARY:
                return "LI"
            case iEpsWhich.SMALLEST_REAL:
                return "SR"
            case iEpsWhich.SMALLEST_IMAGINARY:
                return "SI"
            case iEpsWhich.LARGEST_MAGNITUDE:
                return "LM_abs"
            case _:
                raise ValueError(f"Unsupported type for ARPACK-based eigensolver: {self.name}.")


_HERMITIAN = {iEpsProblemType.HEP, iEpsProblemType.GHEP}

# --------------------------------------------------------------------------- symbolic reuse

_SYM_CACHE_MAX = 4
# (id(A carrier), id(M carrier)) -> weakref of the solver that factored that pencil
_FACTOR_REGISTRY: dict[tuple[int, int], "weakref.ReferenceType[iEpsSolver]"] = {}


def clear_symbolic_cache() -> None:
    """Drop cached symbolic analyses.  A handle (and its device buffers) is released as soon as no solver
    object refers to it any more; handles still in use by live solvers stay valid."""
    _SYM_CACHE.clear()
    _FACTOR_REGISTRY.clear()


def _sample_digest(arr: np.ndarray) -> bytes:
    n = arr.size
    return hashlib.blake2b(arr[:: max(1, n // 4096)].tobytes(), digest_size=8).digest()


def _same_index_arrays(a: np.ndarray, b: np.ndarray) -> bool:
    """Exact equality of two index arrays: identity first, then a strided sample, then the full comparison
    (one streaming pass by a few threads, `lsa_host_equal` -- far cheaper than a cryptographic hash)."""
    if a is b:
        return True
    if a.shape != b.shape:
        return False
    st = max(1, a.size // 4096)
    if not np.array_equal(a[::st], b[::st]):
        return False
    return _lib.host_equal(a, b)


class _SymbolicCache:
    """Symbolic analyses (native handles) keyed by the EXACT sparsity pattern of (A, M) and the analysis options.
    A lookup compares the candidate's index arrays with the stored (canonical) ones, so a freshly loaded matrix of
    a sweep -- a new object with the old pattern -- finds its analysis without hashing ~10^8 indices, and a match
    also proves that the candidate is in canonical CSR form.  dict-like surface for the few callers that iterate."""

    def __init__(self, max_entries: int = 4) -> None:
        self.max_entries = max_entries
        self._entries: list[dict] = []

    @staticmethod
    def _pat(mat):
        return None if mat is None else (mat.shape[0], mat.indptr, mat.indices)

    def lookup(self, A, M, extra: tuple):
        for e in self._entries:
            if e["extra"] != extra or (M is None) != (e["m"] is None) or e["handle"].closed:
                continue
            ok = True
            for mat, pat in ((A, e["a"]), (M, e["m"])):
                if mat is None:
                    continue
                ok = ok and mat.shape[0] == pat[0] and mat.indices.size == pat[2].size and \
                    _same_index_arrays(mat.indptr, pat[1]) and _same_index_arrays(mat.indices, pat[2])
                if not ok:
                    break
            if ok:
                return e["handle"]
        return None

    def store(self, A, M, extra: tuple, handle) -> None:
        while len(self._entries) >= self.max_entries:
            # evicted handles are NOT closed here: live solvers (and adjoint donors) may still hold them; the
            # device buffers go when the last reference does
            self._entries.pop(0)
        self._entries.append({"a": self._pat(A), "m": self._pat(M), "extra": extra, "handle": handle})

    def drop(self, handle) -> None:
        self._entries = [e for e in self._entries if e["handle"] is not handle]

    def clear(self) -> None:
        self._entries.clear()

    def __len__(self) -> int:
        return len(self._entries)


_SYM_CACHE = _SymbolicCache(_SYM_CACHE_MAX)


def _raw_csr(mat):
    """The CSR arrays of a carrier WITHOUT the O(nnz) canonical-form checks (a cache hit proves canonical form)."""
    m = getattr(mat, "_m", None)
    if m is None:
        m = mat.as_scipy_array() if hasattr(mat, "as_scipy_array") else mat
    return m if sp.isspmatrix_csr(m) else None


_ZERO_DIAG_MEMO: dict[int, tuple] = {}


def _zero_diagonal_rows(mat: sp.csr_matrix) -> np.ndarray:
    """Rows whose diagonal entry is zero or absent (memoised per value-array object: M is the same object along a
    sweep; `lsa_host_diag_is_zero`, a binary search per row)."""
    hit = _ZERO_DIAG_MEMO.get(id(mat.data))
    if hit is None or hit[0] is not mat.data or hit[1] != _sample_digest(mat.data):
        rows = np.nonzero(_lib.diag_is_zero(mat))[0].astype(np.int32)
        if len(_ZERO_DIAG_MEMO) > 16:
            _ZERO_DIAG_MEMO.clear()
        hit = (mat.data, _sample_digest(mat.data), rows)
        _ZERO_DIAG_MEMO[id(mat.data)] = hit
    return hit[2]


def _values_token(mat) -> tuple:
    """Identity + cheap content probe of a value array: (the array object, its sampled digest)."""
    return (mat.data, _sample_digest(mat.data))


def _same_values(token: tuple, mat) -> bool:
    return token is not None and token[0] is mat.data and token[1] == _sample_digest(mat.data)


_IDENTITY_MEMO: dict[int, sp.csr_matrix] = {}


def _identity_csr(n: int) -> sp.csr_matrix:
    m = _IDENTITY_MEMO.get(n)
    if m is None:
        if len(_IDENTITY_MEMO) > 8:
            _IDENTITY_MEMO.clear()
        m = sp.identity(n, dtype=np.float64, format="csr")
        _IDENTITY_MEMO[n] = m
    return m


def _as_csr(mat) -> sp.csr_matrix:
    m = mat.as_scipy_array() if hasattr(mat, "as_scipy_array") else mat
    if not sp.isspmatrix_csr(m):
        m = sp.csr_matrix(m)
    if not m.has_canonical_format:  # cached on the matrix object after the first check
        m = m.copy()
        m.sum_duplicates()
    return m


class _RawPC:
    def __init__(self, s: "iEpsSolver") -> None:
        self._s = s

    def getType(self) -> str:  # noqa: N802
        return self._s._pc_type


class _RawKSP:
    def __init__(self, s: "iEpsSolver") -> None:
        self._s = s

    def getPC(self) -> _RawPC:  # noqa: N802
        return _RawPC(self._s)

    def getIterationNumber(self) -> int:  # noqa: N802
        """Number of linear solves of the last eigensolve (each is one direct fwd+bwd solve)."""
        return int(self._s._stats.get("n_op_applies", 0))


class _RawST:
    def __init__(self, s: "iEpsSolver") -> None:
        self._s = s

    def getKSP(self) -> _RawKSP:  # noqa: N802
        return _RawKSP(self._s)

    def getType(self) -> str:  # noqa: N802
        return self._s._st_type.value

    def getShift(self):  # noqa: N802
        return self._s._target


class _RawEPS:
    """Shim exposing the few `SLEPc.EPS` getters that callers and the reference's tests touch
    (`Solver/eigen.py:142`; `tests/unit/Solver/test_eigen.py:97-104,320-322`)."""

    def __init__(self, s: "iEpsSolver") -> None:
        self._s = s

    def getTolerances(self):  # noqa: N802
        return self._s._tol, self._s._max_it

    def getDimensions(self):  # noqa: N802
        return self._s._nev, self._s._ncv_effective(), self._s._ncv_effective()

    def getProblemType(self):  # noqa: N802
        return self._s._problem_type.to_slepc()

    def getST(self) -> _RawST:  # noqa: N802
        return _RawST(self._s)

    def getConverged(self) -> int:  # noqa: N802
        return self._s.get_num_converged()

    def getEigenvalue(self, i: int):  # noqa: N802
        return self._s.get_eigenvalue(i)

    def getTarget(self):  # noqa: N802
        return self._s._target

    def getWhichEigenpairs(self):  # noqa: N802
        return self._s._which_effective().to_slepc()

    def getIterationNumber(self) -> int:  # noqa: N802
        return int(self._s._stats.get("n_restarts", 0))


class iEpsSolver:  # noqa: N801
    """Eigenproblem solver object with the interface of the reference's SLEPc wrapper
    (`Solver/utils.py:190-328`), executing on one B200 through liblsa_b200.so."""

    def __init__(self, A: iPETScMatrix | None = None, M: iPETScMatrix | None = None, comm=None) -> None:
        if M is not None and A is None:
            raise ValueError("Cannot set right-hand operator M without left-hand operator A.")
        self._A = None
        self._M = None
        self._problem_type = iEpsProblemType.NHEP
        self._nev = 1
        self._ncv: int | None = None
        self._tol = 1e-8
        self._max_it = 100
        self._which: iEpsWhich | None = None
        self._target: float | complex = 0.0
        self._st_type = iSTType.SHIFT
        self._pc_type = "lu"
        self._interval = None
        # backend options (extensions; all optional)
        self._opts = dict(leaf_size=64, coords=None, refine_steps=0, tiny_pivot=1e-13, seed=0, device=0,
                          purify=True, nthreads=0, v0=None, force_complex=False, coupled_fraction=0.5,
                          growth_limit=1e6, device_values=None, partition=None, symmetric="auto", b_inner="auto")
        self._adjoint = False
        self._handle: _lib.Handle | None = None
        self._factor_key = None
        self._factor_gen = -1      # generation of the handle's numeric state right after OUR factorisation
        self._result_gen = -1      # ... right after OUR eigensolve (device-side results still ours)
        self._factor_tokens = None  # identity + content probes of the value arrays that were factored
        self._stats: dict = {}
        self._nconv = 0
        self._eigenvalues: np.ndarray = np.zeros(0, dtype=complex)
        self._eigenvectors: np.ndarray | None = None
        self._handed_out: set[int] = set()
        self._complex_mode = False
        if A is not None:
            self.set_operators(A, M)

    # ------------------------------------------------------------------ reference interface
    @property
    def raw(self) -> _RawEPS:
        """Access the underlying solver object (a shim with SLEPc.EPS getter names)."""
        return _RawEPS(self)

    def set_operators(self, A: iPETScMatrix, M: iPETScMatrix | None = None) -> None:
        """Set the matrix operators for the generalized eigenproblem Ax = λMx."""
        self._A, self._M = A, M
        self._factor_key = None

    def set_problem_type(self, problem_type: iEpsProblemType) -> None:
        """Set the eigenproblem type. Refer to iEpsProblemType enum."""
        self._problem_type = problem_type

    def set_dimensions(self, number_eigenpairs: int, subspace_dimension: int | None = None) -> None:
        """Set number of eigenpairs (nev) and subspace dimension (ncv)."""
        self._nev = int(number_eigenpairs)
        self._ncv = None if subspace_dimension is None else int(subspace_dimension)

    def set_tolerances(self, atol: float, max_it: int) -> None:
        """Set convergence tolerance (relative, as SLEPc's `tol`) and maximum number of restarts."""
        self._tol, self._max_it = float(atol), int(max_it)

    def set_which_eigenpairs(self, which: iEpsWhich) -> None:
        """Select which eigenpairs to compute. Refer to iEpsWhich enum."""
        self._which = which

    def set_target(self, sigma: float | complex) -> None:
        """Set spectral transformation shift (target)."""
        self._target = sigma
        self._factor_key = None

    def set_interval(self, a: float, b: float) -> None:
        """Compute eigenvalues in real interval [a, b] (spectrum slicing: not built)."""
        self._interval = (a, b)

    def set_interval_complex(self, a: float, b: float, c: float, d: float) -> None:
        """Compute eigenvalues in complex rectangle [a,b]x[c,d] (contour methods: not built)."""
        self._interval = (a, b, c, d)

    def set_st_type(self, st_type: iSTType) -> None:
        """Set spectral transformation type. Refer to iSTType enum."""
        self._st_type = st_type
        self._factor_key = None

    def set_st_pc_type(self, pc_type: PreconditionerType) -> None:
        """Set the factorisation used inside the spectral transformation."""
        self._pc_type = pc_type.name.lower()

    # ------------------------------------------------------------------ extensions
    def set_backend_options(self, **kw) -> None:
        """B200-backend knobs: leaf_size, coords (n x dim ordering hint), refine_steps, tiny_pivot,
        seed, device, purify (True: through the Krylov-Schur relation; "explicit": one OP apply per vector),
        nthreads, v0 (start vector), force_complex, coupled_fraction, growth_limit (largest LU multiplier
        tolerated before the solve re-analyses / switches iterative refinement on), device_values
        ((A_vals, M_vals) already resident on the GPU, in CSR entry order: torch CUDA tensors or any object
        with `__cuda_array_interface__` / `__dlpack__`; the host copies in A / M are then only used for
        their pattern), partition ("auto": split this factorisation / eigensolve over the GPUs of the initialised
        torch.distributed group, one process per GPU, every rank making the same calls -- lsa_fw_b200/partitioned.py),
        symmetric ("auto": real symmetric problems -- HEP / GHEP with st_pc_type CHOLESKY, real data, real shift -- are
        factored as L D L^T at half the factor storage; True / False force / forbid it), b_inner ("auto": GHEP problems
        run the Krylov-Schur iteration in the M-inner product -- M-orthonormal basis, Hermitian projected matrix, as
        SLEPc does for EPS_GHEP; False keeps Euclidean inner products and only M-normalises the returned vectors)."""
        unknown = set(kw) - set(self._opts)
        if unknown:
            raise TypeError(f"unknown backend option(s): {sorted(unknown)}")
        self._opts.update(kw)

    def set_adjoint(self, flag: bool = True) -> None:
        """Solve the adjoint problem (A^H, M^H) at conj(target) on the factors of (A, M, target):
        left eigenvectors without the second factorisation of `Sensitivity/__init__.py:246-262`."""
        self._adjoint = bool(flag)

    @property
    def stats(self) -> dict:
        """Timings and counters of the last `solve()` (symbolic, factor, eigs, residual data)."""
        return self._stats

    @property
    def handle(self) -> "_lib.Handle | None":
        return self._handle

    # ------------------------------------------------------------------ helpers
    def _ncv_effective(self) -> int:
        n = self._A.shape[0] if self._A is not None else 1 << 30
        # SLEPc's default rule (EPSSetDimensions, ncv = max(2 nev, nev + 15)), capped at the widest basis
        # the device kernels hold (MAX_NCV = 256); an explicit ncv is passed through and checked there
        ncv = self._ncv if self._ncv is not None else min(max(2 * self._nev, self._nev + 15), 256)
        if self._ncv is not None and self._ncv < self._nev:
            # EPSSetDimensions_Default: "The value of ncv must be at least nev"
            raise ValueError(f"The value of ncv ({self._ncv}) must be at least nev ({self._nev})")
        return max(1, min(ncv, n))

    def _partition_world(self) -> tuple[int, int]:
        if not self._opts["partition"]:
            return 0, 1
        from .partitioned import world_info

        return world_info()

    def _which_effective(self) -> iEpsWhich:
        if self._which is not None:
            return self._which
        return iEpsWhich.TARGET_MAGNITUDE if self._st_type == iSTType.SINVERT else iEpsWhich.LARGEST_MAGNITUDE

    def _resolve_adjoint_reuse(self):
        """If (A, M) are `.H` views of a pencil that another live solver factored at conj(target),
        return that solver (drop-in form of the Sensitivity adjoint solve)."""
        a0 = getattr(self._A, "_adjoint_of", None)
        m0 = getattr(self._M, "_adjoint_of", None) if self._M is not None else None
        if a0 is None or (self._M is not None and m0 is None):
            return None
        ref = _FACTOR_REGISTRY.get((id(a0), id(m0) if m0 is not None else 0))
        other = ref() if ref is not None else None
        if other is None or other._handle is None or other._factor_key is None:
            return None
        # the handle may be shared (symbolic cache): it must still hold OUR donor's values and factors, and the
        # donor's matrices must not have been changed since they were factored
        if other._handle.closed or other._handle.gen_factor != other._factor_gen:
            return None
        toks = other._factor_tokens
        if toks is None or not _same_values(toks[0], _as_csr(other._A)):
            return None
        if other._M is not None and not _same_values(toks[1], _as_csr(other._M)):
            return None
        if other._st_type != iSTType.SINVERT or self._st_type != iSTType.SINVERT:
            return None
        if complex(other._target).conjugate() != complex(self._target):
            return None
        return other

    # ------------------------------------------------------------------ solve
    def solve(self) -> None:
        """Run the eigensolver on the configured operators and settings."""
        if self._A is None:
            raise ValueError("Operators must be set before solve().")
        if self._interval is not None or self._which == iEpsWhich.ALL:
            raise NotImplementedError("interval / ALL (spectrum slicing, contour integrals) is not built on the B200 backend")
        if self._st_type not in (iSTType.SHIFT, iSTType.SINVERT):
            raise NotImplementedError(f"spectral transformation {self._st_type.name} is not built on the B200 backend")
        if self._which == iEpsWhich.USER:
            raise NotImplementedError("user-defined eigenvalue ordering is not built on the B200 backend")
        needs_factor = self._st_type == iSTType.SINVERT or self._M is not None
        if needs_factor and self._pc_type not in ("lu", "cholesky"):
            raise NotImplementedError(
                f"st_pc_type '{self._pc_type}' is not built on the B200 backend (direct LU/CHOLESKY only)")
        import time

        t_start = time.perf_counter()
        n = self._A.shape[0]
        sigma = complex(self._target)
        sinvert = self._st_type == iSTType.SINVERT
        adjoint = self._adjoint
        donor = self._resolve_adjoint_reuse()
        stats: dict = {"n": n}

        if donor is not None:
            # adjoint modes on the donor's factors: same handle, trans = H sweeps
            h = donor._handle
            self._handle = h
            adjoint = True
            sigma_fact = complex(donor._target)
            self._complex_mode = donor._complex_mode
            stats["reused_factorisation"] = True
            stats["symbolic_seconds"] = 0.0
            stats["factor_seconds"] = 0.0
        else:
            t0 = time.perf_counter()
            coords = self._opts["coords"]
            # symmetric factorisation L D L^T (the reference's GHEP + CHOLESKY use, Elasticity/utils.py:139-155): real data,
            # real shift, a Hermitian problem type, one GPU
            a0, m0 = _raw_csr(self._A), (_raw_csr(self._M) if self._M is not None else None)
            real_data = (a0 is None or not np.iscomplexobj(a0.data)) and (m0 is None or not np.iscomplexobj(m0.data)) \
                and (a0 is not None or not np.iscomplexobj(_as_csr(self._A).data))
            sym_ok = (real_data and sigma.imag == 0.0 and not self._opts["force_complex"] and self._partition_world()[1] == 1
                      and needs_factor)
            want = self._opts["symmetric"]
            if want is True and not sym_ok:
                raise ValueError("symmetric=True needs real matrices, a real shift and a single GPU")
            use_sym = bool(sym_ok and (want is True or (want == "auto" and self._problem_type in _HERMITIAN
                                                        and self._pc_type == "cholesky")))
            stats["symmetric_factorisation"] = use_sym
            extra_base = (use_sym, self._opts["leaf_size"],
                          None if coords is None else hashlib.blake2b(np.ascontiguousarray(coords).tobytes(), digest_size=16).hexdigest(),
                          self._opts["device"], self._opts["coupled_fraction"], self._st_type.value, self._partition_world(),
                          self._M is None)

            def order_last_of(A, M):
                # zero diagonal of the matrix to be factored (pressure rows): ordered last inside their fronts.
                # Shift-independent form for sinvert (rows whose diagonal vanishes in A AND in M), so that the analysis
                # is valid for every shift of a sweep; part of the cache key together with the transform.
                flags = np.zeros(n, dtype=np.uint8)
                if sinvert and M is not None:
                    rows = _zero_diagonal_rows(M)                      # memoised: M does not change along a sweep
                    flags[rows[_lib.diag_is_zero(A, rows) != 0]] = 1
                elif sinvert:
                    pass                                               # M = I: no zero diagonal in A - sigma I to plan for
                else:
                    flags[_zero_diagonal_rows(M if M is not None else A)] = 1
                return flags

            # fast path: the raw arrays against the cached canonical patterns (a hit proves canonical form)
            h = None
            ht = stats.setdefault("host_timing", {})
            ht["prologue"] = time.perf_counter() - t0
            A, M = _raw_csr(self._A), (_raw_csr(self._M) if self._M is not None else None)
            if A is not None and (self._M is None or M is not None) and len(_SYM_CACHE):
                Mx = M
                if Mx is None and sinvert:
                    Mx = _identity_csr(n)
                t1 = time.perf_counter()
                ol = order_last_of(A, Mx)
                ht["order_last"] = time.perf_counter() - t1
                t1 = time.perf_counter()
                extra = extra_base + (hashlib.blake2b(ol.tobytes(), digest_size=16).hexdigest(),)
                ht["order_last_digest"] = time.perf_counter() - t1
                t1 = time.perf_counter()
                h = _SYM_CACHE.lookup(A, Mx, extra)
                ht["lookup"] = time.perf_counter() - t1
                if h is not None:
                    M = Mx
            if h is None:
                A = _as_csr(self._A)
                M = _as_csr(self._M) if self._M is not None else None
                if M is None and sinvert:
                    # standard problem: the shifted operator is A - sigma I
                    M = _identity_csr(n)
                ol = order_last_of(A, M)
                extra = extra_base + (hashlib.blake2b(ol.tobytes(), digest_size=16).hexdigest(),)
                h = _SYM_CACHE.lookup(A, M, extra)
            data_complex = np.iscomplexobj(A.data) or (M is not None and np.iscomplexobj(M.data))
            use_complex = data_complex or sigma.imag != 0.0 or self._opts["force_complex"]
            self._complex_mode = use_complex
            if h is None:
                rank, world = self._partition_world()
                h = _lib.Handle(n, self._opts["device"], rank, world)
                h.set_option("coupled_fraction", self._opts["coupled_fraction"])
                if use_sym:
                    h.set_option("symmetric", 1)
                h.analyze(A.indptr, A.indices, None if M is None else M.indptr, None if M is None else M.indices,
                          leaf_size=self._opts["leaf_size"], coords=coords, order_last=ol,
                          nthreads=self._opts["nthreads"])
                if h.world > 1:
                    from .partitioned import attach_comm

                    attach_comm(h)
                _SYM_CACHE.store(A, M, extra, h)
                stats["symbolic_cached"] = False
            else:
                stats["symbolic_cached"] = True
            if h.world > 1:
                pi = h.partition_info()
                stats.update(partition_rank=pi.rank, partition_world=pi.world, n_top_fronts=pi.n_top_fronts,
                             n_replicated_rows=int(pi.n_replicated_rows), n_own_rows=int(pi.n_own_rows),
                             partition_weights=(pi.weight_top, pi.weight_max_subtrees, pi.weight_total))
            stats["symbolic_seconds"] = time.perf_counter() - t0
            self._handle = h
            info = h.symbolic_info()
            stats.update(n_fronts=info.n_fronts, n_levels=info.n_levels, nnz_lu=info.nnz_lu,
                         n_decoupled=info.n_decoupled, max_front=info.max_front,
                         symbolic_phases=list(info.seconds))
            t0 = time.perf_counter()
            dv = self._opts["device_values"]
            if dv is not None:
                h.set_values_device(dv[0], None if M is None else dv[1])
            elif M is not None and self._M is not None and _same_values(h.m_token, M):
                ht["m_token_check"] = time.perf_counter() - t0
                h.set_values(A.data, None)          # M unchanged since the last upload to this handle: A only
                stats["m_upload_skipped"] = True
            else:
                h.set_values(A.data, None if M is None else M.data)
                h.m_token = _values_token(M) if (M is not None and self._M is not None) else None
            stats["upload_seconds"] = time.perf_counter() - t0
            ht["set_values_native"] = getattr(h, "set_values_seconds", None)
            sigma_fact = sigma
            # attached nullspace (constant pressure of an enclosed flow, FEM/operators.py:534-545): projected out of
            # every operator application; the vanishing pivot of the singular shifted operator is replaced
            ns = getattr(self._A, "get_nullspace", lambda: None)()
            if isinstance(ns, iPETScNullSpace):
                ns_arr = np.asarray(ns.as_array(n))     # a constant-only nullspace takes its size from the operator
            else:
                ns_arr = None if ns is None else np.asarray(ns.as_array() if hasattr(ns, "as_array") else ns)
            if ns_arr is not None or h.ns_count:
                h.set_nullspace(ns_arr)
            stats["nullspace_dimension"] = 0 if ns_arr is None else int(ns_arr.reshape(n, -1).shape[1])
            if needs_factor:
                scalar = _lib.LSA_C128 if use_complex else _lib.LSA_F64
                if sinvert:
                    fs = h.factor(1.0, -sigma, scalar, self._opts["tiny_pivot"])
                else:
                    fs = h.factor(0.0, 1.0, scalar, 0.0)  # plain M^-1: an exactly singular M must raise
                if use_sym and (fs.n_perturbed > stats["nullspace_dimension"] or fs.max_multiplier > self._opts["growth_limit"]):
                    # L D L^T without pivoting met a vanishing pivot / element growth (indefinite shift): general LU
                    logger.warning("symmetric factorisation unstable (%d pivots replaced, growth %.2e); using the general LU",
                                   fs.n_perturbed, fs.max_multiplier)
                    self._opts["symmetric"] = False
                    self._factor_key = None
                    return self.solve()
                if sinvert and fs.n_perturbed > stats["nullspace_dimension"] and self._opts["coupled_fraction"] < 1.0:
                    # tiny pivots were replaced: the cheap placement of the zero-diagonal unknowns was not
                    # enough for this pencil -> redo the analysis with the robust rule and factor again
                    logger.warning("%d tiny pivots replaced; re-analysing with coupled_fraction = 1", fs.n_perturbed)
                    self._opts["coupled_fraction"] = 1.0
                    self._factor_key = None
                    return self.solve()
                refine_steps = int(self._opts["refine_steps"])
                if sinvert and fs.max_multiplier > self._opts["growth_limit"]:
                    # element growth of the restricted pivoting (candidates = the current 128-row block): the
                    # factors are inaccurate.  First the robust placement of the zero-diagonal unknowns, then
                    # iterative refinement inside every operator application.
                    if self._opts["coupled_fraction"] < 1.0:
                        logger.warning("LU multiplier growth %.2e > %.1e; re-analysing with coupled_fraction = 1",
                                       fs.max_multiplier, self._opts["growth_limit"])
                        self._opts["coupled_fraction"] = 1.0
                        self._factor_key = None
                        return self.solve()
                    refine_steps = max(refine_steps, 2)
                    logger.warning("LU multiplier growth %.2e > %.1e; %d iterative-refinement steps per solve enabled",
                                   fs.max_multiplier, self._opts["growth_limit"], refine_steps)
                    stats["refine_steps_forced"] = refine_steps
                self._refine_steps_effective = refine_steps
                stats.update(factor_seconds=fs.seconds, factor_flops=fs.flops, n_perturbed=fs.n_perturbed,
                             n_row_swaps=fs.n_row_swaps, min_pivot=fs.min_pivot, max_pivot=fs.max_pivot,
                             max_multiplier=fs.max_multiplier,
                             factor_kernels=fs.n_kernels)
                self._factor_key = (sigma_fact, self._st_type)
                self._factor_gen = h.gen_factor
                self._factor_tokens = (_values_token(A), None if self._M is None else _values_token(M))
                _FACTOR_REGISTRY[(id(self._A), id(self._M) if self._M is not None else 0)] = weakref.ref(self)

        # Hermitian-definite problem (EPS_GHEP): M-inner products on one GPU, M-normalised vectors from the device
        b_mode = 0
        if self._problem_type == iEpsProblemType.GHEP and self._M is not None and h.world == 1:
            b_mode = 2 if (self._opts["b_inner"] in ("auto", True) and not adjoint) else 1
        self._b_mode = b_mode
        stats["b_mode"] = b_mode
        which = self._which_effective()
        res = h.eigs(nev=self._nev, ncv=self._ncv_effective(), tol=self._tol, max_restarts=self._max_it,
                     which=which.value, transform=_lib.LSA_ST_SINVERT if sinvert else _lib.LSA_ST_SHIFT,
                     sigma=sigma_fact, adjoint=adjoint, purify=(2 if self._opts["purify"] == "explicit" else int(bool(self._opts["purify"]))) if sinvert else 0,
                     refine_steps=(donor._refine_steps_effective if donor is not None
                                   else getattr(self, "_refine_steps_effective", self._opts["refine_steps"])),
                     seed=self._opts["seed"], v0=self._opts["v0"], b_mode=b_mode)
        self._result_gen = h.gen_result
        self._nconv = res.nconv
        self._eigenvalues = h.eigenvalues(res.nconv)
        # the handle (and its device buffers) may be shared with other solver objects through the
        # symbolic cache / adjoint reuse: bring the vectors to the host now
        self._eigenvectors = None
        self._handed_out = set()
        t0 = time.perf_counter()
        if res.nconv > 0:
            # the pairs `EigenSolver.solve()` hands out (min(nconv, nev)) come to the host now: the handle and its
            # device buffers may be shared; converged extras are fetched by the same call (one block, one size)
            self._fetch_vectors()
        stats["fetch_seconds"] = time.perf_counter() - t0
        stats.update(nconv=res.nconv, n_restarts=res.n_restarts, n_op_applies=res.n_op_applies,
                     breakdown=res.breakdown, n_reorth=res.n_reorth, eigs_seconds=res.seconds, solve_seconds=res.seconds_solve,
                     spmv_seconds=res.seconds_spmv, ortho_seconds=res.seconds_ortho, rr_seconds=res.seconds_rr,
                     restart_seconds=res.seconds_restart, total_seconds=time.perf_counter() - t_start)
        self._stats = stats

    # ------------------------------------------------------------------ results
    def get_num_converged(self) -> int:
        """Return number of converged eigenpairs."""
        return int(self._nconv)

    def get_eigenvalue(self, idx: int) -> float | complex:
        """Get the eigenvalue at index idx (real number for Hermitian problem types)."""
        if not 0 <= idx < self._nconv:
            raise IndexError(f"eigenvalue index {idx} out of range (converged: {self._nconv})")
        lam = complex(self._eigenvalues[idx])
        if self._problem_type in _HERMITIAN:
            return float(lam.real)
        return lam

    def _fetch_vectors(self) -> np.ndarray:
        if self._eigenvectors is None:
            if self._handle.closed or self._handle.gen_result != self._result_gen:
                raise RuntimeError("the device-side results of this solver were overwritten by a later solve on the "
                                   "shared handle; call solve() again")
            X = self._handle.eigenvectors(self._nconv, capacity=max(self._nconv, self._nev))
            # (the arbitrary phase is fixed on the device: largest component real positive)
            if self._problem_type in (iEpsProblemType.GHEP,) and self._M is not None and getattr(self, "_b_mode", 0) == 0:
                # SLEPc normalises GHEP eigenvectors to unit B-norm (done on the device, lsa_eigs_params.b_mode, except
                # for the partitioned solve, whose GPUs hold row blocks of M)
                Mh = _as_csr(self._M)
                for i in range(X.shape[1]):
                    bn = np.sqrt(abs(np.vdot(X[:, i], Mh @ X[:, i])))
                    if bn > 0:
                        X[:, i] /= bn
            self._eigenvectors = X
        return self._eigenvectors

    def get_eigenvector(self, idx: int) -> iComplexPETScVector:
        """Get the eigenvector at index idx (unit 2-norm; B-norm for GHEP).

        Complex mode (complex shift or data; the reference's complex PETSc build): one complex vector,
        `.imag is None` (`Solver/utils.py:293-297`).  Real mode (the real build): `(vr, vi)` with `vi`
        dropped when `||vi|| <= 1e-6` (`Solver/utils.py:280-291`).
        """
        if not 0 <= idx < self._nconv:
            raise IndexError(f"eigenvector index {idx} out of range (converged: {self._nconv})")
        x = self._fetch_vectors()[:, idx]
        from .carriers import _RawVec

        first = idx not in self._handed_out
        self._handed_out.add(idx)

        def wrap(a: np.ndarray) -> iPETScVector:
            # the first request for an index hands out the solver's own slice of the result buffer (no
            # copy of n numbers); later requests for the same index get fresh copies, as the reference does
            return iPETScVector(_RawVec(a if first and a.flags.c_contiguous else np.array(a, copy=True)))

        if self._complex_mode:
            return iComplexPETScVector(wrap(x))
        vi = x.imag
        if np.linalg.norm(vi) <= 1e-6:
            return iComplexPETScVector(wrap(x.real))
        return iComplexPETScVector(wrap(x.real), wrap(vi))

    def get_eigenpair(self, idx: int) -> tuple[float | complex, iComplexPETScVector]:
        """Get (eigenvalue, eigenvector) tuple at index idx."""
        return self.get_eigenvalue(idx), self.get_eigenvector(idx)

    def get_all_eigenpairs_up_to(self, num: int) -> Iterator[tuple[float | complex, iComplexPETScVector]]:
        """Lazily yield up to `num` converged eigenpairs."""
        limit = min(self.get_num_converged(), num)
        for i in range(limit):
            yield self.get_eigenpair(i)

    def get_residuals(self) -> np.ndarray:
        """||A x - λ M x|| / (||A||_F ||x||) of every converged pair, evaluated on the device."""
        if self._nconv == 0:
            return np.zeros(0)
        if self._handle.closed or self._handle.gen_result != self._result_gen:
            raise RuntimeError("the device-side results of this solver were overwritten by a later solve on the "
                               "shared handle (same sparsity pattern); call solve() again before get_residuals()")
        return self._handle.residuals(self._nconv)

    def release(self) -> None:
        """Free the device memory behind this solver now (the handle also leaves the symbolic cache).  Other
        solver objects that share the handle need a new `solve()` afterwards."""
        h = self._handle
        if h is None:
            return
        _SYM_CACHE.drop(h)
        h.close()
        self._handle = None
        self._factor_key = None
