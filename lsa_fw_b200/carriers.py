"""PETSc-free carriers with the surface of LSA-FW's `FEM/utils.py` wrappers used on the eigen path.

Only what crosses the eigensolver boundary is mirrored:

* `iPETScMatrix` (reference `FEM/utils.py:104-659`): the INPUT carrier.  Here it wraps a SciPy CSR
  matrix; `shape`, `raw`, `as_scipy_array()`, `from_path` (MatrixMarket), `from_matrix`, `zeros`,
  `norm`, `nonzero_entries`, `is_numerically_hermitian`, `T`, `H`, item access, `pin_dof`, `axpy`,
  `duplicate`, `export` keep the reference's names and argument meaning.
* `iPETScVector` / `iComplexPETScVector` (reference `FEM/utils.py:662-908`, `:911-1244`): the OUTPUT
  carriers.  `.real`, `.imag`, `.norm()`, `.dot()`, `.scale()`, `.copy()`, `.as_array()`,
  `.real.raw.getArray(readonly=True)` behave as in the reference, including its two build flavours:
  in "complex mode" a single complex vector is returned and `.imag` is None
  (`Solver/utils.py:293-297`), in "real mode" a (real, imag) pair with the imaginary part dropped
  when its norm is <= 1e-6 (`Solver/utils.py:280-291`).

Any object exposing `shape` and `as_scipy_array()` (e.g. the reference's own iPETScMatrix when
petsc4py is installed) is accepted by the eigensolver; these classes exist so that the package is
usable, and testable, without PETSc.
"""

from __future__ import annotations

from pathlib import Path

import numpy as np
import scipy.io
import scipy.sparse as sp


class _RawVec:
    """Stand-in for `PETSc.Vec`: the handful of methods callers of the eigen path touch."""

    def __init__(self, array: np.ndarray) -> None:
        self._a = array

    def getArray(self, readonly: bool = False) -> np.ndarray:  # noqa: N802 (PETSc spelling)
        if readonly:
            v = self._a.view()
            v.flags.writeable = False
            return v
        return self._a

    def getSize(self) -> int:  # noqa: N802
        return int(self._a.size)

    def norm(self) -> float:
        return float(np.linalg.norm(self._a))

    def copy(self) -> "_RawVec":
        return _RawVec(self._a.copy())


class iPETScVector:  # noqa: N801 (reference spelling)
    """Dense vector (reference `FEM/utils.py:662-908`), NumPy-backed."""

    def __init__(self, vec) -> None:
        if isinstance(vec, iPETScVector):
            vec = vec._raw
        if isinstance(vec, _RawVec):
            self._raw = vec
        else:
            self._raw = _RawVec(np.array(vec, copy=True).ravel())

    @classmethod
    def zeros(cls, size: int, comm=None, dtype=np.float64) -> "iPETScVector":
        return cls(np.zeros(size, dtype=dtype))

    @classmethod
    def from_array(cls, array: np.ndarray, comm=None) -> "iPETScVector":
        return cls(np.asarray(array))

    @property
    def raw(self) -> _RawVec:
        return self._raw

    @property
    def size(self) -> int:
        return self._raw.getSize()

    @property
    def norm(self) -> float:
        return self._raw.norm()

    def as_array(self) -> np.ndarray:
        return self._raw.getArray().copy()

    def copy(self) -> "iPETScVector":
        return iPETScVector(self._raw.copy())

    def scale(self, alpha) -> None:
        a = self._raw._a
        if np.iscomplexobj(alpha) and not np.iscomplexobj(a) and complex(alpha).imag != 0.0:
            self._raw._a = a.astype(np.complex128) * alpha
        else:
            a *= alpha.real if (np.iscomplexobj(alpha) and not np.iscomplexobj(a)) else alpha

    def dot(self, other: "iPETScVector"):
        """PETSc `VecDot(x, y) = y^H x` (`FEM/utils.py:894-896`): the ARGUMENT is conjugated."""
        return np.vdot(other._raw._a, self._raw._a)

    def axpy(self, alpha, other: "iPETScVector") -> None:
        self._raw._a += alpha * other._raw._a

    def __getitem__(self, i: int):
        return self._raw._a[i]

    def __setitem__(self, i: int, v) -> None:
        self._raw._a[i] = v

    def __mul__(self, other):
        if isinstance(other, iPETScVector):
            return self.dot(other)
        return iPETScVector(self._raw._a * other)

    __rmul__ = __mul__

    def __add__(self, other: "iPETScVector") -> "iPETScVector":
        return iPETScVector(self._raw._a + other._raw._a)

    def __sub__(self, other: "iPETScVector") -> "iPETScVector":
        return iPETScVector(self._raw._a - other._raw._a)


class iComplexPETScVector:  # noqa: N801
    """Complex vector with optional imaginary part (reference `FEM/utils.py:911-1244`)."""

    def __init__(self, real, imag=None) -> None:
        self._real = real if isinstance(real, iPETScVector) else iPETScVector(real)
        self._imag = None if imag is None else (imag if isinstance(imag, iPETScVector) else iPETScVector(imag))

    @classmethod
    def from_array(cls, data: np.ndarray, comm=None) -> "iComplexPETScVector":
        data = np.asarray(data).ravel()
        if np.iscomplexobj(data):
            return cls(data.real.copy(), data.imag.copy())
        return cls(data.copy())

    @property
    def real(self) -> iPETScVector:
        return self._real

    @property
    def imag(self) -> iPETScVector | None:
        return self._imag

    @property
    def is_complex(self) -> bool:
        return self._imag is not None or np.iscomplexobj(self._real.raw.getArray())

    @property
    def size(self) -> int:
        return self._real.size

    def as_array(self) -> np.ndarray:
        r = self._real.raw.getArray()
        if self._imag is None:
            return r.copy()
        return r.astype(np.complex128) + 1j * self._imag.raw.getArray()

    def norm(self) -> float:
        """Euclidean norm (`FEM/utils.py:1183-1192`)."""
        if self._imag is None:
            return self._real.norm
        return float(np.hypot(self._real.norm, self._imag.norm))

    def dot(self, other) -> complex:
        """Hermitian inner product conjugating `self` (`FEM/utils.py:1194-1212`, real-build branch).

        In the reference's complex build the call falls through to `VecDot(self, other)`, which
        conjugates the ARGUMENT instead (`FEM/utils.py:1205-1206`); that branch is reproduced when this
        vector carries complex data in its single `real` part.
        """
        if isinstance(other, iPETScVector):
            other = iComplexPETScVector(other)
        if not isinstance(other, iComplexPETScVector):
            raise TypeError("Dot product requires a iComplexPETScVector.")
        if self._imag is None and np.iscomplexobj(self._real.raw.getArray()):
            return self._real.dot(other.real)  # complex-build semantics
        return complex(np.vdot(self.as_array(), other.as_array()))

    def scale(self, scalar) -> None:
        """In-place scaling by a real or complex scalar (`FEM/utils.py:1214-1238`)."""
        arr = self._real.raw.getArray()
        if self._imag is None and np.iscomplexobj(arr):
            arr *= scalar
            return
        z = self.as_array() * scalar
        if np.iscomplexobj(z) and (self._imag is not None or np.any(z.imag != 0.0)):
            self._real = iPETScVector(z.real.copy())
            self._imag = iPETScVector(z.imag.copy())
        else:
            self._real = iPETScVector(np.real(z).copy())

    def copy(self) -> "iComplexPETScVector":
        return iComplexPETScVector(self._real.copy(), None if self._imag is None else self._imag.copy())

    def __getitem__(self, i: int):
        if self._imag is None:
            return self._real[i]
        return complex(self._real[i], self._imag[i])


class _RawMat:
    """Stand-in for `PETSc.Mat` exposing the calls the eigen path and its callers use."""

    def __init__(self, owner: "iPETScMatrix") -> None:
        self._o = owner

    def getSize(self):  # noqa: N802
        return self._o._m.shape

    def getValuesCSR(self):  # noqa: N802
        m = self._o._csr()
        return m.indptr, m.indices, m.data

    def norm(self) -> float:
        return self._o.norm

    def createVecRight(self):  # noqa: N802
        return _RawVec(np.zeros(self._o.shape[1], dtype=self._o._m.dtype))


class iPETScMatrix:  # noqa: N801
    """Sparse matrix carrier (reference `FEM/utils.py:104-659`), SciPy-CSR-backed."""

    def __init__(self, mat) -> None:
        if isinstance(mat, iPETScMatrix):
            mat = mat._m
        self._m = sp.csr_matrix(mat)
        self._adjoint_of: "iPETScMatrix | None" = None
        self._nullspace: "iPETScNullSpace | None" = None

    # -- constructors
    @classmethod
    def from_path(cls, path: Path, comm=None) -> "iPETScMatrix":
        """MatrixMarket ingest (`FEM/utils.py:143-147`), without the O(nnz) setValue loop."""
        return cls(scipy.io.mmread(str(path)).tocsr())

    @classmethod
    def load(cls, path: Path, comm=None) -> "iPETScMatrix":
        """PETSc binary ingest (`FEM/utils.py:222-230`: `PETSc.Mat().load(viewer)`) without PETSc: the MATAIJ binary
        layout is big-endian int32 {classid 1211216, rows, cols, nnz}, per-row counts, column indices, then the
        values (float64, or interleaved re/im float64 for the complex build -- told apart by the file size).  The
        arrays go straight into CSR: no per-entry `setValue` loop."""
        raw = np.fromfile(str(path), dtype=np.uint8)
        hdr = raw[:16].view(">i4")
        if len(raw) < 16 or int(hdr[0]) != 1211216:
            raise ValueError(f"{path}: not a PETSc binary matrix (MAT_FILE_CLASSID missing)")
        m, n, nnz = int(hdr[1]), int(hdr[2]), int(hdr[3])
        if nnz < 0:
            raise ValueError(f"{path}: dense / special PETSc binary layouts are not supported")
        off = 16
        counts = raw[off: off + 4 * m].view(">i4").astype(np.int64)
        off += 4 * m
        cols = raw[off: off + 4 * nnz].view(">i4").astype(np.int32)
        off += 4 * nnz
        rest = len(raw) - off
        if rest == 8 * nnz:
            vals = raw[off:].view(">f8").astype(np.float64)
        elif rest == 16 * nnz:
            vals = raw[off:].view(">f8").astype(np.float64).view(np.complex128)
        else:
            raise ValueError(f"{path}: value block of {rest} bytes fits neither float64 nor complex128 for nnz = {nnz} "
                             "(64-bit-index PETSc builds are not supported)")
        indptr = np.concatenate([[0], np.cumsum(counts)])
        if int(indptr[-1]) != nnz:
            raise ValueError(f"{path}: row counts do not add up to nnz")
        out = cls(sp.csr_matrix((vals, cols, indptr), shape=(m, n)))
        out._m.sort_indices()
        return out

    @classmethod
    def from_matrix(cls, matrix, comm=None) -> "iPETScMatrix":
        """From a dense array or SciPy sparse matrix (`FEM/utils.py:183-220`)."""
        if sp.issparse(matrix):
            return cls(matrix.tocsr())
        arr = np.asarray(matrix)
        if arr.ndim != 2:
            raise ValueError("Input array must be 2D.")
        return cls(sp.csr_matrix(arr.astype(np.complex128 if np.iscomplexobj(arr) else np.float64)))

    @classmethod
    def zeros(cls, shape: tuple[int, int], comm=None, nnz=None) -> "iPETScMatrix":
        return cls(sp.csr_matrix(shape, dtype=np.float64))

    # -- properties
    @property
    def raw(self) -> _RawMat:
        return _RawMat(self)

    @property
    def shape(self) -> tuple[int, int]:
        return self._m.shape

    @property
    def nonzero_entries(self) -> int:
        return int(self._m.nnz)

    @property
    def norm(self) -> float:
        """Frobenius norm (`FEM/utils.py:400-403`)."""
        return float(np.sqrt(np.sum(np.abs(self._m.data) ** 2)))

    @property
    def type(self) -> str:
        return "seqaij"

    @property
    def T(self) -> "iPETScMatrix":  # noqa: N802
        return iPETScMatrix(self._m.T.tocsr())

    @property
    def H(self) -> "iPETScMatrix":  # noqa: N802
        """Hermitian transpose.  The result remembers its origin, which lets the eigensolver run the
        adjoint problem (`Sensitivity/__init__.py:246-262`) on the factors of the direct one."""
        out = iPETScMatrix(self._m.conj().T.tocsr())
        out._adjoint_of = self
        out._nullspace = self._nullspace
        return out

    def is_numerically_symmetric(self, tol: float = 1e-6) -> bool:
        d = self._m - self._m.T
        return float(np.sqrt(np.sum(np.abs(d.data) ** 2))) < tol

    def is_numerically_hermitian(self, tol: float = 1e-4) -> bool:
        """`||A - A^H||_F < tol` (`FEM/utils.py:436-448`)."""
        d = self._m - self._m.conj().T
        return float(np.sqrt(np.sum(np.abs(d.data) ** 2))) < tol

    # -- element access / mutation (used by the reference's tests to build tiny matrices)
    def __getitem__(self, idx: tuple[int, int]):
        return self._m[idx]

    def __setitem__(self, idx: tuple[int, int], value) -> None:
        lil = self._m.tolil()
        if np.iscomplexobj(value) and not np.iscomplexobj(lil):
            lil = lil.astype(np.complex128)
        lil[idx] = value
        self._m = lil.tocsr()

    def assemble(self) -> None:
        self._m.sum_duplicates()

    def zero_all_entries(self) -> None:
        self._m = sp.csr_matrix(self._m.shape, dtype=self._m.dtype)

    def add_value(self, row: int, col: int, value) -> None:
        self[row, col] = self._m[row, col] + value

    def get_value(self, row: int, col: int):
        return self._m[row, col]

    def scale(self, alpha) -> "iPETScMatrix":
        self._m = (self._m * alpha).tocsr()
        return self

    def axpy(self, alpha, other: "iPETScMatrix") -> None:
        self._m = (self._m + alpha * other._m).tocsr()

    def duplicate(self, copy: bool = False) -> "iPETScMatrix":
        out = iPETScMatrix(self._m.copy())
        if not copy:
            out._m.data[:] = 0
        return out

    def zero_row_columns(self, rows, diag=0.0) -> None:
        rows = np.asarray(list(rows), dtype=np.int64)
        keep = np.ones(self._m.shape[0])
        keep[rows] = 0.0
        D = sp.diags(keep)
        self._m = (D @ self._m @ D + sp.csr_matrix((np.full(len(rows), diag), (rows, rows)), shape=self._m.shape)).tocsr()

    def attach_nullspace(self, nullspace: "iPETScNullSpace") -> None:
        """Attach a nullspace (`FEM/utils.py:604-607`); the eigensolver then projects it out of every operator
        application instead of requiring a pinned DOF."""
        self._nullspace = nullspace

    def get_nullspace(self) -> "iPETScNullSpace | None":
        return self._nullspace

    def pin_dof(self, index: int) -> None:
        """Zero row and column `index`, unit diagonal (`FEM/utils.py:596-602`)."""
        self.zero_row_columns([index], diag=1.0)

    def as_array(self) -> np.ndarray:
        return self._m.toarray()

    def _csr(self) -> sp.csr_matrix:
        m = self._m
        if not m.has_sorted_indices:
            m.sort_indices()
        return m

    def as_scipy_array(self) -> sp.csr_matrix:
        """CSR view `(data, indices, indptr)` (`FEM/utils.py:585-588`)."""
        return self._csr()

    def export(self, path: Path) -> None:
        """Export (`FEM/utils.py:616-659`): `*.mtx` -> MatrixMarket, anything else -> PETSc binary (MATAIJ layout,
        32-bit indices; real or complex values as the data are), readable by `PETSc.Mat().load` of the matching
        build and by `iPETScMatrix.load`."""
        path = Path(path)
        path.parent.mkdir(parents=True, exist_ok=True)
        if path.suffix.lower() == ".mtx":
            scipy.io.mmwrite(str(path), self._m)
            return
        m = self._csr()
        with open(path, "wb") as f:
            np.array([1211216, m.shape[0], m.shape[1], m.nnz], dtype=">i4").tofile(f)
            np.diff(m.indptr).astype(">i4").tofile(f)
            m.indices.astype(">i4").tofile(f)
            if np.iscomplexobj(m.data):
                np.ascontiguousarray(m.data, dtype=np.complex128).view(np.float64).astype(">f8").tofile(f)
            else:
                np.asarray(m.data, dtype=np.float64).astype(">f8").tofile(f)

    def __matmul__(self, other):
        if isinstance(other, iPETScVector):
            return iPETScVector(self._m @ other.raw.getArray())
        if isinstance(other, iPETScMatrix):
            return iPETScMatrix(self._m @ other._m)
        return NotImplemented

    def __str__(self) -> str:
        return f"iPETScMatrix(shape={self.shape}, nnz={self.nonzero_entries})"


class iPETScNullSpace:  # noqa: N801
    """Nullspace carrier (reference `FEM/utils.py:1247-1380`): a list of basis vectors, orthonormalised on creation."""

    def __init__(self, vectors: list, constant: bool = False, size: int | None = None) -> None:
        cols = [np.asarray(v.raw.getArray() if hasattr(v, "raw") else v) for v in vectors]
        if constant:
            if size is None and not cols:
                raise ValueError("a constant nullspace needs a size")
            cols.insert(0, np.ones(size if size is not None else len(cols[0])))
        if not cols:
            raise ValueError("Cannot create NullSpace from empty vector list")
        B = np.stack([np.asarray(c, dtype=np.complex128) for c in cols], axis=1)
        Q, R = np.linalg.qr(B)
        if np.min(np.abs(np.diag(R))) < 1e-12 * max(1.0, np.max(np.abs(np.diag(R)))):
            raise ValueError("nullspace basis vectors are linearly dependent")
        self._Q = Q
        self._constant = constant

    @classmethod
    def from_vectors(cls, vectors: list) -> "iPETScNullSpace":
        return cls(list(vectors))

    @classmethod
    def create_constant(cls, size: int, comm=None) -> "iPETScNullSpace":
        return cls([], constant=True, size=size)

    @property
    def dimension(self) -> int:
        return self._Q.shape[1]

    @property
    def basis(self) -> list[iPETScVector]:
        return [iPETScVector.from_array(self._Q[:, i].copy()) for i in range(self._Q.shape[1])]

    def has_constant(self) -> bool:
        return self._constant

    def as_array(self) -> np.ndarray:
        """(n, dimension) orthonormal basis."""
        return self._Q

    def test_matrix(self, mat: iPETScMatrix, tol: float = 1e-12) -> tuple[bool, float]:
        nrm = float(np.max(np.linalg.norm(mat.as_scipy_array() @ self._Q, axis=0)))
        return nrm < tol, nrm

    def remove(self, vec: iPETScVector) -> None:
        a = vec.raw.getArray()
        a[...] = a - (self._Q @ (self._Q.conj().T @ a)).astype(a.dtype if np.iscomplexobj(a) else np.float64)

    def attach_to(self, mat: iPETScMatrix) -> None:
        mat.attach_nullspace(self)

    def detach_from(self, mat: iPETScMatrix) -> None:
        mat._nullspace = None
