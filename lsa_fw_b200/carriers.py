"""PETSc-free carriers with the surface of LSA-FW's `FEM/utils.py` wrappers used on the eigen path.

The classes carry the full public surface of the reference's wrappers (every method and property of
`iPETScMatrix`, `iPETScVector`, `iComplexPETScVector`, `iPETScNullSpace`, `iPETScBlockMatrix`; the reference's own
carrier tests, `tests/unit/FEM/test_utils.py`, are restated in `tests/test_carriers.py`), on a SciPy / NumPy backing:

* `iPETScMatrix` (reference `FEM/utils.py:104-659`): the INPUT carrier of the eigen path.
  It wraps a SciPy CSR matrix; `shape`, `raw`, `as_scipy_array()`, `from_path` (MatrixMarket), `load` / `export` (PETSc
  binary), `from_matrix`, `from_nested`, `norm`, `T`, `H`, `to_aij`, arithmetic, row / column access, `pin_dof`,
  nullspace attachment keep the reference's names, argument meaning and error behaviour.  `raw` is a small stand-in for
  `PETSc.Mat` (`mult`, `createVecRight`, `getValuesCSR`, ...: the calls `Sensitivity/__init__.py:281-283` makes).
* `iPETScVector` / `iComplexPETScVector` (reference `FEM/utils.py:662-908`, `:911-1244`): the OUTPUT carriers (row a8).
  `.real`, `.imag`, `.norm()`, `.dot()`, `.scale()`, `.copy()`, `.as_array()`,
  `.real.raw.getArray(readonly=True)` behave as in the reference, including its two build flavours:
  in "complex mode" a single complex vector is returned and `.imag` is None
  (`Solver/utils.py:293-297`), in "real mode" a (real, imag) pair with the imaginary part dropped
  when its norm is <= 1e-6 (`Solver/utils.py:280-291`).
* one process: `.comm` is a serial communicator stand-in; a solve that is split over GPUs shards inside the CUDA library.

Any object exposing `shape` and `as_scipy_array()` (e.g. the reference's own iPETScMatrix when
petsc4py is installed) is accepted by the eigensolver; these classes exist so that the package is
usable, and testable, without PETSc.
"""

from __future__ import annotations

from pathlib import Path

import numpy as np
import scipy.io
import scipy.sparse as sp


class _SerialComm:
    """Stand-in for `PETSc.Comm` of a one-process run (`.comm` of the reference's wrappers): the carriers live in one
    process; a solve that is split over GPUs shards inside the CUDA library, not through these objects."""

    size = 1
    rank = 0

    def getSize(self) -> int:  # noqa: N802 (PETSc spelling)
        return 1

    def getRank(self) -> int:  # noqa: N802
        return 0

    Get_size = getSize
    Get_rank = getRank

    def barrier(self) -> None:
        return None

    def __repr__(self) -> str:
        return "<serial communicator>"


COMM_SELF = _SerialComm()

_VEC_FILE_CLASSID = 1211214   # PETSc binary Vec: big-endian int32 {classid, n}, then n scalars (big-endian)
_MAT_FILE_CLASSID = 1211216


class _RawVec:
    """Stand-in for `PETSc.Vec`: the handful of methods callers of the eigen path touch."""

    comm = COMM_SELF

    def __init__(self, array: np.ndarray) -> None:
        self._a = array

    def getComm(self) -> _SerialComm:  # noqa: N802
        return COMM_SELF

    def getValue(self, i: int):  # noqa: N802
        return self._a[i]

    def setValue(self, i: int, value) -> None:  # noqa: N802
        if np.iscomplexobj(value) and not np.iscomplexobj(self._a):
            if complex(value).imag != 0.0:
                self._a = self._a.astype(np.complex128)
            else:
                value = complex(value).real
        self._a[i] = value

    def duplicate(self) -> "_RawVec":
        return _RawVec(np.zeros_like(self._a))

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

    @classmethod
    def create_seq(cls, size: int, comm=None) -> "iPETScVector":
        """Sequential vector of the given size (`FEM/utils.py:700-705`), zero-filled here."""
        return cls.zeros(size)

    @classmethod
    def from_file(cls, path: Path, comm=None) -> "iPETScVector":
        """PETSc binary Vec ingest (`FEM/utils.py:707-717`) without PETSc: real or complex scalars, told apart by the
        file size."""
        raw = np.fromfile(str(path), dtype=np.uint8)
        hdr = raw[:8].view(">i4")
        if len(raw) < 8 or int(hdr[0]) != _VEC_FILE_CLASSID:
            raise ValueError(f"{path}: not a PETSc binary vector (VEC_FILE_CLASSID missing)")
        n, rest = int(hdr[1]), len(raw) - 8
        if rest == 8 * n:
            return cls(raw[8:].view(">f8").astype(np.float64))
        if rest == 16 * n:
            return cls(raw[8:].view(">f8").astype(np.float64).view(np.complex128))
        raise ValueError(f"{path}: value block of {rest} bytes fits neither float64 nor complex128 for n = {n}")

    @property
    def raw(self) -> _RawVec:
        return self._raw

    @property
    def comm(self) -> _SerialComm:
        return COMM_SELF

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

    def duplicate(self) -> "iPETScVector":
        """New vector of the same size and scalar type (`FEM/utils.py:842-845`); PETSc leaves the values
        unspecified, here they are zero."""
        return iPETScVector(self._raw.duplicate())

    def assemble(self) -> None:
        return None

    def ghost_update(self, addv=None, mode=None) -> None:
        """No ghost entries in a one-process vector (`FEM/utils.py:871-877`)."""
        return None

    def zero_all_entries(self) -> None:
        self._raw._a[...] = 0

    def get_value(self, i: int):
        return self._raw.getValue(i)

    def set_value(self, i: int, value) -> None:
        self._raw.setValue(i, value)

    def set_array(self, array: np.ndarray) -> None:
        """Replace the values (`FEM/utils.py:883-885`); the size must match, as with `VecSetArray`."""
        array = np.asarray(array).ravel()
        if array.size != self.size:
            raise ValueError(f"array of size {array.size} for a vector of size {self.size}")
        if np.iscomplexobj(array) and not np.iscomplexobj(self._raw._a):
            self._raw._a = array.astype(np.complex128)
        else:
            self._raw._a[...] = array

    def set_random(self, rng=None) -> None:
        """Uniform values in [0, 1) (+ i [0, 1) for complex data), PETSc's default `PetscRandom` interval
        (`FEM/utils.py:887-892`); `rng`: a `numpy.random.Generator` or a seed."""
        g = rng if isinstance(rng, np.random.Generator) else np.random.default_rng(rng)
        a = self._raw._a
        if np.iscomplexobj(a):
            a[...] = g.random(a.size) + 1j * g.random(a.size)
        else:
            a[...] = g.random(a.size)

    def print(self) -> None:
        print(f"iPETScVector(size={self.size})\n{self._raw._a}")

    def export(self, path: Path) -> None:
        """PETSc binary Vec (`FEM/utils.py:901-908`), readable by `PETSc.Vec().load` of the matching build and by
        `iPETScVector.from_file`."""
        path = Path(path)
        path.parent.mkdir(parents=True, exist_ok=True)
        a = self._raw._a
        with open(path, "wb") as f:
            np.array([_VEC_FILE_CLASSID, a.size], dtype=">i4").tofile(f)
            if np.iscomplexobj(a):
                np.ascontiguousarray(a, dtype=np.complex128).view(np.float64).astype(">f8").tofile(f)
            else:
                np.asarray(a, dtype=np.float64).astype(">f8").tofile(f)

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
        if not isinstance(other, iPETScVector):
            return NotImplemented
        if self.size != other.size:
            raise ValueError(f"Incompatible vector sizes: {self.size} vs {other.size}")
        return iPETScVector(self._raw._a + other._raw._a)

    __radd__ = __add__

    def __sub__(self, other: "iPETScVector") -> "iPETScVector":
        if not isinstance(other, iPETScVector):
            return NotImplemented
        if self.size != other.size:
            raise ValueError(f"Incompatible vector sizes: {self.size} vs {other.size}")
        return iPETScVector(self._raw._a - other._raw._a)

    def __matmul__(self, other):
        """Outer product `x y^T` (no conjugation) as a dense-pattern matrix (`FEM/utils.py:773-795`).  Any other right
        operand is left to its `__rmatmul__` (a matrix there gives `A^T x`, `FEM/utils.py:301-313`)."""
        if not isinstance(other, iPETScVector):
            return NotImplemented
        return iPETScMatrix(sp.csr_matrix(np.outer(self._raw._a, other._raw._a)))

    def __eq__(self, other: object):
        """`||x - y|| < 1e-12` (`FEM/utils.py:806-816`)."""
        if not isinstance(other, iPETScVector):
            return NotImplemented
        if self.size != other.size:
            return False
        return bool(np.linalg.norm(self._raw._a - other._raw._a) < 1e-12)

    __hash__ = object.__hash__


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

    def get_value(self, index: int):
        """`FEM/utils.py:1005-1009`."""
        if self._imag is None:
            return self._real[index]
        return complex(self._real[index], self._imag[index])

    def set_value(self, index: int, value) -> None:
        """`FEM/utils.py:1011-1031`: a vector that holds complex data in one part takes the value as it is; a
        (real, imag) pair grows an imaginary part when a value needs one and zeroes it for a real value."""
        if self._imag is None and np.iscomplexobj(self._real.raw.getArray()):
            self._real.set_value(index, value)
            return
        z = complex(value)
        self._real.set_value(index, z.real)
        if z.imag != 0.0 or self._imag is not None:
            if self._imag is None:
                self._imag = self._real.duplicate()
            self._imag.set_value(index, z.imag)

    def __getitem__(self, i: int):
        return self.get_value(i)

    def __setitem__(self, i: int, value) -> None:
        self.set_value(i, value)

    def assemble(self) -> None:
        return None

    def _like(self, z: np.ndarray, partner: "iComplexPETScVector | None" = None) -> "iComplexPETScVector":
        """Result of an arithmetic operation in this vector's flavour: one complex part (complex build), or a
        (real, imag) pair whose imaginary part exists when an operand had one or the result needs one."""
        if self._imag is None and np.iscomplexobj(self._real.raw.getArray()):
            return iComplexPETScVector(np.asarray(z, dtype=np.complex128))
        had_imag = self._imag is not None or (partner is not None and partner.is_complex)
        if np.iscomplexobj(z) and (had_imag or np.any(z.imag != 0.0)):
            return iComplexPETScVector(z.real.copy(), z.imag.copy())
        return iComplexPETScVector(np.real(z).copy())

    def __add__(self, other):
        """`FEM/utils.py:1041-1061`."""
        if isinstance(other, iPETScVector):
            other = iComplexPETScVector(other)
        if not isinstance(other, iComplexPETScVector):
            return NotImplemented
        if self.size != other.size:
            raise ValueError(f"Incompatible vector sizes: {self.size} vs {other.size}")
        return self._like(self.as_array() + other.as_array(), other)

    def __sub__(self, other):
        """`FEM/utils.py:1063-1081`."""
        if isinstance(other, iPETScVector):
            other = iComplexPETScVector(other)
        if not isinstance(other, iComplexPETScVector):
            return NotImplemented
        if self.size != other.size:
            raise ValueError(f"Incompatible vector sizes: {self.size} vs {other.size}")
        return self._like(self.as_array() - other.as_array(), other)

    def __mul__(self, scalar):
        """Scalar multiple as a new vector (`FEM/utils.py:1083-1108`)."""
        if isinstance(scalar, (bool, np.bool_)) or not isinstance(scalar, (int, float, complex, np.number)):
            return NotImplemented
        return self._like(self.as_array() * scalar)

    __rmul__ = __mul__

    def __matmul__(self, other):
        """`x @ A = A^T x`, part by part (`FEM/utils.py:1123-1142`); vector-vector products are not defined there."""
        if isinstance(other, (iPETScVector, iComplexPETScVector)):
            raise NotImplementedError("Vector-vector cross product not implemented yet.")
        if isinstance(other, iPETScMatrix):
            return self._like(other.as_scipy_array().T @ self.as_array())
        raise NotImplementedError(f"Cannot multiply complex vector by {type(other)}")

    def __rmatmul__(self, other):
        """`A @ x`, part by part (`FEM/utils.py:1152-1168`)."""
        if isinstance(other, (iPETScVector, iComplexPETScVector)):
            return NotImplemented
        if isinstance(other, iPETScMatrix):
            return self._like(other.as_scipy_array() @ self.as_array())
        raise NotImplementedError(f"Cannot multiply {type(other)} by complex vector")

    def __eq__(self, other: object):
        """`||x - y|| < 1e-12` (`FEM/utils.py:1170-1175`); False for any other type."""
        if not isinstance(other, iComplexPETScVector):
            return False
        if self.size != other.size:
            return False
        return bool(np.linalg.norm(self.as_array() - other.as_array()) < 1e-12)

    __hash__ = object.__hash__


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

    def createVecLeft(self):  # noqa: N802
        return _RawVec(np.zeros(self._o.shape[0], dtype=self._o._m.dtype))

    @staticmethod
    def _put(out: _RawVec, y: np.ndarray) -> None:
        if np.iscomplexobj(y) and not np.iscomplexobj(out._a):
            out._a = y.astype(np.complex128)
        else:
            out._a[...] = y

    def mult(self, x: _RawVec, y: _RawVec) -> None:
        """`MatMult`: y = A x, as `Sensitivity/__init__.py:281-283` calls it on `M.raw`."""
        self._put(y, self._o._csr() @ x._a)

    def multTranspose(self, x: _RawVec, y: _RawVec) -> None:  # noqa: N802
        self._put(y, self._o._csr().T @ x._a)

    def multHermitian(self, x: _RawVec, y: _RawVec) -> None:  # noqa: N802
        self._put(y, self._o._csr().conj().T @ x._a)

    def getType(self) -> str:  # noqa: N802
        return self._o.type

    @property
    def comm(self) -> _SerialComm:
        return COMM_SELF


class iPETScMatrix:  # noqa: N801
    """Sparse matrix carrier (reference `FEM/utils.py:104-659`), SciPy-CSR-backed."""

    def __init__(self, mat) -> None:
        if isinstance(mat, iPETScMatrix):
            mat = mat._m
        self._m = sp.csr_matrix(mat)
        self._adjoint_of: "iPETScMatrix | None" = None
        self._nullspace: "iPETScNullSpace | None" = None
        self._kind = "seqaij"                       # PETSc type name: "nest" / "transpose" / "hermitiantranspose" as made
        self._blocks: "list[list[iPETScMatrix | None]] | None" = None

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
        """From a dense array or SciPy sparse matrix (`FEM/utils.py:183-220`); a carrier is returned as it is, the raw
        handle of one is wrapped back into its carrier, anything else raises `TypeError`."""
        if isinstance(matrix, iPETScMatrix):
            return matrix
        if isinstance(matrix, _RawMat):
            return matrix._o
        if sp.issparse(matrix):
            return cls(matrix.tocsr())
        if not isinstance(matrix, (np.ndarray, list, tuple)):
            raise TypeError(f"Unsupported matrix type: {type(matrix)}")
        arr = np.asarray(matrix)
        if arr.ndim != 2:
            raise ValueError("Input array must be 2D.")
        out = cls(sp.csr_matrix(arr.astype(np.complex128 if np.iscomplexobj(arr) else np.float64)))
        out._kind = "seqdense"      # the reference makes a dense PETSc matrix of an array: rows and columns read back in full
        return out

    @classmethod
    def zeros(cls, shape: tuple[int, int], comm=None, nnz=None) -> "iPETScMatrix":
        return cls(sp.csr_matrix(shape, dtype=np.float64))

    @classmethod
    def create_aij(cls, shape: tuple[int, int], comm=None, nnz=None) -> "iPETScMatrix":
        """Empty sparse matrix of the given shape, ready for `add_value` / item assignment (`FEM/utils.py:161-181`;
        the preallocation hint `nnz` has no meaning for the CSR backing)."""
        return cls(sp.csr_matrix(shape, dtype=np.float64))

    @classmethod
    def from_nested(cls, blocks: list, comm=None) -> "iPETScMatrix":
        """Block matrix `[[A, G], [D, None]]` (`FEM/utils.py:118-141`); `None` = empty block.  The blocks are kept for
        `sub(i, j)`, the type reads "nest", and the flat CSR form every consumer of this package works on is built
        once (the reference needs `to_aij()` for that)."""
        if not blocks or not all(isinstance(row, (list, tuple)) for row in blocks):
            raise ValueError("`blocks` must be a non-empty 2D list")
        if any(len(row) != len(blocks[0]) for row in blocks):
            raise ValueError("All block rows must have the same length")
        wrapped = [[None if b is None else cls.from_matrix(b) for b in row] for row in blocks]
        out = cls(sp.bmat([[None if b is None else b._m for b in row] for row in wrapped], format="csr"))
        out._kind = "nest"
        out._blocks = wrapped
        return out

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
        return self._kind

    @property
    def comm(self) -> _SerialComm:
        return COMM_SELF

    @property
    def T(self) -> "iPETScMatrix":  # noqa: N802
        out = iPETScMatrix(self._m.T.tocsr())
        out._kind = "transpose"
        return out

    @property
    def H(self) -> "iPETScMatrix":  # noqa: N802
        """Hermitian transpose.  The result remembers its origin, which lets the eigensolver run the
        adjoint problem (`Sensitivity/__init__.py:246-262`) on the factors of the direct one."""
        out = iPETScMatrix(self._m.conj().T.tocsr())
        out._adjoint_of = self
        out._nullspace = self._nullspace
        out._kind = "hermitiantranspose"
        return out

    @property
    def is_symmetric(self) -> bool:
        """Exact symmetry, pattern and values (`MatIsSymmetric` with zero tolerance, `FEM/utils.py:410-413`)."""
        d = (self._m - self._m.T).tocsr()
        return d.nnz == 0 or not np.any(d.data != 0)

    def is_hermitian(self) -> bool:
        """Exact Hermitian symmetry (`MatIsHermitian`, `FEM/utils.py:429-434`)."""
        d = (self._m - self._m.conj().T).tocsr()
        return d.nnz == 0 or not np.any(d.data != 0)

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

    def shift(self, alpha) -> None:
        """A <- A + alpha I (`MatShift`, `FEM/utils.py:463-465`)."""
        self._m = (self._m + alpha * sp.eye(*self._m.shape, format="csr")).tocsr()

    def sub(self, row: int, col: int) -> "iPETScMatrix | None":
        """(row, col) block of a nested matrix, None for an empty block (`FEM/utils.py:467-477`)."""
        if self._kind != "nest" or self._blocks is None:
            raise NotImplementedError("Submatrix access is only available for nested matrices.")
        return self._blocks[row][col]

    def to_aij(self) -> "iPETScMatrix":
        """Flat AIJ form of a nested / transposed / Hermitian-transposed matrix (`FEM/utils.py:556-565`; plain AIJ
        matrices raise there as well).  The CSR backing is flat already: a copy with the plain type."""
        if self._kind not in ("nest", "hermitiantranspose", "transpose"):
            raise NotImplementedError("Only MatNest matrices can be flattened with `to_aij()`.")
        out = iPETScMatrix(self._m.copy())
        out._nullspace = self._nullspace
        return out

    def get_row(self, row: int) -> tuple[list[int], list]:
        """Column indices and values of one row (`FEM/utils.py:491-505`)."""
        m = self._csr()
        if self._kind == "seqdense":
            return list(range(m.shape[1])), m[[row], :].toarray().ravel().tolist()
        lo, hi = int(m.indptr[row]), int(m.indptr[row + 1])
        return m.indices[lo:hi].tolist(), m.data[lo:hi].tolist()

    def get_column(self, col: int) -> tuple[list[int], list]:
        """Row indices and values of one column (`FEM/utils.py:507-527`), in increasing row order."""
        m = self._csr()
        if self._kind == "seqdense":
            return list(range(m.shape[0])), m[:, [col]].toarray().ravel().tolist()
        hit = np.nonzero(m.indices == col)[0]
        rows = np.searchsorted(m.indptr, hit, side="right") - 1
        return rows.tolist(), m.data[hit].tolist()

    def create_vector_right(self) -> iPETScVector:
        return iPETScVector(self.raw.createVecRight())

    def create_vector_left(self) -> iPETScVector:
        return iPETScVector(self.raw.createVecLeft())

    def print(self) -> None:
        print(f"{self}\n{self._m}")

    def axpy(self, alpha, other: "iPETScMatrix") -> None:
        """this <- alpha * other + this (`FEM/utils.py:529-541`)."""
        if not isinstance(other, iPETScMatrix):
            raise NotImplementedError(f"Cannot add iPETScMatrix with {type(other)}")
        if self.shape != other.shape:
            raise ValueError(f"Incompatible matrix shapes: {self.shape} vs {other.shape}")
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
        if nullspace is not None and hasattr(nullspace, "as_array"):
            nullspace.as_array(self.shape[0])      # a constant-only nullspace takes its size here

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
        """Matrix-vector / matrix-matrix product (`FEM/utils.py:271-299`); a complex vector carrier answers through
        its own `__rmatmul__`."""
        if isinstance(other, iPETScVector):
            if self.shape[1] != other.size:
                raise ValueError(f"Incompatible matrix-vector shapes: {self.shape[1]} vs {other.size}")
            return iPETScVector(self._m @ other.raw.getArray())
        if isinstance(other, iPETScMatrix):
            if self.shape[1] != other.shape[0]:
                raise ValueError(f"Incompatible matrix-matrix shapes: {self.shape} vs {other.shape}")
            return iPETScMatrix(self._m @ other._m)
        return NotImplemented

    def __rmatmul__(self, other):
        """`x @ A = A^T x`, without conjugation (`MatMultTranspose`, `FEM/utils.py:301-313`)."""
        if not isinstance(other, iPETScVector):
            raise NotImplementedError(f"Cannot multiply iPETScMatrix with {type(other)}")
        if self.shape[0] != other.size:
            raise ValueError(f"Incompatible matrix and vector sizes: {self.shape[0]} vs {other.size}")
        return iPETScVector(self._m.T @ other.raw.getArray())

    def _binary(self, other: object, sign: float) -> "iPETScMatrix":
        if not isinstance(other, iPETScMatrix):
            raise NotImplementedError(f"Cannot add iPETScMatrix with {type(other)}")
        if self.shape != other.shape:
            raise ValueError(f"Incompatible matrix shapes: {self.shape} vs {other.shape}")
        return iPETScMatrix((self._m + sign * other._m).tocsr())

    def __add__(self, other: object) -> "iPETScMatrix":
        """`FEM/utils.py:235-246`."""
        return self._binary(other, 1.0)

    def __sub__(self, other: object) -> "iPETScMatrix":
        """`FEM/utils.py:248-259`."""
        return self._binary(other, -1.0)

    def __radd__(self, other: object) -> "iPETScMatrix":
        return self._binary(other, 1.0)

    def __eq__(self, other: object):
        """`||A - B||_F < 1e-12` (`FEM/utils.py:337-351`)."""
        if not isinstance(other, iPETScMatrix):
            return NotImplemented
        if self.shape != other.shape:
            return False
        d = (self._m - other._m).tocsr()
        return bool(np.sqrt(np.sum(np.abs(d.data) ** 2)) < 1e-12)

    __hash__ = object.__hash__       # identity: the solver keys its factor registry by carrier object

    def __str__(self) -> str:
        return f"iPETScMatrix(shape={self.shape}, nnz={self.nonzero_entries})"


class iPETScNullSpace:  # noqa: N801
    """Nullspace carrier (reference `FEM/utils.py:1247-1380`): basis vectors and, optionally, the constant vector.
    The constant vector has no size of its own in PETSc (`create_constant(comm)`); here it takes the size of the
    other vectors, of `size`, or of the first matrix / vector it is used with.  The orthonormal basis the
    eigensolver projects with is built once the size is known."""

    def __init__(self, vectors: list, constant: bool = False, size: int | None = None) -> None:
        self._cols = [np.asarray(v.raw.getArray() if hasattr(v, "raw") else v).ravel().copy() for v in vectors]
        if not self._cols and not constant:
            raise ValueError("Cannot create NullSpace from empty vector list")
        if len({c.size for c in self._cols}) > 1:
            raise ValueError("nullspace basis vectors differ in size")
        self._constant = constant
        self._Q: np.ndarray | None = None
        n = self._cols[0].size if self._cols else size
        if n is not None:
            self._build(int(n))

    def _build(self, n: int) -> np.ndarray:
        if self._Q is not None and self._Q.shape[0] == n:
            return self._Q
        if self._cols and self._cols[0].size != n:
            raise ValueError(f"nullspace vectors of size {self._cols[0].size} used with size {n}")
        cols = ([np.ones(n)] if self._constant else []) + self._cols
        B = np.stack([np.asarray(c, dtype=np.complex128) for c in cols], axis=1)
        Q, R = np.linalg.qr(B)
        d = np.abs(np.diag(R))
        keep = d >= 1e-12 * max(1.0, float(d.max()))
        if not self._constant and not np.all(keep):
            raise ValueError("nullspace basis vectors are linearly dependent")
        # with the constant present, a further vector may repeat it (PETSc accepts that): dependent columns are dropped
        self._Q = np.ascontiguousarray(Q[:, keep])
        return self._Q

    @classmethod
    def from_vectors(cls, vectors: list) -> "iPETScNullSpace":
        """Basis vectors, no constant (`FEM/utils.py:1291-1305`)."""
        if not vectors:
            raise ValueError("Cannot create NullSpace from empty vector list")
        if not all(isinstance(v, iPETScVector) for v in vectors):
            raise TypeError("from_vectors requires a list of iPETScVector")
        return cls(list(vectors))

    @classmethod
    def create_constant(cls, size=None, comm=None) -> "iPETScNullSpace":
        """Constant vector only (`FEM/utils.py:1307-1311`: `create_constant(comm)`); `size` is optional -- without it
        the nullspace sizes itself on first use.  A communicator passed first (the reference's positional form) is
        accepted."""
        if size is not None and not isinstance(size, (int, np.integer)):
            size = None
        return cls([], constant=True, size=size)

    @classmethod
    def create_constant_and_vectors(cls, comm=None, vectors: list | None = None, size: int | None = None) -> "iPETScNullSpace":
        """Constant vector + further basis vectors (`FEM/utils.py:1313-1330`)."""
        if not vectors:
            return cls.create_constant(size)
        if not all(isinstance(v, iPETScVector) for v in vectors):
            raise TypeError("create_constant_and_vectors requires a list of iPETScVector or None")
        return cls(list(vectors), constant=True, size=size)

    def __repr__(self) -> str:
        info = f"{self.dimension}-vector"
        if self._constant:
            info = "constant + " + info
        return f"<iPETScNullSpace {info}, comm={COMM_SELF}>"

    @property
    def raw(self) -> "iPETScNullSpace":
        """There is no PETSc object underneath; the carrier stands for it."""
        return self

    @property
    def comm(self) -> _SerialComm:
        return COMM_SELF

    @property
    def dimension(self) -> int:
        """Number of basis vectors including the constant (`FEM/utils.py:1280-1283`); once the orthonormal basis
        exists, its rank (a vector that repeats the constant does not count twice)."""
        if self._Q is not None:
            return self._Q.shape[1]
        return len(self._cols) + (1 if self._constant else 0)

    @property
    def basis(self) -> list[iPETScVector]:
        """The vectors the nullspace was made from, without the constant (`FEM/utils.py:1285-1289`)."""
        return [iPETScVector.from_array(c.copy()) for c in self._cols]

    def has_constant(self) -> bool:
        return self._constant

    def as_array(self, size: int | None = None) -> np.ndarray:
        """(n, dimension) orthonormal basis, constant included; `size` settles a constant-only nullspace."""
        if self._Q is None:
            if size is None:
                raise ValueError("the size of a constant-only nullspace is not known yet: pass `size`")
            return self._build(int(size))
        return self._Q if size is None else self._build(int(size))

    def test_vector(self, mat: "iPETScMatrix", vec: iPETScVector, tol: float = 1e-12) -> tuple[bool, float]:
        """`||A x|| < tol` for one vector (`FEM/utils.py:1336-1344`)."""
        if not isinstance(vec, iPETScVector):
            raise TypeError("test_vector requires an iPETScVector")
        nrm = float((mat @ vec).norm)
        return nrm < tol, nrm

    def test_matrix(self, mat: "iPETScMatrix", tol: float = 1e-12) -> tuple[bool, float]:
        """`||A x|| < tol` for every basis vector, the constant included (`FEM/utils.py:1346-1355`)."""
        Q = self._build(mat.shape[1])
        nrm = float(np.max(np.linalg.norm(mat.as_scipy_array() @ Q, axis=0)))
        return nrm < tol, nrm

    def remove(self, vec: iPETScVector) -> None:
        """Project the nullspace out of `vec` in place (`FEM/utils.py:1357-1365`)."""
        if not isinstance(vec, iPETScVector):
            raise TypeError("remove requires an iPETScVector")
        Q = self._build(vec.size)
        a = vec.raw.getArray()
        p = a - Q @ (Q.conj().T @ a)
        a[...] = p if np.iscomplexobj(a) else p.real

    def attach_to(self, mat: "iPETScMatrix") -> None:
        mat.attach_nullspace(self)

    def detach_from(self, mat: "iPETScMatrix") -> None:
        mat._nullspace = None

    def destroy(self) -> None:
        return None


class iPETScBlockMatrix:  # noqa: N801
    """Block matrix of `iPETScMatrix` blocks or None (reference `FEM/utils.py:1385-1489`): block access by index, flat
    form through `to_aij()`."""

    def __init__(self, blocks: list, comm=None) -> None:
        if not blocks or not all(isinstance(row, list) for row in blocks):
            raise ValueError("`blocks` must be a non-empty 2D list")
        if any(len(row) != len(blocks[0]) for row in blocks):
            raise ValueError("All block rows must have the same length")
        for row in blocks:
            for b in row:
                if b is not None and not isinstance(b, iPETScMatrix):
                    raise TypeError(f"Block entries must be iPETScMatrix or None, got {type(b)}")
        self._blocks = blocks
        self._mat = iPETScMatrix.from_nested(blocks)

    @classmethod
    def from_nested(cls, raw: iPETScMatrix, blocks: list) -> "iPETScBlockMatrix":
        obj = cls.__new__(cls)
        obj._mat = raw
        obj._blocks = blocks
        return obj

    @property
    def raw(self) -> iPETScMatrix:
        return self._mat

    @property
    def comm(self) -> _SerialComm:
        return COMM_SELF

    @property
    def shape(self) -> tuple[int, int]:
        return self._mat.shape

    def __getitem__(self, idx: tuple[int, int]) -> "iPETScMatrix | None":
        i, j = idx
        try:
            return self._blocks[i][j]
        except IndexError:
            raise IndexError(f"Block index out of range: {idx}") from None

    def sub(self, i: int, j: int) -> "iPETScMatrix | None":
        return self[i, j]

    def to_aij(self) -> iPETScMatrix:
        return self._mat.to_aij()

    def assemble(self) -> None:
        """Rebuild the flat form from the blocks (they may have been modified in place)."""
        self._mat = iPETScMatrix.from_nested(self._blocks)

    def __str__(self) -> str:
        return f"iPETScBlockMatrix(shape={self.shape}, blocks={len(self._blocks)}x{len(self._blocks[0])})"
