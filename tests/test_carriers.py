"""Carrier classes against the expectations of the reference's own carrier tests (`tests/unit/FEM/test_utils.py`):
the cases below restate what those tests assert -- TestVector `:26-166`, TestMatrix `:168-566`, TestNullSpace
`:569-615`, TestComplexVector `:618-844` -- for the PETSc-free carriers of this package (rows a7 / a8).  The NumPy
backing serves both PETSc build flavours at once: `iPETScVector` / `iPETScMatrix` hold real or complex data (the
reference's complex-build cases), `iComplexPETScVector.from_array` splits into (real, imag) parts (its real-build
cases, lazy imaginary part included)."""

import numpy as np
import pytest
import scipy.sparse as sp

from lsa_fw_b200.carriers import (iComplexPETScVector, iPETScBlockMatrix, iPETScMatrix, iPETScNullSpace,
                                  iPETScVector)


# ------------------------------------------------------------------------------------------------ vectors (:26-166)
def test_vector_creation_arithmetic_and_products():
    z = iPETScVector.zeros(5)
    assert z.size == 5 and np.all(z.as_array() == 0.0)
    arr = np.array([1.0, 2.0, 3.0])
    np.testing.assert_array_equal(iPETScVector.from_array(arr).as_array(), arr)
    v1, v2 = iPETScVector.from_array(np.array([1.0, 2.0])), iPETScVector.from_array(np.array([3.0, 4.0]))
    np.testing.assert_array_equal((v1 + v2).as_array(), [4.0, 6.0])
    np.testing.assert_array_equal((v2 + v1).as_array(), [4.0, 6.0])
    np.testing.assert_array_equal((v2 - v1).as_array(), [2.0, 2.0])
    v = iPETScVector.from_array(np.array([2.0, -1.0]))
    scaled = v * 3.0
    assert isinstance(scaled, iPETScVector)
    np.testing.assert_allclose(scaled.as_array(), [6.0, -3.0])
    np.testing.assert_allclose((2.0 * v).as_array(), [4.0, -2.0])
    assert v * iPETScVector.from_array(np.array([4.0, 5.0])) == pytest.approx(3.0)
    outer = v1 @ v2
    assert isinstance(outer, iPETScMatrix) and outer.shape == (2, 2)
    assert outer[0, 0] == pytest.approx(3.0) and outer[1, 1] == pytest.approx(8.0)
    with pytest.raises(ValueError):
        _ = v1 + z


def test_vector_complex_scale_values_and_norm():
    arr = np.array([1 + 1j, 2 - 2j])
    vec = iPETScVector.from_array(arr)
    vec.scale(0.5 + 0.25j)
    np.testing.assert_allclose(vec.as_array(), arr * (0.5 + 0.25j), atol=1e-12)
    z = iPETScVector.zeros(3)
    z.set_value(0, 1 + 2j)
    z.set_value(2, -3 - 4j)
    assert z.get_value(0) == pytest.approx(1 + 2j) and z.get_value(1) == pytest.approx(0j)
    assert z.get_value(2) == pytest.approx(-3 - 4j)
    c = np.array([3 + 4j, -1 + 2j, 0 - 1j])
    assert iPETScVector.from_array(c).norm == pytest.approx(np.linalg.norm(c), abs=1e-12)
    assert iPETScVector.from_array(np.array([3.0, 4.0])).norm == pytest.approx(5.0)


def test_vector_items_copy_random_export(tmp_path, capsys):
    v = iPETScVector.zeros(3)
    v[1] = 5.0
    assert v[1] == pytest.approx(5.0)
    v.set_value(0, 42)
    assert v.get_value(0) == pytest.approx(42)
    v1 = iPETScVector.from_array(np.array([1.0, 2.0]))
    v2 = v1.copy()
    v2.scale(2.0)
    np.testing.assert_allclose(v2.as_array(), [2.0, 4.0])
    np.testing.assert_allclose(v1.as_array(), [1.0, 2.0])
    assert v1 == iPETScVector.from_array(np.array([1.0, 2.0])) and v1 != v2 and not (v1 == v)
    r = iPETScVector.zeros(3)
    r.set_random()
    assert not np.allclose(r.as_array(), 0.0) and np.all((r.as_array() >= 0) & (r.as_array() < 1))
    r.zero_all_entries()
    assert np.all(r.as_array() == 0.0)
    d = v1.duplicate()
    assert d.size == 2 and d is not v1
    s = iPETScVector.create_seq(4)
    s.set_array(np.arange(4.0))
    s.assemble()
    s.ghost_update()
    np.testing.assert_array_equal(s.as_array(), np.arange(4.0))
    assert s.comm.size == 1 and s.raw.getComm().getRank() == 0
    for data in (np.array([1.0, 2.0]), np.array([1.0 - 2j, 0.5j])):
        path = tmp_path / "sub" / "vector_export.bin"
        iPETScVector.from_array(data).export(path)
        assert path.stat().st_size == 8 + data.size * data.itemsize
        np.testing.assert_array_equal(iPETScVector.from_file(path).as_array(), data)
    v1.print()
    assert "iPETScVector" in capsys.readouterr().out


# ------------------------------------------------------------------------------------------------ matrices (:168-566)
def test_matrix_constructors():
    m = iPETScMatrix.create_aij((3, 4), comm=None, nnz=2)
    assert isinstance(m, iPETScMatrix) and m.shape == (3, 4) and m.nonzero_entries == 0 and "aij" in m.type.lower()
    z = iPETScMatrix.zeros((4, 5))
    assert z.shape == (4, 5) and z.nonzero_entries == 0
    a = iPETScMatrix.create_aij((2, 2), nnz=1)
    assert iPETScMatrix.from_matrix(a) is a
    assert iPETScMatrix.from_matrix(a.raw) is a                       # a raw handle wraps back into its carrier
    arr = np.array([[1.0, 0.0, 2.0], [0.0, -3.5, 0.0]])
    M = iPETScMatrix.from_matrix(arr)
    assert M.shape == arr.shape and M[0, 0] == 1.0 and M[0, 2] == 2.0 and M[1, 1] == -3.5
    data, rows, cols = np.array([10, 20, 30]), np.array([0, 1, 2]), np.array([2, 0, 1])
    S = iPETScMatrix.from_matrix(sp.coo_matrix((data, (rows, cols)), shape=(4, 4)))
    S.assemble()
    assert S.shape == (4, 4) and S.nonzero_entries == 3
    assert all(S[i, j] == v for i, j, v in zip(rows, cols, data))
    with pytest.raises(TypeError):
        iPETScMatrix.from_matrix("not a matrix")


def test_matrix_nested_and_block():
    A = iPETScMatrix.create_aij((2, 2))
    A[0, 0] = 1.0
    A[1, 1] = 2.0
    A.assemble()
    Z = A.duplicate()
    Z.zero_all_entries()
    nested = iPETScMatrix.from_nested([[A.raw, Z.raw], [Z.raw, None]])
    assert isinstance(nested, iPETScMatrix) and nested.shape == (4, 4) and nested.type == "nest"
    aij = nested.to_aij()
    assert aij[0, 0] == pytest.approx(1.0) and aij[1, 1] == pytest.approx(2.0) and aij[2, 2] == 0.0
    assert "aij" in aij.type and nested.sub(0, 0) is A and nested.sub(1, 1) is None
    with pytest.raises(NotImplementedError):
        A.sub(0, 0)
    with pytest.raises(NotImplementedError):
        A.to_aij()                                                   # plain AIJ matrices raise in the reference too
    B = iPETScBlockMatrix([[A, None], [None, A]])
    assert B.shape == (4, 4) and B[0, 0] is A and B.sub(0, 1) is None and "2x2" in str(B)
    assert B.to_aij()[3, 3] == pytest.approx(2.0) and B.raw.type == "nest"
    with pytest.raises(IndexError):
        _ = B[2, 0]
    with pytest.raises(TypeError):
        iPETScBlockMatrix([[A, np.eye(2)]])
    with pytest.raises(ValueError):
        iPETScBlockMatrix([[A, None], [A]])


def test_matrix_arithmetic_products_and_equality():
    A = iPETScMatrix.create_aij((2, 2))
    A[0, 0] = 1.0
    B = A.duplicate(copy=True)
    assert (A + B)[0, 0] == pytest.approx(2.0) and (B + A)[0, 0] == pytest.approx(2.0)
    assert (A - B).norm == 0.0 and A == B and not (A == iPETScMatrix.zeros((2, 2))) and A != iPETScMatrix.zeros((3, 3))
    with pytest.raises(NotImplementedError):
        _ = A + 1.0
    with pytest.raises(ValueError):
        _ = A + iPETScMatrix.zeros((3, 3))
    D = iPETScMatrix.from_matrix(np.array([[2.0, 0.0], [0.0, 3.0]]))
    np.testing.assert_allclose((D @ iPETScVector.from_array(np.array([1.0, 2.0]))).as_array(), [2.0, 6.0])
    assert isinstance(D @ D, iPETScMatrix) and (D @ D)[0, 0] == pytest.approx(4.0)
    np.testing.assert_allclose((iPETScVector.from_array(np.array([3.0, 4.0])) @ D).as_array(), [6.0, 12.0])
    U = iPETScMatrix.from_matrix(np.array([[0.0, 1.0], [0.0, 0.0]]))
    np.testing.assert_allclose((iPETScVector.from_array(np.array([1.0, 0.0])) @ U).as_array(), [0.0, 1.0])   # A^T x
    with pytest.raises(ValueError):
        _ = D @ iPETScVector.zeros(3)
    # the raw handle multiplies as Sensitivity/__init__.py:281-283 uses it
    x, y = iPETScVector.from_array(np.array([1.0, 2.0])), D.create_vector_left()
    D.raw.mult(x.raw, y.raw)
    np.testing.assert_allclose(y.as_array(), [2.0, 6.0])


def test_matrix_transposes_and_symmetry():
    A = iPETScMatrix.from_matrix(np.array([[0.0, 1.0, 0.0], [0.0, 0.0, 2.0]]))
    t = A.T.to_aij()
    assert A.shape == (2, 3) and A.T.shape == (3, 2) and t[2, 1] == 2.0 and t[1, 0] == 1.0
    C = iPETScMatrix.from_matrix(np.array([[1.0 + 0.0j, 4.0 - 2.0j], [3.0 + 5.0j, 2.0 + 0.0j]]))
    H = C.H.to_aij()
    for i in range(2):
        for j in range(2):
            assert H[i, j] == pytest.approx(np.conj(C[j, i]), abs=1e-12)
    R = iPETScMatrix.from_matrix(np.array([[1.0, 4.0], [3.0, 2.0]]))
    assert R.T.to_aij() == R.H.to_aij()
    S = iPETScMatrix.from_matrix(np.array([[1.0, 2.0], [2.0, 2.0]]))
    assert S.is_symmetric and S.is_numerically_symmetric(tol=1e-12) and S.is_hermitian()
    assert not R.is_symmetric and not C.is_hermitian()
    assert iPETScMatrix.from_matrix(np.array([[1.0, 2 + 1j], [2 - 1j, 3.0]])).is_hermitian()


def test_matrix_in_place_operations(tmp_path, capsys):
    A = iPETScMatrix.create_aij((1, 1))
    A[0, 0] = 1.0
    A.shift(2.0)
    assert A[0, 0] == pytest.approx(3.0)
    A.scale(0.5)
    assert A[0, 0] == pytest.approx(1.5)
    A.axpy(2.0, A)
    assert A[0, 0] == pytest.approx(4.5)
    A.zero_all_entries()
    assert A[0, 0] == 0.0
    C = iPETScMatrix.create_aij((1, 1))
    C[0, 0] = 2.0 + 0.0j
    C.scale(1.0j)
    assert C[0, 0] == pytest.approx(2.0j)
    E = iPETScMatrix.create_aij((1, 1))
    E[0, 0] = 1.0 + 0.0j
    E.axpy(2.0 + 3.0j, E)
    assert E[0, 0] == pytest.approx(3.0 + 3.0j)
    W = iPETScMatrix.create_aij((3, 2))
    assert W.create_vector_right().size == 2 and W.create_vector_left().size == 3
    P = iPETScMatrix.from_matrix(np.array([[1.0, 2.0], [3.0, 4.0]]))
    P.pin_dof(0)
    assert (P[0, 0], P[0, 1], P[1, 0], P[1, 1]) == (1.0, 0.0, 0.0, 4.0)
    X = iPETScMatrix.create_aij((2, 2))
    X[0, 1] = 4.2
    X.export(tmp_path / "matrix_output.bin")
    assert (tmp_path / "matrix_output.bin").stat().st_size > 0 and iPETScMatrix.load(tmp_path / "matrix_output.bin") == X
    X.print()
    assert "iPETScMatrix(shape=(2, 2), nnz=1)" in capsys.readouterr().out


def test_matrix_rows_columns_and_accumulation():
    M = iPETScMatrix.from_matrix(np.array([[1.0, 0.0, 2.0], [0.0, 3.0, 0.0]]))
    assert M.get_row(0) == ([0, 1, 2], [1.0, 0.0, 2.0]) and M.get_row(1) == ([0, 1, 2], [0.0, 3.0, 0.0])   # dense origin
    for c, v in zip(*M.get_row(1)):
        assert M.get_value(1, c) == pytest.approx(v)
    Z = iPETScMatrix.zeros((3, 3))
    Z[0, 1] = 42.0
    assert Z.get_row(0) == ([1], [42.0]) and Z.get_row(1) == ([], [])
    K = iPETScMatrix.from_matrix(np.array([[5.0, 0.0], [0.0, 7.0], [8.0, 9.0]]))
    assert K.get_column(0) == ([0, 1, 2], [5.0, 0.0, 8.0]) and K.get_column(1) == ([0, 1, 2], [0.0, 7.0, 9.0])
    Ks = iPETScMatrix.from_matrix(sp.csr_matrix(K.as_array()))
    assert Ks.get_column(0) == ([0, 2], [5.0, 8.0]) and Ks.get_row(2) == ([0, 1], [8.0, 9.0])
    O = iPETScMatrix.zeros((2, 3))
    O[0, 0] = 1.0
    assert O.get_column(0) == ([0], [1.0])
    Q = iPETScMatrix.zeros((3, 3))
    Q.add_value(1, 2, 4.2)
    Q.add_value(1, 2, 1.3)
    Q.assemble()
    assert Q[1, 2] == pytest.approx(5.5, abs=1e-12)
    cols, vals = Q.get_row(1)
    assert cols == [2] and vals == pytest.approx([5.5])


# ------------------------------------------------------------------------------------------------ nullspaces (:569-615)
def test_nullspace_carrier():
    A = iPETScMatrix.from_matrix(np.array([[1.0, -1.0, 0.0], [-1.0, 2.0, -1.0], [0.0, -1.0, 1.0]]))
    ones = iPETScVector.from_array(np.ones(3))
    ns = iPETScNullSpace.from_vectors([ones])
    assert ns.test_matrix(A)[0] and ns.test_vector(A, ones)[0] and not ns.has_constant() and ns.dimension == 1
    nc = iPETScNullSpace.create_constant(comm=A.comm)                     # no size: PETSc's constant has none either
    assert nc.has_constant() and nc.test_matrix(A)[0] and "constant" in repr(nc)
    v = iPETScVector.from_array(np.array([2.0, 3.0, 4.0]))
    iPETScNullSpace.create_constant(comm=v.comm).remove(v)
    np.testing.assert_allclose(v.as_array(), [-1.0, 0.0, 1.0], atol=1e-12)
    both = iPETScNullSpace.create_constant_and_vectors(A.comm, [iPETScVector.from_array(np.full(3, 2.0))])
    assert both.has_constant() and both.test_matrix(A)[0]
    B = iPETScMatrix.create_aij((3, 3))
    B.attach_nullspace(iPETScNullSpace.create_constant_and_vectors(B.comm, [ones]))
    got = B.get_nullspace()
    assert got is not None and got.test_matrix(B)[0] and got.as_array().shape[0] == 3
    assert not ns.test_matrix(iPETScMatrix.from_matrix(np.eye(3)))[0]
    with pytest.raises(ValueError):
        iPETScNullSpace.from_vectors([])
    with pytest.raises(TypeError):
        iPETScNullSpace.from_vectors([np.ones(3)])
    with pytest.raises(ValueError):
        iPETScNullSpace.from_vectors([ones, iPETScVector.from_array(2 * np.ones(3))])
    nc.destroy()
    assert nc.raw is nc and nc.comm.size == 1


# ------------------------------------------------------------------------------------------------ complex vectors (:618-844)
def test_complex_vector_parts_and_lazy_imaginary_part():
    vec = iPETScVector.zeros(3)
    assert iComplexPETScVector(vec).imag is None and iComplexPETScVector(real=vec, imag=vec).imag is not None
    c = iComplexPETScVector(iPETScVector.zeros(2))
    c[0] = 1.0
    c[1] = -2.5
    c.assemble()
    assert c.imag is None
    c[1] = 3.0 + 4.0j
    assert c.imag is not None and c[1] == 3.0 + 4.0j and isinstance(c[0], complex) and c[0] == 1.0 + 0j
    c[1] = 7.0                                                           # a real value zeroes the imaginary entry
    assert c[1] == 7.0 + 0j
    assert not iComplexPETScVector.from_array(np.array([1.0, 0.0, 3.0])).is_complex
    assert iComplexPETScVector.from_array(np.array([1.0 + 0j, 2j, 3.0])).is_complex


def test_complex_vector_arithmetic():
    v1 = iComplexPETScVector.from_array(np.array([1 + 2j, 3 + 4j]))
    v2 = iComplexPETScVector.from_array(np.array([5 + 6j, 7 + 8j]))
    s, d = v1 + v2, v2 - v1
    assert s[0] == pytest.approx(6 + 8j) and s[1] == pytest.approx(10 + 12j)
    assert d[0] == pytest.approx(4 + 4j) and d[1] == pytest.approx(4 + 4j)
    r = iComplexPETScVector.from_array(np.array([1, 3])) + iComplexPETScVector.from_array(np.array([5, 7]))
    assert r.imag is None and r[0] == pytest.approx(6) and r[1] == pytest.approx(10)
    mixed = v1 + iPETScVector.from_array(np.array([1.0, 1.0]))
    assert mixed[0] == pytest.approx(2 + 2j)
    comp = iComplexPETScVector.from_array(np.array([2.0, 0.5]))
    t = comp * 2.0
    assert t.imag is None and t[0] == pytest.approx(4.0) and t[1] == pytest.approx(1.0)
    u = comp * (1 + 1j)
    assert u.imag is not None and u[0] == pytest.approx(2 + 2j) and u[1] == pytest.approx(0.5 + 0.5j)
    assert ((1 + 1j) * comp)[0] == pytest.approx(2 + 2j)
    comp.scale(1 + 1j)
    assert comp.imag is not None and comp[0] == pytest.approx(2 + 2j) and comp[1] == pytest.approx(0.5 + 0.5j)
    # one complex part (the reference's complex build): results stay in one part
    one = iComplexPETScVector(np.array([2.0 + 0j, 0.5]))
    w = one * (1 + 1j)
    assert one.imag is None and w.imag is None and w[0] == pytest.approx(2 + 2j)


def test_complex_vector_matrix_products_norm_dot_equality():
    mat = iPETScMatrix.from_matrix(np.array([[1.0, 2.0], [3.0, 4.0]]))
    vec = iComplexPETScVector.from_array(np.array([1 + 1j, 2 + 2j]))
    res = mat @ vec
    assert isinstance(res, iComplexPETScVector) and res.is_complex
    assert res[0] == pytest.approx((1 + 1j) + 2 * (2 + 2j)) and res[1] == pytest.approx(3 * (1 + 1j) + 4 * (2 + 2j))
    res = vec @ mat
    assert isinstance(res, iComplexPETScVector) and res.is_complex
    assert res[0] == pytest.approx((1 + 1j) + (2 + 2j) * 3) and res[1] == pytest.approx((1 + 1j) * 2 + (2 + 2j) * 4)
    with pytest.raises(NotImplementedError):
        _ = vec @ vec
    v = iComplexPETScVector.from_array(np.array([3 + 4j, 0]))
    w = iComplexPETScVector.from_array(np.array([3 - 4j, 0]))
    assert v.norm() == pytest.approx(5.0) and w.norm() == pytest.approx(5.0)
    assert v.dot(v) == pytest.approx(25.0) and w.dot(w) == pytest.approx(25.0)
    assert v.dot(w) == pytest.approx(-7 - 24j) and w.dot(v) == pytest.approx(-7 + 24j)
    data = np.array([1.0, 2 + 3j, -4.5])
    a = iComplexPETScVector.from_array(data)
    b = a.copy()
    assert b == a
    a[0] = 9.0
    assert b != a and not (a == "something else")


def test_example_sweep_script_loads_its_cases(tmp_path):
    """`examples/eigenvalues_sweep.py --synthetic --dry-run`: writes MatrixMarket cases in the reference's directory
    layout and reads them back through `iPETScMatrix.from_path` (no GPU needed up to the solve)."""
    import os
    import subprocess
    import sys

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, os.path.join(root, "examples", "eigenvalues_sweep.py"), str(tmp_path / "cases"),
                          "--synthetic", "--dry-run"], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    assert out.stderr.count("A: shape=") == 11 and "All cases processed." in out.stderr
    assert (tmp_path / "cases" / "reynolds_40.0" / "matrices" / "M.mtx").exists()
