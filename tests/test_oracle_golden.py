"""CPU tests: the SciPy oracle against the reference's own known answers (tests/golden/kat_eigen.json)."""

import numpy as np
import pytest
import scipy.sparse as sp

from lsa_fw_b200 import pencils
from oracle import eigen_oracle as O


def _standard_ks(A, nev, tol, which="LARGEST_MAGNITUDE", M=None):
    A = np.asarray(A, dtype=float)
    n = A.shape[0]
    if M is not None:
        Minv = np.linalg.inv(np.asarray(M, dtype=float))
        op = lambda x: Minv @ (A @ x)  # noqa: E731
    else:
        op = lambda x: A @ x  # noqa: E731
    key = lambda th: O.which_key(which, th)  # noqa: E731
    v0 = np.random.default_rng(0).standard_normal(n)
    theta, X, _, _ = O.krylov_schur(op, n, nev, min(80, n), tol, 100, key_fn=key, v0=v0)
    return theta, X


def test_kat_diag_standard_and_generalized(golden):
    for name in ("diag3_standard", "diag3_generalized_identity"):
        c = golden["cases"][name]
        theta, X = _standard_ks(c["A"], c["num_eig"], c["atol"], M=c["M"])
        assert sorted(theta.real) == pytest.approx(c["expected_sorted"], abs=c["abs_tol"])
        assert np.linalg.norm(X, axis=0) == pytest.approx(1.0, abs=1e-12)  # test_eigen.py:231-239


def test_kat_jordan_block(golden):
    c = golden["cases"]["jordan2"]
    theta, _ = _standard_ks(c["A"], 2, c["atol"])
    assert sorted(theta.real) == pytest.approx(c["expected_sorted_real"], abs=c["abs_tol"])


def test_kat_complex_pair(golden):
    c = golden["cases"]["complex_pair"]
    theta, X = _standard_ks(c["A"], 2, c["atol"])
    order = np.argsort(theta.imag)
    for pos, idx in enumerate(order):
        assert theta[idx] == pytest.approx(complex(*c["expected_by_imag"][pos]), abs=c["abs_tol"])
        assert X[0, idx] / X[1, idx] == pytest.approx(complex(*c["ratios_by_imag"][pos]), abs=c["abs_tol"])


def test_kat_smallest_magnitude_alias(golden):
    c = golden["cases"]["smallest_magnitude_alias"]
    theta, _ = _standard_ks(c["A"], 3, c["atol"], which="SMALLEST_MAGNITUDE")  # aliased to LARGEST_REAL
    assert sorted(theta.real[:2]) == pytest.approx(c["first_two_sorted"], abs=c["abs_tol"])


def test_kat_random_spd(golden):
    c = golden["cases"]["random_spd5"]
    theta, _ = _standard_ks(c["A"], 5, c["atol"])
    assert sorted(theta.real) == pytest.approx(c["expected_sorted"], rel=c["rel_tol"])


def test_kat_shift_invert_epsilon(golden):
    c = golden["cases"]["shift_invert_epsilon"]
    A = sp.csr_matrix(np.asarray(c["A"]))
    r = O.shift_invert_krylov_schur(A, None, c["target"], 3, ncv=3, tol=c["atol"], force_complex=False)
    assert sorted(r.eigenvalues.real) == pytest.approx(c["expected_sorted"], rel=c["rel_tol"])


def test_kat_singular_m(golden):
    c = golden["cases"]["singular_m_raises"]
    with pytest.raises(Exception):
        import scipy.sparse.linalg as spla

        spla.splu(sp.csc_matrix(np.asarray(c["M"]))).solve(np.ones(3))


def test_kat_repeated(golden):
    c = golden["cases"]["repeated_223"]
    w, V = O.dense_pencil_eigs(np.asarray(c["A"]))
    assert sorted(w.real) == pytest.approx(c["expected_sorted"], abs=c["abs_tol"])
    assert np.linalg.matrix_rank(V) == c["rank"]


def test_membrane_table_of_the_reference(golden):
    """P2 membrane, (a, b) = (2, 4), 32 x 32: the relative errors published by the reference
    (tests/benchmark/vibrating_membrane.md:102-110) are reproduced to their printed digits."""
    g = golden["membrane"]
    pm = pencils.membrane_pencil(*g["mesh"], g["a"], g["b"])
    r = O.shift_invert_krylov_schur(pm.A, pm.M, 15.0, 22, ncv=80, tol=1e-12, force_complex=False)
    lam = np.sort(r.eigenvalues.real)
    lam = lam[np.abs(lam - 1.0) > 1e-6][: g["modes"]]  # spurious Dirichlet modes, vibrating_membrane.py:169-173
    ana = pencils.membrane_analytic(g["modes"], g["a"], g["b"])
    err = np.abs(lam - ana) / ana
    assert lam[:3] == pytest.approx(g["lambda_num"], abs=2e-6)
    assert ana[:3] == pytest.approx(g["lambda_ana"], abs=1e-6)
    assert err[:3] == pytest.approx(g["rel_err_first3"], rel=2e-2)
    assert err.mean() == pytest.approx(g["rel_err_mean"], rel=2e-2)


def test_oracle_matches_dense_eig_on_ns_pencil():
    pc = pencils.assemble_pencil((12, 8), (6.0, 2.0), re=40.0, baseflow=pencils.wake_profile(0.9, 1.2, 1.0))
    w, _ = O.dense_pencil_eigs(pc.A, pc.M)
    sigma = 0.5j
    ref = w[np.argsort(abs(w - sigma))][:5]
    for fn in (O.shift_invert_krylov_schur, O.shift_invert_arpack):
        r = fn(pc.A, pc.M, sigma, 6, ncv=30, tol=1e-12)
        assert np.sort_complex(r.eigenvalues[:5]) == pytest.approx(np.sort_complex(ref), rel=1e-8)
        assert r.residuals[:5].max() < 1e-10
    ra = O.shift_invert_krylov_schur(pc.A, pc.M, sigma, 6, ncv=30, tol=1e-12, adjoint=True)
    assert np.sort_complex(np.conj(ra.eigenvalues[:5])) == pytest.approx(np.sort_complex(ref), rel=1e-8)


def test_pencil_structure_matches_reference_assembly():
    """Structural facts pinned by the reference's FEM tests."""
    pc = pencils.assemble_pencil((6, 4), (3.0, 2.0), re=20.0, baseflow=pencils.wake_profile(0.9, 1.2, 1.0))
    A, M = pc.A.toarray(), pc.M.toarray()
    assert pc.n == pencils.th_dofs((6, 4))
    assert np.abs(A[np.ix_(pc.dofs_p, pc.dofs_p)]).max() == 0.0          # zero pp block
    assert np.abs(M[pc.dofs_p]).max() == 0.0 and np.abs(M[:, pc.dofs_p]).max() == 0.0
    for d in pc.dirichlet:                                               # identity rows in A and M
        assert A[d, d] == 1.0 and M[d, d] == 1.0
        assert np.count_nonzero(A[d]) == 1 and np.count_nonzero(A[:, d]) == 1
    Muu = M[np.ix_(pc.dofs_u, pc.dofs_u)]
    assert np.allclose(Muu, Muu.T) and np.linalg.eigvalsh(Muu).min() > 0
    assert pencils.th_dofs((110, 50)) == 50303 and pencils.th_dofs((54, 54, 54)) == 4051462


def test_oracle_reproduces_committed_lns_eigenvalues():
    """The oracle against the committed eigenvalues of the small linearised Navier-Stokes test pencils
    (`tests/golden/lns_eigs.json`, written by `tests/golden/make_lns_golden.py`): a SciPy / LAPACK change that moves the
    checker shows up here, on CPU, before it can blur a GPU parity statement.  Tolerance: 1e-9 relative, or what double
    precision leaves of an eigenvalue with condition number kappa (1e-13 kappa)."""
    import importlib.util
    import json
    import os

    here = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
    spec = importlib.util.spec_from_file_location("make_lns_golden", os.path.join(here, "make_lns_golden.py"))
    G = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(G)
    gold = json.load(open(os.path.join(here, "lns_eigs.json")))
    assert set(gold) == set(G.KINDS)
    for kind in G.KINDS:
        pc, sigma = G.pencil(kind)
        g = gold[kind]
        assert pc.n == g["n"] and complex(*g["sigma"]) == complex(sigma)
        orc = O.shift_invert_krylov_schur(pc.A, pc.M, sigma, g["nev"], ncv=g["ncv"], tol=g["tol"])
        lam = orc.eigenvalues[: g["nev"]]
        for z, kappa in zip(g["eigenvalues"], g["kappa"]):
            z = complex(*z)
            tol = max(1e-9, 1e-13 * kappa) if kappa is not None else 1e-9
            assert min(abs(lam - z)) <= tol * abs(z), (kind, z, kappa)
