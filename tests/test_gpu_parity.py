"""GPU parity tests (run on the B200 box): the CUDA path, called through the reference-facing API and
the C ABI, against the SciPy oracle, the reference's golden known answers and size-independent
properties.  Tolerances are the north star's: eigenvalues 1e-8 relative, residuals
||A x - lambda M x|| / (||A||_F ||x||) <= 1e-10."""

import numpy as np
import pytest
import scipy.sparse as sp

import lsa_fw_b200 as L
from lsa_fw_b200 import _lib, pencils
from oracle import eigen_oracle as O

pytestmark = pytest.mark.gpu

EIG_RTOL = 1e-8
RESID_BAR = 1e-10


def _mat(a):
    return L.iPETScMatrix.from_matrix(np.asarray(a, dtype=float)) if a is not None else None


def _solver(c, **kw):
    cfg = L.EigensolverConfig(num_eig=c["num_eig"], problem_type=L.iEpsProblemType[c["problem_type"]], atol=c["atol"],
                              max_it=c["max_it"])
    return L.EigenSolver(cfg, A=_mat(c["A"]), M=_mat(c["M"]), **kw)


def _vec(v):
    a = v.real.as_array()
    return a if v.imag is None else a + 1j * v.imag.as_array()


# ------------------------------------------------------------------ reference known-answer tests
@pytest.mark.parametrize("name", ["diag3_standard", "diag3_generalized_identity", "random_spd5", "repeated_223"])
def test_kat_sorted_eigenvalues(golden, name):
    c = golden["cases"][name]
    pairs = _solver(c).solve()
    found = sorted(val for val, _ in pairs)          # floats for Hermitian problem types
    tol = dict(abs=c["abs_tol"]) if "abs_tol" in c else dict(rel=c["rel_tol"])
    assert found == pytest.approx(c["expected_sorted"], **tol)
    for _, vec in pairs:
        assert vec.norm() == pytest.approx(1.0, abs=1e-12)       # test_eigen.py:231-239
        assert vec.real.size == 3 or vec.real.size == 5
    if "rank" in c:
        V = np.vstack([_vec(v) for _, v in pairs]).T
        assert np.linalg.matrix_rank(V) == c["rank"]


def test_kat_jordan_block(golden):
    c = golden["cases"]["jordan2"]
    vals = sorted(val.real for val, _ in _solver(c).solve())
    assert vals == pytest.approx(c["expected_sorted_real"], abs=c["abs_tol"])


def test_kat_complex_pair_in_real_mode(golden):
    c = golden["cases"]["complex_pair"]
    pairs = _solver(c).solve()
    (val1, vec1), (val2, vec2) = sorted(pairs, key=lambda p: p[0].imag)
    assert val1 == pytest.approx(complex(*c["expected_by_imag"][0]), abs=c["abs_tol"])
    assert val2 == pytest.approx(complex(*c["expected_by_imag"][1]), abs=c["abs_tol"])
    for vec, ratio in ((vec1, c["ratios_by_imag"][0]), (vec2, c["ratios_by_imag"][1])):
        assert vec.imag is not None                   # real mode returns (vr, vi), Solver/utils.py:280-291
        arr = _vec(vec)
        assert arr[0] / arr[1] == pytest.approx(complex(*ratio), abs=c["abs_tol"])


def test_kat_smallest_magnitude_alias(golden):
    c = golden["cases"]["smallest_magnitude_alias"]
    es = _solver(c)
    es.solver.set_which_eigenpairs(L.iEpsWhich.SMALLEST_MAGNITUDE)
    es.solve()
    found = sorted(val for val, _ in es.solver.get_all_eigenpairs_up_to(2))
    assert found == pytest.approx(c["first_two_sorted"], abs=c["abs_tol"])


def test_kat_shift_invert_epsilon(golden):
    c = golden["cases"]["shift_invert_epsilon"]
    es = _solver(c)
    es.solver.set_st_type(L.iSTType.SINVERT)
    es.solver.set_target(c["target"])
    found = sorted(float(v.real) for v, _ in es.solve())
    assert found == pytest.approx(c["expected_sorted"], rel=c["rel_tol"])


def test_kat_singular_m_raises(golden):
    c = golden["cases"]["singular_m_raises"]
    with pytest.raises(L.LsaError):
        _solver(c).solve()


def test_real_eigenvector_has_no_imaginary_part(golden):
    c = dict(golden["cases"]["diag3_standard"], problem_type="GNHEP", num_eig=2)
    for _, vec in _solver(c).solve():
        assert isinstance(vec, L.iComplexPETScVector) and vec.imag is None and vec.real.size == 3


def test_membrane_table_of_the_reference(golden):
    g = golden["membrane"]
    pm = pencils.membrane_pencil(*g["mesh"], g["a"], g["b"])
    cfg = L.EigensolverConfig(num_eig=20, problem_type=L.iEpsProblemType.GHEP, atol=1e-12, max_it=200)
    es = L.EigenSolver(L.iPETScMatrix(pm.A), L.iPETScMatrix(pm.M), cfg, check_hermitian=False)
    es.solver.set_st_type(L.iSTType.SINVERT)
    es.solver.set_target(17.5)   # the 15 lowest modes (3.08 .. 32.08) lie closer to 17.5 than the spurious 1
    lam = np.sort([v for v, _ in es.solve()])
    lam = lam[np.abs(lam - 1.0) > 1e-6][: g["modes"]]
    ana = pencils.membrane_analytic(g["modes"], g["a"], g["b"])
    err = np.abs(lam - ana) / ana
    assert err[:3] == pytest.approx(g["rel_err_first3"], rel=2e-2)
    assert err.mean() == pytest.approx(g["rel_err_mean"], rel=2e-2)


def test_wide_subspace_hundred_modes():
    """nev = 100 with ncv = 2 nev = 200 (SLEPc's default rule): the orthogonalisation, Rayleigh-Ritz and restart kernels
    run past their 128-column chunk (dots in two column groups, RR out of shared memory)."""
    pm = pencils.membrane_pencil(24, 24, 1.0, 0.83)
    sigma = 1000.0   # 100 nearest modes span 180 .. 1820, clear of the spurious Dirichlet eigenvalue 1
    cfg = L.EigensolverConfig(num_eig=100, problem_type=L.iEpsProblemType.GHEP, atol=1e-11, max_it=300, ncv=200)
    es = L.EigenSolver(L.iPETScMatrix(pm.A), L.iPETScMatrix(pm.M), cfg, check_hermitian=False)
    es.solver.set_st_type(L.iSTType.SINVERT)
    es.solver.set_target(sigma)
    pairs = es.solve()
    assert len(pairs) >= 100
    assert es.solver.raw.getDimensions()[1] == 200
    lam = np.array([v for v, _ in pairs][:100])
    ref = O.shift_invert_arpack(pm.A, pm.M, sigma, 104, tol=1e-12).eigenvalues
    ref = ref[np.argsort(np.abs(ref - sigma))][:100]
    assert np.abs(np.sort(lam.real) - np.sort(ref.real)).max() <= 1e-8 * np.abs(ref).max()
    X = np.stack([_vec(v) for _, v in pairs[:100]], axis=1)
    assert O.north_star_residuals(pm.A, pm.M, lam, X).max() < RESID_BAR


# ------------------------------------------------------------------ linearised Navier-Stokes pencils
def _ns(kind):
    if kind == "th2d":
        return pencils.assemble_pencil((24, 12), (8.0, 3.0), re=50.0, baseflow=pencils.wake_profile(0.9, 1.2, 1.5)), 0.05 + 0.6j
    if kind == "mini2d":
        return pencils.assemble_pencil((20, 12), (8.0, 3.0), re=50.0, baseflow=pencils.wake_profile(0.9, 1.2, 1.5), space="MINI"), 0.05 + 0.6j
    if kind == "th3d":
        return pencils.cavity_3d(6), 0.1 + 0.3j
    if kind == "th2d_real":
        return pencils.assemble_pencil((24, 12), (8.0, 3.0), re=50.0, baseflow=pencils.wake_profile(0.9, 1.2, 1.5)), -0.3
    raise ValueError(kind)


def _run(pc, sigma, nev=6, ncv=40, tol=1e-11, adjoint=False, **opts):
    cfg = L.EigensolverConfig(num_eig=nev, atol=tol, max_it=200, ncv=ncv)
    es = L.EigenSolver(L.iPETScMatrix(pc.A), L.iPETScMatrix(pc.M), cfg, check_hermitian=False)
    es.solver.set_st_type(L.iSTType.SINVERT)
    es.solver.set_target(sigma)
    es.solver.set_st_pc_type(L.PreconditionerType.LU)
    es.solver.set_backend_options(leaf_size=32, **opts)
    es.solver.set_adjoint(adjoint)
    return es, es.solve()


def _match(lam, ref):
    return max(min(abs(l - ref)) / abs(l) for l in lam)


@pytest.mark.parametrize("kind", ["th2d", "mini2d", "th3d", "th2d_real"])
@pytest.mark.parametrize("use_coords", [False, True], ids=["graph", "geometric"])
def test_eigenpairs_match_oracle(kind, use_coords):
    pc, sigma = _ns(kind)
    es, pairs = _run(pc, sigma, coords=pc.coords if use_coords else None)
    assert len(pairs) == 6
    lam = np.array([p[0] for p in pairs])
    orc = O.shift_invert_krylov_schur(pc.A, pc.M, sigma, 8, ncv=40, tol=1e-12)
    # 1e-8 relative, or the accuracy attainable in double precision for a non-normal eigenvalue:
    # |d lambda| ~ kappa(lambda) * eps * ||A||, kappa = ||y|| ||x|| / |y^H M x| from the oracle's left/right
    # vectors (the real-shift case has kappa ~ 1e7: even LAPACK's dense QZ only agrees to ~1e-9 there)
    adj = O.shift_invert_krylov_schur(pc.A, pc.M, sigma, 8, ncv=40, tol=1e-12, adjoint=True)
    for l in lam:
        j = int(np.argmin(abs(orc.eigenvalues - l)))
        ja = int(np.argmin(abs(np.conj(adj.eigenvalues) - orc.eigenvalues[j])))
        x, yv = orc.eigenvectors[:, j], adj.eigenvectors[:, ja]
        kappa = np.linalg.norm(x) * np.linalg.norm(yv) / max(abs(np.vdot(yv, pc.M @ x)), 1e-300)
        print(f"{kind}: lambda = {l:.12g}  kappa = {kappa:.3e}  |d lambda|/|lambda| = {abs(l - orc.eigenvalues[j]) / abs(l):.2e}")
        assert 1e-14 * kappa <= 1e-6, ("eigenvalue too ill-conditioned for a parity statement", l, kappa)
        assert abs(l - orc.eigenvalues[j]) / abs(l) < max(EIG_RTOL, 1e-14 * kappa), (l, orc.eigenvalues[j], kappa)
    # ... and against the COMMITTED oracle values of the same pencil (tests/golden/lns_eigs.json; the CPU suite checks
    # that the live oracle still reproduces them)
    import json
    import os

    g = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "lns_eigs.json")))[kind]
    gl = np.array([complex(*z) for z in g["eigenvalues"]])
    assert g["n"] == pc.n and complex(*g["sigma"]) == complex(sigma)
    for l in lam:
        j = int(np.argmin(abs(gl - l)))
        kg = g["kappa"][j] or 0.0
        assert abs(l - gl[j]) / abs(l) < max(EIG_RTOL, 1e-14 * kg), (l, gl[j], kg)
    # `which` order: increasing |lambda - sigma| (TARGET_MAGNITUDE default under sinvert)
    assert np.all(np.diff(np.abs(lam - sigma)) >= -1e-9 * np.abs(lam[:-1] - sigma))
    X = np.stack([_vec(v) for _, v in pairs], axis=1)
    assert np.linalg.norm(X, axis=0) == pytest.approx(1.0, abs=1e-12)
    assert O.north_star_residuals(pc.A, pc.M, lam, X).max() < RESID_BAR
    assert es.solver.get_residuals()[:6].max() < RESID_BAR
    complex_mode = isinstance(sigma, complex)
    assert all((v.imag is None) for _, v in pairs) if complex_mode else True
    assert es.solver.stats["n_perturbed"] == 0
    # eigenvectors agree with the oracle up to phase (simple eigenvalues only)
    for i, l in enumerate(lam):
        j = int(np.argmin(abs(orc.eigenvalues - l)))
        others = np.delete(orc.eigenvalues, j)
        if min(abs(others - l)) > 1e-6:
            assert abs(np.vdot(orc.eigenvectors[:, j], X[:, i])) == pytest.approx(1.0, abs=1e-6)


def test_adjoint_modes_reuse_the_factorisation():
    """Sensitivity/__init__.py:230-311: left eigenvectors at conj(sigma); here on the same LU."""
    pc, sigma = _ns("th2d")
    A, M = L.iPETScMatrix(pc.A), L.iPETScMatrix(pc.M)
    cfg = L.EigensolverConfig(num_eig=5, atol=1e-11, max_it=200, ncv=40)
    es = L.EigenSolver(A, M, cfg, check_hermitian=False)
    es.solver.set_st_type(L.iSTType.SINVERT)
    es.solver.set_target(sigma)
    es.solver.set_st_pc_type(L.PreconditionerType.LU)
    pairs = es.solve()
    lam, v = min(pairs, key=lambda p: abs(p[0] - sigma))
    # drop-in form: explicit Hermitian transposes, target conj(sigma), TARGET_REAL (Sensitivity :246-262)
    es_adj = L.EigenSolver(A.H, M.H, cfg, check_hermitian=False)
    es_adj.solver.set_st_type(L.iSTType.SINVERT)
    es_adj.solver.set_st_pc_type(L.PreconditionerType.LU)
    es_adj.solver.set_target(np.conj(sigma))
    es_adj.solver.set_which_eigenpairs(L.iEpsWhich.TARGET_REAL)
    pairs_adj = es_adj.solve()
    assert es_adj.solver.stats.get("reused_factorisation") is True
    lam_adj, a = min(pairs_adj, key=lambda p: abs(p[0] - np.conj(lam)))
    assert lam_adj == pytest.approx(np.conj(lam), rel=EIG_RTOL)
    av, vv = _vec(a), _vec(v)
    AH, MH = pc.A.conj().T.tocsr(), pc.M.conj().T.tocsr()
    assert np.linalg.norm(AH @ av - lam_adj * (MH @ av)) / (np.sqrt((pc.A.data**2).sum()) * np.linalg.norm(av)) < RESID_BAR
    # bi-orthonormalisation a^H M v = 1 (Sensitivity :280-287)
    prod = a.dot(L.iPETScVector(pc.M @ vv))
    assert abs(prod) > 1e-8
    a.scale(1.0 / prod)
    assert a.dot(L.iPETScVector(pc.M @ vv)) == pytest.approx(1.0, abs=1e-10)
    # explicit-matrix route without reuse gives the same eigenvalue
    es2 = L.EigenSolver(L.iPETScMatrix(AH), L.iPETScMatrix(MH), cfg, check_hermitian=False)
    es2.solver.set_st_type(L.iSTType.SINVERT)
    es2.solver.set_target(np.conj(sigma))
    lam2 = min((p[0] for p in es2.solve()), key=lambda z: abs(z - np.conj(lam)))
    assert lam2 == pytest.approx(lam_adj, rel=EIG_RTOL)


def test_direct_and_adjoint_modes_are_biorthonormal():
    """Sensitivity/__init__.py:171-311 in one call: direct mode, adjoint mode on the same LU, a^H M v = 1, and
    the classical bi-orthogonality a_i^H M v_j = 0 for i != j of a non-normal pencil."""
    pc = pencils.assemble_pencil((24, 12), (8.0, 3.0), re=50.0, baseflow=pencils.wake_profile(0.9, 1.2, 1.5))
    sigma = 0.05 + 0.6j
    A, M = L.iPETScMatrix(pc.A), L.iPETScMatrix(pc.M)
    cfg = L.EigensolverConfig(num_eig=4, atol=1e-11, max_it=200, ncv=40)
    (lam, v), (lam_adj, a) = L.direct_and_adjoint_modes(A, M, sigma, cfg, backend_options=dict(leaf_size=32))
    assert abs(lam_adj - np.conj(lam)) < 1e-8 * abs(lam)
    va, aa = _vec(v), _vec(a)
    assert np.vdot(aa, pc.M @ va) == pytest.approx(1.0, abs=1e-10)
    norm_a = float(np.sqrt((np.abs(pc.A.data) ** 2).sum()))
    assert np.linalg.norm(pc.A.conj().T @ aa - lam_adj * (pc.M.conj().T @ aa)) / (norm_a * np.linalg.norm(aa)) < RESID_BAR
    # the next direct mode is M-orthogonal to this adjoint mode
    es = L.EigenSolver(A, M, cfg, check_hermitian=False)
    es.solver.set_st_type(L.iSTType.SINVERT)
    es.solver.set_target(sigma)
    es.solver.set_backend_options(leaf_size=32)
    others = [(l, _vec(x)) for l, x in es.solve() if abs(l - lam) > 1e-6]
    assert others and all(abs(np.vdot(aa, pc.M @ x)) < 1e-7 * np.linalg.norm(aa) for _, x in others)


def test_symbolic_analysis_is_reused_across_shifts_and_reynolds():
    L.clear_symbolic_cache()
    base = dict(shape=(24, 12), lengths=(8.0, 3.0), baseflow=pencils.wake_profile(0.9, 1.2, 1.5))
    cached = []
    for re_, sigma in ((40.0, 0.02 + 0.55j), (50.0, 0.05 + 0.6j), (60.0, 0.06 + 0.62j)):
        pc = pencils.assemble_pencil(base["shape"], base["lengths"], re=re_, baseflow=base["baseflow"])
        es, pairs = _run(pc, sigma, nev=4, ncv=30)
        cached.append(es.solver.stats["symbolic_cached"])
        lam = np.array([p[0] for p in pairs])
        orc = O.shift_invert_arpack(pc.A, pc.M, sigma, 6, ncv=30, tol=1e-12)
        assert _match(lam, orc.eigenvalues) < EIG_RTOL
    assert cached == [False, True, True]


def test_gram_schmidt_refinement_if_needed_matches_always():
    """SLEPc's default orthogonalisation refines only when the first pass removed more than half of the squared
    norm (BV_ORTHOG_REFINE_IFNEEDED, eta = 1/sqrt 2); the device takes that decision per column.  Same pairs as
    with unconditional CGS2, far fewer second passes, basis-dependent quantities still at the parity bars."""
    pc = pencils.assemble_pencil((24, 12), (8.0, 3.0), re=50.0, baseflow=pencils.wake_profile(0.9, 1.2, 1.5))
    sigma = 0.05 + 0.6j
    h = _lib.Handle(pc.n, 0)
    flag = ((pc.A.diagonal() - sigma * pc.M.diagonal()) == 0).astype(np.uint8)
    h.analyze(pc.A.indptr, pc.A.indices, pc.M.indptr, pc.M.indices, leaf_size=32, order_last=flag)
    h.set_values(pc.A.data, pc.M.data)
    h.factor(1.0, -sigma, _lib.LSA_C128, 1e-13)
    out = {}
    for fuse in (1, 0):         # fused "update of pass 1 + dots of pass 2" kernel and the separate kernels
        h.set_option("fuse_ortho", fuse)
        for always in (1, 0):
            h.set_option("ortho_refine_always", always)
            r = h.eigs(nev=6, ncv=40, tol=1e-11, max_restarts=100, which="TARGET_MAGNITUDE", transform=_lib.LSA_ST_SINVERT,
                       sigma=sigma, seed=1)
            assert r.nconv >= 6
            X = h.eigenvectors(6)
            out[fuse, always] = (r, h.eigenvalues(6), h.residuals(6), X)
        ra, la, resa, _ = out[fuse, 1]
        ri, li, resi, Xi = out[fuse, 0]
        assert ra.n_reorth > 0 and ri.n_reorth < ra.n_reorth
        assert np.abs(li - la).max() < EIG_RTOL * np.abs(la).max()
        assert resi.max() < RESID_BAR and resa.max() < RESID_BAR
        assert np.linalg.norm(Xi, axis=0) == pytest.approx(1.0, abs=1e-12)
    # both kernel organisations compute the same thing
    assert out[1, 0][0].n_reorth == out[0, 0][0].n_reorth and out[1, 0][0].n_op_applies == out[0, 0][0].n_op_applies
    assert np.abs(out[1, 0][1] - out[0, 0][1]).max() < EIG_RTOL * np.abs(out[0, 0][1]).max()
    h.close()


def test_large_front_paths_real_factor():
    """Same 3-D cavity with a REAL shift: real FP64 factor (the reference's real PETSc build): whole-block inverses +
    triangular GEMVs on every level (the streamed kernel is complex-only), and with `invert_max_k` lowered the
    multi-step kernels -- sliced / chunked cluster sweeps, deferred contribution rows, per-step launches."""
    import scipy.sparse.linalg as spla

    pc = pencils.cavity_3d(8)
    sigma = 0.1
    h = _lib.Handle(pc.n, 0)
    flag = ((pc.A.diagonal() - sigma * pc.M.diagonal()) == 0).astype(np.uint8)
    info = h.analyze(pc.A.indptr, pc.A.indices, pc.M.indptr, pc.M.indices, leaf_size=64, order_last=flag)
    assert info.max_pivots > 512
    h.set_values(pc.A.data, pc.M.data)
    C = (pc.A - sigma * pc.M).tocsc()
    lu = spla.splu(C)
    b = np.random.default_rng(5).standard_normal(pc.n) + 1j * np.random.default_rng(6).standard_normal(pc.n)
    xs = lu.solve(b.real) + 1j * lu.solve(b.imag)
    for opts in (dict(), dict(invert_max_k=256), dict(invert_max_k=0), dict(invert_max_k=0, cluster_slices=0),
                 dict(invert_max_k=0, cluster_slices=0, defer_cb=0), dict(invert_max_k=0, use_clusters=0)):
        for opt, val in dict(dict(cluster_slices=1, defer_cb=1, use_clusters=1, invert_max_k=4096), **opts).items():
            h.set_option(opt, val)
        fs = h.factor(1.0, -sigma, _lib.LSA_F64, 1e-13)
        assert fs.scalar == _lib.LSA_F64 and fs.n_perturbed == 0
        x = h.solve(b)
        assert np.linalg.norm(x - xs) / np.linalg.norm(xs) < 1e-9
        assert np.linalg.norm(C @ x - b) / np.linalg.norm(b) < 1e-12
        xh = h.solve(b, _lib.LSA_OP_H)
        assert np.linalg.norm(C.conj().T @ xh - b) / np.linalg.norm(b) < 1e-12
    h.close()


def test_result_buffers_are_page_locked_and_recycled():
    import gc

    _lib.pinned_pool_clear()
    a = _lib.pinned_empty((1 << 17, 2), np.complex128)            # 4 MB: page-locked
    assert isinstance(a.base, np.ndarray) or a.base is not None
    addr = a.ctypes.data
    a[:] = 1.0 + 2.0j
    view = a[:, 1]
    del a
    gc.collect()
    assert view[5] == 1.0 + 2.0j                                   # a live view keeps the block
    assert not _lib._PINNED_POOL.get(1 << 22)
    del view
    gc.collect()
    assert _lib._PINNED_POOL.get(1 << 22) == [addr]                # ... and the last one returns it to the pool
    b = _lib.pinned_empty((1 << 18,), np.complex128)               # same size: same block again
    assert b.ctypes.data == addr
    small = _lib.pinned_empty((8,), np.float64)                    # small arrays stay pageable
    assert small.base is None
    del b
    gc.collect()
    _lib.pinned_pool_clear()
    assert not _lib._PINNED_POOL


# ------------------------------------------------------------------ C-ABI level kernels
@pytest.fixture(scope="module")
def factored():
    pc = pencils.assemble_pencil((30, 16), (8.0, 3.0), re=50.0, baseflow=pencils.wake_profile(0.9, 1.2, 1.5))
    sigma = 0.05 + 0.6j
    h = _lib.Handle(pc.n, 0)
    flag = ((pc.A.diagonal() - sigma * pc.M.diagonal()) == 0).astype(np.uint8)
    h.analyze(pc.A.indptr, pc.A.indices, pc.M.indptr, pc.M.indices, leaf_size=32, order_last=flag)
    h.set_values(pc.A.data, pc.M.data)
    fs = h.factor(1.0, -sigma, _lib.LSA_C128, 1e-13)
    yield pc, sigma, h, fs
    h.close()


def test_triangular_solves_all_modes(factored):
    pc, sigma, h, fs = factored
    assert fs.n_perturbed == 0 and fs.min_pivot > 0
    C = (pc.A - sigma * pc.M).tocsc()
    rng = np.random.default_rng(5)
    b = rng.standard_normal(pc.n) + 1j * rng.standard_normal(pc.n)
    for trans, op in ((_lib.LSA_OP_N, C), (_lib.LSA_OP_T, C.T), (_lib.LSA_OP_H, C.conj().T)):
        x = h.solve(b, trans)
        assert np.linalg.norm(op @ x - b) / np.linalg.norm(b) < 1e-12
        x = h.solve(b, trans, refine_steps=1)
        assert np.linalg.norm(op @ x - b) / np.linalg.norm(b) < 1e-13
    import scipy.sparse.linalg as spla

    xs = spla.splu(C).solve(b)
    assert np.linalg.norm(h.solve(b) - xs) / np.linalg.norm(xs) < 1e-10


def test_spmv_all_operators(factored):
    pc, _, h, _ = factored
    rng = np.random.default_rng(6)
    x = rng.standard_normal(pc.n) + 1j * rng.standard_normal(pc.n)
    for which, mat in ((_lib.LSA_MAT_A, pc.A), (_lib.LSA_MAT_M, pc.M)):
        for trans, op in ((_lib.LSA_OP_N, mat), (_lib.LSA_OP_T, mat.T), (_lib.LSA_OP_H, mat.conj().T)):
            y = h.spmv(which, x, trans)
            assert np.linalg.norm(y - op @ x) <= 1e-14 * np.linalg.norm(op @ x)


def test_spmv_row_blocks_with_empty_stretches_and_long_rows():
    """The streamed SpMV cuts the rows into blocks of <= 2048 entries / <= 1024 rows: stretches of empty rows (pressure
    rows of M), rows longer than one block (handled by a block-wide reduction) and complex values, N / T / H."""
    import scipy.sparse as sp

    n = 7000
    rng = np.random.default_rng(11)
    A = sp.random(n, n, density=4e-3, random_state=3, format="lil", dtype=np.float64)
    A.setdiag(4.0)
    A[17, :] = 0.0
    A[17, ::2] = rng.standard_normal(n // 2)             # 3500 entries in one row
    A[:, 4001] = rng.standard_normal((n, 1))             # ... and, transposed, in one row of A^T
    A = A.tocsr().astype(np.complex128)
    A.data = A.data + 1j * rng.standard_normal(A.nnz)
    A.sort_indices()
    M = sp.random(n, n, density=2e-3, random_state=4, format="lil", dtype=np.float64)
    M[2000:4500, :] = 0.0                                 # 2500 consecutive empty rows
    M[5000, 100:5100] = 1.5                               # a long real row
    M = M.tocsr()
    M.eliminate_zeros()
    M.sort_indices()
    h = _lib.Handle(n, 0)
    try:
        h.analyze(A.indptr, A.indices, M.indptr, M.indices, leaf_size=32)
        h.set_values(A.data, M.data)
        x = rng.standard_normal(n) + 1j * rng.standard_normal(n)
        for which, mat in ((_lib.LSA_MAT_A, A), (_lib.LSA_MAT_M, M)):
            for trans, op in ((_lib.LSA_OP_N, mat), (_lib.LSA_OP_T, mat.T), (_lib.LSA_OP_H, mat.conj().T)):
                y = h.spmv(which, x, trans)
                ref = op @ x
                assert np.abs(y - ref).max() <= 1e-13 * max(1.0, np.abs(ref).max())
    finally:
        h.close()


def test_real_factor_with_complex_vectors():
    pc, _ = _ns("th2d")
    h = _lib.Handle(pc.n, 0)
    h.analyze(pc.A.indptr, pc.A.indices, pc.M.indptr, pc.M.indices, leaf_size=32,
              order_last=((pc.A.diagonal() + 0.3 * pc.M.diagonal()) == 0).astype(np.uint8))
    h.set_values(pc.A.data, pc.M.data)
    fs = h.factor(1.0, 0.3, _lib.LSA_F64, 1e-13)
    assert fs.scalar == _lib.LSA_F64
    C = (pc.A + 0.3 * pc.M).tocsc()
    rng = np.random.default_rng(7)
    b = rng.standard_normal(pc.n) + 1j * rng.standard_normal(pc.n)
    for trans, op in ((_lib.LSA_OP_N, C), (_lib.LSA_OP_H, C.T)):
        assert np.linalg.norm(op @ h.solve(b, trans) - b) / np.linalg.norm(b) < 1e-12
    with pytest.raises(_lib.LsaError):
        h.factor(1.0, 0.3j, _lib.LSA_F64, 1e-13)
    h.close()


def test_dense_schur_kernel_and_gemm_tiles():
    h = _lib.Handle(4, 0)
    rng = np.random.default_rng(8)
    for m in (1, 2, 7, 40, 80, 120):
        S0 = rng.standard_normal((m, m)) + 1j * rng.standard_normal((m, m))
        T, Q = h.dense_schur(S0, "SMALLEST_REAL")
        assert np.abs(Q @ T @ Q.conj().T - S0).max() < 1e-12 * m
        assert np.abs(np.tril(T, -1)).max() == 0 and np.all(np.diff(np.diag(T).real) >= -1e-10)
    for scalar in (_lib.LSA_F64, _lib.LSA_C128):
        for (m, n, k) in ((1, 1, 1), (64, 64, 32), (65, 63, 33), (200, 130, 70), (33, 257, 19)):
            _, err = h.gemm_bench(scalar, m, n, k, 1)
            assert 0 <= err < 1e-12 * k
    h.close()


def test_tiny_pivot_replacement_counts_and_zero_pivot_error():
    A = sp.csr_matrix(np.array([[0.0, 1.0, 0.0], [1.0, 0.0, 0.0], [0.0, 0.0, 0.0]]))
    h = _lib.Handle(3, 0)
    h.analyze(A.indptr, A.indices)
    h.set_values(A.data, None)
    with pytest.raises(_lib.LsaError) as e:
        h.factor(1.0, 0.0, _lib.LSA_F64, 0.0)
    assert e.value.status == -3
    fs = h.factor(1.0, 0.0, _lib.LSA_F64, 1e-10)
    assert fs.n_perturbed >= 1
    h.close()


# ------------------------------------------------------------------ BASELINE config 1 at full size
def test_config1_full_size_properties_and_oracle():
    """2-D wake pencil, 50 303 DOFs, complex shift, nev = 10, ncv = 80 (.examples/eigenvalues.py path)."""
    pc = pencils.cylinder_wake_2d()
    sigma = 0.05 + 0.74j
    es, pairs = _run(pc, sigma, nev=10, ncv=80, tol=1e-10)
    assert len(pairs) == 10
    lam = np.array([p[0] for p in pairs])
    X = np.stack([_vec(v) for _, v in pairs], axis=1)
    assert O.north_star_residuals(pc.A, pc.M, lam, X).max() < RESID_BAR
    assert es.solver.get_residuals()[:10].max() < RESID_BAR
    orc = O.shift_invert_arpack(pc.A, pc.M, sigma, 12, ncv=80, tol=1e-12)
    assert _match(lam, orc.eigenvalues) < EIG_RTOL
    # the adjoint spectrum is the conjugate spectrum (size-independent property)
    es2, pairs2 = _run(pc, sigma, nev=10, ncv=80, tol=1e-10, adjoint=True)
    lam2 = np.array([p[0] for p in pairs2])
    adj = O.shift_invert_arpack(pc.A, pc.M, sigma, 12, ncv=80, tol=1e-12, adjoint=True)
    assert _match(lam2, adj.eigenvalues) < EIG_RTOL          # the adjoint modes against the oracle's adjoint run
    assert _match(np.conj(lam2[:8]), lam) < EIG_RTOL         # and the 8 leading ones against the direct spectrum


# ------------------------------------------------------------------ data-format and edge cases
def test_explicit_zero_pattern_of_m_gives_identical_results():
    """dolfinx assembles M on the full mixed-space pattern (explicit zeros in the pressure blocks)."""
    pc, sigma = _ns("th2d")
    Mz = sp.csr_matrix((np.zeros(pc.A.nnz), pc.A.indices.copy(), pc.A.indptr.copy()), shape=pc.A.shape)
    Mz.data[:] = 0.0
    Mc = pc.M.tocoo()
    lookup = {(i, j): v for i, j, v in zip(Mc.row, Mc.col, Mc.data)}
    rows = np.repeat(np.arange(pc.n), np.diff(pc.A.indptr))
    Mz.data[:] = [lookup.get((i, j), 0.0) for i, j in zip(rows, pc.A.indices)]
    assert Mz.nnz == pc.A.nnz and abs(Mz - pc.M).max() == 0
    pz = pencils.Pencil(A=pc.A, M=Mz, dofs_u=pc.dofs_u, dofs_p=pc.dofs_p, dirichlet=pc.dirichlet, coords=pc.coords)
    _, pairs = _run(pc, sigma)
    _, pairs_z = _run(pz, sigma)
    lam, lam_z = np.array([p[0] for p in pairs]), np.array([p[0] for p in pairs_z])
    assert _match(lam_z, lam) < 1e-10


def test_complex_valued_operators():
    pc, sigma = _ns("th2d")
    Ac = (pc.A + 0.03j * pc.M).tocsr()          # complex data: scatter, SpMV and residual paths in c128
    pcx = pencils.Pencil(A=Ac, M=pc.M, dofs_u=pc.dofs_u, dofs_p=pc.dofs_p, dirichlet=pc.dirichlet, coords=pc.coords)
    es, pairs = _run(pcx, sigma)
    lam = np.array([p[0] for p in pairs])
    orc = O.shift_invert_arpack(Ac, pc.M, sigma, 8, ncv=40, tol=1e-12)
    assert _match(lam, orc.eigenvalues) < EIG_RTOL
    assert es.solver.get_residuals()[:6].max() < RESID_BAR
    # eigenvalues of A + 0.03j M are those of (A, M) shifted by 0.03j
    _, base = _run(pc, sigma - 0.03j)
    assert _match(lam, np.array([p[0] for p in base]) + 0.03j) < EIG_RTOL


@pytest.mark.parametrize("n", [1, 2])
def test_tiny_orders(n):
    A = L.iPETScMatrix.from_matrix(np.diag(np.arange(1.0, n + 1.0)) + (np.eye(n, k=1) * 0.5 if n > 1 else 0))
    pairs = L.EigenSolver(A, cfg=L.EigensolverConfig(num_eig=5, atol=1e-10)).solve()
    assert sorted(p[0].real for p in pairs) == pytest.approx(np.arange(1.0, n + 1.0), abs=1e-10)
    es = L.EigenSolver(A, cfg=L.EigensolverConfig(num_eig=5, atol=1e-10))
    es.solver.set_st_type(L.iSTType.SINVERT)
    es.solver.set_target(0.3)
    assert sorted(p[0].real for p in es.solve()) == pytest.approx(np.arange(1.0, n + 1.0), abs=1e-10)


def test_non_convergence_returns_fewer_pairs_without_raising():
    pc, sigma = _ns("th2d")
    cfg = L.EigensolverConfig(num_eig=30, atol=1e-14, max_it=1, ncv=32)
    es = L.EigenSolver(L.iPETScMatrix(pc.A), L.iPETScMatrix(pc.M), cfg, check_hermitian=False)
    es.solver.set_st_type(L.iSTType.SINVERT)
    es.solver.set_target(sigma)
    pairs = es.solve()                      # Solver/utils.py:325-328: min(nconv, num) pairs, no exception
    assert len(pairs) == min(es.solver.get_num_converged(), 30) < 30
    assert es.solver.stats["n_restarts"] == 1


def test_singular_pencil_is_perturbed_not_crashed(caplog):
    """P1/P1 ("SIMPLE", not inf-sup stable) with these BCs has an empty pressure row: A - sigma M is exactly
    singular.  The backend replaces the tiny pivots (eigen2.py:135-136 reaches for MUMPS CNTL(3) for the same reason),
    retries with the robust pressure placement and reports the count instead of failing."""
    import logging

    caplog.set_level(logging.WARNING)
    pc = pencils.assemble_pencil((24, 16), (6.0, 2.0), re=40.0, baseflow=pencils.wake_profile(0.9, 1.2, 1.0), space="SIMPLE")
    es, pairs = _run(pc, 0.1 + 0.6j, nev=4, ncv=30, tol=1e-8)
    assert es.solver.stats["n_perturbed"] >= 1
    assert any("tiny pivots replaced" in r.getMessage() for r in caplog.records)


def test_large_front_paths_match_scipy_solve():
    """3-D cavity with fronts of several hundred pivots: whole-block inverses + triangular GEMVs (default), and with
    `invert_max_k` lowered the multi-step cluster sweeps; streamed (bulk-copy) sweeps in every tile geometry / ring
    depth / CTA shape, per-step launches, rank-128 deferred updates."""
    import scipy.sparse.linalg as spla

    pc = pencils.cavity_3d(8)
    sigma = 0.1 + 0.3j
    h = _lib.Handle(pc.n, 0)
    flag = ((pc.A.diagonal() - sigma * pc.M.diagonal()) == 0).astype(np.uint8)
    info = h.analyze(pc.A.indptr, pc.A.indices, pc.M.indptr, pc.M.indices, leaf_size=64, order_last=flag)
    assert info.max_pivots > 512
    h.set_values(pc.A.data, pc.M.data)
    fs = h.factor(1.0, -sigma, _lib.LSA_C128, 1e-13)
    assert fs.n_perturbed == 0 and fs.max_multiplier < 1e3
    C = (pc.A - sigma * pc.M).tocsc()
    b = np.random.default_rng(3).standard_normal(pc.n) + 1j * np.random.default_rng(4).standard_normal(pc.n)
    xs = spla.splu(C).solve(b)
    base = dict(use_clusters=1, use_graphs=1, use_stream=1, stream_min_fronts=96, stream_flags=3, invert_max_k=0,
                stream_small_rows=192, stream_stages=0, cluster_max_rows=8192, cluster_max_width=16, cluster_slices=1, defer_cb=1)
    variants = (dict(invert_max_k=4096), dict(invert_max_k=4096, use_graphs=0), dict(invert_max_k=4096, use_stream=0),
                dict(invert_max_k=300), dict(invert_max_k=300, use_stream=0), dict(invert_max_k=128, stream_min_fronts=1),
                dict(), dict(use_clusters=0, use_graphs=0),
                dict(use_stream=0), dict(use_stream=0, use_clusters=0, cluster_max_rows=0),
                dict(stream_min_fronts=1), dict(stream_min_fronts=1, stream_flags=0, stream_stages=2),
                dict(stream_min_fronts=1, stream_flags=7, stream_small_rows=0),
                dict(stream_min_fronts=4, stream_flags=1, stream_small_rows=100000, stream_stages=12),
                dict(cluster_max_rows=1280, cluster_max_width=4), dict(cluster_slices=0), dict(cluster_slices=0, defer_cb=0),
                dict(use_stream=0, use_graphs=0),
                dict(use_stream=0, cluster_max_width=2), dict(use_stream=0, cluster_max_width=1),
                dict(use_stream=0, cluster_max_width=8, use_graphs=0))
    for var in variants:
        opts = dict(base, **var)
        for opt, val in opts.items():
            h.set_option(opt, val)
        h.factor(1.0, -sigma, _lib.LSA_C128, 1e-13)      # re-factor: drops the captured sweep graphs
        x = h.solve(b)
        assert np.linalg.norm(x - xs) / np.linalg.norm(xs) < 1e-9
        assert np.linalg.norm(C @ x - b) / np.linalg.norm(b) < 1e-12
        xh = h.solve(b, _lib.LSA_OP_H)
        assert np.linalg.norm(C.conj().T @ xh - b) / np.linalg.norm(b) < 1e-12
    h.close()


# ------------------------------------------------------------------ the benchmarked pencil families against the oracle
def _kappa(pc, d, a, l):
    """Condition number of eigenvalue l from the oracle's right (d) and left (a) eigenvectors:
    kappa = ||x|| ||y|| / |y^H M x|."""
    j = int(np.argmin(abs(d.eigenvalues - l)))
    ja = int(np.argmin(abs(np.conj(a.eigenvalues) - d.eigenvalues[j])))
    x, y = d.eigenvectors[:, j], a.eigenvectors[:, ja]
    return j, np.linalg.norm(x) * np.linalg.norm(y) / max(abs(np.vdot(y, pc.M @ x)), 1e-300)


@pytest.mark.parametrize("shape", [(84, 21), (167, 42)], ids=["16k", "64k"])
def test_backward_step_pencil_against_oracle(shape):
    """BASELINE config 2 family (channel with a step shear layer, Re = 500, sigma next to the least stable modes),
    direct and adjoint modes, nev = 20, against the oracle's Krylov-Schur on the same matrices.

    The pencil is strongly non-normal (kappa(lambda) from 4e7 for the leading pair to > 1e15 for the 20th mode: the
    oracle's OWN direct, adjoint and ARPACK runs agree only to kappa * 1e-16 there), so eigenvalue parity is a
    statement about the modes whose conditioning admits one (kappa <= 1e9; every such mode must match to
    max(1e-8, 1e-14 kappa) <= 1e-5), kappa is printed per mode, and ALL modes must meet the backward-error bar
    ||A x - lambda M x|| / (||A||_F ||x||) <= 1e-10 -- each returned pair is an exact eigenpair of a pencil
    1e-10-close to the given one, which is all any backend can deliver for kappa ~ 1e15."""
    pc = pencils.backward_step_2d(*shape, re=500.0)
    sigma = -0.35 + 0.1j
    d = O.shift_invert_krylov_schur(pc.A, pc.M, sigma, 20, ncv=80, tol=1e-12)
    a = O.shift_invert_krylov_schur(pc.A, pc.M, sigma, 20, ncv=80, tol=1e-12, adjoint=True)
    es, pairs = _run(pc, sigma, nev=20, ncv=80, tol=1e-11)
    assert len(pairs) == 20
    lam = np.array([p[0] for p in pairs])
    X = np.stack([_vec(v) for _, v in pairs], axis=1)
    assert O.north_star_residuals(pc.A, pc.M, lam, X).max() < RESID_BAR
    ea, pairs_adj = _run(pc, sigma, nev=20, ncv=80, tol=1e-11, adjoint=True)
    assert len(pairs_adj) == 20
    lam_adj = np.array([p[0] for p in pairs_adj])
    Y = np.stack([_vec(v) for _, v in pairs_adj], axis=1)
    AH, MH = pc.A.conj().T.tocsr(), pc.M.conj().T.tocsr()
    assert O.north_star_residuals(AH, MH, lam_adj, Y).max() < RESID_BAR
    well = 0
    for tag, ls, conj in (("direct", lam, False), ("adjoint", lam_adj, True)):
        for l in ls:
            lt = np.conj(l) if conj else l
            j, kappa = _kappa(pc, d, a, lt)
            err = abs(lt - d.eigenvalues[j]) / abs(lt)
            print(f"step {shape} {tag}: lambda = {lt:.10g}  kappa = {kappa:.2e}  |d lambda|/|lambda| = {err:.1e}")
            if kappa <= 1e9:
                well += 1
                assert err < max(EIG_RTOL, 1e-14 * kappa), (tag, l, d.eigenvalues[j], kappa)
    assert well >= 2          # the leading conjugate pair is always well enough conditioned
    # the set of wanted modes: every oracle mode that is both well conditioned and among its 10 nearest to sigma is found
    for l in d.eigenvalues[:10]:
        j, kappa = _kappa(pc, d, a, l)
        if kappa <= 1e9:
            assert min(abs(lam - l)) / abs(l) < max(EIG_RTOL, 1e-14 * kappa)


def test_reynolds_sweep_mini_config3_against_oracle():
    """BASELINE config 3 in small: graded wake mesh, three (Re, sigma) pairs of the sweep on ONE symbolic analysis
    (the values change, the pattern does not), nev = 10, every pair against the oracle: the SET of 10 eigenvalues
    must agree (both inclusions), 1e-8 relative."""
    import scipy.sparse as sp

    from bench import SWEEP

    L.clear_symbolic_cache()
    pc = pencils.adapted_wake_2d(96, 24, re=SWEEP[0][0], split_viscous=True)
    M_c = L.iPETScMatrix(pc.M)
    cached = []
    for re_, sigma in (SWEEP[0], SWEEP[4], SWEEP[7]):
        A = sp.csr_matrix((pc.a_data_at(re_), pc.A.indices, pc.A.indptr), shape=pc.A.shape)
        cfg = L.EigensolverConfig(num_eig=10, atol=1e-11, max_it=200, ncv=80)
        es = L.EigenSolver(L.iPETScMatrix(A), M_c, cfg, check_hermitian=False)
        es.solver.set_st_type(L.iSTType.SINVERT)
        es.solver.set_target(sigma)
        es.solver.set_st_pc_type(L.PreconditionerType.LU)
        pairs = es.solve()
        cached.append(es.solver.stats["symbolic_cached"])
        assert len(pairs) == 10
        lam = np.array([p[0] for p in pairs])
        orc = O.shift_invert_krylov_schur(A, pc.M, sigma, 12, ncv=80, tol=1e-12)
        assert _match(lam, orc.eigenvalues) < EIG_RTOL                 # every GPU mode is an oracle mode
        assert _match(orc.eigenvalues[:10], lam) < EIG_RTOL            # and the oracle's 10 nearest are all found
        X = np.stack([_vec(v) for _, v in pairs], axis=1)
        assert O.north_star_residuals(A, pc.M, lam, X).max() < RESID_BAR
        assert es.solver.get_residuals()[:10].max() < RESID_BAR
    assert cached == [False, True, True]


# ------------------------------------------------------------------ device-pointer half of the C ABI (on_device = 1)
def test_device_pointer_entry_points():
    """Values, right-hand sides, SpMV operands and eigenvectors handed over as device memory (torch CUDA tensors ->
    `data_ptr()`, a `__cuda_array_interface__` exporter, and a DLPack capsule): same results as the host route."""
    import torch

    pc, sigma = _ns("th2d")
    dev = torch.device("cuda", 0)
    h = _lib.Handle(pc.n, 0)
    flag = ((pc.A.diagonal() == 0) & (pc.M.diagonal() == 0)).astype(np.uint8)
    h.analyze(pc.A.indptr, pc.A.indices, pc.M.indptr, pc.M.indices, leaf_size=32, order_last=flag)
    a_d, m_d = torch.from_numpy(pc.A.data).to(dev), torch.from_numpy(pc.M.data).to(dev)

    class CAI:      # a bare __cuda_array_interface__ exporter (what CuPy / Numba arrays look like)
        def __init__(self, t):
            self._t = t
            self.__cuda_array_interface__ = t.__cuda_array_interface__

    h.set_values_device(CAI(a_d), m_d)
    h.factor(1.0, -sigma, _lib.LSA_C128, 1e-13)
    C = (pc.A - sigma * pc.M).tocsr()
    rng = np.random.default_rng(11)
    b = rng.standard_normal(pc.n) + 1j * rng.standard_normal(pc.n)
    b_d = torch.from_numpy(b).to(dev)
    x_d = torch.empty_like(b_d)
    h.solve_device(b_d, x_d)
    x = x_d.cpu().numpy()
    assert np.linalg.norm(C @ x - b) / np.linalg.norm(b) < 1e-12
    assert np.allclose(x, h.solve(b), rtol=1e-12, atol=1e-14)
    h.solve_device(b_d, x_d, _lib.LSA_OP_H)
    assert np.linalg.norm(C.conj().T @ x_d.cpu().numpy() - b) / np.linalg.norm(b) < 1e-12
    h.solve_device(b_d, b_d)                                        # in place
    assert np.allclose(b_d.cpu().numpy(), x, rtol=1e-12, atol=1e-14)
    y_d = torch.empty_like(x_d)
    xin = torch.from_numpy(b).to(dev)
    for which, mat in ((_lib.LSA_MAT_A, pc.A), (_lib.LSA_MAT_M, pc.M)):
        h.spmv_device(which, xin, y_d)
        assert np.allclose(y_d.cpu().numpy(), mat @ b, rtol=1e-12, atol=1e-12)
        h.spmv_device(which, xin, y_d, _lib.LSA_OP_H)
        assert np.allclose(y_d.cpu().numpy(), mat.conj().T @ b, rtol=1e-12, atol=1e-12)
    r = h.eigs(nev=4, ncv=40, tol=1e-11, max_restarts=100, which="TARGET_MAGNITUDE", transform=_lib.LSA_ST_SINVERT, sigma=sigma)
    assert r.nconv >= 4
    X_d = torch.empty((4, pc.n), dtype=torch.complex128, device=dev)
    assert h.eigenvectors_device(X_d, 4) == 4
    assert np.allclose(X_d.cpu().numpy().T, h.eigenvectors(4), rtol=0, atol=1e-15)
    # a DLPack exporter that is not a torch tensor
    class DL:
        def __init__(self, t):
            self._t = t

        def __dlpack__(self, **kw):
            return self._t.__dlpack__(**kw)

        def __dlpack_device__(self):
            return self._t.__dlpack_device__()

    h.set_values_device(DL(a_d), DL(m_d))
    h.factor(1.0, -sigma, _lib.LSA_C128, 1e-13)
    assert np.allclose(h.solve(b), x, rtol=1e-12, atol=1e-14)
    # the reference-facing route: device_values backend option
    es, pairs = _run(pc, sigma, nev=4, device_values=(a_d, m_d))
    es2, pairs2 = _run(pc, sigma, nev=4)
    assert np.allclose([p[0] for p in pairs], [p[0] for p in pairs2], rtol=1e-12)
    with pytest.raises(ValueError):
        h.set_values_device(torch.from_numpy(pc.A.data), m_d)      # host tensor
    with pytest.raises(TypeError):
        h.set_values_device(a_d.float(), m_d)
    h.close()


# ------------------------------------------------------------------ robustness paths named by the round-1 review
def test_nonfinite_start_vector_is_reported_not_converged():
    """A NaN in the Krylov basis is LSA_ERR_NONFINITE, not a 'breakdown' that marks everything converged."""
    pc, sigma = _ns("th2d")
    v0 = np.random.default_rng(0).standard_normal(pc.n).astype(np.complex128)
    v0[pc.dofs_u[pc.dofs_u.size // 2]] = np.nan       # a velocity unknown (M has no entries in pressure columns)
    with pytest.raises(L.LsaError) as e:
        _run(pc, sigma, v0=v0)
    assert e.value.status == -4


def test_ncv_above_kernel_limit_is_an_argument_error():
    pc, sigma = _ns("th2d")
    with pytest.raises(L.LsaError) as e:
        _run(pc, sigma, nev=4, ncv=300)
    assert e.value.status == -1 and "256" in str(e.value)


def test_purification_through_the_krylov_schur_relation_equals_an_explicit_apply():
    """x = V y + v_next (b . y) / theta  ==  OP (V y) / theta: same eigenvectors, nev fewer operator applications."""
    pc, sigma = _ns("th2d")
    es1, p1 = _run(pc, sigma, nev=6, purify=True)
    es2, p2 = _run(pc, sigma, nev=6, purify="explicit")
    es0, p0 = _run(pc, sigma, nev=6, purify=False)
    assert es2.solver.stats["n_op_applies"] == es1.solver.stats["n_op_applies"] + es2.solver.stats["nconv"]
    for (l1, v1), (l2, v2), (l0, v0_) in zip(p1, p2, p0):
        assert l1 == pytest.approx(l2, rel=1e-12)
        assert np.linalg.norm(_vec(v1) - _vec(v2)) < 1e-9          # phase is fixed on the device
    # purified vectors have no component outside range(OP): pressure-free M makes M x reproduce lambda-scaled A x
    X = np.stack([_vec(v) for _, v in p1], axis=1)
    assert O.north_star_residuals(pc.A, pc.M, np.array([p[0] for p in p1]), X).max() < RESID_BAR


def test_adjoint_reuse_is_refused_when_the_shared_handle_moved_on():
    """Two solvers on one sparsity pattern share a native handle (symbolic cache).  After the second one has
    factored ITS shift, the adjoint solve that would reuse the first one's factors must notice and re-factor."""
    L.clear_symbolic_cache()
    pc, sigma = _ns("th2d")
    A, M = L.iPETScMatrix(pc.A), L.iPETScMatrix(pc.M)
    cfg = L.EigensolverConfig(num_eig=4, atol=1e-11, max_it=200, ncv=40)

    def make(a, m, s):
        es = L.EigenSolver(a, m, cfg, check_hermitian=False)
        es.solver.set_st_type(L.iSTType.SINVERT)
        es.solver.set_target(s)
        es.solver.set_st_pc_type(L.PreconditionerType.LU)
        return es

    es1 = make(A, M, sigma)
    lam1 = np.array([p[0] for p in es1.solve()])
    es2 = make(A, M, sigma + 0.2)             # same pattern -> same handle, other factors
    es2.solve()
    assert es2.solver.handle is es1.solver.handle
    with pytest.raises(RuntimeError):
        es1.solver.get_residuals()            # device-side results of es1 are gone
    ea = make(A.H, M.H, np.conj(sigma))
    lam_adj = np.array([p[0] for p in ea.solve()])
    assert not ea.solver.stats.get("reused_factorisation")
    assert _match(np.conj(lam_adj), lam1) < EIG_RTOL
    # in-place edit of the donor's matrix between the direct and the adjoint solve
    es3 = make(A, M, sigma)
    es3.solve()
    A.scale(1.0)                              # new value array object: the factors no longer belong to `A`
    eb = make(A.H, M.H, np.conj(sigma))
    eb.solve()
    assert not eb.solver.stats.get("reused_factorisation")
    # untouched donor: reuse
    es4 = make(A, M, sigma)
    es4.solve()
    ec = make(A.H, M.H, np.conj(sigma))
    ec.solve()
    assert ec.solver.stats.get("reused_factorisation") is True


def test_multiplier_growth_triggers_reanalysis_then_refinement(caplog):
    """`max_multiplier` above `growth_limit`: robust pressure placement first, iterative refinement second."""
    import logging

    L.clear_symbolic_cache()
    pc, sigma = _ns("th2d")
    with caplog.at_level(logging.WARNING):
        es, pairs = _run(pc, sigma, nev=4, growth_limit=1.0)      # every LU has multipliers > 1 somewhere
    assert es.solver.stats.get("refine_steps_forced") == 2
    assert any("re-analysing" in r.message for r in caplog.records) and any("refinement" in r.message for r in caplog.records)
    es0, pairs0 = _run(pc, sigma, nev=4)
    assert np.allclose([p[0] for p in pairs], [p[0] for p in pairs0], rtol=1e-9)
    assert es.solver.get_residuals()[:4].max() < RESID_BAR


# ------------------------------------------------------------------ one solve split over several GPUs (row e)
@pytest.mark.parametrize("world", [2, 4, 8])
def test_partitioned_solve_over_gpus(world):
    """Sub-trees per GPU + replicated top, NCCL broadcast / all-reduce inside the CUDA library: solves to 1e-12,
    eigenvalues identical to the single-GPU run, same results on every rank (tests/workers/partitioned_gpu_worker.py)."""
    import os
    import subprocess
    import sys

    import torch

    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
                          "--master-addr", "127.0.0.1", "--master-port", str(29650 + world),
                          os.path.join(root, "tests", "workers", "partitioned_gpu_worker.py")],
                         capture_output=True, text=True, timeout=900)
    assert out.returncode == 0, (out.stdout[-2000:], out.stderr[-4000:])
    assert "PARTITIONED_OK" in out.stdout
    print(out.stdout[-1500:])


# ------------------------------------------------------------------ rows f2 / f3: the callers on either side of the path
def test_linear_solver_seam_runs_on_the_lu_kernels():
    """`Solver/linear.py:38-87` / `Solver/nonlinear2.py:61-70`: direct (PREONLY + LU) and LU-preconditioned GMRES solves
    of Jacobian-like systems (sparsity of A, real FP64), symbolic analysis reused from one Newton step to the next."""
    import scipy.sparse.linalg as spla

    L.clear_symbolic_cache()
    pc = pencils.assemble_pencil((24, 12), (8.0, 3.0), re=50.0, baseflow=pencils.wake_profile(0.9, 1.2, 1.5), split_viscous=True)
    rng = np.random.default_rng(8)
    b = rng.standard_normal(pc.n)
    cached = []
    for re_ in (50.0, 60.0, 75.0):                         # "Newton steps": new values on the same pattern
        J = sp.csr_matrix((pc.a_data_at(re_), pc.A.indices, pc.A.indptr), shape=pc.A.shape)
        ref = spla.splu(J.tocsc()).solve(b)
        ksp = L.iKSP(L.iPETScMatrix(J))
        ksp.set_type(L.KSPType.PREONLY)
        ksp.set_preconditioner(L.PreconditionerType.LU)
        x = ksp.solve(L.iPETScVector.from_array(b.copy()))
        xa = x.as_array()
        assert not np.iscomplexobj(xa) and ksp.stats["scalar"] == "f64"
        assert np.linalg.norm(J @ xa - b) / np.linalg.norm(b) < 1e-11
        assert np.linalg.norm(xa - ref) / np.linalg.norm(ref) < 1e-8
        assert ksp.get_iteration_number() == 1 and ksp.raw.getType() == "preonly"
        cached.append(ksp.stats["symbolic_cached"])
    assert cached == [False, True, True]
    J = sp.csr_matrix((pc.a_data_at(50.0), pc.A.indices, pc.A.indptr), shape=pc.A.shape)
    sol = L.LinearSolver.solve(L.iPETScMatrix(J), L.iPETScVector.from_array(b.copy()), ksp_type=L.KSPType.GMRES, tol=1e-14, rtol=1e-13)
    assert np.linalg.norm(J @ sol.as_array() - b) / np.linalg.norm(b) < 1e-12
    # complex system, solution written into a caller-owned vector
    Jc = (J + 0.3j * pc.M).tocsr()
    bc = b + 1j * rng.standard_normal(pc.n)
    ksp = L.iKSP(L.iPETScMatrix(Jc))
    ksp.set_type(L.KSPType.GMRES)
    ksp.set_tolerances(tol=1e-14, rtol=1e-13, max_it=5)
    out = L.iPETScVector.from_array(np.zeros(pc.n, dtype=complex))
    ksp.solve(L.iPETScVector.from_array(bc.copy()), out)
    assert np.linalg.norm(Jc @ out.as_array() - bc) / np.linalg.norm(bc) < 1e-12 and ksp.get_residual_norm() < 1e-10
    assert ksp.get_solution() is out
    with pytest.raises(ValueError):
        L.LinearSolver.solve(L.iPETScMatrix(J), L.iPETScVector.from_array(b), ksp_type=L.KSPType.CG)
    bad = L.iKSP(L.iPETScMatrix(J))
    bad.set_preconditioner(L.PreconditionerType.JACOBI)
    with pytest.raises(NotImplementedError):
        bad.solve(L.iPETScVector.from_array(b))


def test_eigenvalue_sensitivity_contraction_matches_finite_differences():
    """`Sensitivity/__init__.py:354-385` with pre-assembled operators: d lambda / d Re = a^H (dA/dRe) v / (a^H M v),
    dA/dRe = -(1/Re^2) dA/d(1/Re), contracted on the device; checked against centred differences of the
    eigenvalue itself (fixed base flow: the explicit term is the whole derivative)."""
    pc = pencils.assemble_pencil((24, 12), (8.0, 3.0), re=50.0, baseflow=pencils.wake_profile(0.9, 1.2, 1.5), split_viscous=True)
    re0, sigma = 50.0, 0.05 + 0.6j
    M_c = L.iPETScMatrix(pc.M)

    def modes(re_, adjoint=False):
        A = sp.csr_matrix((pc.a_data_at(re_), pc.A.indices, pc.A.indptr), shape=pc.A.shape)
        cfg = L.EigensolverConfig(num_eig=4, atol=1e-12, max_it=300, ncv=40)
        es = L.EigenSolver(L.iPETScMatrix(A), M_c, cfg, check_hermitian=False)
        es.solver.set_st_type(L.iSTType.SINVERT)
        es.solver.set_target(sigma)
        es.solver.set_st_pc_type(L.PreconditionerType.LU)
        es.solver.set_backend_options(leaf_size=32)
        es.solver.set_adjoint(adjoint)
        return es, es.solve()

    es, pairs = modes(re0)
    lam, v = L.select_mode(pairs, sigma)
    ea, pairs_adj = modes(re0, adjoint=True)
    lam_adj, a = L.select_mode(pairs_adj, np.conj(lam))
    assert abs(np.conj(lam_adj) - lam) < 1e-9 * abs(lam)
    dl = L.eigenvalue_sensitivity(es, lam, v, a, -(1.0 / re0**2) * pc.a_visc)
    # host check of the contraction itself
    va, aa = _vec(v), _vec(a)
    K = sp.csr_matrix((pc.a_visc, pc.A.indices, pc.A.indptr), shape=pc.A.shape)
    assert dl == pytest.approx(-(1.0 / re0**2) * np.vdot(aa, K @ va) / np.vdot(aa, pc.M @ va), rel=1e-10)
    # finite differences of the eigenvalue
    d = 0.05
    lp = L.select_mode(modes(re0 + d)[1], lam)[0]
    lm = L.select_mode(modes(re0 - d)[1], lam)[0]
    fd = (lp - lm) / (2 * d)
    print(f"d lambda / d Re: contraction {dl:.8e}  finite differences {fd:.8e}")
    assert abs(dl - fd) < 1e-5 * abs(fd) + 1e-9


def test_attached_pressure_nullspace_is_projected_not_pinned():
    """Enclosed flow WITHOUT a pinned pressure DOF: A and M share the constant-pressure nullspace, A - sigma M is singular
    for every shift.  With the nullspace attached to A (`FEM/operators.py:534-545`, `FEM/utils.py:604-607`) the solver
    replaces the vanishing pivot and projects the nullspace out of every operator application (KSP's behaviour;
    `Solver/eigen2.py:171-176`): same spectrum as the pinned formulation (`FEM/utils.py:596-602`)."""
    pinned = pencils.cavity_3d(6)
    free = pencils.cavity_3d(6, pin_pressure=False)
    sigma = -1.2 + 0.1j          # next to the least damped physical modes, away from the spurious lambda = 1 cluster
    nvec = np.zeros(free.n)
    nvec[free.dofs_p] = 1.0
    ns = L.iPETScNullSpace.from_vectors([L.iPETScVector.from_array(nvec)])
    A = L.iPETScMatrix(free.A)
    ok, nrm = ns.test_matrix(A, tol=1e-10)
    assert ok, nrm
    ns.attach_to(A)
    assert A.get_nullspace() is ns and ns.dimension == 1
    cfg = L.EigensolverConfig(num_eig=6, atol=1e-11, max_it=200, ncv=40)

    def run(Ac, Mc):
        es = L.EigenSolver(Ac, Mc, cfg, check_hermitian=False)
        es.solver.set_st_type(L.iSTType.SINVERT)
        es.solver.set_target(sigma)
        es.solver.set_st_pc_type(L.PreconditionerType.LU)
        es.solver.set_backend_options(leaf_size=32)
        return es, es.solve()

    es, pairs = run(A, L.iPETScMatrix(free.M))
    assert es.solver.stats["nullspace_dimension"] == 1 and es.solver.stats["n_perturbed"] >= 1
    es_p, pairs_p = run(L.iPETScMatrix(pinned.A), L.iPETScMatrix(pinned.M))
    lam, lam_p = np.array([p[0] for p in pairs]), np.array([p[0] for p in pairs_p])
    # the pinned formulation carries one more spurious eigenvalue 1 (the pinned row); physical modes agree
    phys = lam_p[np.abs(lam_p - 1.0) > 1e-6]
    mine = lam[np.abs(lam - 1.0) > 1e-6]
    assert len(mine) >= 3 and _match(mine[:3], phys) < EIG_RTOL
    X = np.stack([_vec(v) for _, v in pairs], axis=1)
    assert O.north_star_residuals(free.A, free.M, lam, X).max() < RESID_BAR
    assert np.abs(nvec @ X).max() / np.linalg.norm(nvec) < 1e-9          # no constant-pressure component in the modes
    # the adjoint modes on the same (singular) factorisation
    ea = L.EigenSolver(A.H, L.iPETScMatrix(free.M).H, cfg, check_hermitian=False)
    ea.solver.set_st_type(L.iSTType.SINVERT)
    ea.solver.set_target(np.conj(sigma))
    ea.solver.set_st_pc_type(L.PreconditionerType.LU)
    ea.solver.set_backend_options(leaf_size=32)
    lam_a = np.array([p[0] for p in ea.solve()])
    mine_a = lam_a[np.abs(lam_a - 1.0) > 1e-6]
    assert _match(np.conj(mine_a[:3]), phys) < EIG_RTOL


def _poisson3d(n1: int) -> sp.csr_matrix:
    t = sp.diags([-np.ones(n1 - 1), 2.0 * np.ones(n1), -np.ones(n1 - 1)], [-1, 0, 1])
    e = sp.identity(n1)
    return (sp.kron(sp.kron(t, e), e) + sp.kron(sp.kron(e, t), e) + sp.kron(sp.kron(e, e), t)).tocsr()


def test_symmetric_ldlt_factor_and_sweeps():
    """Row f4: the symmetric L D L^T variant (no U blocks, no pivoting; `Elasticity/utils.py:139-155` asks PETSc for
    CHOLESKY).  3-D Poisson + mass shift: half the factor storage of the LU, same solutions, N = T = H."""
    import scipy.sparse.linalg as spla

    K = _poisson3d(22)
    n = K.shape[0]
    Mm = sp.identity(n, format="csr") * 0.5
    Mm = sp.csr_matrix((np.full(K.nnz, 0.0), K.indices, K.indptr), shape=K.shape) + Mm   # same pattern family, diagonal mass
    Mm.sort_indices()
    rng = np.random.default_rng(12)
    b = rng.standard_normal(n)
    stats = {}
    for symm in (0, 1):
        h = _lib.Handle(n)
        if symm:
            h.set_option("symmetric", 1)
        info = h.analyze(K.indptr, K.indices, Mm.indptr, Mm.indices, leaf_size=32)
        h.set_values(K.data, Mm.data)
        for sigma in (-0.3, 0.013):          # definite, and a shift just above the lowest modes (indefinite, no growth here)
            fs = h.factor(1.0, -sigma, _lib.LSA_F64, 0.0)
            assert fs.n_perturbed == 0
            C = (K - sigma * Mm).tocsc()
            ref = spla.splu(C).solve(b)
            for trans in (_lib.LSA_OP_N, _lib.LSA_OP_T, _lib.LSA_OP_H):
                x = h.solve(b.astype(complex), trans)
                assert np.abs(x.imag).max() == 0.0
                assert np.linalg.norm(C @ x.real - b) / np.linalg.norm(b) < 1e-12
                assert np.linalg.norm(x.real - ref) / np.linalg.norm(ref) < 1e-10
            bc = b + 1j * rng.standard_normal(n)     # real factor, complex right-hand side
            xc = h.solve(bc)
            assert np.linalg.norm(C @ xc - bc) / np.linalg.norm(bc) < 1e-12
        stats[symm] = (info.nnz_lu, fs.flops, info.max_front)
        if symm:
            with pytest.raises(Exception):
                h.factor(1.0, 0.1j, _lib.LSA_C128, 0.0)      # the symmetric mode is real FP64 only
        h.close()
    assert stats[1][0] < 0.62 * stats[0][0] and stats[1][1] < 0.62 * stats[0][1]
    assert stats[1][2] > 256                                  # root separator beyond one 128-pivot panel


def test_ghep_with_cholesky_uses_the_symmetric_factor_and_matches_lu():
    """GHEP + SINVERT at 0 + CHOLESKY (the reference's elasticity modal solve, `Elasticity/utils.py:139-155`) runs on the
    L D L^T factor and in the M-inner product; eigenpairs equal those of the general LU path and of the oracle.  An
    indefinite matrix that breaks the unpivoted factorisation falls back to LU."""
    L.clear_symbolic_cache()
    pm = pencils.membrane_pencil(40, 40, 1.0, 0.83, m_bc_diagonal=1e-6)      # spurious Dirichlet eigenvalue at 1e6
    ref = np.sort(O.shift_invert_arpack(pm.A, pm.M, 0.0, 16, tol=1e-13).eigenvalues.real)[:12]
    out = {}
    for pc in (L.PreconditionerType.LU, L.PreconditionerType.CHOLESKY):
        cfg = L.EigensolverConfig(num_eig=12, problem_type=L.iEpsProblemType.GHEP, atol=1e-12, max_it=200)
        es = L.EigenSolver(L.iPETScMatrix(pm.A), L.iPETScMatrix(pm.M), cfg, check_hermitian=False)
        es.solver.set_st_type(L.iSTType.SINVERT)
        es.solver.set_target(0.0)
        es.solver.set_st_pc_type(pc)
        pairs = es.solve()
        assert all(isinstance(v, float) for v, _ in pairs)
        lam = np.array([v for v, _ in pairs][:12])
        X = np.stack([_vec(v) for _, v in pairs[:12]], axis=1)
        assert O.north_star_residuals(pm.A, pm.M, lam, X).max() < RESID_BAR
        G = X.conj().T @ (pm.M @ X)                      # M-orthonormal vectors (SLEPc's GHEP normalisation)
        assert np.abs(G - np.eye(12)).max() < 1e-8
        out[pc] = (np.sort(lam.real), dict(es.solver.stats))
        assert np.abs(out[pc][0] - ref).max() <= EIG_RTOL * np.abs(ref).max()
    assert out[L.PreconditionerType.CHOLESKY][1]["symmetric_factorisation"] is True
    assert out[L.PreconditionerType.LU][1]["symmetric_factorisation"] is False
    assert out[L.PreconditionerType.CHOLESKY][1]["b_mode"] == 2 and out[L.PreconditionerType.LU][1]["b_mode"] == 2
    assert out[L.PreconditionerType.CHOLESKY][1]["nnz_lu"] < 0.7 * out[L.PreconditionerType.LU][1]["nnz_lu"]
    # linear seam: PREONLY + CHOLESKY
    K = (pm.A + 0.5 * pm.M).tocsr()
    rhs = np.random.default_rng(3).standard_normal(pm.n)
    ksp = L.iKSP(L.iPETScMatrix(K))
    ksp.set_type(L.KSPType.PREONLY)
    ksp.set_preconditioner(L.PreconditionerType.CHOLESKY)
    x = ksp.solve(L.iPETScVector.from_array(rhs.copy())).as_array()
    assert ksp.stats["symmetric_factorisation"] is True
    assert np.linalg.norm(K @ x - rhs) / np.linalg.norm(rhs) < 1e-12
    # a saddle-point matrix (zero diagonal block): whatever the unpivoted factorisation makes of it, the answer is right
    pc = pencils.assemble_pencil((10, 6), (4.0, 2.0), re=20.0, baseflow=pencils.zero_flow())
    S = ((pc.A + pc.A.T) * 0.5).tocsr()
    S.sort_indices()
    rhs = np.random.default_rng(4).standard_normal(pc.n)
    ksp = L.iKSP(L.iPETScMatrix(S))
    ksp.set_type(L.KSPType.PREONLY)
    ksp.set_preconditioner(L.PreconditionerType.CHOLESKY)
    x = ksp.solve(L.iPETScVector.from_array(rhs.copy())).as_array()
    assert np.linalg.norm(S @ x - rhs) / np.linalg.norm(rhs) < 1e-10


def test_elasticity_modal_solve_in_the_mass_inner_product():
    """The reference's elasticity modal problem (`Elasticity/utils.py:139-155`: GHEP, target 0, SINVERT, CHOLESKY, 24
    pairs) on a deep plane-stress cantilever (NAFEMS-like free vibration): L D L^T factor, M-orthonormal Lanczos-type
    Krylov-Schur (`b_mode = 2`), eigenvalues against the ARPACK oracle, mass-normalised modes as `process_modes`
    (`Elasticity/utils.py:85-96`) expects them, and the Euclidean-inner-product variant for comparison."""
    L.clear_symbolic_cache()
    pc = pencils.elasticity_pencil(96, 24, m_bc_diagonal=1e-14)
    nev = 24
    ref = np.sort(O.shift_invert_arpack(pc.A, pc.M, 0.0, nev + 4, tol=1e-13).eigenvalues.real)[:nev]
    res = {}
    for b_inner in ("auto", False):
        cfg = L.EigensolverConfig(num_eig=25, atol=1e-11, max_it=300)
        es = L.EigenSolver(L.iPETScMatrix(pc.A), L.iPETScMatrix(pc.M), cfg, check_hermitian=False)
        sol = es.solver
        sol.set_problem_type(L.iEpsProblemType.GHEP)
        sol.set_target(0.0)
        sol.set_st_type(L.iSTType.SINVERT)
        sol.set_st_pc_type(L.PreconditionerType.CHOLESKY)
        sol.set_dimensions(number_eigenpairs=nev)
        sol.set_backend_options(b_inner=b_inner, coords=pc.coords)
        pairs = es.solve()
        assert len(pairs) >= nev
        lam = np.array([v for v, _ in pairs][:nev])
        X = np.stack([_vec(v) for _, v in pairs[:nev]], axis=1)
        assert not np.iscomplexobj(X) or np.abs(X.imag).max() < 1e-9
        order = np.argsort(lam)
        lam, X = lam[order], X[:, order].real
        assert np.abs(lam - ref).max() <= EIG_RTOL * np.abs(ref).max()
        G = X.T @ (pc.M @ X)
        assert np.abs(np.diag(G) - 1.0).max() < 1e-9                     # v^H M v = 1 (mass_chk of the reference)
        Kp = X.T @ (pc.A @ X)
        assert np.abs(np.diag(Kp) - lam).max() <= 1e-7 * np.abs(lam).max()   # Rayleigh quotient = omega^2
        if b_inner == "auto":
            assert np.abs(G - np.eye(nev)).max() < 1e-8                  # the whole set is M-orthonormal
        st = sol.stats
        assert st["symmetric_factorisation"] is True and st["b_mode"] == (2 if b_inner == "auto" else 1)
        assert O.north_star_residuals(pc.A, pc.M, lam, X).max() < RESID_BAR
        res[b_inner] = (lam, st["n_op_applies"])
    assert np.abs(res["auto"][0] - res[False][0]).max() <= EIG_RTOL * np.abs(ref).max()
    hz = np.sqrt(res["auto"][0][0]) / (2 * np.pi)
    assert 14.5 < hz < 16.5                                              # Euler-Bernoulli: 16.15 Hz, lower for a deep beam
