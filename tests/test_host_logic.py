"""CPU tests of the host side: C-ABI surface, symbolic phase, Rayleigh-Ritz core, Python boundary."""

import ctypes
import logging
import os
import re
import subprocess
import sys

import numpy as np
import pytest
import scipy.sparse as sp

import lsa_fw_b200 as L
from lsa_fw_b200 import _lib, pencils
from mf_emulator import Emulator

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


# ------------------------------------------------------------------------------- C ABI
def test_library_loads_and_exports_every_declared_symbol():
    lib = _lib.load()
    header = open(os.path.join(ROOT, "include", "lsa_b200.h")).read()
    declared = set(re.findall(r"\b(lsa_[a-z_0-9]+)\s*\(", header))
    assert declared >= set(_lib.EXPORTS)
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/lsa_b200.h but not exported"
    assert b"sm_100a" in lib.lsa_version()


def test_struct_layouts_match_header():
    assert ctypes.sizeof(_lib.SymbolicInfo) == 4 * 8 + 8 * 7 + 8 + 32
    assert ctypes.sizeof(_lib.EigsParams) == 8 * 4 + 8 * 3 + 8 + 8 + 8
    assert ctypes.sizeof(_lib.FactorStats) == 8 * 4 + 8 + 24
    assert ctypes.sizeof(_lib.EigsResult) == 4 * 4 + 6 * 8 + 4 * 4 + 8


def test_numeric_entry_points_fail_loudly_without_device():
    h = _lib.Handle(3, device=-1)
    a = sp.identity(3, format="csr")
    h.analyze(a.indptr, a.indices)
    with pytest.raises(_lib.LsaError):
        h.set_values(a.data, None)


# ------------------------------------------------------------------------------- symbolic phase
CASES = [
    ("th2d", dict(shape=(12, 8), lengths=(6.0, 2.0)), {}),
    ("th3d", dict(shape=(5, 4, 3), lengths=(2.0, 1.0, 1.0)), {}),
    ("cavity3d", dict(shape=(5, 5, 5), lengths=(1.0, 1.0, 1.0)),
     dict(dirichlet_faces=("x0", "x1", "y0", "y1", "z0", "z1"), pin_pressure=True)),
    ("mini2d", dict(shape=(16, 10), lengths=(6.0, 2.0)), dict(space="MINI")),
]


@pytest.mark.parametrize("name,geo,kw", CASES, ids=[c[0] for c in CASES])
@pytest.mark.parametrize("use_coords", [False, True], ids=["graph", "geometric"])
def test_symbolic_structures_drive_a_correct_multifrontal_lu(name, geo, kw, use_coords):
    pc = pencils.assemble_pencil(geo["shape"], geo["lengths"], re=40.0, baseflow=pencils.wake_profile(0.9, 1.2, 1.0), **kw)
    n, sigma = pc.n, 0.1 + 0.6j
    h = _lib.Handle(n, device=-1)
    flag = ((pc.A.diagonal() - sigma * pc.M.diagonal()) == 0).astype(np.uint8)
    info = h.analyze(pc.A.indptr, pc.A.indices, pc.M.indptr, pc.M.indices, leaf_size=24,
                     coords=pc.coords if use_coords else None, order_last=flag)
    perm = h.symbolic_array("perm")
    assert sorted(perm.tolist()) == list(range(n))
    parent, level = h.symbolic_array("parent"), h.symbolic_array("level")
    has_parent = parent >= 0
    assert np.all(level[has_parent] == level[parent[has_parent]] + 1) and np.all(level[~has_parent] == 0)
    assert np.all(parent[has_parent] > np.nonzero(has_parent)[0])          # postorder
    st_ptr, st_idx, sn_ptr = h.symbolic_array("st_ptr"), h.symbolic_array("st_idx"), h.symbolic_array("sn_ptr")
    for s in range(info.n_fronts):
        rows = st_idx[st_ptr[s]:st_ptr[s + 1]]
        assert np.all(np.diff(rows) > 0) and (len(rows) == 0 or rows[0] >= sn_ptr[s + 1])
    assert info.n_decoupled == len(pc.dirichlet) + len(pc.meta["pinned"])
    em = Emulator(h, n)
    em.factor(pc.A.data, pc.M.data, 1.0, -sigma)
    C = (pc.A - sigma * pc.M).tocsc()
    b = np.random.default_rng(0).standard_normal(n) + 1j * np.random.default_rng(1).standard_normal(n)
    assert np.linalg.norm(C @ em.solve(b) - b) / np.linalg.norm(b) < 1e-11
    assert np.linalg.norm(C.conj().T @ em.solve(b, "H") - b) / np.linalg.norm(b) < 1e-11


def test_symmetric_layout_drives_a_correct_ldlt():
    """Row f4 host logic: option "symmetric" lays the factor store out without Q blocks (nnz_lu = sum k^2 + k r), the
    scatter maps drop the U12 entries, and the structures drive a correct L D L^T factorisation + solve (membrane
    stiffness / mass pencil, the reference's GHEP + Cholesky use, `Elasticity/utils.py:139-155`)."""
    from mf_emulator import SymEmulator

    pm = pencils.membrane_pencil(14, 10)
    n = pm.n
    sizes = {}
    for symm in (0, 1):
        h = _lib.Handle(n, device=-1)
        h.set_option("symmetric", symm)
        info = h.analyze(pm.A.indptr, pm.A.indices, pm.M.indptr, pm.M.indices, leaf_size=24)
        k, r = h.symbolic_array("front_k").astype(float), h.symbolic_array("front_r").astype(float)
        sizes[symm] = (info.nnz_lu, info.factor_entries, info.flops_real)
        assert info.nnz_lu == int((k * k + (1 if symm else 2) * k * r).sum()) + info.n_decoupled
    assert sizes[1][1] < 0.75 * sizes[0][1] and sizes[1][2] < sizes[0][2]
    a_dst = h.symbolic_array("a_dst")
    assert (a_dst < 0).sum() > 0 and (a_dst >= 0).sum() >= pm.A.nnz // 2
    sigma = 0.0
    em = SymEmulator(h, n)
    em.factor(pm.A.data, pm.M.data, 1.0, -sigma)
    C = (pm.A - sigma * pm.M).tocsc()
    b = np.random.default_rng(0).standard_normal(n)
    x = em.solve(b)
    assert np.linalg.norm(C @ x - b) / np.linalg.norm(b) < 1e-11
    with pytest.raises(_lib.LsaError):
        h.set_option("symmetric", 0)          # after the analysis: too late


def test_symbolic_counters_and_pattern_reuse():
    pc = pencils.assemble_pencil((20, 10), (6.0, 2.0), re=40.0)
    h = _lib.Handle(pc.n, device=-1)
    info = h.analyze(pc.A.indptr, pc.A.indices, pc.M.indptr, pc.M.indices, leaf_size=32)
    k, r = h.symbolic_array("front_k").astype(float), h.symbolic_array("front_r").astype(float)
    assert info.nnz_lu == int((k * k + 2 * k * r).sum()) + info.n_decoupled
    assert info.flops_real == pytest.approx((2 / 3 * k**3 + 2 * k * k * r + 2 * k * r * r).sum())
    assert k.sum() + info.n_decoupled == pc.n
    a_dst, m_dst = h.symbolic_array("a_dst"), h.symbolic_array("m_dst")
    assert len(np.unique(a_dst)) == len(a_dst) == pc.A.nnz and len(m_dst) == pc.M.nnz
    assert a_dst.max() < info.factor_entries


def test_options_and_pressure_placement_rule():
    pc = pencils.assemble_pencil((20, 10), (6.0, 2.0), re=40.0)
    flag = (pc.A.diagonal() == 0).astype(np.uint8)
    entries = {}
    for frac in (0.0, 0.5, 1.0):
        h = _lib.Handle(pc.n, device=-1)
        h.set_option("coupled_fraction", frac)
        info = h.analyze(pc.A.indptr, pc.A.indices, pc.M.indptr, pc.M.indices, leaf_size=32, order_last=flag)
        entries[frac] = info.nnz_lu
        # a flagged unknown never precedes the required share of its regular neighbours
        iperm, sn_ptr = h.symbolic_array("iperm"), h.symbolic_array("sn_ptr")
        sn_of = np.searchsorted(sn_ptr, iperm, side="right") - 1
        G = (pc.A + pc.A.T).tocsr()
        for v in np.nonzero(flag)[0][:200]:
            nb = G.indices[G.indptr[v]:G.indptr[v + 1]]
            nb = nb[(flag[nb] == 0) & (nb != v)]
            if len(nb) == 0 or iperm[v] < info.n_decoupled:
                continue
            need = max(1, int(np.ceil(frac * len(nb))))
            assert np.sum(sn_of[nb] <= sn_of[v]) >= need
    assert entries[0.0] <= entries[0.5] <= entries[1.0]
    h = _lib.Handle(3, device=-1)
    with pytest.raises(_lib.LsaError):
        h.set_option("coupled_fraction", 1.5)
    with pytest.raises(_lib.LsaError):
        h.set_option("no_such_option", 1.0)
    # sweep tuning knobs (include/lsa_b200.h): accepted without a device, value-checked where it matters
    for name, val in (("use_graphs", 0), ("use_clusters", 1), ("use_stream", 1), ("stream_min_fronts", 48),
                      ("stream_small_rows", 0), ("stream_stages", 6), ("stream_flags", 7), ("cluster_max_rows", 4096),
                      ("cluster_max_width", 8), ("cluster_slices", 0), ("invert_max_k", 1024), ("defer_cb", 0),
                      ("ortho_refine_always", 1)):
        h.set_option(name, val)
    for name, val in (("cluster_max_width", 3), ("stream_stages", 1), ("stream_stages", 13), ("invert_max_k", -1),
                      ("use_subtrees", 1), ("cluster_lookahead", 1)):   # the last two were removed with their kernels
        with pytest.raises(_lib.LsaError):
            h.set_option(name, val)
    h.set_option("use_graphs", 0)
    h.set_option("use_clusters", 0)


# ------------------------------------------------------------------------------- Rayleigh-Ritz core
@pytest.fixture(scope="module")
def rr_lib():
    out = os.path.join(ROOT, "tests", "_build")
    os.makedirs(out, exist_ok=True)
    so = os.path.join(out, "librr_host.so")
    src = os.path.join(ROOT, "tests", "csrc", "rr_host.cpp")
    hdr = os.path.join(ROOT, "lsa_fw_b200", "csrc", "rr_core.h")
    if not os.path.exists(so) or os.path.getmtime(so) < max(os.path.getmtime(src), os.path.getmtime(hdr)):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-I/usr/local/cuda/include",
                               "-x", "c++", src, "-o", so])
    return ctypes.CDLL(so)


def _p(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_double))


@pytest.mark.parametrize("m", [1, 2, 3, 5, 17, 80])
def test_rr_schur_form(rr_lib, m):
    rng = np.random.default_rng(m)
    S0 = rng.standard_normal((m, m)) + 1j * rng.standard_normal((m, m))
    S, Q = np.asfortranarray(S0.copy()), np.zeros((m, m), complex, order="F")
    assert rr_lib.rr_host_schur(m, m, 0, _p(S), _p(Q)) == 0
    assert np.abs(Q @ S @ Q.conj().T - S0).max() < 1e-12 * max(1, np.abs(S0).max()) * m
    assert np.abs(np.tril(S, -1)).max() == 0.0
    assert np.abs(Q.conj().T @ Q - np.eye(m)).max() < 1e-13 * m
    assert np.sort_complex(np.diag(S)) == pytest.approx(np.sort_complex(np.linalg.eigvals(S0)), abs=1e-10)


def test_rr_defective_and_repeated(rr_lib):
    for S0 in (np.array([[1, 1], [0, 1]], complex), np.diag([2.0, 2.0, 3.0]).astype(complex)):
        m = S0.shape[0]
        S, Q = np.asfortranarray(S0.copy()), np.zeros((m, m), complex, order="F")
        assert rr_lib.rr_host_schur(m, m, 0, _p(S), _p(Q)) == 0
        assert np.abs(Q @ S @ Q.conj().T - S0).max() < 1e-14


@pytest.mark.parametrize("which,key", [(1, lambda l, s: -abs(l)), (3, lambda l, s: -l.real), (4, lambda l, s: l.real),
                                       (7, lambda l, s: abs(l - s)), (8, lambda l, s: abs(l.real - s.real))])
def test_rr_full_ordering_locking_and_restart(rr_lib, which, key):
    rng = np.random.default_rng(7)
    m, ld, nconv, nev, sigma = 30, 31, 3, 6, 0.2 + 0.5j
    S = np.zeros((ld, m), complex, order="F")
    S[:m, :] = np.triu(rng.standard_normal((m, m)) + 1j * rng.standard_normal((m, m)), -1)
    S[nconv:, :nconv] = 0.0                       # locked block is triangular and decoupled
    S[:nconv, :nconv] = np.triu(S[:nconv, :nconv])
    S[m, m - 1] = 0.37
    S0 = S.copy()
    Q, theta, resid, out = np.zeros((m, m), complex, order="F"), np.zeros(m, complex), np.zeros(m), np.zeros(4, np.int32)
    rr_lib.rr_host_full(m, ld, nconv, nev, which, 1, 0, ctypes.c_double(1e-30), ctypes.c_double(sigma.real),
                        ctypes.c_double(sigma.imag), ctypes.c_double(1.0), _p(S), _p(Q), _p(theta), _p(resid),
                        out.ctypes.data_as(ctypes.POINTER(ctypes.c_int)))
    k, keep, status = out[:3]
    assert status == 0 and k == nconv and keep == nconv + (m - nconv) // 2
    assert np.abs(Q[:nconv, :nconv] - np.eye(nconv)).max() == 0 and np.abs(Q[:nconv, nconv:]).max() == 0
    lam = sigma + 1.0 / theta[nconv:]
    keys = np.array([key(l, sigma) for l in lam])
    assert np.all(np.diff(keys) >= -1e-12)        # active block ordered by `which` on lambda = sigma + 1/theta
    T = Q.conj().T @ S0[:m, :m] @ Q
    assert np.abs(T[:keep, :keep] - S[:keep, :keep]).max() < 1e-12
    assert np.abs(S[keep, nconv:keep] - 0.37 * Q[m - 1, nconv:keep]).max() < 1e-13   # coupling row
    assert np.abs(S[keep, :nconv]).max() == 0 and np.abs(S[:, keep:]).max() == 0 and np.abs(S[keep + 1:, :]).max() == 0
    # residual estimate of the first active Ritz pair: beta |e_m^T Q y|
    w, Y = np.linalg.eig(T[:nconv + 1, :nconv + 1])
    y = Y[:, np.argmin(abs(w - theta[nconv]))]
    assert resid[nconv] == pytest.approx(abs(0.37 * Q[m - 1, :nconv + 1] @ y) / np.linalg.norm(y), rel=1e-8)


# ------------------------------------------------------------------------------- Python boundary
def test_enums_keep_reference_names_and_alias():
    assert L.iEpsWhich.SMALLEST_MAGNITUDE is L.iEpsWhich.LARGEST_REAL      # Solver/utils.py:157-158
    assert [m.name for m in L.iEpsProblemType] == ["HEP", "NHEP", "GHEP", "GNHEP", "PGNHEP", "GHIEP"]
    assert L.iEpsProblemType.from_string("gnhep") is L.iEpsProblemType.GNHEP
    with pytest.raises(ValueError):
        L.iEpsProblemType.from_string("nope")
    assert L.iEpsWhich.LARGEST_REAL.to_arpack() == "LR"
    with pytest.raises(ValueError):
        L.iEpsWhich.TARGET_REAL.to_arpack()
    assert {t.name for t in L.iSTType} == {"SHELL", "SHIFT", "SINVERT", "CAYLEY", "PRECOND", "FILTER"}
    assert [t.name for t in L.KSPType] == ["CG", "GMRES", "BICG", "BICGSTAB", "RICHARDSON", "CHEBYSHEV", "PREONLY", "QCG",
                                           "CGS", "GCR", "LSQR", "LGMRES", "FGMRES"]          # Solver/utils.py:96-124
    assert L.KSPType.FGMRES.to_petsc() == "fgmres"
    ksp = L.iKSP()
    ksp.set_type(L.KSPType.LGMRES)
    assert ksp.get_type() == "lgmres" and ksp.raw.getType() == "lgmres"
    ksp.reset()                                                           # Solver/utils.py:417-419
    assert ksp.get_iteration_number() == 0


def test_config_defaults_and_solver_properties():
    cfg = L.EigensolverConfig()
    assert (cfg.num_eig, cfg.problem_type, cfg.atol, cfg.max_it, cfg.ncv) == (5, L.iEpsProblemType.GNHEP, 1e-6, 500, 80)
    A = L.iPETScMatrix.from_matrix(np.diag([1.0, 1.5, -42.0]))
    cfg = L.EigensolverConfig(num_eig=3, problem_type=L.iEpsProblemType.GHEP, atol=1e-3, max_it=100)
    for es in (L.EigenSolver(cfg, A=A), L.EigenSolver(A, None, cfg), L.EigenSolver(A, cfg=cfg), L.EigenSolver(cfg, A)):
        assert es.config is cfg and isinstance(es.solver, L.iEpsSolver)      # test_eigen.py:87-104
        assert es.solver.raw.getTolerances() == (cfg.atol, cfg.max_it)
        assert es.solver.raw.getDimensions()[0] == cfg.num_eig
        assert es.solver.raw.getProblemType() == cfg.problem_type.to_slepc()


def test_constructor_validation():
    A = L.iPETScMatrix.from_matrix(np.ones((2, 3)))
    with pytest.raises(ValueError, match="must be square"):
        L.EigenSolver(A)
    with pytest.raises(ValueError, match="does not match"):
        L.EigenSolver(L.iPETScMatrix.from_matrix(np.eye(3)), L.iPETScMatrix.from_matrix(np.eye(2)))
    with pytest.raises(ValueError):
        L.iEpsSolver(M=L.iPETScMatrix.from_matrix(np.eye(3)))                  # test_eigen.py:81-84
    with pytest.raises(TypeError):
        L.EigenSolver()


def test_hermitian_warning(caplog):
    caplog.set_level(logging.WARNING)
    A = L.iPETScMatrix.from_matrix(np.diag([1.0, 1.5, -42.0]))
    A[0, 1] = 0.1
    A.assemble()
    L.EigenSolver(L.EigensolverConfig(problem_type=L.iEpsProblemType.GHEP), A=A)
    assert any("assumes Hermitian A" in r.getMessage() for r in caplog.records)   # test_eigen.py:188-200


@pytest.mark.parametrize("pc_type", list(L.PreconditionerType))
def test_set_st_pc_type_round_trip(pc_type):
    s = L.iEpsSolver(A=L.iPETScMatrix.from_matrix(np.diag([1.0, 1.5, -42.0])))
    s.set_problem_type(L.iEpsProblemType.HEP)
    s.set_dimensions(number_eigenpairs=3)
    s.set_tolerances(atol=1e-8, max_it=50)
    s.set_st_type(L.iSTType.SINVERT)
    s.set_target(2.0)
    s.set_st_pc_type(pc_type)
    assert s.raw.getST().getKSP().getPC().getType() == pc_type.name.lower()       # test_eigen.py:307-322


def test_dimension_rules_follow_slepc():
    big = L.iPETScMatrix.from_matrix(np.diag(np.arange(1.0, 601.0)))
    s = L.iEpsSolver(A=big)
    s.set_dimensions(number_eigenpairs=5)
    assert s.raw.getDimensions()[1] == 20            # max(2 nev, nev + 15)
    s.set_dimensions(number_eigenpairs=100)
    assert s.raw.getDimensions()[1] == 200
    s.set_dimensions(number_eigenpairs=200)
    assert s.raw.getDimensions()[1] == 256           # widest basis of the device kernels
    s.set_dimensions(number_eigenpairs=100, subspace_dimension=80)
    with pytest.raises(ValueError, match="must be at least nev"):
        s.raw.getDimensions()
    small = L.iEpsSolver(A=L.iPETScMatrix.from_matrix(np.diag([1.0, 1.5, -42.0])))
    small.set_dimensions(number_eigenpairs=3, subspace_dimension=80)
    assert small.raw.getDimensions()[1] == 3         # never wider than the problem


def test_unsupported_combinations_raise_not_implemented():
    A = L.iPETScMatrix.from_matrix(np.diag([1.0, 1.5, -42.0]))
    s = L.iEpsSolver(A)
    s.set_st_type(L.iSTType.CAYLEY)
    with pytest.raises(NotImplementedError):
        s.solve()
    s = L.iEpsSolver(A)
    s.set_st_type(L.iSTType.SINVERT)
    s.set_st_pc_type(L.PreconditionerType.ILU)
    with pytest.raises(NotImplementedError):
        s.solve()
    s = L.iEpsSolver(A)
    s.set_interval(0.0, 1.0)
    with pytest.raises(NotImplementedError):
        s.solve()


def test_solve_without_gpu_raises_instead_of_falling_back():
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    A = L.iPETScMatrix.from_matrix(np.diag([1.0, 1.5, -42.0]))
    with pytest.raises(L.LsaError):
        L.EigenSolver(A).solve()


def test_carriers_follow_reference_semantics(tmp_path):
    v = L.iComplexPETScVector(np.array([3.0, 0.0]), np.array([0.0, 4.0]))
    assert v.norm() == pytest.approx(5.0) and v.imag is not None
    w = L.iComplexPETScVector(np.array([1.0, 1.0]))
    assert v.dot(w) == pytest.approx(np.vdot([3, 4j], [1, 1]))                  # conjugates self
    v.scale(1j)
    assert v.as_array() == pytest.approx([3j, -4.0])
    z = L.iComplexPETScVector(L.iPETScVector(np.array([1 + 1j, 2.0])))          # complex-build flavour
    assert z.imag is None and z.real.raw.getArray(readonly=True).dtype == np.complex128
    assert z.dot(L.iPETScVector(np.array([1j, 1.0]))) == pytest.approx(np.vdot([1j, 1.0], [1 + 1j, 2.0]))
    with pytest.raises(ValueError):
        z.real.raw.getArray(readonly=True)[0] = 0
    M = L.iPETScMatrix.from_matrix(np.array([[2.0, 1.0], [0.0, 3.0]]))
    assert M.shape == (2, 2) and M.nonzero_entries == 3 and M.norm == pytest.approx(np.sqrt(14))
    assert not M.is_numerically_hermitian() and (M.H.as_array() == M.as_array().T).all()
    csr = M.as_scipy_array()
    assert csr.indices.dtype == np.int32 and csr.has_sorted_indices
    M.export(tmp_path / "m.mtx")
    assert (L.iPETScMatrix.from_path(tmp_path / "m.mtx").as_array() == M.as_array()).all()
    M.pin_dof(0)
    assert (M.as_array() == [[1.0, 0.0], [0.0, 3.0]]).all()


# ------------------------------------------------------------------------------- replicas (gloo, world 2)
def test_petsc_binary_and_matrix_market_round_trip(tmp_path):
    """Row f1: MatrixMarket (`FEM/utils.py:143-147`) and PETSc binary (`:222-230`, `:616-659`) ingest / export without
    PETSc, real and complex, incl. the explicit zeros dolfinx leaves in the pattern."""
    rng = np.random.default_rng(4)
    A = sp.random(40, 40, 0.1, random_state=5, format="csr") + sp.eye(40, format="csr")
    A.data[3] = 0.0                                   # an explicit zero stays part of the pattern
    for mat in (A.tocsr(), (A + 1j * sp.random(40, 40, 0.05, random_state=6, format="csr")).tocsr()):
        c = L.iPETScMatrix(mat)
        c.export(tmp_path / "m.bin")
        back = L.iPETScMatrix.load(tmp_path / "m.bin")
        b = back.as_scipy_array()
        assert b.dtype == mat.dtype and np.array_equal(b.indptr, mat.indptr) and np.array_equal(b.indices, mat.indices)
        assert np.array_equal(b.data, mat.data)
        c.export(tmp_path / "m.mtx")
        mm = L.iPETScMatrix.from_path(tmp_path / "m.mtx").as_scipy_array()
        assert abs(mm - mat).max() == 0
    raw = np.fromfile(tmp_path / "m.bin", dtype=">i4", count=4)
    assert raw[0] == 1211216 and raw[1] == 40 and raw[2] == 40          # the header PETSc's MatLoad expects
    (tmp_path / "junk.bin").write_bytes(b"\0" * 64)
    with pytest.raises(ValueError):
        L.iPETScMatrix.load(tmp_path / "junk.bin")


def test_shard_tasks():
    from lsa_fw_b200.replicas import shard_tasks

    assert shard_tasks(8, 0, 2) == [0, 2, 4, 6] and shard_tasks(8, 1, 2) == [1, 3, 5, 7]
    assert sorted(sum((shard_tasks(11, r, 4) for r in range(4)), [])) == list(range(11))


def test_run_sharded_gloo_world_size_2(tmp_path):
    script = tmp_path / "w.py"
    script.write_text(
        "import os, sys\nsys.path.insert(0, %r)\nimport torch.distributed as dist\n"
        "from lsa_fw_b200.replicas import run_sharded\n"
        "dist.init_process_group('gloo')\n"
        "res = run_sharded(list(range(7)), lambda t: (t * t, dist.get_rank()))\n"
        "assert [r[0] for r in res] == [t * t for t in range(7)], res\n"
        "assert [r[1] for r in res] == [t %% 2 for t in range(7)], res\n"
        "dist.destroy_process_group()\nprint('ok', flush=True)\n" % ROOT)
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                          "--master-addr", "127.0.0.1", "--master-port", "29631", str(script)],
                         capture_output=True, text=True, timeout=240)
    assert out.returncode == 0, out.stderr[-2000:]
    assert out.stdout.count("ok") == 2


@pytest.mark.parametrize("world", [2, 3, 4])
def test_partitioned_solve_gloo(world):
    """Row e (multi-GPU) host logic on CPU: sub-tree -> rank mapping, rank-local structures (ghost roots, cut pool,
    local levels), broadcast of the sub-tree roots' contribution blocks, ONE all-reduce over the replicated rows
    per solve -- emulated in NumPy per rank, collectives over gloo; N and H solves of a 2-D and a 3-D pencil must
    reproduce SciPy on every rank."""
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
                          "--master-addr", "127.0.0.1", "--master-port", str(29640 + world),
                          os.path.join(ROOT, "tests", "workers", "partition_emulator_worker.py")],
                         capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-3000:]
    assert out.stdout.count("resid N") == 2 * world


def test_partition_mapping_properties():
    """Every front has exactly one owner or is replicated; owned sub-trees are closed under descendants; the
    replicated rows + own rows of all ranks tile [0, n); local level lists put the replicated top above the
    rank's sub-trees; the value scatter maps of the ranks cover every matrix entry exactly once below the top."""
    pc = pencils.cavity_3d(6)
    flag = ((pc.A.diagonal() == 0) & (pc.M.diagonal() == 0)).astype(np.uint8)
    world = 4
    hg = _lib.Handle(pc.n, device=-1)
    hg.analyze(pc.A.indptr, pc.A.indices, pc.M.indptr, pc.M.indices, leaf_size=32, order_last=flag)
    parent_g, k_g = hg.symbolic_array("parent"), hg.symbolic_array("front_k")
    covered = np.zeros(pc.n, dtype=int)
    entry_owner_count = np.zeros(pc.A.nnz, dtype=int)
    owners = None
    for rank in range(world):
        h = _lib.Handle(pc.n, device=-1, rank=rank, world=world)
        info = h.analyze(pc.A.indptr, pc.A.indices, pc.M.indptr, pc.M.indices, leaf_size=32, order_last=flag)
        pi = h.partition_info()
        owner, g2l, l2g = h.symbolic_array("owner"), h.symbolic_array("g2l"), h.symbolic_array("l2g")
        if owners is None:
            owners = owner
        assert np.array_equal(owners, owner)                       # the mapping is the same on every rank
        assert pi.n_fronts_global == len(owner) and pi.n_fronts_local == len(l2g) == info.n_fronts
        for s in range(len(owner)):                                # sub-trees are closed under descendants
            if parent_g[s] >= 0 and owner[parent_g[s]] >= 0:
                assert owner[s] == owner[parent_g[s]]
            if owner[s] == -1 and parent_g[s] >= 0:
                assert owner[parent_g[s]] == -1                    # everything above a top front is top
        flags, level, k = h.symbolic_array("front_flags"), h.symbolic_array("level"), h.symbolic_array("front_k")
        for l, s in enumerate(l2g):
            if owner[s] == -1:
                assert flags[l] == 0 and level[l] < pi.n_top_levels
            elif owner[s] == rank:
                assert level[l] >= pi.n_top_levels and k[l] == k_g[s]
            else:
                assert flags[l] == 1 and k[l] == 0                 # ghost: another rank's sub-tree root
        for lo, hi in zip(h.symbolic_array("own_lo"), h.symbolic_array("own_hi")):
            covered[lo:hi] += 1
        if rank == 0:
            for lo, hi in zip(h.symbolic_array("top_lo"), h.symbolic_array("top_hi")):
                covered[lo:hi] += 1
        a_dst = h.symbolic_array("a_dst")
        entry_owner_count += (a_dst >= 0)
        assert pi.n_own_rows + pi.n_replicated_rows <= pc.n and pi.weight_mine <= pi.weight_max_subtrees * (1 + 1e-12)
    assert np.all(covered == 1)
    # entries of replicated fronts are assembled on every rank, everything else exactly once
    assert entry_owner_count.min() >= 1 and set(np.unique(entry_owner_count)) <= {1, world}


# ------------------------------------------------------------------------------- bench contract (CPU arm)
def test_bench_reference_arm_contract():
    """`bench.py --impl reference` runs without a GPU and prints ONE JSON line with the keys the driver reads
    (the oracle port is what it times; config 1 is the reference's own CPU-runnable case)."""
    import json
    import subprocess
    import sys

    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", "cfg1",
                          "--steps", "1", "--warmup", "0"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.strip().startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "s" and d["higher_is_better"] is False
    assert d["metric"].startswith("shift-invert eigensolve") and d["n_gpus"] == 1 and d["steps"] == 1
    assert d["value"] > 0 and d["e2e"] == {"value": d["value"], "unit": "s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and "SuperLU" in cb["sample"]
    assert "workload" in d["config"] and "model" not in d["config"]
    assert abs(complex(*cb["sample_eig0"]) - (-0.030502986 + 0.738731056j)) < 1e-6     # same leading mode as the GPU path


# ------------------------------------------------------------------------------- sensitivity helpers (row f2)
def test_select_mode_and_biorthonormal_scaling():
    rng = np.random.default_rng(0)
    n = 60
    Md = sp.random(n, n, 0.15, random_state=3) + sp.eye(n)
    M = L.iPETScMatrix(sp.csr_matrix(Md))
    v = rng.standard_normal(n) + 1j * rng.standard_normal(n)
    a = rng.standard_normal(n) + 1j * rng.standard_normal(n)
    # complex build (one complex vector, VecDot conjugates the argument): exactly a^H M v = 1
    ac = L.iComplexPETScVector(L.iPETScVector.from_array(a.copy()))
    prod = L.normalize_adjoint(ac, M, v)
    assert prod == pytest.approx(np.conj(np.vdot(a, Md @ v)))
    assert np.vdot(ac.as_array(), Md @ v) == pytest.approx(1.0, abs=1e-13)
    # real build ((real, imag) pair, the dot conjugates self, Sensitivity/__init__.py:280-287 as written): modulus 1
    ar = L.iComplexPETScVector.from_array(a.copy())
    L.normalize_adjoint(ar, M, L.iComplexPETScVector.from_array(v))
    assert abs(np.vdot(ar.as_array(), Md @ v)) == pytest.approx(1.0, abs=1e-13)
    with pytest.raises(RuntimeError, match="Bi-orthonormal"):
        L.normalize_adjoint(L.iComplexPETScVector(L.iPETScVector.from_array(a.copy())), M, np.zeros(n))
    pairs = [(1 + 1j, "x"), (0.2 - 0.5j, "y"), (0.3 + 0.5j, "z")]
    assert L.select_mode(pairs, 0.25 - 0.4j)[1] == "y" and L.select_mode(pairs, np.conj(0.25 - 0.4j))[1] == "z"
    with pytest.raises(RuntimeError):
        L.select_mode([], 0.0)


def test_page_locked_arrays_survive_the_scipy_constructor_without_a_copy():
    """Arrays handed out by `pinned_empty` must not look like small views of a larger base: SciPy's CSR constructor
    prunes (copies) those, which silently moves the values back to pageable memory (5x slower uploads)."""
    import ctypes as C

    from lsa_fw_b200 import _lib

    keep = []

    class FakeLib:      # page-locked allocation stands in: plain host memory
        @staticmethod
        def lsa_host_alloc(nbytes, ref):
            buf = C.create_string_buffer(int(nbytes.value))
            keep.append(buf)
            ref._obj.value = C.addressof(buf)
            return _lib.LSA_OK

        @staticmethod
        def lsa_host_free(ptr):
            return _lib.LSA_OK

    n = 1 << 18
    for dtype in (np.float64, np.complex128):
        a = _lib.pinned_empty((n,), dtype, lib=FakeLib)
        assert a.ctypes.data == C.addressof(keep[-1]) and a.dtype == dtype and a.shape == (n,)
        a[...] = 1.0
        m = sp.csr_matrix((a, np.arange(n, dtype=np.int32) % 1000, np.arange(0, n + 1, n // 1000, dtype=np.int32)), shape=(1000, 1000))
        assert m.data.ctypes.data == a.ctypes.data
        carrier = L.iPETScMatrix(m)
        assert carrier.as_scipy_array().data.ctypes.data == a.ctypes.data
    b = _lib.pinned_empty((n // 8, 8), np.complex128, lib=FakeLib)
    assert b.shape == (n // 8, 8) and b.flags.c_contiguous


def test_host_equal_is_an_exact_comparison():
    from lsa_fw_b200 import _lib

    rng = np.random.default_rng(0)
    a = rng.integers(0, 1 << 30, size=(1 << 22) + 12345, dtype=np.int32)
    b = a.copy()
    assert _lib.host_equal(a, b) and _lib.host_equal(a, a)
    for pos in (0, 1 << 20, a.size - 1):
        b[pos] ^= 1
        assert not _lib.host_equal(a, b)
        b[pos] ^= 1
    assert _lib.host_equal(a, b)
    assert not _lib.host_equal(a, b[:-1]) and not _lib.host_equal(a, b.astype(np.int64))
    assert _lib.host_equal(a[::2], b[::2]) and _lib.host_equal(a[:0], b[:0])


def test_symbolic_analysis_is_pinned_and_thread_count_independent():
    """The host analysis (parallel graph build with atomically claimed slots, task-parallel nested dissection on
    compact per-node subgraphs, first-touch arrays) returns the SAME ordering, fronts and scatter maps (a) as the
    analysis all GPU records of the round were taken on (`tests/golden/symbolic_digests.json`, written before the
    host-side rework), (b) whatever the number of host threads, including on a graph large enough to spawn the
    dissection tasks and the task loops of the sub-graph copies."""
    import json

    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
    import make_symbolic_digests as G

    gold = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "symbolic_digests.json")))
    gold_hub = gold.pop("hub")
    assert G.compute() == gold

    # 9-point grid graph, 640 x 420 = 268 800 vertices (> 200 000: task loops; > 20 000: dissection tasks),
    # structurally unsymmetric input (strict upper triangle dropped on every third row), one isolated vertex
    nx, ny = 640, 420
    ex = sp.diags([1.0, 1.0, 1.0], [-1, 0, 1], shape=(nx, nx))
    ey = sp.diags([1.0, 1.0, 1.0], [-1, 0, 1], shape=(ny, ny))
    a = sp.kron(ex, ey, format="csr")
    keep = np.ones(a.nnz, dtype=bool)
    rows = np.repeat(np.arange(a.shape[0]), np.diff(a.indptr))
    keep[(rows % 3 == 0) & (a.indices > rows)] = False
    iso = 12345
    keep[(rows == iso) | (a.indices == iso)] = False
    keep[(rows == iso) & (a.indices == iso)] = True
    a = sp.csr_matrix((a.data[keep], (rows[keep], a.indices[keep])), shape=a.shape)
    a.sort_indices()
    n = a.shape[0]
    ref = None
    try:
        for nthreads in (1, 3, 8):
            h = _lib.Handle(n, device=-1)
            info = h.analyze(a.indptr, a.indices, leaf_size=48, nthreads=nthreads)
            got = {nm: h.symbolic_array(nm) for nm in ("perm", "sn_ptr", "st_idx", "ea_map", "a_dst", "lvl_front")}
            h.close()
            assert info.n_decoupled == 1
            if ref is None:
                ref = got
                assert sorted(got["perm"].tolist()) == list(range(n))
            else:
                for nm, v in got.items():
                    assert np.array_equal(v, ref[nm]), (nm, nthreads)
    finally:
        _lib.Handle(3, device=-1).analyze(np.array([0, 1, 2, 3]), np.array([0, 1, 2]), nthreads=os.cpu_count() or 1)

    # a graph with hub vertices (one dense row, one dense column) peels one vertex per dissection step: ~n steps deep.
    # The dissection loops on the larger child, so this must neither overflow the stack nor change the ordering
    # (digest of the plain recursion, written by tests/golden/make_symbolic_digests.py before the rework).
    hub = G.hub_pattern()
    h = _lib.Handle(hub.shape[0], device=-1)
    info = h.analyze(hub.indptr, hub.indices, leaf_size=32)
    assert info.n_fronts > 3000
    assert G.digest(h) == gold_hub
    h.close()


def _random_pattern(seed: int, n: int, density: float, n_hubs: int, n_decoupled: int, n_blocks: int):
    """Random sparse test matrix: `n_blocks` diagonal blocks without coupling between them, hub rows / columns, rows that
    hold only their diagonal entry (decoupled 1x1 pivots), structurally unsymmetric, diagonally dominant values."""
    rng = np.random.default_rng(seed)
    a = sp.lil_matrix((n, n))
    cuts = np.sort(rng.choice(np.arange(1, n), size=min(n_blocks - 1, n - 1), replace=False)) if n_blocks > 1 else []
    lo = 0
    for hi in list(cuts) + [n]:
        m = hi - lo
        blk = sp.random(m, m, density=min(1.0, density), random_state=int(rng.integers(1 << 30)), format="lil")
        a[lo:hi, lo:hi] = blk
        lo = hi
    for _ in range(n_hubs):
        v = int(rng.integers(n))
        a[v, rng.random(n) < 0.5] = 1.0
        if rng.random() < 0.5:
            a[rng.random(n) < 0.5, v] = 1.0
    dec = rng.choice(n, size=min(n_decoupled, n), replace=False)
    for v in dec:
        a[v, :] = 0.0
        a[:, v] = 0.0
    a = a.tocsr()
    a.data[:] = rng.standard_normal(a.nnz)
    a = (a + sp.diags(np.asarray(abs(a).sum(axis=1)).ravel() + np.asarray(abs(a).sum(axis=0)).ravel() + 1.0)).tocsr()
    a.sort_indices()
    return a


def test_symbolic_analysis_on_random_patterns_property():
    """Property test of the host analysis on patterns no FE mesh produces (disconnected blocks, hub rows and columns,
    decoupled rows, tiny orders): the permutation is a permutation, the assembly tree is a postordered forest whose fronts
    cover the pattern (the scatter maps are complete), and the structures drive an LU whose N and H solves are exact."""
    from hypothesis import HealthCheck, given, settings
    from hypothesis import strategies as st

    @settings(max_examples=60, deadline=None, derandomize=True, suppress_health_check=list(HealthCheck))
    @given(seed=st.integers(0, 2**20), n=st.integers(1, 160), density=st.sampled_from([0.0, 0.01, 0.05, 0.3]),
           n_hubs=st.integers(0, 2), n_decoupled=st.integers(0, 5), n_blocks=st.integers(1, 4),
           leaf=st.sampled_from([1, 2, 8, 32]))
    def check(seed, n, density, n_hubs, n_decoupled, n_blocks, leaf):
        a = _random_pattern(seed, n, density, n_hubs, n_decoupled, n_blocks)
        h = _lib.Handle(n, device=-1)
        try:
            info = h.analyze(a.indptr, a.indices, leaf_size=leaf)
            perm = h.symbolic_array("perm")
            assert sorted(perm.tolist()) == list(range(n))
            parent = h.symbolic_array("parent")
            has_parent = parent >= 0
            assert np.all(parent[has_parent] > np.nonzero(has_parent)[0])
            assert np.all(h.symbolic_array("a_dst") >= 0)
            offdiag = a - sp.diags(a.diagonal())
            offdiag.eliminate_zeros()
            lonely = (np.diff(offdiag.tocsr().indptr) == 0) & (np.diff(offdiag.tocsc().indptr) == 0)
            assert info.n_decoupled == int(lonely.sum())
            em = Emulator(h, n)
            em.factor(a.data.astype(np.complex128), None, 1.0, 0.0)
            rng = np.random.default_rng(seed)
            b = rng.standard_normal(n) + 1j * rng.standard_normal(n)
            assert np.linalg.norm(a @ em.solve(b) - b) <= 1e-11 * np.linalg.norm(b)
            assert np.linalg.norm(a.conj().T @ em.solve(b, "H") - b) <= 1e-11 * np.linalg.norm(b)
        finally:
            h.close()

    check()
