"""Staged GPU bring-up check: every stage is independent and reports its own verdict.

Usage on a GPU box:  python tools/gpu_check.py [--big]   (writes gpurun_out/gpu_check.json)
The oracle (SciPy) is used here only as the checker.
"""

from __future__ import annotations

import json
import os
import sys
import time
import traceback

import numpy as np



ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from lsa_fw_b200 import _lib, pencils  # noqa: E402
from oracle import eigen_oracle as O  # noqa: E402

OUT: dict = {}


def stage(name):
    def deco(fn):
        def run(*a, **k):
            t0 = time.time()
            try:
                res = fn(*a, **k)
                OUT[name] = {"ok": True, "seconds": time.time() - t0, **(res or {})}
            except Exception as e:  # noqa: BLE001
                OUT[name] = {"ok": False, "error": f"{type(e).__name__}: {e}", "trace": traceback.format_exc()[-1500:]}
            print(name, json.dumps(OUT[name], default=str)[:1500], flush=True)
        return run
    return deco


@stage("fp64_peak_cublas")
def fp64_peak():
    import torch

    n = 8192
    a = torch.randn(n, n, dtype=torch.float64, device="cuda")
    b = torch.randn(n, n, dtype=torch.float64, device="cuda")
    for _ in range(2):
        a @ b
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        a @ b
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    tf = 2 * n**3 / (best * 1e-3) / 1e12
    az = torch.randn(4096, 4096, dtype=torch.complex128, device="cuda")
    bz = torch.randn(4096, 4096, dtype=torch.complex128, device="cuda")
    az @ bz
    torch.cuda.synchronize()
    bestz = 1e9
    for _ in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        az @ bz
        e1.record()
        torch.cuda.synchronize()
        bestz = min(bestz, e0.elapsed_time(e1))
    tfz = 8 * 4096**3 / (bestz * 1e-3) / 1e12
    return {"dgemm_tflops": tf, "zgemm_tflops_real_equiv": tfz, "gpu": torch.cuda.get_device_name(0)}


@stage("gemm_dmma")
def gemm():
    h = _lib.Handle(4, 0)
    res = {}
    for sc, name in ((0, "real"), (1, "cplx")):
        for (m, n, k) in ((100, 70, 50), (513, 257, 129), (64, 64, 32)):
            ms, err = h.gemm_bench(sc, m, n, k, 2)
            res[f"{name}_{m}x{n}x{k}_err"] = err
        for (m, n, k) in ((4096, 4096, 4096), (8192, 8192, 32), (8192, 8192, 512)):
            ms, err = h.gemm_bench(sc, m, n, k, 5)
            fl = (8 if sc else 2) * m * n * k
            res[f"{name}_{m}x{n}x{k}_tflops"] = fl / (ms * 1e-3) / 1e12
    h.close()
    return res


@stage("dense_schur")
def dense_schur():
    h = _lib.Handle(4, 0)
    rng = np.random.default_rng(0)
    res = {}
    for m in (2, 5, 33, 80):
        S0 = rng.standard_normal((m, m)) + 1j * rng.standard_normal((m, m))
        t0 = time.time()
        T, Q = h.dense_schur(S0, "LARGEST_MAGNITUDE")
        res[f"m{m}_recon"] = float(np.abs(Q @ T @ Q.conj().T - S0).max())
        res[f"m{m}_lower"] = float(np.abs(np.tril(T, -1)).max())
        res[f"m{m}_sorted"] = bool(np.all(np.diff(np.abs(np.diag(T))) <= 1e-9))
        res[f"m{m}_sec"] = time.time() - t0
    h.close()
    return res


def _pencil(kind):
    if kind == "tiny2d":
        return pencils.assemble_pencil((12, 8), (6.0, 2.0), re=40.0, baseflow=pencils.wake_profile(0.9, 1.2, 1.0))
    if kind == "small2d":
        return pencils.assemble_pencil((40, 20), (12.0, 4.0), re=50.0, baseflow=pencils.wake_profile(0.9, 1.2, 2.0))
    if kind == "cfg1":
        return pencils.cylinder_wake_2d()
    if kind == "tiny3d":
        return pencils.cavity_3d(5)
    if kind == "small3d":
        return pencils.cavity_3d(10)
    if kind == "mid3d":
        return pencils.cavity_3d(16)
    if kind == "cfg2_quarter":
        return pencils.backward_step_2d(334, 84)
    if kind == "cfg2":
        return pencils.backward_step_2d()
    raise ValueError(kind)


def run_pencil(kind, sigma, complex_factor=True, eig=True, nev=6, ncv=40, leaf=64, use_coords=False):
    pc = _pencil(kind)
    n = pc.n
    res = {"n": n}
    h = _lib.Handle(n, 0)
    dF = pc.A.diagonal() - sigma * pc.M.diagonal()
    t0 = time.time()
    info = h.analyze(pc.A.indptr, pc.A.indices, pc.M.indptr, pc.M.indices, leaf_size=leaf,
                     coords=pc.coords if use_coords else None, order_last=(dF == 0).astype(np.uint8))
    res.update(analyze_s=time.time() - t0, fronts=info.n_fronts, levels=info.n_levels, nnz_lu=info.nnz_lu,
               max_front=info.max_front, flops_real=info.flops_real)
    h.set_values(pc.A.data, pc.M.data)
    rng = np.random.default_rng(1)
    x = rng.standard_normal(n) + 1j * rng.standard_normal(n)
    y = h.spmv(_lib.LSA_MAT_M, x)
    res["spmv_M_err"] = float(np.linalg.norm(y - pc.M @ x) / np.linalg.norm(pc.M @ x))
    y = h.spmv(_lib.LSA_MAT_A, x, _lib.LSA_OP_H)
    res["spmv_AH_err"] = float(np.linalg.norm(y - pc.A.conj().T @ x) / np.linalg.norm(pc.A.T @ x))
    fs = h.factor(1.0, -sigma, _lib.LSA_C128 if complex_factor else _lib.LSA_F64, 1e-13)
    res.update(factor_s=fs.seconds, factor_tflops=fs.flops / fs.seconds / 1e12, perturbed=fs.n_perturbed,
               swaps=fs.n_row_swaps, min_piv=fs.min_pivot, max_piv=fs.max_pivot, factor_kernels=fs.n_kernels)
    C = (pc.A - sigma * pc.M).tocsc()
    b = rng.standard_normal(n) + 1j * rng.standard_normal(n)
    xs = h.solve(b)
    res["solve_N_resid"] = float(np.linalg.norm(C @ xs - b) / np.linalg.norm(b))
    xs = h.solve(b, _lib.LSA_OP_H)
    res["solve_H_resid"] = float(np.linalg.norm(C.conj().T @ xs - b) / np.linalg.norm(b))
    xs = h.solve(b, _lib.LSA_OP_T)
    res["solve_T_resid"] = float(np.linalg.norm(C.T @ xs - b) / np.linalg.norm(b))
    xs = h.solve(b, _lib.LSA_OP_N, 1)
    res["solve_N_refined_resid"] = float(np.linalg.norm(C @ xs - b) / np.linalg.norm(b))
    t0 = time.time()
    for _ in range(5):
        h.solve(b)
    res["solve_host_roundtrip_ms"] = (time.time() - t0) / 5 * 1e3
    if eig:
        v0 = rng.standard_normal(n).astype(complex)
        r = h.eigs(nev=nev, ncv=ncv, tol=1e-10, max_restarts=100, which="TARGET_MAGNITUDE",
                   transform=_lib.LSA_ST_SINVERT, sigma=sigma, v0=v0)
        lam = h.eigenvalues(r.nconv)
        res.update(nconv=r.nconv, restarts=r.n_restarts, applies=r.n_op_applies, eigs_s=r.seconds,
                   t_solve=r.seconds_solve, t_spmv=r.seconds_spmv, t_ortho=r.seconds_ortho, t_rr=r.seconds_rr,
                   t_restart=r.seconds_restart, lam=[complex(z) for z in lam[:nev]])
        resid = h.residuals(r.nconv)
        res["resid_max"] = float(resid[:nev].max()) if r.nconv else None
        X = h.eigenvectors(r.nconv)
        if r.nconv:
            res["resid_host_max"] = float(O.north_star_residuals(pc.A, pc.M, lam, X)[:nev].max())
        if n <= 60000 or kind == "cfg2_quarter":
            orc = O.shift_invert_arpack(pc.A, pc.M, sigma, nev, ncv=ncv, tol=1e-10, seed=3)
            k = min(nev, r.nconv, len(orc.eigenvalues))
            d = [min(abs(l - orc.eigenvalues[:k + 2])) / abs(l) for l in lam[:k]]
            res["eig_rel_err_vs_oracle"] = float(max(d)) if d else None
            res["oracle_applies"] = orc.n_op_applies
            res["oracle_s"] = orc.seconds
        # adjoint on the same factors
        r2 = h.eigs(nev=nev, ncv=ncv, tol=1e-10, max_restarts=100, which="TARGET_MAGNITUDE",
                    transform=_lib.LSA_ST_SINVERT, sigma=sigma, adjoint=True, v0=v0)
        lam2 = h.eigenvalues(r2.nconv)
        k = min(nev, r.nconv, r2.nconv)
        if k:
            res["adjoint_conj_err"] = float(max(min(abs(np.conj(l) - lam2)) / abs(l) for l in lam[:k]))
            res["lam_adjoint"] = [complex(z) for z in lam2[:nev]]
            res["adjoint_resid_max"] = float(h.residuals(r2.nconv)[:k].max())
    h.close()
    return res


def main():
    big = "--big" in sys.argv
    fp64_peak()
    gemm()
    dense_schur()
    for kind, sigma in (("tiny2d", 0.0 + 0.5j), ("small2d", 0.05 + 0.6j), ("tiny3d", 0.1 + 0.2j)):
        stage(f"pencil_{kind}_cplx")(run_pencil)(kind, sigma)
    stage("pencil_tiny2d_real")(run_pencil)("tiny2d", -0.3, complex_factor=False)
    stage("pencil_small2d_geo")(run_pencil)("small2d", 0.05 + 0.6j, use_coords=True)
    stage("pencil_cfg1")(run_pencil)("cfg1", 0.05 + 0.74j, nev=10, ncv=80)
    stage("pencil_small3d")(run_pencil)("small3d", 0.1 + 0.2j, nev=6, ncv=40)
    if big:
        stage("pencil_mid3d")(run_pencil)("mid3d", 0.1 + 0.2j, nev=6, ncv=40)
        stage("pencil_cfg2_quarter")(run_pencil)("cfg2_quarter", 1.0j, nev=20, ncv=80)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "gpu_check.json"), "w") as f:
        json.dump(OUT, f, indent=1, default=str)
    bad = [k for k, v in OUT.items() if not v.get("ok")]
    print("FAILED STAGES:", bad)


if __name__ == "__main__":
    main()
