"""Export a benchmark pencil and this backend's eigenvalues so that someone WITH a PETSc/SLEPc install
can close the parity loop against the reference's own path.

    python tools/export_for_slepc.py cfg1 out_dir        # on a GPU box

Writes out_dir/A.mtx, out_dir/M.mtx (MatrixMarket, the interchange format of
.examples/eigenvalues.py:74-77) and out_dir/eigs_b200.json (sigma, nev, eigenvalues, residuals).
With the reference installed:  A = iPETScMatrix.from_path("A.mtx") ...  EigenSolver(A, M, cfg) with
SINVERT / target sigma / LU, then compare `get_eigenvalue(i)` with eigs_b200.json.
"""
import json
import os
import sys

import scipy.io

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import lsa_fw_b200 as L  # noqa: E402
from bench import TOL, build_pencil  # noqa: E402


def main() -> None:
    name = sys.argv[1] if len(sys.argv) > 1 else "cfg1"
    out = sys.argv[2] if len(sys.argv) > 2 else f"export_{name}"
    os.makedirs(out, exist_ok=True)
    pc, sigma, nev, ncv, desc = build_pencil(name)
    scipy.io.mmwrite(os.path.join(out, "A.mtx"), pc.A)
    scipy.io.mmwrite(os.path.join(out, "M.mtx"), pc.M)
    cfg = L.EigensolverConfig(num_eig=nev, atol=TOL, max_it=200, ncv=ncv)
    es = L.EigenSolver(L.iPETScMatrix(pc.A), L.iPETScMatrix(pc.M), cfg, check_hermitian=False)
    es.solver.set_st_type(L.iSTType.SINVERT)
    es.solver.set_target(sigma)
    es.solver.set_st_pc_type(L.PreconditionerType.LU)
    pairs = es.solve()
    res = es.solver.get_residuals()[: len(pairs)]
    json.dump({"workload": desc, "sigma": [sigma.real, sigma.imag], "nev": nev, "ncv": ncv, "tol": TOL,
               "eigenvalues": [[complex(l).real, complex(l).imag] for l, _ in pairs],
               "residuals": [float(r) for r in res]}, open(os.path.join(out, "eigs_b200.json"), "w"), indent=1)
    print(f"wrote {out}/A.mtx, M.mtx, eigs_b200.json ({len(pairs)} pairs)")


if __name__ == "__main__":
    main()
