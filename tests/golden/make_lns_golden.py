"""Writes tests/golden/lns_eigs.json: eigenvalues of the small linearised Navier-Stokes test pencils next to their shifts,
as the CPU oracle (`oracle/eigen_oracle.py`: SuperLU + Krylov-Schur restatement of `Solver/eigen2.py`) computes them, with
the condition number kappa = ||x|| ||y|| / |y^H M x| of each eigenvalue from the oracle's right and left vectors.

The reference holds no golden values for Navier-Stokes pencils ("parity unpinned" there); these
fixtures pin the ORACLE, so that (a) a change of SciPy / LAPACK that moves the checker is noticed on CPU
(`tests/test_oracle_golden.py::test_oracle_reproduces_committed_lns_eigenvalues`) and (b) the GPU parity tests compare the
CUDA path with committed numbers as well as with the live oracle.

    python tests/golden/make_lns_golden.py
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

NEV, NCV, TOL = 8, 40, 1e-12


def pencil(kind):
    from lsa_fw_b200 import pencils

    wake = pencils.wake_profile(0.9, 1.2, 1.5)
    if kind == "th2d":
        return pencils.assemble_pencil((24, 12), (8.0, 3.0), re=50.0, baseflow=wake), 0.05 + 0.6j
    if kind == "mini2d":
        return pencils.assemble_pencil((20, 12), (8.0, 3.0), re=50.0, baseflow=wake, space="MINI"), 0.05 + 0.6j
    if kind == "th3d":
        return pencils.cavity_3d(6), 0.1 + 0.3j
    if kind == "th2d_real":
        return pencils.assemble_pencil((24, 12), (8.0, 3.0), re=50.0, baseflow=wake), -0.3
    raise ValueError(kind)


KINDS = ("th2d", "mini2d", "th3d", "th2d_real")


def compute(kind) -> dict:
    from oracle import eigen_oracle as O

    pc, sigma = pencil(kind)
    orc = O.shift_invert_krylov_schur(pc.A, pc.M, sigma, NEV, ncv=NCV, tol=TOL)
    adj = O.shift_invert_krylov_schur(pc.A, pc.M, sigma, NEV, ncv=NCV, tol=TOL, adjoint=True)
    lam, kappa = [], []
    for j, l in enumerate(orc.eigenvalues[:NEV]):     # in `which` order: increasing |lambda - sigma|
        x = orc.eigenvectors[:, j]
        # left vector of the same eigenvalue; inside a cluster of equal eigenvalues (the Dirichlet rows' lambda = 1) the one
        # that pairs with x
        near = np.nonzero(abs(np.conj(adj.eigenvalues) - l) <= 1e-6 * abs(l) + abs(np.conj(adj.eigenvalues) - l).min())[0]
        ja = int(max(near, key=lambda q: abs(np.vdot(adj.eigenvectors[:, q], pc.M @ x))))
        y = adj.eigenvectors[:, ja]
        lam.append([float(np.real(l)), float(np.imag(l))])
        if abs(np.conj(adj.eigenvalues[ja]) - l) > 1e-6 * abs(l):
            kappa.append(None)                              # the adjoint run did not return this eigenvalue
        else:
            kappa.append(float(np.linalg.norm(x) * np.linalg.norm(y) / max(abs(np.vdot(y, pc.M @ x)), 1e-300)))
    s = complex(sigma)
    return {"n": int(pc.n), "sigma": [s.real, s.imag], "nev": NEV, "ncv": NCV, "tol": TOL, "eigenvalues": lam, "kappa": kappa}


if __name__ == "__main__":
    out = {k: compute(k) for k in KINDS}
    json.dump(out, open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "lns_eigs.json"), "w"), indent=1)
    for k, v in out.items():
        print(k, v["n"], ["%.6f%+.6fj" % tuple(z) for z in v["eigenvalues"][:3]], "kappa max %.2e" % max(k for k in v["kappa"] if k is not None))
