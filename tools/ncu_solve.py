"""Factor a named workload once and run a few un-graphed triangular-solve sweeps: the target of the
`ncu --set full -k regex:...` captures of the sweep kernels (profiles/*_ncu_sweep_*).
    LSA_NO_GRAPHS=1 python tools/ncu_solve.py cfg2 [n_solves] [N|H]
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from lsa_fw_b200 import _lib, pencils  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "cfg2"
nsolve = int(sys.argv[2]) if len(sys.argv) > 2 else 2
trans = sys.argv[3] if len(sys.argv) > 3 else "N"
pc, sigma = {"cfg2_quarter": (lambda: (pencils.backward_step_2d(334, 84), -0.35 + 0.1j)),
             "cfg2": (lambda: (pencils.backward_step_2d(), -0.35 + 0.1j)),
             "cfg1": (lambda: (pencils.cylinder_wake_2d(), 0.05 + 0.74j)),
             "cfg3": (lambda: (pencils.adapted_wake_2d(re=100.0), 0.135 + 0.727j)),
             "cfg3_quarter": (lambda: (pencils.adapted_wake_2d(578, 145, re=100.0), 0.135 + 0.727j)),
             "cav3d": (lambda: (pencils.cavity_3d(16), 0.1 + 0.3j)),
             "cav3d24": (lambda: (pencils.cavity_3d(24), 0.1 + 0.3j))}[name]()
h = _lib.Handle(pc.n, 0)
flag = ((pc.A.diagonal() == 0) & (pc.M.diagonal() == 0)).astype(np.uint8)
h.analyze(pc.A.indptr, pc.A.indices, pc.M.indptr, pc.M.indices, leaf_size=64, order_last=flag)
h.set_values(pc.A.data, pc.M.data)
h.factor(1.0, -sigma, _lib.LSA_C128, 1e-13)
b = np.random.default_rng(0).standard_normal(pc.n).astype(complex)
for _ in range(nsolve):
    x = h.solve(b, trans={"N": _lib.LSA_OP_N, "H": _lib.LSA_OP_H}[trans])
F = (pc.A - sigma * pc.M).tocsr()
r = (F @ x - b) if trans == "N" else (F.conj().T @ x - b)
print(f"{name}: n = {pc.n}, {nsolve} solves ({trans}), residual {np.linalg.norm(r) / np.linalg.norm(b):.2e}", flush=True)
h.close()
