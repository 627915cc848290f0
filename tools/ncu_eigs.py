"""Factor a named workload and run a short Krylov-Schur: the target of `ncu --set full -k regex:k_update_dots|k_dots|k_update`
captures of the Gram-Schmidt kernels.
    python tools/ncu_eigs.py cfg3_quarter
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from lsa_fw_b200 import _lib, pencils  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "cfg3_quarter"
pc, sigma = {"cfg3": (lambda: (pencils.adapted_wake_2d(re=100.0), 0.135 + 0.727j)),
             "cfg3_quarter": (lambda: (pencils.adapted_wake_2d(578, 145, re=100.0), 0.135 + 0.727j)),
             "cfg1": (lambda: (pencils.cylinder_wake_2d(), 0.05 + 0.74j))}[name]()
h = _lib.Handle(pc.n, 0)
flag = ((pc.A.diagonal() == 0) & (pc.M.diagonal() == 0)).astype(np.uint8)
h.analyze(pc.A.indptr, pc.A.indices, pc.M.indptr, pc.M.indices, leaf_size=64, order_last=flag)
h.set_values(pc.A.data, pc.M.data)
h.factor(1.0, -sigma, _lib.LSA_C128, 1e-13)
r = h.eigs(nev=10, ncv=80, tol=1e-11, max_restarts=1, which="TARGET_MAGNITUDE", transform=_lib.LSA_ST_SINVERT, sigma=sigma, seed=1)
print(f"{name}: n = {pc.n}, {r.n_op_applies} applies, ortho {r.seconds_ortho * 1e3:.1f} ms, {r.n_reorth} second passes", flush=True)
h.close()
