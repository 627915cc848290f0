"""Small driver for ncu launch lists: one factorisation + one (short) eigensolve of a named workload.

    python tools/profile_phases.py cfg2_quarter [nev] [ncv]
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from lsa_fw_b200 import _lib, pencils  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "cfg2_quarter"
nev = int(sys.argv[2]) if len(sys.argv) > 2 else 4
ncv = int(sys.argv[3]) if len(sys.argv) > 3 else 16
if name == "cfg2_quarter":
    pc, sigma = pencils.backward_step_2d(334, 84), 1.0j
elif name == "cfg2":
    pc, sigma = pencils.backward_step_2d(), 1.0j
elif name == "cfg1":
    pc, sigma = pencils.cylinder_wake_2d(), 0.05 + 0.74j
elif name == "cav3d":
    pc, sigma = pencils.cavity_3d(16), 0.1 + 0.3j
else:
    raise SystemExit("unknown workload")
h = _lib.Handle(pc.n, 0)
flag = ((pc.A.diagonal() - sigma * pc.M.diagonal()) == 0).astype(np.uint8)
info = h.analyze(pc.A.indptr, pc.A.indices, pc.M.indptr, pc.M.indices, leaf_size=64, order_last=flag)
h.set_values(pc.A.data, pc.M.data)
fs = h.factor(1.0, -sigma, _lib.LSA_C128, 1e-13)
r = h.eigs(nev=nev, ncv=ncv, tol=1e-8, max_restarts=2, which="TARGET_MAGNITUDE", transform=_lib.LSA_ST_SINVERT,
           sigma=sigma, seed=1)
c = h.counters()
print(f"{name}: n={pc.n} fronts={info.n_fronts} levels={info.n_levels} nnz_lu={info.nnz_lu} "
      f"factor={fs.seconds * 1e3:.2f} ms ({fs.flops / fs.seconds / 1e12:.2f} TFLOP/s, {fs.n_kernels} kernels) "
      f"solve/apply={r.seconds_solve / r.n_op_applies * 1e3:.3f} ms "
      f"({c.bytes_solve / (r.seconds_solve / r.n_op_applies) / 1e9:.0f} GB/s) applies={r.n_op_applies} nconv={r.nconv}")
h.close()
