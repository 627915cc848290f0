"""One traced triangular-solve sweep (per-launch device times) of a named workload.
    LSA_TRACE=1 LSA_NO_GRAPHS=1 python tools/trace_solve.py cfg2_quarter 2> trace.txt
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from lsa_fw_b200 import _lib, pencils  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "cfg2_quarter"
pc, sigma = {"cfg2_quarter": (lambda: (pencils.backward_step_2d(334, 84), 1.0j)),
             "cfg2": (lambda: (pencils.backward_step_2d(), 1.0j)),
             "cfg1": (lambda: (pencils.cylinder_wake_2d(), 0.05 + 0.74j)),
             "cav3d": (lambda: (pencils.cavity_3d(16), 0.1 + 0.3j))}[name]()
os.environ.pop("LSA_TRACE", None)
h = _lib.Handle(pc.n, 0)
flag = ((pc.A.diagonal() - sigma * pc.M.diagonal()) == 0).astype(np.uint8)
h.analyze(pc.A.indptr, pc.A.indices, pc.M.indptr, pc.M.indices, leaf_size=64, order_last=flag)
h.set_values(pc.A.data, pc.M.data)
h.factor(1.0, -sigma, _lib.LSA_C128, 1e-13)
if "--factor" in sys.argv:
    os.environ["LSA_TRACE"] = "1"
    h.factor(1.0, -sigma, _lib.LSA_C128, 1e-13)
    os.environ.pop("LSA_TRACE")
b = np.random.default_rng(0).standard_normal(pc.n).astype(complex)
for _ in range(3):
    h.solve(b)
os.environ["LSA_TRACE"] = "1"
h.solve(b)
os.environ.pop("LSA_TRACE")
h.close()
