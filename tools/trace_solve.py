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
             "cfg3": (lambda: (pencils.adapted_wake_2d(re=100.0), 0.135 + 0.727j)),
             "cfg3_quarter": (lambda: (pencils.adapted_wake_2d(578, 145, re=100.0), 0.135 + 0.727j)),
             "cav3d": (lambda: (pencils.cavity_3d(16), 0.1 + 0.3j)),
             "cav3d24": (lambda: (pencils.cavity_3d(24), 0.1 + 0.3j))}[name]()
os.environ.pop("LSA_TRACE", None)
h = _lib.Handle(pc.n, 0)
flag = ((pc.A.diagonal() == 0) & (pc.M.diagonal() == 0)).astype(np.uint8)
h.analyze(pc.A.indptr, pc.A.indices, pc.M.indptr, pc.M.indices, leaf_size=64, order_last=flag)
h.set_values(pc.A.data, pc.M.data)
h.factor(1.0, -sigma, _lib.LSA_C128, 1e-13)
if "--factor" in sys.argv:
    os.environ["LSA_TRACE"] = "1"
    h.factor(1.0, -sigma, _lib.LSA_C128, 1e-13)
    os.environ.pop("LSA_TRACE")
b = np.random.default_rng(0).standard_normal(pc.n).astype(complex)
import time
for _ in range(3):
    h.solve(b)
t0 = time.perf_counter()
for _ in range(20):
    h.solve(b)
print(f"{name}: host-roundtrip solve {(time.perf_counter() - t0) / 20 * 1e3:.3f} ms (incl. 2 x {pc.n * 16 / 1e6:.0f} MB PCIe copies)", flush=True)
r = h.eigs(nev=4, ncv=24, tol=1e-8, max_restarts=2, which="TARGET_MAGNITUDE", transform=_lib.LSA_ST_SINVERT, sigma=sigma, seed=1)
print(f"{name}: device sweep {r.seconds_solve / r.n_op_applies * 1e3:.3f} ms/apply over {r.n_op_applies} applies "
      f"(INVERT_MAX_K={os.environ.get('LSA_INVERT_MAX_K')}, NO_GRAPHS={os.environ.get('LSA_NO_GRAPHS')})", flush=True)
os.environ["LSA_TRACE"] = "1"
h.solve(b, _lib.LSA_OP_H if "--H" in sys.argv else _lib.LSA_OP_N)
os.environ.pop("LSA_TRACE")
if "--spmv-variants" in sys.argv:
    # streamed SpMV: entries per row block (device time per M v from the eigensolver's event timer)
    bytes_m = pc.M.nnz * 12 + 8 * (pc.n + 1) + 32 * pc.n
    for blk in (512, 1024, 2048):
        h.set_option("spmv_block", blk)
        r = h.eigs(nev=4, ncv=24, tol=1e-8, max_restarts=2, which="TARGET_MAGNITUDE", transform=_lib.LSA_ST_SINVERT, sigma=sigma, seed=1)
        t = r.seconds_spmv / r.n_op_applies
        print(f"{name}: spmv_block {blk}: {t * 1e6:.1f} us per M v, {bytes_m / t / 1e9:.0f} GB/s algorithmic", flush=True)
h.close()
