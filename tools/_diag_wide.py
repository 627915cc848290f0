import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import lsa_fw_b200 as L
from lsa_fw_b200 import pencils
pm = pencils.membrane_pencil(24, 24, 1.0, 0.83)
for nev, ncv in ((50, 100), (60, 126), (60, 130), (60, 160), (100, 200), (100, 256)):
    cfg = L.EigensolverConfig(num_eig=nev, problem_type=L.iEpsProblemType.GHEP, atol=1e-11, max_it=300, ncv=ncv)
    es = L.EigenSolver(L.iPETScMatrix(pm.A), L.iPETScMatrix(pm.M), cfg, check_hermitian=False)
    es.solver.set_st_type(L.iSTType.SINVERT)
    es.solver.set_target(1000.0)
    pairs = es.solve()
    st = es.solver.stats
    lam = np.array([v for v, _ in pairs])
    d = np.abs(lam - 1000.0)
    print(f"nev {nev} ncv {ncv}: nconv {len(pairs)} restarts {st['n_restarts']} applies {st['n_op_applies']} breakdown {st['breakdown']} "
          f"ordered {bool(np.all(np.diff(d) >= -1e-9))} dmax {d.max() if len(d) else None:.3f} rr_s {st['rr_seconds']:.3f} eigs_s {st['eigs_seconds']:.3f} "
          f"resid {es.solver.get_residuals().max():.2e}", flush=True)
