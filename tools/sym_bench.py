"""Symmetric L D L^T against the general LU on the same symmetric matrix (3-D Poisson + mass shift), real FP64:
factor storage, factor time, sweep time, residuals.  usage: python tools/sym_bench.py [N]"""
import json
import os
import sys
import time

import numpy as np
import scipy.sparse as sp
import torch

sys.path.insert(0, ".")
from lsa_fw_b200 import _lib  # noqa: E402

n1 = int(sys.argv[1]) if len(sys.argv) > 1 else 80
t = sp.diags([-np.ones(n1 - 1), 2.0 * np.ones(n1), -np.ones(n1 - 1)], [-1, 0, 1])
e = sp.identity(n1)
K = (sp.kron(sp.kron(t, e), e) + sp.kron(sp.kron(e, t), e) + sp.kron(sp.kron(e, e), t)).tocsr()
K.sort_indices()
n = K.shape[0]
g = np.arange(n1, dtype=np.float64)
coords = np.stack(np.meshgrid(g, g, g, indexing="ij"), axis=-1).reshape(n, 3)
b = np.random.default_rng(0).standard_normal(n).astype(complex)
out = {"n": n}
for symm in (0, 1):
    h = _lib.Handle(n)
    if symm:
        h.set_option("symmetric", 1)
    t0 = time.perf_counter()
    info = h.analyze(K.indptr, K.indices, None, None, leaf_size=64, coords=coords)
    t_sym = time.perf_counter() - t0
    h.set_values(K.data, None)
    secs = []
    for _ in range(4):
        fs = h.factor(1.0, 0.05, _lib.LSA_F64, 0.0)
        secs.append(fs.seconds)
    print("MODE", "ldlt" if symm else "lu", file=sys.stderr, flush=True)
    os.environ["LSA_TRACE"] = "1"
    h.factor(1.0, 0.05, _lib.LSA_F64, 0.0)
    os.environ.pop("LSA_TRACE")
    x = h.solve(b)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(5):
        x = h.solve(b)
    t_solve = (time.perf_counter() - t0) / 5
    C = K      # no M attached: the factored matrix is alpha A
    out["ldlt" if symm else "lu"] = dict(
        analyze_s=t_sym, nnz_factor=int(info.nnz_lu), factor_bytes=int(info.nnz_lu) * 8, factor_s=min(secs), flops=fs.flops,
        tflops=fs.flops / min(secs) / 1e12, solve_host_s=t_solve, resid=float(np.linalg.norm(C @ x - b) / np.linalg.norm(b)),
        max_front=int(info.max_front), n_perturbed=int(fs.n_perturbed), max_multiplier=fs.max_multiplier)
    h.close()
print(json.dumps(out))
