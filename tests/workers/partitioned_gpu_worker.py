"""Worker of tests/test_gpu_parity.py::test_partitioned_solve_over_gpus (one process per GPU, NCCL): the partitioned
factorisation, N / H solves, SpMV and eigensolve of a 2-D and a 3-D pencil against SciPy and against the same
computation on ONE GPU (rank 0 runs it alone)."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import numpy as np
import scipy.sparse.linalg as spla
import torch
import torch.distributed as dist

import lsa_fw_b200 as L
from lsa_fw_b200 import _lib, pencils
from lsa_fw_b200.partitioned import attach_comm, make_handle

rank = int(os.environ["RANK"])
local = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
world = dist.get_world_size()
out = {}
for kind in ("2d", "3d"):
    pc = pencils.adapted_wake_2d(96, 24, re=100.0) if kind == "2d" else pencils.cavity_3d(8)
    sigma = 0.135 + 0.727j if kind == "2d" else 0.1 + 0.3j
    flag = ((pc.A.diagonal() == 0) & (pc.M.diagonal() == 0)).astype(np.uint8)
    h = make_handle(pc.n, local)
    h.analyze(pc.A.indptr, pc.A.indices, pc.M.indptr, pc.M.indices, leaf_size=64, order_last=flag)
    attach_comm(h)
    pi = h.partition_info()
    assert pi.world == world and pi.rank == rank
    h.set_values(pc.A.data, pc.M.data)
    fs = h.factor(1.0, -sigma, _lib.LSA_C128, 1e-13)
    assert fs.n_perturbed == 0
    C = (pc.A - sigma * pc.M).tocsc()
    rng = np.random.default_rng(3)
    b = rng.standard_normal(pc.n) + 1j * rng.standard_normal(pc.n)
    x = h.solve(b)
    rn = np.linalg.norm(C @ x - b) / np.linalg.norm(b)
    xh = h.solve(b, _lib.LSA_OP_H)
    rh = np.linalg.norm(C.conj().T @ xh - b) / np.linalg.norm(b)
    assert rn < 1e-12 and rh < 1e-12, (kind, rn, rh)
    for which, mat in ((_lib.LSA_MAT_A, pc.A), (_lib.LSA_MAT_M, pc.M)):
        assert np.allclose(h.spmv(which, b), mat @ b, rtol=1e-12, atol=1e-12)
    v0 = np.random.default_rng(5).standard_normal(pc.n).astype(np.complex128)
    r = h.eigs(nev=6, ncv=40, tol=1e-11, max_restarts=200, which="TARGET_MAGNITUDE", transform=_lib.LSA_ST_SINVERT, sigma=sigma, v0=v0)
    lam = h.eigenvalues(6)
    X = h.eigenvectors(6)
    res = h.residuals(6)
    assert r.nconv >= 6 and res.max() < 1e-10, (kind, r.nconv, res)
    ra = h.eigs(nev=6, ncv=40, tol=1e-11, max_restarts=200, which="TARGET_MAGNITUDE", transform=_lib.LSA_ST_SINVERT, sigma=sigma,
                adjoint=True, v0=v0)
    lam_adj = h.eigenvalues(6)
    assert ra.nconv >= 6 and h.residuals(6).max() < 1e-10
    # every rank holds the same complete results
    t = torch.from_numpy(np.concatenate([lam.view(np.float64), X[:, 0].view(np.float64)]).copy()).cuda()
    t0 = t.clone()
    dist.broadcast(t0, 0)
    assert torch.equal(t, t0), "ranks disagree on the results"
    h.close()
    if rank == 0:
        # the same on one GPU
        h1 = _lib.Handle(pc.n, local)
        h1.analyze(pc.A.indptr, pc.A.indices, pc.M.indptr, pc.M.indices, leaf_size=64, order_last=flag)
        h1.set_values(pc.A.data, pc.M.data)
        h1.factor(1.0, -sigma, _lib.LSA_C128, 1e-13)
        x1 = h1.solve(b)
        assert np.linalg.norm(x - x1) / np.linalg.norm(x1) < 1e-12
        r1 = h1.eigs(nev=6, ncv=40, tol=1e-11, max_restarts=200, which="TARGET_MAGNITUDE", transform=_lib.LSA_ST_SINVERT, sigma=sigma, v0=v0)
        lam1 = h1.eigenvalues(6)
        assert max(min(abs(l - lam1)) / abs(l) for l in lam) < 1e-8
        assert max(min(abs(np.conj(l) - lam1)) / abs(l) for l in lam_adj) < 1e-8
        h1.close()
        out[kind] = dict(n=pc.n, resid_N=rn, resid_H=rh, nconv=int(r.nconv), eig_resid=float(res.max()), top_fronts=pi.n_top_fronts,
                         replicated_rows=int(pi.n_replicated_rows), applies=int(r.n_op_applies), applies_1gpu=int(r1.n_op_applies))
# the reference-facing route
pc = pencils.adapted_wake_2d(96, 24, re=100.0)
cfg = L.EigensolverConfig(num_eig=6, atol=1e-11, max_it=200, ncv=40)
es = L.EigenSolver(L.iPETScMatrix(pc.A), L.iPETScMatrix(pc.M), cfg, check_hermitian=False)
es.solver.set_st_type(L.iSTType.SINVERT)
es.solver.set_target(0.135 + 0.727j)
es.solver.set_st_pc_type(L.PreconditionerType.LU)
es.solver.set_backend_options(device=local, partition="auto")
pairs = es.solve()
assert len(pairs) == 6 and es.solver.stats["partition_world"] == world
assert es.solver.get_residuals()[:6].max() < 1e-10
if rank == 0:
    print("PARTITIONED_OK " + json.dumps(out), flush=True)
dist.destroy_process_group()
