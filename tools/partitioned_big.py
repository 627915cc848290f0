"""One factorisation + eigensolve of a 3-D cavity pencil that does NOT fit one GPU, split over the GPUs of the node
(lsa_fw_b200/partitioned.py).  Launch with torchrun, one process per GPU:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 tools/partitioned_big.py 40

Prints one JSON line on rank 0: per-rank memory, factor / eigensolve seconds (max over ranks), residuals.
"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import torch.distributed as dist

from lsa_fw_b200 import _lib, pencils
from lsa_fw_b200.partitioned import attach_comm, make_handle

n_cells = int(sys.argv[1]) if len(sys.argv) > 1 else 40
rank, local = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
device = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=device)
world = dist.get_world_size()
t0 = time.perf_counter()
pc = pencils.cavity_3d(n_cells)
t_asm = time.perf_counter() - t0
sigma, nev, ncv = 0.1 + 0.3j, 10, 40
flags = ((pc.A.diagonal() == 0) & (pc.M.diagonal() == 0)).astype(np.uint8)
h = make_handle(pc.n, local)
t0 = time.perf_counter()
info = h.analyze(pc.A.indptr, pc.A.indices, pc.M.indptr, pc.M.indices, leaf_size=64, order_last=flags)
t_sym = time.perf_counter() - t0
attach_comm(h)
pi = h.partition_info()
need = (info.factor_entries + sum(info.pool_entries) + pi.cut_pool_entries) * 16
free, total = torch.cuda.mem_get_info(device)
ok = torch.tensor([1.0 if need < 0.92 * free else 0.0], device=device)
dist.all_reduce(ok, op=dist.ReduceOp.MIN)
out = {"workload": f"3D lid-driven-cavity surrogate, Taylor-Hood {n_cells}^3 x 6 tets ({pc.n} DOFs), complex shift, {world} GPUs",
       "nnz_A": int(pc.A.nnz), "factor_bytes_global": int(pi.nnz_lu_global) * 16,
       "need_bytes_this_rank": int(need), "free_bytes_this_rank": int(free)}
if ok.item() < 0.5:
    out["skipped"] = "a rank would not fit its part (factor store + pools + cut pool) into free device memory"
else:
    h.set_values(pc.A.data, pc.M.data)
    dist.barrier()
    t0 = time.perf_counter()
    fs = h.factor(1.0, -sigma, _lib.LSA_C128, 1e-13)
    t_f = time.perf_counter() - t0
    b = np.random.default_rng(1).standard_normal(pc.n) + 0j
    t0 = time.perf_counter()
    x = h.solve(b)
    t_s = time.perf_counter() - t0
    resid = float(np.linalg.norm((pc.A - sigma * pc.M) @ x - b) / np.linalg.norm(b))
    v0 = np.random.default_rng(2).standard_normal(pc.n).astype(np.complex128)
    t0 = time.perf_counter()
    r = h.eigs(nev=nev, ncv=ncv, tol=1e-10, max_restarts=50, which="TARGET_MAGNITUDE", transform=_lib.LSA_ST_SINVERT, sigma=sigma, v0=v0)
    t_e = time.perf_counter() - t0
    lam = h.eigenvalues(min(nev, r.nconv))
    res = h.residuals(min(nev, r.nconv))
    t = torch.tensor([t_f, t_e, fs.seconds, r.seconds_solve / max(1, r.n_op_applies), float(need)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    tt = [float(v) for v in t.tolist()]
    out.update(factor_s=tt[0], eigs_s=tt[1], factor_device_s=tt[2], sweep_s_per_apply=tt[3], max_need_bytes_per_gpu=tt[4],
               factor_tflops_aggregate=float(pi.flops_real_global) * 4 / tt[2] / 1e12, solve_resid=resid, n_perturbed=int(fs.n_perturbed),
               nconv=int(r.nconv), op_applies=int(r.n_op_applies), eig_resid_max=float(res.max()) if len(res) else None,
               eigenvalues=[[z.real, z.imag] for z in lam[:4]], top_fronts=pi.n_top_fronts, cut_roots=pi.n_cut_roots,
               replicated_rows=int(pi.n_replicated_rows), assemble_s=t_asm, symbolic_s=t_sym, host_solve_roundtrip_s=t_s,
               model_speedup_bound=pi.weight_total / max(1e-300, pi.weight_top + pi.weight_max_subtrees))
if rank == 0:
    print(json.dumps(out), flush=True)
h.close()
dist.destroy_process_group()
