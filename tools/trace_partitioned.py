"""Per-rank timers and one traced sweep of the partitioned solve (torchrun, one process per GPU):
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 tools/trace_partitioned.py 20
"""
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

n_cells = int(sys.argv[1]) if len(sys.argv) > 1 else 20
rank, local = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
world = dist.get_world_size()
pc = pencils.cavity_3d(n_cells)
sigma = 0.1 + 0.3j
flags = ((pc.A.diagonal() == 0) & (pc.M.diagonal() == 0)).astype(np.uint8)
h = make_handle(pc.n, local)
info = h.analyze(pc.A.indptr, pc.A.indices, pc.M.indptr, pc.M.indices, leaf_size=64, order_last=flags)
attach_comm(h)
pi = h.partition_info()
h.set_values(pc.A.data, pc.M.data)
h.factor(1.0, -sigma, _lib.LSA_C128, 1e-13)
fs = h.factor(1.0, -sigma, _lib.LSA_C128, 1e-13)
v0 = np.random.default_rng(2).standard_normal(pc.n).astype(np.complex128)
for rep in range(2):
    dist.barrier()
    t0 = time.perf_counter()
    r = h.eigs(nev=10, ncv=80, tol=1e-11, max_restarts=100, which="TARGET_MAGNITUDE", transform=_lib.LSA_ST_SINVERT, sigma=sigma, v0=v0)
    wall = time.perf_counter() - t0
    msg = (f"rank {rank}/{world} rep {rep}: own rows {pi.n_own_rows} replicated {pi.n_replicated_rows} local fronts {pi.n_fronts_local} "
           f"nnz_lu local {info.nnz_lu:.3g}: factor {fs.seconds:.3f} s | eigs wall {wall:.3f} device {r.seconds:.3f}: solve {r.seconds_solve:.3f} "
           f"spmv {r.seconds_spmv:.3f} ortho {r.seconds_ortho:.3f} rr {r.seconds_rr:.3f} restart {r.seconds_restart:.3f} applies {r.n_op_applies} kernels {r.n_kernels}")
    for q in range(world):
        dist.barrier()
        if q == rank:
            print(msg, flush=True)
b = np.random.default_rng(1).standard_normal(pc.n) + 0j
for _ in range(2):
    h.solve(b)
dist.barrier()
os.environ["LSA_TRACE"] = "1"
h.solve(b)
os.environ.pop("LSA_TRACE")
h.close()
dist.destroy_process_group()
