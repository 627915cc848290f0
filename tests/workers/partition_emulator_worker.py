"""Worker of tests/test_host_logic.py::test_partitioned_solve_gloo: one rank of the partitioned multifrontal solve,
emulated in NumPy on the rank-local symbolic structures, collectives over torch.distributed (gloo, CPU)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch, torch.distributed as dist
from lsa_fw_b200 import _lib, pencils
from mf_emulator import PartEmulator
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
class Comm:
    def bcast(self, a, root):
        t = torch.from_numpy(a.view(np.float64)); dist.broadcast(t, root)
    def allreduce(self, a):
        t = torch.from_numpy(a.view(np.float64)); dist.all_reduce(t)
for kind in ("2d", "3d"):
    pc = pencils.assemble_pencil((24, 12), (8.0, 3.0), re=50.0, baseflow=pencils.wake_profile(0.9, 1.2, 1.5)) if kind == "2d" else pencils.cavity_3d(5)
    sigma = 0.1 + 0.6j
    h = _lib.Handle(pc.n, -1, rank, world)
    flag = ((pc.A.diagonal() == 0) & (pc.M.diagonal() == 0)).astype(np.uint8)
    h.analyze(pc.A.indptr, pc.A.indices, pc.M.indptr, pc.M.indices, leaf_size=24, order_last=flag)
    em = PartEmulator(h, pc.n, Comm())
    em.factor(pc.A.data, pc.M.data, 1.0, -sigma)
    C = (pc.A - sigma * pc.M).tocsc()
    b = np.random.default_rng(0).standard_normal(pc.n) + 1j * np.random.default_rng(1).standard_normal(pc.n)
    rn = np.linalg.norm(C @ em.solve(b) - b) / np.linalg.norm(b)
    rh = np.linalg.norm(C.conj().T @ em.solve(b, "H") - b) / np.linalg.norm(b)
    pi = h.partition_info()
    print(f"rank {rank}/{world} {kind}: n={pc.n} local fronts {pi.n_fronts_local}/{pi.n_fronts_global} top {pi.n_top_fronts} resid N {rn:.1e} H {rh:.1e}", flush=True)
    assert rn < 1e-11 and rh < 1e-11
dist.destroy_process_group()
