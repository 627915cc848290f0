"""Where the per-solve upload time goes: lsa_set_values from page-locked and from pageable host arrays (692 MB of FP64
values, the config-3 size), plus the host helpers a cached solve calls.  usage: python tools/h2d_probe.py"""
import sys
import time

import numpy as np
import scipy.sparse as sp

sys.path.insert(0, ".")
from lsa_fw_b200 import _lib  # noqa: E402

n, bands = 3_000_000, 29
offs = [o * 37 for o in range(-(bands // 2), bands // 2 + 1)]
A = sp.diags([np.ones(n - abs(o)) for o in offs], offs, format="csr")
A.sort_indices()
print("n", n, "nnz", A.nnz, "bytes", A.nnz * 8)
h = _lib.Handle(n)
t0 = time.perf_counter()
h.analyze(A.indptr, A.indices, None, None, leaf_size=64)
print("analyze %.2f s" % (time.perf_counter() - t0))
lib = _lib.load()
pinned = _lib.pinned_empty(A.data.shape, np.float64)
import ctypes as C
p = C.c_void_p()
rc = lib.lsa_host_alloc(C.c_uint64(A.nnz * 8), C.byref(p))
print("lsa_host_alloc rc", rc, "ptr", hex(p.value or 0))
import torch
print("pinned_empty is page-locked according to torch:", torch.from_numpy(pinned).is_pinned())
pinned[...] = A.data
pageable = A.data.copy()
for name, arr in (("pinned", pinned), ("pageable", pageable), ("pinned", pinned), ("pageable", pageable)):
    t0 = time.perf_counter()
    h.set_values(arr, None)
    dt = time.perf_counter() - t0
    print("set_values from %-8s %.1f ms  (%.1f GB/s)" % (name, dt * 1e3, A.nnz * 8 / dt / 1e9))
d = torch.empty(A.nnz, dtype=torch.float64, device="cuda")
for name, arr in (("pinned", pinned), ("pageable", pageable)):
    src = torch.from_numpy(arr)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    d.copy_(src, non_blocking=True)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    print("torch copy from %-8s %.1f ms  (%.1f GB/s)" % (name, dt * 1e3, A.nnz * 8 / dt / 1e9))
rows = np.arange(0, n, 7, dtype=np.int32)
for _ in range(3):
    t0 = time.perf_counter()
    _lib.diag_is_zero(A, rows)
    print("diag_is_zero(%d rows) %.2f ms" % (len(rows), (time.perf_counter() - t0) * 1e3))
# several page-locked arrays, as the Reynolds sweep of bench.py holds them (one per pair)
import subprocess
arrs = []
for q in range(9):
    a = _lib.pinned_empty(A.data.shape, np.float64)
    a[...] = A.data
    arrs.append(a)
print("pinned fallbacks", _lib.pinned_fallbacks)
for rep in range(2):
    line = []
    for a in arrs:
        t0 = time.perf_counter()
        h.set_values(a, None)
        line.append("%.1f" % ((time.perf_counter() - t0) * 1e3))
    print("set_values ms from 9 page-locked arrays:", " ".join(line))
big = [np.random.default_rng(0).standard_normal(50_000_000) for _ in range(40)]     # 16 GB of ordinary host memory in use
line = []
for a in arrs:
    t0 = time.perf_counter()
    h.set_values(a, None)
    line.append("%.1f" % ((time.perf_counter() - t0) * 1e3))
print("with 16 GB of pageable arrays alive:", " ".join(line))
for cmd in (["nvidia-smi", "topo", "-m"], ["numactl", "-H"], ["grep", "-E", "MemTotal|MemFree|MemAvailable", "/proc/meminfo"], ["nproc"]):
    try:
        print(" ".join(cmd), "->", subprocess.run(cmd, capture_output=True, text=True, timeout=20).stdout[:1500])
    except Exception as ex:
        print(cmd, "failed", ex)
