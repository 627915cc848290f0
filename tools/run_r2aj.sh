# round 2, GPU call aj (1 GPU, the last 4 GPU-minutes of the round): the GPU parity suite on the library with the reworked
# host analysis (ordering bit-identical by the CPU test; this run covers the device uploads of the re-typed arrays), then,
# if time is left, the host analysis of config 3 timed on the GPU box's cores.
mkdir -p gpurun_out
timeout -k 5 185 python -m pytest tests -q -m gpu -x -p no:cacheprovider > gpurun_out/r2aj_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2aj_pytest_gpu.log | cut -c1-300
LSA_TRACE_ANALYZE=1 timeout -k 5 45 python - > gpurun_out/r2aj_analyze_cfg3.txt 2>&1 <<'PY'
import time, os
import bench
from lsa_fw_b200 import _lib
t0 = time.perf_counter()
pc, w = bench.build_pencil("cfg3")
print("assemble %.2f s, host cores %s" % (time.perf_counter() - t0, os.cpu_count()), flush=True)
h = _lib.Handle(pc.n, device=-1)
t0 = time.perf_counter()
info = h.analyze(pc.A.indptr, pc.A.indices, pc.M.indptr, pc.M.indices, leaf_size=bench.LEAF, order_last=bench.order_last_flags(pc))
print("analyze %.3f s (host only), nnz_lu %d, fronts %d" % (time.perf_counter() - t0, info.nnz_lu, info.n_fronts), flush=True)
PY
echo "analyze rc=$?"; tail -12 gpurun_out/r2aj_analyze_cfg3.txt
