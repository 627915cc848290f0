# 2 GPUs: partitioned sweeps from CUDA graphs (NCCL captured), whole-block inverses up to 8192 pivots (3-D tops)
mkdir -p gpurun_out
timeout -k 5 900 python -m pytest tests -q -m gpu -x -k "partitioned or large_front" > gpurun_out/r2k_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r2k_pytest.log | cut -c1-500
timeout -k 5 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29751 tools/trace_partitioned.py 20 > gpurun_out/r2k_trace_partitioned_n20_2gpus.out 2> gpurun_out/r2k_trace_partitioned_n20_2gpus.err; echo rc=$?
grep "^rank" gpurun_out/r2k_trace_partitioned_n20_2gpus.out | cut -c1-400; tail -3 gpurun_out/r2k_trace_partitioned_n20_2gpus.err | cut -c1-300
LSA_PARTITION_GRAPHS=0 timeout -k 5 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29752 tools/trace_partitioned.py 20 2>/dev/null | grep "rep 1" | cut -c1-400
timeout -k 5 600 python bench.py --workload cav3d --steps 3 --warmup 2 --no-extras --no-cpu-baseline > gpurun_out/r2k_bench_cav3d.json 2> gpurun_out/r2k_bench_cav3d.err; echo "cav3d rc=$?"
python - <<'PY'
import json
d = json.loads([l for l in open("gpurun_out/r2k_bench_cav3d.json") if l.startswith("{")][-1])
print("cav3d value", d["value"], "phases", d["phases_s_per_step"], "sweep frac", d["roofline"]["frac"], "lu frac", d["roofline_lu"]["frac"], "resid", d["parity"]["solve_resid_N"], d["parity"]["solve_resid_H"])
PY
