# round 2, GPU call h (8 GPUs): partitioned solve at world 8 (GPUs without a sub-tree), bench launch path at N = 8
mkdir -p gpurun_out
nvidia-smi -L | wc -l; free -g | head -2
timeout -k 5 600 python -m pytest tests -q -m gpu -x -k "partitioned and 8" > gpurun_out/r2h_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r2h_pytest.log | cut -c1-600
LSA_BENCH_PART3D_N=20 timeout -k 5 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29731 bench.py --gpus 8 --steps 8 --warmup 1 --workload cfg3_ref > gpurun_out/r2h_bench_n8_cfg3ref.json 2> gpurun_out/r2h_bench_n8.err; echo "bench n8 rc=$?"; tail -3 gpurun_out/r2h_bench_n8.err | cut -c1-400
python - <<'PY'
import json
try:
    d = json.loads([l for l in open("gpurun_out/r2h_bench_n8_cfg3ref.json") if l.startswith("{")][-1])
    print("value", d["value"], "e2e", d["e2e"]["value"], "partitioned:", json.dumps(d.get("partitioned"))[:1500])
except Exception as e:
    print("no bench line", e)
PY
free -g | head -2
