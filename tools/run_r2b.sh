# round 2, GPU call b (2 GPUs): partitioned solve on hardware, strong-scaling bench path, fixed tests
mkdir -p gpurun_out
nvidia-smi -L
timeout -k 5 900 python -m pytest tests -q -m gpu -x -k "partitioned or backward_step or purification or adjoint_reuse or multiplier" > gpurun_out/r2b_pytest.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/r2b_pytest.log | cut -c1-400
LSA_BENCH_PART3D_N=20 timeout -k 5 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29711 bench.py --gpus 2 --steps 4 --warmup 1 --workload cfg3_ref > gpurun_out/r2b_bench_n2_cfg3ref.json 2> gpurun_out/r2b_bench_n2.err; echo "bench n2 rc=$?"; tail -5 gpurun_out/r2b_bench_n2.err | cut -c1-400
python - <<'PY'
import json
try:
    d = json.loads([l for l in open("gpurun_out/r2b_bench_n2_cfg3ref.json") if l.startswith("{")][-1])
    print("value", d["value"], "e2e", d["e2e"]["value"], "partitioned:", json.dumps(d.get("partitioned"))[:1500])
except Exception as e:
    print("no bench line", e)
PY
