# round 2, GPU call d (4 GPUs): partitioned solve at world 2 and 4, strong-scaling bench, a 3-D case that does not fit one GPU
mkdir -p gpurun_out
nvidia-smi -L | head -8
timeout -k 5 900 python -m pytest tests -q -m gpu -x -k "partitioned" > gpurun_out/r2d_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r2d_pytest.log | cut -c1-300
timeout -k 5 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29721 bench.py --gpus 4 --steps 8 --warmup 2 > gpurun_out/r2d_bench_n4_cfg3.json 2> gpurun_out/r2d_bench_n4.err; echo "bench n4 rc=$?"; tail -3 gpurun_out/r2d_bench_n4.err | cut -c1-300
timeout -k 5 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29722 tools/partitioned_big.py 40 > gpurun_out/r2d_partitioned_n40_4gpus.json 2> gpurun_out/r2d_partitioned_n40.err; echo "big rc=$?"; tail -3 gpurun_out/r2d_partitioned_n40.err | cut -c1-400; cat gpurun_out/r2d_partitioned_n40_4gpus.json | cut -c1-1500
python - <<'PY'
import json
try:
    d = json.loads([l for l in open("gpurun_out/r2d_bench_n4_cfg3.json") if l.startswith("{")][-1])
    print("value", d["value"], "e2e", d["e2e"]["value"], "partitioned:", json.dumps(d.get("partitioned"))[:1800])
except Exception as e:
    print("no bench line", e)
PY
