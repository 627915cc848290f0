# round 2, GPU call ah (2 GPUs): final code -- partitioned-solve tests, bench line at N = 2 (config-3 sweep dealt over the
# ranks + ONE 3-D cavity solve split over the two GPUs)
mkdir -p gpurun_out
timeout -k 5 600 python -m pytest tests -q -m gpu -x -k "partitioned" > gpurun_out/r2ah_pytest_2gpus.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2ah_pytest_2gpus.log | cut -c1-300
timeout -k 5 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29711 bench.py --gpus 2 --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/r2ah_bench_n2_cfg3.json 2> gpurun_out/r2ah_bench_n2.err; echo "bench n2 rc=$?"; tail -3 gpurun_out/r2ah_bench_n2.err | cut -c1-300
python - <<'PY'
import json
try:
    d = json.loads([l for l in open("gpurun_out/r2ah_bench_n2_cfg3.json") if l.startswith("{")][-1])
    print("value", d["value"], "n_gpus", d["n_gpus"], "scaling", d["scaling"], "e2e", d["e2e"]["value"], "partitioned:", json.dumps(d.get("partitioned"))[:1200])
except Exception as e:
    print("no bench line", e)
PY
