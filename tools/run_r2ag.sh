# round 2, GPU call ag (1 GPU): final code -- full parity suite, smoke, default bench
mkdir -p gpurun_out
timeout -k 5 1500 python -m pytest tests -q -m gpu > gpurun_out/r2ag_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r2ag_pytest_gpu.log | cut -c1-300
timeout -k 5 300 python __graft_entry__.py --smoke 2>&1 | tail -2
timeout -k 5 900 python bench.py --steps 8 --warmup 3 > gpurun_out/r2ag_bench_cfg3.json 2> gpurun_out/r2ag_bench_cfg3.err; echo "bench rc=$?"; tail -3 gpurun_out/r2ag_bench_cfg3.err | cut -c1-300
python - <<'PY'
import json
d = json.loads([l for l in open("gpurun_out/r2ag_bench_cfg3.json") if l.startswith("{")][-1])
print("value", d["value"], "e2e", d["e2e"]["value"], "cold", d["e2e_cold"]["value"], "roofline", d["roofline"]["frac"], "lu", d["roofline_lu"]["frac"], "ortho", d["roofline_ortho"]["frac"], "spmv", d["roofline_spmv"].get("frac"), "lu3d", d.get("roofline_lu_3d", {}).get("frac"))
print("phases", d["phases_s_per_step"])
PY
