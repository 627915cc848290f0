# round 2, GPU call ac (1 GPU): final code -- full parity suite, default bench, bounded ncu launch list (graphs off: ncu
# cannot profile kernels inside captured graphs)
mkdir -p gpurun_out
timeout -k 5 1500 python -m pytest tests -q -m gpu > gpurun_out/r2ac_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/r2ac_pytest_gpu.log | cut -c1-300
timeout -k 5 900 python bench.py --steps 8 --warmup 3 > gpurun_out/r2ac_bench_cfg3.json 2> gpurun_out/r2ac_bench_cfg3.err; echo "bench rc=$?"; tail -3 gpurun_out/r2ac_bench_cfg3.err | cut -c1-300
LSA_NO_GRAPHS=1 timeout -k 5 240 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/r2ac_launches_bench_cfg3.csv python bench.py --steps 1 --warmup 0 --no-extras --no-cpu-baseline > gpurun_out/r2ac_ncu_launches.log 2>&1; echo "ncu launches rc=$?"
gzip -f gpurun_out/r2ac_launches_bench_cfg3.csv
python - <<'PY'
import json
d = json.loads([l for l in open("gpurun_out/r2ac_bench_cfg3.json") if l.startswith("{")][-1])
print("value", d["value"], "e2e", d["e2e"]["value"], "cold", d["e2e_cold"]["value"], "roofline", d["roofline"]["frac"], "lu", d["roofline_lu"]["frac"], "ortho", d["roofline_ortho"]["frac"], "spmv", d["roofline_spmv"].get("frac"), "lu3d", d.get("roofline_lu_3d", {}).get("frac"))
print("phases", d["phases_s_per_step"])
PY
