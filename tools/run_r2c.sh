# round 2, GPU call c (1 GPU): full parity suite, config-3 bench line with all extras, per-launch trace
mkdir -p gpurun_out
timeout -k 5 1200 python -m pytest tests -q -m gpu -x > gpurun_out/r2c_pytest.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/r2c_pytest.log | cut -c1-300
timeout -k 5 900 python bench.py --steps 8 --warmup 3 > gpurun_out/r2c_bench_cfg3.json 2> gpurun_out/r2c_bench_cfg3.err; echo "bench rc=$?"; tail -3 gpurun_out/r2c_bench_cfg3.err | cut -c1-300
LSA_TRACE=1 LSA_NO_GRAPHS=1 timeout -k 5 400 python tools/trace_solve.py cfg3 > gpurun_out/r2c_trace_cfg3.out 2> gpurun_out/r2c_trace_cfg3_solve_N.txt; cat gpurun_out/r2c_trace_cfg3.out
python - <<'PY'
import json
d = json.loads([l for l in open("gpurun_out/r2c_bench_cfg3.json") if l.startswith("{")][-1])
print("value", d["value"], "e2e", d["e2e"]["value"], "cold", d["e2e_cold"]["value"], "roofline", d["roofline"]["frac"], "lu", d["roofline_lu"]["frac"], "ortho", d["roofline_ortho"]["frac"])
print("phases", d["phases_s_per_step"]); print("e2e phases", d["e2e_phases"])
PY
