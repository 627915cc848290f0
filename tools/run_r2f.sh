# round 2, GPU call f (1 GPU): parity suite after the GEMM epilogue / panel / dot-grid changes, bench, factor trace
mkdir -p gpurun_out
timeout -k 5 1500 python -m pytest tests -q -m gpu > gpurun_out/r2f_pytest.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/r2f_pytest.log | cut -c1-300
timeout -k 5 900 python bench.py --steps 8 --warmup 3 > gpurun_out/r2f_bench_cfg3.json 2> gpurun_out/r2f_bench_cfg3.err; echo "bench rc=$?"; tail -3 gpurun_out/r2f_bench_cfg3.err | cut -c1-300
LSA_TRACE=1 LSA_NO_GRAPHS=1 timeout -k 5 400 python tools/trace_solve.py cfg3 --factor > gpurun_out/r2f_trace_cfg3.out 2> gpurun_out/r2f_trace_cfg3_factor_and_solve_N.txt; cat gpurun_out/r2f_trace_cfg3.out
python - <<'PY'
import json
d = json.loads([l for l in open("gpurun_out/r2f_bench_cfg3.json") if l.startswith("{")][-1])
print("value", d["value"], "e2e", d["e2e"]["value"], "cold", d["e2e_cold"]["value"], "roofline", d["roofline"]["frac"], "lu", d["roofline_lu"]["frac"], "ortho", d["roofline_ortho"]["frac"], "lu3d", d.get("roofline_lu_3d", {}).get("frac"))
print("phases", d["phases_s_per_step"])
PY
grep "TRACE total" gpurun_out/r2f_trace_cfg3_factor_and_solve_N.txt
