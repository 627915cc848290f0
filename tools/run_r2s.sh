# round 2, GPU call s (1 GPU): register-resident panel kernel + streamed SpMV: parity suite, quick bench, factor trace
mkdir -p gpurun_out
timeout -k 5 1500 python -m pytest tests -q -m gpu -x > gpurun_out/r2s_pytest.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/r2s_pytest.log | cut -c1-300
timeout -k 5 900 python bench.py --steps 3 --warmup 2 --no-extras --no-cpu-baseline > gpurun_out/r2s_bench_cfg3_quick.json 2> gpurun_out/r2s_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d = json.loads([l for l in open("gpurun_out/r2s_bench_cfg3_quick.json") if l.startswith("{")][-1])
print("value", d["value"], "e2e", d["e2e"]["value"], "roofline", d["roofline"]["frac"], "lu", d["roofline_lu"]["frac"], "ortho", d["roofline_ortho"]["frac"], "spmv", d["roofline_spmv"])
print("phases", d["phases_s_per_step"]); print("parity", {k: d["parity"][k] for k in ("resid_direct_max", "n_perturbed", "solve_resid_N", "max_multiplier")})
PY
tail -3 gpurun_out/r2s_bench.err | cut -c1-300
LSA_TRACE=1 LSA_NO_GRAPHS=1 timeout -k 5 400 python tools/trace_solve.py cfg3 --factor > gpurun_out/r2s_trace_cfg3.out 2> gpurun_out/r2s_trace_cfg3_factor_and_solve_N.txt; cat gpurun_out/r2s_trace_cfg3.out
grep "TRACE total" gpurun_out/r2s_trace_cfg3_factor_and_solve_N.txt
