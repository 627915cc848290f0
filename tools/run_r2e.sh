# round 2, GPU call e (1 GPU): full parity suite on the final code, bench lines (GPU arm + reference arm), ncu evidence
mkdir -p gpurun_out
timeout -k 5 1500 python -m pytest tests -q -m gpu > gpurun_out/r2e_pytest.log 2>&1; echo "pytest rc=$?"; tail -8 gpurun_out/r2e_pytest.log | cut -c1-300
timeout -k 5 900 python bench.py --steps 8 --warmup 3 > gpurun_out/r2e_bench_cfg3.json 2> gpurun_out/r2e_bench_cfg3.err; echo "bench rc=$?"; tail -3 gpurun_out/r2e_bench_cfg3.err | cut -c1-300
timeout -k 5 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2e_bench_reference_arm.json 2> gpurun_out/r2e_bench_reference_arm.err; echo "ref arm rc=$?"
LSA_NO_GRAPHS=1 timeout -k 5 500 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:'k_front_stream|k_tri_gemv|k_up_off|k_down_off|k_up_gather|k_solve_decoupled|k_level_unpermute|k_sweep|k_step' --csv --log-file gpurun_out/r2e_dram_solve_cfg3.csv python tools/ncu_solve.py cfg3 2 N > gpurun_out/r2e_ncu_dram.log 2>&1; echo "ncu dram rc=$?"
LSA_NO_GRAPHS=1 timeout -k 5 500 ncu --set full --clock-control none --import-source on -k regex:'k_tri_gemv|k_up_off|k_down_off' --launch-skip 26 --launch-count 10 -o gpurun_out/r2e_top_levels -f python tools/ncu_solve.py cfg3 2 N > gpurun_out/r2e_ncu_top.log 2>&1; echo "ncu top rc=$?"
LSA_NO_GRAPHS=1 timeout -k 5 500 ncu --set full --clock-control none --import-source on -k regex:'k_front_stream' --launch-skip 24 --launch-count 6 -o gpurun_out/r2e_stream -f python tools/ncu_solve.py cfg3 2 N > gpurun_out/r2e_ncu_stream.log 2>&1; echo "ncu stream rc=$?"
timeout -k 5 500 ncu --set full --clock-control none --import-source on -k regex:'k_update_dots|k_dots|k_update' --launch-skip 150 --launch-count 9 -o gpurun_out/r2e_ortho -f python tools/ncu_eigs.py cfg3_quarter > gpurun_out/r2e_ncu_ortho.log 2>&1; echo "ncu ortho rc=$?"
gzip -f gpurun_out/r2e_dram_solve_cfg3.csv
python - <<'PY'
import json
d = json.loads([l for l in open("gpurun_out/r2e_bench_cfg3.json") if l.startswith("{")][-1])
print("value", d["value"], "e2e", d["e2e"]["value"], "cold", d["e2e_cold"]["value"], "roofline", d["roofline"]["frac"], "lu", d["roofline_lu"]["frac"], "ortho", d["roofline_ortho"]["frac"])
print("e2e phases", d["e2e_phases"])
r = json.loads([l for l in open("gpurun_out/r2e_bench_reference_arm.json") if l.startswith("{")][-1])
print("reference arm value", r["value"], r["cpu_baseline"]["sample"][:600])
PY
ls -la gpurun_out/*.ncu-rep
