# round 2, GPU call z (1 GPU): full parity suite, default bench (GPU arm + reference arm), ncu evidence for the final kernels
mkdir -p gpurun_out
timeout -k 5 1500 python -m pytest tests -q -m gpu > gpurun_out/r2z_pytest.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/r2z_pytest.log | cut -c1-300
timeout -k 5 900 python bench.py --steps 8 --warmup 3 > gpurun_out/r2z_bench_cfg3.json 2> gpurun_out/r2z_bench_cfg3.err; echo "bench rc=$?"; tail -3 gpurun_out/r2z_bench_cfg3.err | cut -c1-300
timeout -k 5 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2z_bench_reference_arm.json 2> gpurun_out/r2z_bench_reference_arm.err; echo "ref arm rc=$?"
LSA_NO_GRAPHS=1 timeout -k 5 500 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:'k_front_stream|k_tri_gemv|k_up_off|k_down_off|k_up_gather|k_solve_decoupled|k_level_unpermute|k_sweep|k_step' --csv --log-file gpurun_out/r2z_dram_solve_cfg3.csv python tools/ncu_solve.py cfg3 2 N > gpurun_out/r2z_ncu_dram.log 2>&1; echo "ncu dram rc=$?"
timeout -k 5 500 ncu --set full --clock-control none --import-source on -k regex:'k_spmv' --launch-skip 20 --launch-count 3 -o gpurun_out/r2z_spmv -f python tools/ncu_eigs.py cfg3 > gpurun_out/r2z_ncu_spmv.log 2>&1; echo "ncu spmv rc=$?"
timeout -k 5 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 14000 --csv --log-file gpurun_out/r2z_launches_bench_cfg3.csv python bench.py --steps 1 --warmup 1 --no-extras --no-cpu-baseline > gpurun_out/r2z_ncu_launches.log 2>&1; echo "ncu launches rc=$?"
gzip -f gpurun_out/r2z_dram_solve_cfg3.csv gpurun_out/r2z_launches_bench_cfg3.csv
python - <<'PY'
import json
d = json.loads([l for l in open("gpurun_out/r2z_bench_cfg3.json") if l.startswith("{")][-1])
print("value", d["value"], "e2e", d["e2e"]["value"], "cold", d["e2e_cold"]["value"], "roofline", d["roofline"]["frac"], "lu", d["roofline_lu"]["frac"], "ortho", d["roofline_ortho"]["frac"], "spmv", d["roofline_spmv"].get("achieved"), "lu3d", d.get("roofline_lu_3d", {}).get("frac"))
print("phases", d["phases_s_per_step"])
r = json.loads([l for l in open("gpurun_out/r2z_bench_reference_arm.json") if l.startswith("{")][-1])
print("reference arm value", r["value"], r["cpu_baseline"]["sample"][:400])
PY
ls -la gpurun_out/*.ncu-rep
