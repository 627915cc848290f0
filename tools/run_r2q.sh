mkdir -p gpurun_out
timeout -k 5 900 python bench.py --steps 3 --warmup 2 --no-extras --no-cpu-baseline > gpurun_out/r2q_bench_cfg3_quick.json 2> gpurun_out/r2q_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d = json.loads([l for l in open("gpurun_out/r2q_bench_cfg3_quick.json") if l.startswith("{")][-1])
print("value", d["value"], "e2e", d["e2e"], "uploads", d.get("e2e_upload_s_by_step"), "not pinned", d.get("e2e_inputs_not_page_locked"), d.get("e2e_host_timing"))
print("e2e phases", d["e2e_phases"]); print("phases", d["phases_s_per_step"])
PY
tail -3 gpurun_out/r2q_bench.err | cut -c1-300
timeout -k 5 600 python tools/h2d_probe.py 2>&1 | grep "set_values from\|page-locked arrays" | head -6
