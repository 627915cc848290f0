# Gram-Schmidt grid experiment: waves of resident blocks per dot kernel
mkdir -p gpurun_out
for w in 1 2 3; do
  LSA_GS_WAVES=$w timeout -k 5 600 python bench.py --steps 4 --warmup 2 --no-extras --no-cpu-baseline > gpurun_out/r2g_bench_w$w.json 2> gpurun_out/r2g_bench_w$w.err; echo "waves $w rc=$?"
  python - <<PY
import json
d = json.loads([l for l in open("gpurun_out/r2g_bench_w$w.json") if l.startswith("{")][-1])
print("waves $w value", d["value"], "ortho s/step", d["phases_s_per_step"]["ortho"], "frac", d["roofline_ortho"]["frac"])
PY
done
timeout -k 5 300 python -m pytest tests -q -m gpu -k "nonfinite or gram_schmidt" 2>&1 | tail -3
