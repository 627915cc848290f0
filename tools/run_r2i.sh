# leaf-size experiment on config 3 (nested-dissection leaf size 32 / 48 / 96 vs the default 64) + validation of the masked-GEMM K ranges
mkdir -p gpurun_out
timeout -k 5 600 python -m pytest tests -q -m gpu -x -k "large_front or dense_schur or triangular or eigenpairs_match" 2>&1 | tail -3
for leaf in 64 32 48 96; do
  LSA_BENCH_LEAF=$leaf timeout -k 5 600 python bench.py --steps 4 --warmup 2 --no-extras --no-cpu-baseline > gpurun_out/r2i_bench_leaf$leaf.json 2> gpurun_out/r2i_bench_leaf$leaf.err; echo "leaf $leaf rc=$?"
  python - <<PY
import json
d = json.loads([l for l in open("gpurun_out/r2i_bench_leaf$leaf.json") if l.startswith("{")][-1])
print("leaf $leaf value", round(d["value"],4), "e2e", round(d["e2e"]["value"],4), "phases", {k: round(v,4) for k,v in d["phases_s_per_step"].items()}, "sweep frac", round(d["roofline"]["frac"],3), "bytes", d["roofline"]["algorithmic_bytes_per_launch"], "nnz_lu", d["symbolic"]["nnz_lu"], "fronts", d["symbolic"]["fronts"])
PY
done
