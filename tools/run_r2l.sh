# symmetric L D L^T mode (row f4): full GPU suite + LDL^T vs LU on a 3-D Poisson matrix
mkdir -p gpurun_out
timeout -k 5 1200 python -m pytest tests -q -m gpu -x > gpurun_out/r2l_pytest.log 2>&1; echo "pytest rc=$?"; tail -12 gpurun_out/r2l_pytest.log | cut -c1-600
timeout -k 5 600 python tools/sym_bench.py 80 > gpurun_out/r2l_sym_bench_n80.json 2> gpurun_out/r2l_sym_bench.err; echo "sym rc=$?"; cat gpurun_out/r2l_sym_bench_n80.json; python - <<'PY'
import collections
mode, acc = None, {}
for line in open("gpurun_out/r2l_sym_bench.err"):
    w = line.split()
    if line.startswith("MODE"):
        mode = w[1]; acc[mode] = collections.Counter()
    elif line.startswith("TRACE level") and mode:
        acc[mode][w[3]] += float(w[-2])
for m, c in acc.items():
    print(m, {k: round(v / 1e3, 2) for k, v in c.most_common()}, "ms")
PY
