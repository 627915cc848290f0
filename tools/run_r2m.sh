# symmetric mode after the lower-triangle Schur + deferred scaling; upload probe
mkdir -p gpurun_out
timeout -k 5 600 python -m pytest tests -q -m gpu -x -k "symmetric or cholesky" > gpurun_out/r2m_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r2m_pytest.log | cut -c1-600
timeout -k 5 600 python tools/sym_bench.py 80 > gpurun_out/r2m_sym_bench_n80.json 2> gpurun_out/r2m_sym_bench.err; echo "sym rc=$?"; cat gpurun_out/r2m_sym_bench_n80.json
python - <<'PY'
import collections
mode, acc = None, {}
for line in open("gpurun_out/r2m_sym_bench.err"):
    w = line.split()
    if line.startswith("MODE"):
        mode = w[1]; acc[mode] = collections.Counter()
    elif line.startswith("TRACE level") and mode:
        acc[mode][w[3]] += float(w[-2])
for m, c in acc.items():
    print(m, {k: round(v / 1e3, 2) for k, v in c.most_common()}, "ms")
PY
timeout -k 5 600 python tools/h2d_probe.py > gpurun_out/r2m_h2d_probe.txt 2>&1; echo "probe rc=$?"; cat gpurun_out/r2m_h2d_probe.txt | cut -c1-300
