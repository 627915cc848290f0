timeout -k 5 300 python -m pytest tests -q -m gpu -x -k "gram_schmidt or eigenpairs or adjoint or kat or membrane or wide" > gpurun_out/pytest_quick.log 2>&1; echo "quick rc=$?"; tail -5 gpurun_out/pytest_quick.log | cut -c1-250
timeout -k 5 600 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_quick.json 2> gpurun_out/bench_quick.err; tail -c 300 gpurun_out/bench_quick.err; python -c "
import json; d=json.loads(open('gpurun_out/bench_quick.json').read().strip().splitlines()[-1]); print(d['value'], d['e2e']['value'], d['roofline']['frac'], d['phases_s_per_step'], d['op_applies_per_step'], d.get('reorthogonalised_columns_per_step'), d['parity']['resid_direct_max'], d['parity']['resid_adjoint_max'], d['parity']['nconv_direct'], d['parity']['nconv_adjoint'])"
LSA_TRACE=1 LSA_NO_GRAPHS=1 timeout -k 5 300 python tools/trace_solve.py cfg2 --H > gpurun_out/trace_cfg2_H.txt 2>&1; grep -c TRACE gpurun_out/trace_cfg2_H.txt
export LSA_NO_GRAPHS=1
python tools/ncu_solve.py cfg2 2 N 2>&1 | tail -1 && \
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_sweep_cluster' --launch-skip 14 --launch-count 3 -o gpurun_out/r1h_cluster_cfg2 -f python tools/ncu_solve.py cfg2 2 N > gpurun_out/ncu_cluster.log 2>&1
tail -2 gpurun_out/ncu_cluster.log; ls -la gpurun_out/*.ncu-rep
