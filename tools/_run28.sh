timeout -k 5 900 python -m pytest tests -q -m gpu -x > gpurun_out/pytest_gpu.log 2>&1; echo "full rc=$?"; tail -6 gpurun_out/pytest_gpu.log | cut -c1-250
for sr in 0 128 192 320; do echo "SMALL_ROWS=$sr"; LSA_STREAM_SMALL_ROWS=$sr timeout -k 5 300 python tools/trace_solve.py cfg2 2>/dev/null | tail -1; done
echo cfg1; timeout -k 5 300 python tools/trace_solve.py cfg1 2>/dev/null | tail -1
echo cav3d; timeout -k 5 300 python tools/trace_solve.py cav3d 2>/dev/null | tail -1
LSA_TRACE=1 LSA_NO_GRAPHS=1 timeout -k 5 300 python tools/trace_solve.py cfg2 > gpurun_out/trace_cfg2_default.txt 2>&1; grep -c TRACE gpurun_out/trace_cfg2_default.txt
timeout -k 5 600 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_quick.json 2> gpurun_out/bench_quick.err; tail -c 400 gpurun_out/bench_quick.err; python -c "
import json; d=json.loads(open('gpurun_out/bench_quick.json').read().strip().splitlines()[-1]); print(d['value'], d['e2e']['value'], d['roofline']['achieved'], d['roofline']['frac'], d['parity']['resid_direct_max'], d['parity']['resid_adjoint_max'], d['parity']['solve_resid_N'], d['parity']['solve_resid_H'])"
