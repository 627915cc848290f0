timeout -k 5 180 python -m pytest tests -q -m gpu -x -k "triangular or large_front or spmv" > gpurun_out/pytest_quick.log 2>&1; echo "quick rc=$?"; tail -5 gpurun_out/pytest_quick.log | cut -c1-250
timeout -k 5 600 python -m pytest tests -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "full rc=$?"; tail -8 gpurun_out/pytest_gpu.log | cut -c1-250
for mf in 1000000 48 24 12 6; do echo "STREAM_MIN_FRONTS=$mf"; LSA_STREAM_MIN_FRONTS=$mf LSA_CLUSTER_MAX_ROWS=4096 timeout -k 5 300 python tools/trace_solve.py cfg2 2>/dev/null | tail -1; done
for mf in 1000000 48 12; do echo "STREAM_MIN_FRONTS=$mf"; LSA_STREAM_MIN_FRONTS=$mf LSA_CLUSTER_MAX_ROWS=8192 timeout -k 5 300 python tools/trace_solve.py cav3d 2>/dev/null | tail -1; done
for mf in 1000000 48 12; do echo "STREAM_MIN_FRONTS=$mf"; LSA_STREAM_MIN_FRONTS=$mf LSA_CLUSTER_MAX_ROWS=4096 timeout -k 5 300 python tools/trace_solve.py cfg1 2>/dev/null | tail -1; done
LSA_TRACE=1 LSA_NO_GRAPHS=1 LSA_CLUSTER_MAX_ROWS=4096 timeout -k 5 300 python tools/trace_solve.py cfg2 > gpurun_out/trace_cfg2_stream2.txt 2>&1; grep -c TRACE gpurun_out/trace_cfg2_stream2.txt
