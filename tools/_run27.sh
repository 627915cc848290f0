export LSA_CLUSTER_MAX_ROWS=4096
timeout -k 5 180 python -m pytest tests -q -m gpu -x -k "triangular or large_front or spmv or eigenpairs" > gpurun_out/pytest_quick.log 2>&1; echo "quick rc=$?"; tail -3 gpurun_out/pytest_quick.log | cut -c1-250
for mf in 1000000 200 100 48 24 12; do echo "MIN_FRONTS=$mf"; LSA_STREAM_MIN_FRONTS=$mf timeout -k 5 300 python tools/trace_solve.py cfg2 2>/dev/null | tail -1; done
for sg in 3 4 6; do echo "STAGES=$sg"; LSA_STREAM_STAGES=$sg timeout -k 5 300 python tools/trace_solve.py cfg2 2>/dev/null | tail -1; done
echo cfg1; timeout -k 5 300 python tools/trace_solve.py cfg1 2>/dev/null | tail -1
echo cav3d; LSA_CLUSTER_MAX_ROWS=8192 timeout -k 5 300 python tools/trace_solve.py cav3d 2>/dev/null | tail -1
echo cav3d mf48; LSA_STREAM_MIN_FRONTS=48 LSA_CLUSTER_MAX_ROWS=8192 timeout -k 5 300 python tools/trace_solve.py cav3d 2>/dev/null | tail -1
LSA_TRACE=1 LSA_NO_GRAPHS=1 LSA_STREAM_MIN_FRONTS=12 timeout -k 5 300 python tools/trace_solve.py cfg2 > gpurun_out/trace_cfg2_stream5_mf12.txt 2>&1; grep -c TRACE gpurun_out/trace_cfg2_stream5_mf12.txt
timeout -k 5 600 python -m pytest tests -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "full rc=$?"; tail -4 gpurun_out/pytest_gpu.log | cut -c1-250
