export LSA_CLUSTER_MAX_ROWS=4096
timeout -k 5 180 env LSA_STREAM_FLAGS=7 python -m pytest tests -q -m gpu -x -k "triangular or large_front or spmv or eigenpairs" > gpurun_out/pytest_quick.log 2>&1; echo "quick(flags7) rc=$?"; tail -3 gpurun_out/pytest_quick.log | cut -c1-250
for fl in 3 7 5; do echo "FLAGS=$fl"; LSA_STREAM_FLAGS=$fl timeout -k 5 300 python tools/trace_solve.py cfg2 2>/dev/null | tail -1; done
for mf in 200 48 24 12; do echo "FLAGS=7 MIN_FRONTS=$mf"; LSA_STREAM_FLAGS=7 LSA_STREAM_MIN_FRONTS=$mf timeout -k 5 300 python tools/trace_solve.py cfg2 2>/dev/null | tail -1; done
for sg in 2 4; do echo "FLAGS=7 STAGES=$sg"; LSA_STREAM_FLAGS=7 LSA_STREAM_STAGES=$sg timeout -k 5 300 python tools/trace_solve.py cfg2 2>/dev/null | tail -1; done
echo cfg1; LSA_STREAM_FLAGS=7 timeout -k 5 300 python tools/trace_solve.py cfg1 2>/dev/null | tail -1
echo cav3d; LSA_STREAM_FLAGS=7 LSA_CLUSTER_MAX_ROWS=8192 timeout -k 5 300 python tools/trace_solve.py cav3d 2>/dev/null | tail -1
echo cav3d mf48; LSA_STREAM_FLAGS=7 LSA_STREAM_MIN_FRONTS=48 LSA_CLUSTER_MAX_ROWS=8192 timeout -k 5 300 python tools/trace_solve.py cav3d 2>/dev/null | tail -1
LSA_STREAM_FLAGS=7 LSA_TRACE=1 LSA_NO_GRAPHS=1 timeout -k 5 300 python tools/trace_solve.py cfg2 > gpurun_out/trace_cfg2_stream4.txt 2>&1; grep -c TRACE gpurun_out/trace_cfg2_stream4.txt
LSA_STREAM_FLAGS=7 LSA_TRACE=1 LSA_NO_GRAPHS=1 LSA_STREAM_MIN_FRONTS=12 timeout -k 5 300 python tools/trace_solve.py cfg2 > gpurun_out/trace_cfg2_stream4_mf12.txt 2>&1; grep -c TRACE gpurun_out/trace_cfg2_stream4_mf12.txt
