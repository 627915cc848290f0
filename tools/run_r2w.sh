# round 2, GPU call w (1 GPU): sweep variants on config 3 -- span-split triangular GEMVs, stream_min_fronts
mkdir -p gpurun_out
timeout -k 5 600 python -m pytest tests -q -m gpu -x -k "triangular or sweep or option or real_factor or kat" > gpurun_out/r2w_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2w_pytest.log | cut -c1-300
for v in "512 96" "0 96" "256 96" "512 192" "512 300"; do
  set -- $v
  echo "== LSA_TRI_SPAN=$1 LSA_STREAM_MIN_FRONTS=$2"
  LSA_TRI_SPAN=$1 LSA_STREAM_MIN_FRONTS=$2 timeout -k 5 400 python tools/trace_solve.py cfg3 2> gpurun_out/r2w_trace_cfg3_solve_span$1_min$2.txt | grep "device sweep"
done
