export LSA_CLUSTER_MAX_ROWS=4096 LSA_STREAM_MIN_FRONTS=48 LSA_NO_GRAPHS=1
python tools/ncu_solve.py cfg2 2 N 2>&1 | tail -1 && \
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_front_stream' --launch-skip 26 --launch-count 4 -o gpurun_out/r1h_stream_cfg2 -f python tools/ncu_solve.py cfg2 2 N > gpurun_out/ncu_stream.log 2>&1
tail -2 gpurun_out/ncu_stream.log; ls -la gpurun_out/*.ncu-rep
