python tools/_diag_wide.py > gpurun_out/diag_wide.log 2>&1; tail -8 gpurun_out/diag_wide.log
python -m pytest tests -q -m gpu -x > gpurun_out/pytest_gpu.log 2>&1; tail -30 gpurun_out/pytest_gpu.log | cut -c1-250
for w in 16 8; do for mr in 1280 4096; do echo "WIDTH=$w MAX_ROWS=$mr"; LSA_CLUSTER_MAX_WIDTH=$w LSA_CLUSTER_MAX_ROWS=$mr python tools/trace_solve.py cfg2 2> gpurun_out/trace_cfg2_w${w}_r${mr}.txt | tail -1; done; done
for w in 16 8; do for mr in 1280 8192; do echo "WIDTH=$w MAX_ROWS=$mr"; LSA_CLUSTER_MAX_WIDTH=$w LSA_CLUSTER_MAX_ROWS=$mr python tools/trace_solve.py cav3d 2> gpurun_out/trace_cav3d_w${w}_r${mr}.txt | tail -1; done; done
LSA_CLUSTER_MAX_WIDTH=16 LSA_CLUSTER_MAX_ROWS=4096 python tools/trace_solve.py cfg1 2>/dev/null | tail -1
LSA_CLUSTER_MAX_ROWS=4096 LSA_NO_GRAPHS=1 python tools/ncu_solve.py cfg2 2 N 2>&1 | tail -1 && \
LSA_CLUSTER_MAX_ROWS=4096 LSA_NO_GRAPHS=1 timeout 900 ncu --set full --clock-control none -k regex:'k_step|k_down_off|k_up_gather|k_sweep_cluster' --launch-skip 63 --launch-count 63 -o /tmp/r1g_sweep_cfg2 -f python tools/ncu_solve.py cfg2 2 N > gpurun_out/ncu_sweep.log 2>&1
ncu -i /tmp/r1g_sweep_cfg2.ncu-rep --page raw --csv > gpurun_out/r1g_ncu_sweep_cfg2_raw.csv 2>/dev/null; ls -la /tmp/r1g_sweep_cfg2.ncu-rep gpurun_out/
