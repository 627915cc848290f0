set -x
python -m pytest tests -q -m gpu -k "wide_subspace or membrane" 2>&1 | tail -3
LSA_NO_GRAPHS=1 python tools/ncu_solve.py cfg2 2 N 2>&1 | tail -2 && \
LSA_NO_GRAPHS=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_step|k_down_off|k_up_gather|k_sweep_cluster' --launch-skip 63 --launch-count 63 -o gpurun_out/r1g_sweep_cfg2 -f python tools/ncu_solve.py cfg2 2 N > gpurun_out/ncu_sweep.log 2>&1; tail -3 gpurun_out/ncu_sweep.log
ls -la gpurun_out/*.ncu-rep
