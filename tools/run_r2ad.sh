# round 2, GPU call ad (1 GPU): ncu --set full of the panel-chain kernels at a mid level (config 3, level 10: 1 025 fronts)
mkdir -p gpurun_out
LSA_NO_GRAPHS=1 timeout -k 5 400 ncu --set full --clock-control none --import-source on -k regex:'k_swap_trsm|k_trsm_cols|k_extend_add|k_panel_lu' --launch-skip 61 --launch-count 15 -o gpurun_out/r2ad_panel_chain -f python tools/ncu_solve.py cfg3 1 N > gpurun_out/r2ad_ncu.log 2>&1; echo "ncu rc=$?"; tail -3 gpurun_out/r2ad_ncu.log
ls -la gpurun_out/r2ad_panel_chain.ncu-rep
