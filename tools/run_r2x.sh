# round 2, GPU call x (1 GPU): post-factor (block inverse) trace on config 3
mkdir -p gpurun_out
LSA_TRACE=1 LSA_NO_GRAPHS=1 timeout -k 5 500 python tools/trace_solve.py cfg3 --factor > gpurun_out/r2x_trace_cfg3.out 2> gpurun_out/r2x_trace_cfg3_factor_and_solve_N.txt; cat gpurun_out/r2x_trace_cfg3.out
grep "TRACE total" gpurun_out/r2x_trace_cfg3_factor_and_solve_N.txt
grep "inv_" gpurun_out/r2x_trace_cfg3_factor_and_solve_N.txt
