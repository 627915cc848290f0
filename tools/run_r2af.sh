# round 2, GPU call af (1 GPU): panel kernel with 32 / 64 / 128 threads by panel height
mkdir -p gpurun_out
timeout -k 5 900 python -m pytest tests -q -m gpu -x -k "kat or triangular or spmv or real_factor or symmetric or ghep or zero_pivot or growth or mini_config3 or perturb or partitioned" > gpurun_out/r2af_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r2af_pytest.log | cut -c1-300
LSA_TRACE=1 LSA_NO_GRAPHS=1 timeout -k 5 500 python tools/trace_solve.py cfg3 --factor > gpurun_out/r2af_trace_cfg3.out 2> gpurun_out/r2af_trace_cfg3_factor_and_solve_N.txt; cat gpurun_out/r2af_trace_cfg3.out
grep "TRACE total" gpurun_out/r2af_trace_cfg3_factor_and_solve_N.txt
python tools/summarize_trace.py gpurun_out/r2af_trace_cfg3_factor_and_solve_N.txt 2>/dev/null | head -1
