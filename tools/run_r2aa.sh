# round 2, GPU call aa (1 GPU): fused small-front kernel: tests, factor trace with and without
mkdir -p gpurun_out
timeout -k 5 900 python -m pytest tests -q -m gpu -x -k "small_front or kat or triangular or real_factor or zero_pivot or growth or mini_config3 or perturb or spmv" > gpurun_out/r2aa_pytest.log 2>&1; echo "pytest rc=$?"; tail -12 gpurun_out/r2aa_pytest.log | cut -c1-300
LSA_TRACE=1 LSA_NO_GRAPHS=1 timeout -k 5 500 python tools/trace_solve.py cfg3 --factor > gpurun_out/r2aa_trace_cfg3.out 2> gpurun_out/r2aa_trace_cfg3_factor_and_solve_N.txt; cat gpurun_out/r2aa_trace_cfg3.out
grep "TRACE total" gpurun_out/r2aa_trace_cfg3_factor_and_solve_N.txt
grep "front_small\|inv_" gpurun_out/r2aa_trace_cfg3_factor_and_solve_N.txt | head -30
python tools/summarize_trace.py gpurun_out/r2aa_trace_cfg3_factor_and_solve_N.txt 2>/dev/null | head -1
