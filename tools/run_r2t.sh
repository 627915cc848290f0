# round 2, GPU call t (1 GPU): rolled register panel kernel + composed-permutation swap kernel: parity suite, factor trace
mkdir -p gpurun_out
timeout -k 5 1500 python -m pytest tests -q -m gpu -x > gpurun_out/r2t_pytest.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/r2t_pytest.log | cut -c1-300
LSA_TRACE=1 LSA_NO_GRAPHS=1 timeout -k 5 400 python tools/trace_solve.py cfg3 --factor > gpurun_out/r2t_trace_cfg3.out 2> gpurun_out/r2t_trace_cfg3_factor_and_solve_N.txt; cat gpurun_out/r2t_trace_cfg3.out
grep "TRACE total" gpurun_out/r2t_trace_cfg3_factor_and_solve_N.txt
python tools/summarize_trace.py gpurun_out/r2t_trace_cfg3_factor_and_solve_N.txt | head -1
