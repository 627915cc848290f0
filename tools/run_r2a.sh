# round 2, GPU call a: parity suite, first config-3 bench line, per-launch traces, ncu launch list + DRAM traffic
mkdir -p gpurun_out
timeout -k 5 900 python -m pytest tests -q -m gpu -x > gpurun_out/r2a_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r2a_pytest.log | cut -c1-300
timeout -k 5 900 python bench.py --steps 8 --warmup 3 > gpurun_out/r2a_bench_cfg3.json 2> gpurun_out/r2a_bench_cfg3.err; echo "bench rc=$?"; tail -3 gpurun_out/r2a_bench_cfg3.err | cut -c1-300
LSA_TRACE=1 LSA_NO_GRAPHS=1 timeout -k 5 400 python tools/trace_solve.py cfg3 --factor > gpurun_out/r2a_trace_cfg3.out 2> gpurun_out/r2a_trace_cfg3_factor_N.txt; echo "trace rc=$?"; cat gpurun_out/r2a_trace_cfg3.out
LSA_TRACE=1 LSA_NO_GRAPHS=1 LSA_INVERT_MAX_K=0 timeout -k 5 400 python tools/trace_solve.py cfg3 > gpurun_out/r2a_trace_cfg3_noinv.out 2> gpurun_out/r2a_trace_cfg3_noinv_N.txt; cat gpurun_out/r2a_trace_cfg3_noinv.out
LSA_NO_GRAPHS=1 timeout -k 5 500 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:'k_front_stream|k_tri_gemv|k_up_off|k_down_off|k_up_gather|k_solve_decoupled|k_level_unpermute|k_sweep|k_step' --csv --log-file gpurun_out/r2a_dram_solve_cfg3.csv python tools/ncu_solve.py cfg3 2 N > gpurun_out/r2a_ncu_dram.log 2>&1; echo "ncu dram rc=$?"; tail -2 gpurun_out/r2a_ncu_dram.log
timeout -k 5 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 30000 --csv --log-file gpurun_out/r2a_launches_bench_cfg3.csv python bench.py --steps 1 --warmup 1 --no-extras --no-cpu-baseline > gpurun_out/r2a_ncu_bench.log 2>&1; echo "ncu bench rc=$?"
gzip -f gpurun_out/r2a_launches_bench_cfg3.csv gpurun_out/r2a_dram_solve_cfg3.csv
ls -la gpurun_out | tail -20
