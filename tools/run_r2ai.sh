# round 2, GPU call ai (1 GPU): ncu launch list of ONE COMPLETE step of the bench command (graphs off: ncu cannot profile
# kernels inside captured graphs).  k_extend_add is left out of the capture by name (r2ac: ncu fails on its 40th launch
# while saving/restoring the factor store around the in-place kernel, which ended that list after the first
# factorisation); its share is known from the CUDA-event trace (profiles/r2ae_trace_*: 27.7 ms per factorisation).
mkdir -p gpurun_out
LSA_NO_GRAPHS=1 timeout -k 5 450 ncu --metrics gpu__time_duration.sum --clock-control none \
  --kernel-name 'regex:k_([a-df-z]|e[a-wyz])|Kernel2' -c 9000 --csv \
  --log-file gpurun_out/r2ai_launches_bench_cfg3_step.csv \
  python bench.py --steps 1 --warmup 0 --no-extras --no-cpu-baseline > gpurun_out/r2ai_ncu_launches.log 2>&1; echo "ncu launches rc=$?"
tail -3 gpurun_out/r2ai_ncu_launches.log | cut -c1-400
wc -l gpurun_out/r2ai_launches_bench_cfg3_step.csv
gzip -f gpurun_out/r2ai_launches_bench_cfg3_step.csv
