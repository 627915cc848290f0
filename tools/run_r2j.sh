# per-rank timers and one traced sweep of the partitioned solve on 4 GPUs
mkdir -p gpurun_out
timeout -k 5 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29741 tools/trace_partitioned.py 20 > gpurun_out/r2j_trace_partitioned_n20_4gpus.out 2> gpurun_out/r2j_trace_partitioned_n20_4gpus.err; echo rc=$?
grep "^rank" gpurun_out/r2j_trace_partitioned_n20_4gpus.out | cut -c1-400
grep -c TRACE gpurun_out/r2j_trace_partitioned_n20_4gpus.err
