# round 2, GPU call y (2 GPUs): partitioned solve tests after the panel / swap / SpMV kernel changes
mkdir -p gpurun_out
nvidia-smi -L
timeout -k 5 900 python -m pytest tests -q -m gpu -x -k "partitioned" > gpurun_out/r2y_pytest_2gpus.log 2>&1; echo "pytest rc=$?"; tail -8 gpurun_out/r2y_pytest_2gpus.log | cut -c1-400
