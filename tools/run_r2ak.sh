# round 2, GPU call ak (1 GPU, the last 2.7 GPU-minutes): tests 29..58 of the GPU parity suite on the final library
# (tests 1..28 passed in call aj on the same device code; the 29th, the SpMV test on a hub graph, had overflowed the stack
# of the host analysis there -- fixed and pinned by a CPU test since).
mkdir -p gpurun_out
timeout -k 5 128 python -m pytest -q -x -p no:cacheprovider --durations=8 $(cat tools/r2ak_tests.txt | tr '\n' ' ') > gpurun_out/r2ak_pytest_gpu_29_58.log 2>&1; echo "pytest rc=$?"; tail -14 gpurun_out/r2ak_pytest_gpu_29_58.log | cut -c1-200
