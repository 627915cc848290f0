for mr in 4096 2560 1280 640 0; do echo "LSA_CLUSTER_MAX_ROWS=$mr"; LSA_CLUSTER_MAX_ROWS=$mr python tools/trace_solve.py cfg2 2>/dev/null | tail -1; done
for mr in 4096 1280 0; do echo "LSA_CLUSTER_MAX_ROWS=$mr"; LSA_CLUSTER_MAX_ROWS=$mr python tools/trace_solve.py cfg1 2>/dev/null | tail -1; done
for mr in 8192 4096 1280 0; do echo "LSA_CLUSTER_MAX_ROWS=$mr"; LSA_CLUSTER_MAX_ROWS=$mr python tools/trace_solve.py cav3d 2>/dev/null | tail -1; done
