mkdir -p gpurun_out
timeout -k 5 600 python tools/h2d_probe.py > gpurun_out/r2p_h2d_probe.txt 2>&1; echo "probe rc=$?"; cat gpurun_out/r2p_h2d_probe.txt | cut -c1-400
