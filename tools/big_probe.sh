mkdir -p gpurun_out
timeout 600 python tools/fused_probe.py c3 1 g,d > gpurun_out/r02_probe_c3.txt 2>&1; tail -3 gpurun_out/r02_probe_c3.txt | cut -c1-400
timeout 900 python tools/fused_probe.py t30 1 r > gpurun_out/r02_probe_t30.txt 2>&1; tail -1 gpurun_out/r02_probe_t30.txt | cut -c1-400
