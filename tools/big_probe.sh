mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_fused.py -x -q 2>&1 | tail -2
timeout 100 python tools/fused_probe.py c2 3 r,d 2>&1 | grep "^c2\|rel"
timeout 400 python tools/fused_probe.py c3 1 d > gpurun_out/r02_probe_c3.txt 2>&1; tail -2 gpurun_out/r02_probe_c3.txt
timeout 900 python tools/fused_probe.py t30 1 r > gpurun_out/r02_probe_t30.txt 2>&1; tail -2 gpurun_out/r02_probe_t30.txt
