mkdir -p gpurun_out
timeout 600 ncu --set full --import-source on --clock-control none -k regex:qgt_fused_pipe -s 9 -c 1 -f -o gpurun_out/r02_pipe_c2_main python tools/fused_probe.py c2 1 p > gpurun_out/ncu_pipe.log 2>&1
ls -la gpurun_out | tail -3
