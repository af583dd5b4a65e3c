mkdir -p gpurun_out
timeout 600 ncu --set full --import-source on --clock-control none -k regex:qgt_fused_lean -s 9 -c 1 -f -o gpurun_out/r02_lean_c2_main python tools/fused_probe.py c2 1 l > gpurun_out/ncu_lean.log 2>&1
ls -la gpurun_out | tail -3
