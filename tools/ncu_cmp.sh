mkdir -p gpurun_out
M=smsp__inst_executed.sum,sm__inst_executed_pipe_tensor_subpipe_dmma.sum,gpu__time_duration.sum,sm__warps_active.avg.pct_of_peak_sustained_active,l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum,smsp__inst_executed_pipe_fp64.sum,launch__occupancy_limit_shared_mem,launch__occupancy_limit_registers,launch__grid_size,sm__pipe_tensor_subpipe_dmma_cycles_active.avg.pct_of_peak_sustained_active
timeout 300 ncu --metrics $M --clock-control none -k regex:"qgt_sweep|qgt_fused" --csv --log-file gpurun_out/r02_cmp_g.csv python tools/fused_probe.py c2 1 g > /dev/null 2>&1
timeout 300 ncu --metrics $M --clock-control none -k regex:"qgt_sweep|qgt_fused" --csv --log-file gpurun_out/r02_cmp_d.csv python tools/fused_probe.py c2 1 d > /dev/null 2>&1
timeout 300 ncu --metrics $M --clock-control none -k regex:"qgt_sweep|qgt_fused" --csv --log-file gpurun_out/r02_cmp_d1.csv python tools/fused_probe.py c2 1 d:1 > /dev/null 2>&1
ls -la gpurun_out/*.csv
