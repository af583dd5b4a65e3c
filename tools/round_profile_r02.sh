#!/bin/sh
# Round-2 evidence for profiles/ (one GPU box, through gpurun; outputs land in gpurun_out/):
#   sh tools/round_profile_r02.sh
O=gpurun_out
mkdir -p $O
nvidia-smi --query-gpu=name,driver_version,clocks.max.sm,memory.total --format=csv > $O/r02_gpu.txt
# 1. the driver's command: default workload c3 (28 qubits, 256 parameters)
python bench.py > $O/r02_bench_c3_final.json 2> $O/r02_bench_c3_final.err
# 2. launch list of one evaluation of the same command (cold-cache, serialised: compare shares, not absolutes)
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file $O/r02_launches_c3.csv \
    python bench.py --steps 1 --warmup 1 --explore --no-cpu-baseline > $O/r02_ncu_launches.log 2>&1
# 3. DRAM bytes of the dominant kernel's launches in the same command (single-pass metrics: no kernel replay)
timeout 1200 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none \
    --kernel-name regex:"qgt_fused" --launch-skip 240 -c 220 --csv --log-file $O/r02_dram_c3.csv \
    python bench.py --steps 1 --warmup 1 --explore --no-cpu-baseline > $O/r02_ncu_dram.log 2>&1
# 4. full capture of one direct-kernel launch (28 qubits would need save/restore of a 130 GB arena between replay passes:
#    the same kernel on the 24-qubit sub-problem with the column count capped, so that the blocked fused schedule runs)
timeout 900 ncu --set full --import-source on --clock-control none --kernel-name regex:qgt_fused_direct --launch-skip 30 --launch-count 1 \
    -f -o $O/r02_fused_direct_c3s python tools/fused_probe.py c3s 1 d max_slots=48 > $O/r02_ncu_full.log 2>&1
ls -la $O/r02_*
