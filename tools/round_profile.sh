#!/bin/sh
# One GPU box: the evidence a round commits under profiles/ (run through gpurun; outputs land in gpurun_out/).
#   sh tools/round_profile.sh r01
R=${1:-r01}
O=gpurun_out
mkdir -p $O
nvidia-smi --query-gpu=name,driver_version,clocks.max.sm,memory.total --format=csv > $O/${R}_gpu.txt
tools/_build/peaks > $O/${R}_peaks.json 2> $O/${R}_peaks.err
{ echo "# dmma_ilp"; tools/_build/dmma_ilp; echo "# dual_pipe"; tools/_build/dual_pipe; echo "# subpass_bench"; tools/_build/subpass_bench; } > $O/${R}_microbench.txt 2>&1
tools/_build/lds_pattern > $O/${R}_lds_pattern.txt 2>&1
python -m pytest tests -m gpu -q 2>&1 | tail -3 > $O/${R}_pytest_gpu.txt
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $O/${R}_smoke.txt 2>&1
python bench.py > $O/${R}_bench_c2.json 2> $O/${R}_bench_c2.err
python bench.py --impl reference --steps 2 --warmup 1 > $O/${R}_bench_c2_reference.json 2> $O/${R}_bench_c2_reference.err
QGT_B200_TRACE=1 python tools/trace_run.py c2 > $O/${R}_stats_c2.txt 2> $O/${R}_trace_c2_all.txt
grep "cat=" $O/${R}_trace_c2_all.txt | tail -21 > $O/${R}_trace_c2_launches.txt; rm -f $O/${R}_trace_c2_all.txt
# launch list of one bench step (cold-cache, serialised: compare shares, not absolutes)
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/${R}_launches_c2.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > $O/${R}_ncu_launches.log 2>&1
# DRAM bytes per sweep / Gram launch (source of roofline.traffic; tools/make_traffic.py turns it into profiles/traffic.json)
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none --kernel-name regex:"qgt_(sweep|gram)_kernel" -c 80 --csv --log-file $O/${R}_dram_c2.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline > $O/${R}_ncu_dram.log 2>&1
# full captures (c2: 9 sweep launches + 1 Gram per evaluation): a dense-stage sweep launch (run 4, 102 columns), a
# diagonal-real one (run 8, 160 columns), the Gram
ncu --set full --import-source on --clock-control none --kernel-name regex:qgt_sweep_kernel --launch-skip 13 --launch-count 1 -o $O/${R}_sweep_dense -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline > $O/${R}_ncu_a.log 2>&1
ncu --set full --import-source on --clock-control none --kernel-name regex:qgt_sweep_kernel --launch-skip 17 --launch-count 1 -o $O/${R}_sweep_diagreal -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline > $O/${R}_ncu_b.log 2>&1
ncu --set full --import-source on --clock-control none --kernel-name regex:qgt_gram_kernel --launch-skip 1 --launch-count 1 -o $O/${R}_gram -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline > $O/${R}_ncu_c.log 2>&1
ls -la $O/*.ncu-rep
