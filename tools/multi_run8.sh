N=8
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 tests/multi_gpu_check.py > gpurun_out/r02_multi_check_$N.txt 2>&1
grep -E "n=22 P|passed|FAIL" gpurun_out/r02_multi_check_$N.txt | tail -6
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 3 --warmup 3 > gpurun_out/r02_bench_c3_sharded$N.json 2>gpurun_out/err_s$N.txt
tail -1 gpurun_out/r02_bench_c3_sharded$N.json | cut -c1-1800
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 tools/run_dist.py c5 > gpurun_out/r02_run_c5_8gpu.json 2>gpurun_out/err_c5.txt
tail -2 gpurun_out/err_c5.txt; tail -1 gpurun_out/r02_run_c5_8gpu.json | cut -c1-1800
