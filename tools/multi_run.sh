N=$1
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 2 --warmup 3 > gpurun_out/r02_bench_c3_sharded$N.json 2>gpurun_out/err_s$N.txt
tail -2 gpurun_out/err_s$N.txt | cut -c1-300; tail -1 gpurun_out/r02_bench_c3_sharded$N.json | cut -c1-900
