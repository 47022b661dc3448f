OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29557 bench.py --gpus 8 --steps 10 --warmup 3 --no-cpu > $OUT/c17_bench_8gpu.json 2> $OUT/c17_bench_8gpu.err; tail -c 2200 $OUT/c17_bench_8gpu.json; tail -3 $OUT/c17_bench_8gpu.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29558 bench.py --gpus 4 --steps 10 --warmup 3 --no-cpu --no-checks > $OUT/c17_bench_4gpu.json 2> $OUT/c17_bench_4gpu.err; tail -c 900 $OUT/c17_bench_4gpu.json
