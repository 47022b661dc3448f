OUT=gpurun_out
mkdir -p $OUT
timeout 900 python -m pytest tests/test_multi_rank.py -m gpu -q -rs > $OUT/c14_multi_rank_2gpu.log 2>&1; echo "pytest rc=$?"; tail -8 $OUT/c14_multi_rank_2gpu.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29555 bench.py --gpus 2 --nelem 32 --particles 6.25e7 --steps 5 --warmup 3 --no-cpu > $OUT/c14_bench_2gpu_32.json 2> $OUT/c14_bench_2gpu_32.err; tail -c 1600 $OUT/c14_bench_2gpu_32.json; tail -3 $OUT/c14_bench_2gpu_32.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29556 bench.py --gpus 2 --steps 5 --warmup 3 --no-cpu --no-checks > $OUT/c14_bench_2gpu_64.json 2> $OUT/c14_bench_2gpu_64.err; tail -c 1200 $OUT/c14_bench_2gpu_64.json; tail -3 $OUT/c14_bench_2gpu_64.err
