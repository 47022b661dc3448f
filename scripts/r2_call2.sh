# round 2, GPU call 2 (2 GPUs): the NCCL multi-rank parity tests the driver's 1-GPU box skips + a 2-GPU bench line with checks
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=index,name --format=csv > $OUT/c2_gpus.txt
timeout 900 python -m pytest tests/test_multi_rank.py -m gpu -q -rs > $OUT/c2_multi_rank_2gpu.log 2>&1; echo "pytest rc=$?"; tail -8 $OUT/c2_multi_rank_2gpu.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29555 bench.py --gpus 2 --nelem 32 --particles 6.25e7 --steps 5 --warmup 3 --no-cpu > $OUT/c2_bench_2gpu_32.json 2> $OUT/c2_bench_2gpu_32.err; tail -c 1500 $OUT/c2_bench_2gpu_32.json; tail -3 $OUT/c2_bench_2gpu_32.err
