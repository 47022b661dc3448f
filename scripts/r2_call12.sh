OUT=gpurun_out
PICLAS_GPU_DEBUG=1 timeout 900 python bench.py --steps 3 --warmup 1 --no-cpu --no-e2e > $OUT/c12_dbg.json 2> $OUT/c12_dbg.err; tail -c 700 $OUT/c12_dbg.json; grep "piclas_gpu" $OUT/c12_dbg.err | tail -12; nvidia-smi --query-gpu=memory.used --format=csv
