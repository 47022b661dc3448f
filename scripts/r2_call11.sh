OUT=gpurun_out
PICLAS_GPU_DEBUG=1 timeout 600 python bench.py --nelem 64 --particles 1e8 --steps 4 --warmup 1 --no-cpu --no-e2e > $OUT/c11_dbg.json 2> $OUT/c11_dbg.err; tail -c 600 $OUT/c11_dbg.json; grep -c "re-planning" $OUT/c11_dbg.err; grep "piclas_gpu" $OUT/c11_dbg.err | tail -12
PICLAS_GPU_DEBUG=1 timeout 600 python bench.py --nelem 48 --particles 2.1e8 --steps 4 --warmup 1 --no-cpu --no-e2e > $OUT/c11_dbg2.json 2> $OUT/c11_dbg2.err; tail -c 400 $OUT/c11_dbg2.json; grep "piclas_gpu" $OUT/c11_dbg2.err | tail -8
