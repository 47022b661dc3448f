# round 2, GPU call 1 (1 GPU): the whole -m gpu suite incl. the new parity tests, FP64 peak, secondary variant, 1/8-size baseline
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > $OUT/c1_gpu.txt
timeout 1200 python -m pytest tests -m gpu -q -x --durations=15 > $OUT/c1_gpu_tests.log 2>&1; echo "pytest rc=$?"; tail -25 $OUT/c1_gpu_tests.log
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o $OUT/fp64_peak scripts/fp64_peak.cu && $OUT/fp64_peak | tee $OUT/c1_fp64_peak.json
timeout 300 bash scripts/bench32.sh 2>&1 | tee $OUT/c1_bench32.log
timeout 600 python bench.py --variant ref_sf --particles 2e7 --steps 3 --warmup 3 --no-cpu > $OUT/c1_bench_ref_sf.json 2> $OUT/c1_bench_ref_sf.err; tail -c 1800 $OUT/c1_bench_ref_sf.json; tail -3 $OUT/c1_bench_ref_sf.err
