OUT=gpurun_out
mkdir -p $OUT
timeout 600 python -m pytest tests/test_zz_gpu_properties.py tests/test_gpu_parity.py -m gpu -q -x > $OUT/c10_tests.log 2>&1; echo "pytest rc=$?"; tail -3 $OUT/c10_tests.log
bash scripts/run_variants.sh default $PWD/piclas_b200/libpiclas_gpu_v1.so $PWD/piclas_b200/libpiclas_gpu_v2.so 2>&1 | tee $OUT/c10_variants.log
nvidia-smi --query-gpu=memory.used,memory.total --format=csv
timeout 1200 python bench.py --steps 5 --warmup 3 --no-cpu > $OUT/c10_bench_full.json 2> $OUT/c10_bench_full.err; tail -c 2500 $OUT/c10_bench_full.json; tail -5 $OUT/c10_bench_full.err
