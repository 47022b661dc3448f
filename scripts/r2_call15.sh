OUT=gpurun_out
mkdir -p $OUT
timeout 1500 python -m pytest tests -m gpu -q -x > $OUT/c15_tests.log 2>&1; echo "pytest rc=$?"; tail -5 $OUT/c15_tests.log
bash scripts/run_variants.sh default $PWD/piclas_b200/libpiclas_gpu_a.so $PWD/piclas_b200/libpiclas_gpu_b.so 2>&1 | tee $OUT/c15_variants.log
timeout 900 bash scripts/r2_profile.sh r2g
