OUT=gpurun_out
mkdir -p $OUT
timeout 1200 python -m pytest tests -m gpu -q -x > $OUT/c9_tests.log 2>&1; echo "pytest rc=$?"; tail -5 $OUT/c9_tests.log
bash scripts/run_variants.sh default $PWD/piclas_b200/libpiclas_gpu_mb16.so 2>&1 | tee $OUT/c9_variants.log
timeout 900 bash scripts/r2_profile.sh r2f
