OUT=gpurun_out
mkdir -p $OUT
timeout 900 python -m pytest tests/test_zz_gpu_properties.py tests/test_gpu_parity.py tests/test_gpu_loopback.py tests/test_gpu_errors.py -m gpu -q -x > $OUT/c5_tests.log 2>&1; echo "pytest rc=$?"; tail -5 $OUT/c5_tests.log
bash scripts/run_variants.sh default $PWD/piclas_b200/libpiclas_gpu_mb2.so $PWD/piclas_b200/libpiclas_gpu_mb4.so 2>&1 | tee $OUT/c5_variants.log
timeout 900 bash scripts/r2_profile.sh r2b
