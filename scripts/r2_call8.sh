OUT=gpurun_out
mkdir -p $OUT
timeout 1200 python -m pytest tests/test_zz_gpu_properties.py tests/test_gpu_parity.py tests/test_gpu_loopback.py tests/test_gpu_configs.py tests/test_reference_tracking.py -m gpu -q -x > $OUT/c8_tests.log 2>&1; echo "pytest rc=$?"; tail -5 $OUT/c8_tests.log
bash scripts/run_variants.sh default $PWD/piclas_b200/libpiclas_gpu_nt64.so $PWD/piclas_b200/libpiclas_gpu_nt128.so 2>&1 | tee $OUT/c8_variants.log
timeout 900 bash scripts/r2_profile.sh r2e
