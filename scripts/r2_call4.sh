OUT=gpurun_out
mkdir -p $OUT
timeout 900 python -m pytest tests/test_zz_gpu_properties.py tests/test_gpu_parity.py tests/test_gpu_loopback.py -m gpu -q -x > $OUT/c4_tests.log 2>&1; echo "pytest rc=$?"; tail -5 $OUT/c4_tests.log
timeout 200 bash scripts/bench32.sh 2>&1 | tee $OUT/c4_bench32.log
timeout 900 bash scripts/r2_profile.sh r2a
