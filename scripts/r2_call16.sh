OUT=gpurun_out
mkdir -p $OUT
timeout 1500 python -m pytest tests -m gpu -q -x > $OUT/c16_tests.log 2>&1; echo "pytest rc=$?"; tail -5 $OUT/c16_tests.log
timeout 300 bash scripts/bench32.sh 2>&1 | tee $OUT/c16_bench32.log
timeout 900 bash scripts/r2_profile.sh r2h
