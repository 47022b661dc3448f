OUT=gpurun_out
mkdir -p $OUT
timeout 1500 python -m pytest tests -m gpu -x -q > $OUT/c25_tests.log 2>&1; echo "tests rc=$?"; tail -5 $OUT/c25_tests.log
bash scripts/r2_call22.sh 2>&1 | grep -v '^{"metric"'
