OUT=gpurun_out
mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_emission.py -m gpu -x -q > $OUT/c31_tests.log 2>&1; echo "tests rc=$?"; tail -30 $OUT/c31_tests.log
