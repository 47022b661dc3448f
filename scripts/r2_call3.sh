# round 2, GPU call 3: binned layout — parity suite, then 1/8-size bench (bins vs sorted arrays)
OUT=gpurun_out
mkdir -p $OUT
timeout 1500 python -m pytest tests -m gpu -q -x --durations=8 > $OUT/c3_gpu_tests.log 2>&1; echo "pytest rc=$?"; tail -40 $OUT/c3_gpu_tests.log
echo "== bins"; timeout 300 bash scripts/bench32.sh 2>&1 | tee $OUT/c3_bench32_bins.log
echo "== sorted"; PICLAS_GPU_LAYOUT=sorted timeout 300 bash scripts/bench32.sh 2>&1 | tee $OUT/c3_bench32_sorted.log
