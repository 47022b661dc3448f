# final single-GPU run of round 2 (b): GPU test suite, the driver's bench command for both arms, profiles of the same build
OUT=gpurun_out
mkdir -p $OUT
timeout 1500 python -m pytest tests -m gpu -q -rs > $OUT/r2final_gpu_tests.log 2>&1; echo "tests rc=$?"; tail -4 $OUT/r2final_gpu_tests.log
PICLAS_GPU_DEBUG=1 timeout 1500 python bench.py --gpus 1 --steps 20 --warmup 5 > $OUT/r2final_bench.json 2> $OUT/r2final_bench.err; echo "bench rc=$?"; grep -c "re-planning" $OUT/r2final_bench.err
timeout 900 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > $OUT/r2final_bench_reference.json 2> $OUT/r2final_bench_reference.err; echo "ref rc=$?"; tail -c 700 $OUT/r2final_bench_reference.json
bash scripts/r2_profile.sh r2final
