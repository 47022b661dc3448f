# final single-GPU run of round 2: GPU test suite, both bench arms, profiles, sanitizer on the kernels added last
OUT=gpurun_out
mkdir -p $OUT
timeout 1500 python -m pytest tests -m gpu -q -rs > $OUT/r2final_gpu_tests.log 2>&1; echo "tests rc=$?"; tail -4 $OUT/r2final_gpu_tests.log
PICLAS_GPU_DEBUG=1 timeout 1500 python bench.py > $OUT/r2final_bench.json 2> $OUT/r2final_bench.err; echo "bench rc=$?"; tail -c 600 $OUT/r2final_bench.json; grep -c "re-planning" $OUT/r2final_bench.err
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > $OUT/r2final_bench_reference.json 2> $OUT/r2final_bench_reference.err; echo "ref rc=$?"; tail -c 900 $OUT/r2final_bench_reference.json
bash scripts/r2_profile.sh r2final
SEL="cartesian_box_steps and restructured and 3 or full_regions or many_particles_per_element and restructured or open_boundaries and restructured or degenerate_flights and restructured"
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 7 --print-limit 5 python -m pytest tests/test_gpu_parity.py tests/test_gpu_emission.py -m gpu -q -x -k "$SEL or emission or sin_deviation or cos_distribution or append" > $OUT/r2final_memcheck.log 2>&1; echo "memcheck rc=$?"; grep -E "ERROR SUMMARY|passed|failed" $OUT/r2final_memcheck.log | tail -4
timeout 900 compute-sanitizer --tool racecheck --racecheck-report analysis --error-exitcode 7 --print-limit 5 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "cartesian_box_steps and restructured and 3 or full_regions and replan" > $OUT/r2final_racecheck.log 2>&1; echo "racecheck rc=$?"; grep -E "RACECHECK SUMMARY|passed|failed" $OUT/r2final_racecheck.log | tail -4
