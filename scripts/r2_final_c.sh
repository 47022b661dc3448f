# last single-GPU call of round 2: full GPU suite once more on the final build, sanitizer subset, the driver's bench command
OUT=gpurun_out
mkdir -p $OUT
timeout 1500 python -m pytest tests -m gpu -q -rs > $OUT/r2final_gpu_tests.log 2>&1; echo "tests rc=$?"; tail -4 $OUT/r2final_gpu_tests.log
SEL="cartesian_box_steps and restructured and 3 or full_regions or many_particles_per_element and restructured or open_boundaries and restructured or degenerate_flights and restructured"
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 7 --print-limit 5 python -m pytest tests/test_gpu_parity.py tests/test_gpu_emission.py -m gpu -q -x -k "$SEL or emission or sin_deviation or cos_distribution or append" > $OUT/r2final_memcheck.log 2>&1; echo "memcheck rc=$?"; grep -E "ERROR SUMMARY|passed|failed" $OUT/r2final_memcheck.log | tail -4
timeout 900 compute-sanitizer --tool racecheck --racecheck-report analysis --error-exitcode 7 --print-limit 5 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "cartesian_box_steps and restructured and 3 or full_regions and replan" > $OUT/r2final_racecheck.log 2>&1; echo "racecheck rc=$?"; grep -E "RACECHECK SUMMARY|passed|failed" $OUT/r2final_racecheck.log | tail -4
PICLAS_GPU_DEBUG=1 timeout 1500 python bench.py --gpus 1 --steps 20 --warmup 5 > $OUT/r2final_bench.json 2> $OUT/r2final_bench.err; echo "bench rc=$?"; grep -c "re-planning" $OUT/r2final_bench.err
PICLAS_GPU_DEBUG=1 timeout 1500 python bench.py --gpus 1 --steps 20 --warmup 5 --kick 0.05 --no-cpu > $OUT/r2final_bench_kick5.json 2> $OUT/r2final_bench_kick5.err; echo "bench rc=$?"; grep -c "re-planning" $OUT/r2final_bench_kick5.err
python - <<'PY'
import json
for f in ('r2final_bench','r2final_bench_kick5'):
    d=json.loads(open('gpurun_out/%s.json'%f).read().strip().splitlines()[-1])
    print(f, round(d['ms_per_step'],3), d['value'], {k:round(v,2) for k,v in d['roofline']['phase_ms'].items()}, d['e2e'], d['checks']['ok'], d['roofline']['traffic'], d['roofline']['frac'])
PY
grep "push_track" $OUT/r2final_bench_kick5.err | sed -e 's/.*through the far list, //' -e 's/, [0-9]* walked.*//' | tr '\n' ';' | cut -c1-900
