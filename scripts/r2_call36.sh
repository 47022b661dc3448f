OUT=gpurun_out
mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -q -rs > $OUT/r2final_gpu_tests.log 2>&1; echo "tests rc=$?"; tail -4 $OUT/r2final_gpu_tests.log
timeout 600 python bench.py --variant ref_sf --particles 2e7 --steps 3 --warmup 3 --no-cpu > $OUT/c36_bench_ref_sf.json 2> $OUT/c36_bench_ref_sf.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/c36_bench_ref_sf.json').read().strip().splitlines()[-1])
print('ref_sf', d['ms_per_step'], d['roofline']['phase_ms'], d['checks'].get('ok'))
PY
tail -2 $OUT/c36_bench_ref_sf.err
