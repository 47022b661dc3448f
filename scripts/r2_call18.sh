OUT=gpurun_out
mkdir -p $OUT
timeout 1500 python -m pytest tests -m gpu -q -x > $OUT/c18_tests.log 2>&1; echo "pytest rc=$?"; tail -5 $OUT/c18_tests.log
bash scripts/run_variants.sh default $PWD/piclas_b200/libpiclas_gpu_lv5.so 2>&1 | tee $OUT/c18_variants.log
timeout 600 python bench.py --variant ref_sf --particles 2e7 --steps 3 --warmup 3 --no-cpu > $OUT/c18_bench_ref_sf.json 2> $OUT/c18_bench_ref_sf.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/c18_bench_ref_sf.json').read().strip().splitlines()[-1])
print('ref_sf', d['ms_per_step'], d['roofline']['phase_ms'], d['checks'].get('ok'))
PY
tail -3 $OUT/c18_bench_ref_sf.err
