OUT=gpurun_out
mkdir -p $OUT
# far-hint kernel: parity suites, then the walk share at 32^3 / 6.25e7 and 64^3 / 5e8 with and without it
timeout 1500 python -m pytest tests -m gpu -x -q > $OUT/c21_tests.log 2>&1; echo "tests rc=$?"; tail -5 $OUT/c21_tests.log
for v in hint nohint; do
  if [ $v = nohint ]; then export PICLAS_GPU_NO_FAR_HINT=1; else unset PICLAS_GPU_NO_FAR_HINT; fi
  PICLAS_GPU_DEBUG=1 timeout 600 python bench.py --nelem 32 --particles 6.25e7 --steps 10 --warmup 4 --no-cpu --no-e2e > $OUT/c21_32_$v.json 2> $OUT/c21_32_$v.err
  python - <<PY
import json
d=json.loads(open('gpurun_out/c21_32_$v.json').read().strip().splitlines()[-1])
print('$v 32^3', d['ms_per_step'], d['roofline']['phase_ms'])
PY
  grep "push_track" $OUT/c21_32_$v.err | tail -2
done
unset PICLAS_GPU_NO_FAR_HINT
PICLAS_GPU_DEBUG=1 timeout 1200 python bench.py --steps 6 --warmup 3 --no-cpu --no-e2e > $OUT/c21_full.json 2> $OUT/c21_full.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/c21_full.json').read().strip().splitlines()[-1])
print('64^3', d['ms_per_step'], d['value'], d['roofline']['phase_ms'])
PY
grep "push_track" $OUT/c21_full.err | tail -3
