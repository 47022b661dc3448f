OUT=gpurun_out
mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_loopback.py tests/test_zz_gpu_properties.py -m gpu -x -q > $OUT/c30_tests.log 2>&1; echo "tests rc=$?"; tail -3 $OUT/c30_tests.log
bash scripts/r2_call22.sh 2>&1 | grep -v '^{"metric"'
PICLAS_GPU_DEBUG=1 timeout 1200 python bench.py --steps 6 --warmup 3 --no-cpu > $OUT/c30_full.json 2> $OUT/c30_full.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/c30_full.json').read().strip().splitlines()[-1])
print('64^3', d['ms_per_step'], d['value'], d['roofline']['phase_ms'], d['e2e'])
PY
grep "push_track" $OUT/c30_full.err | tail -3
