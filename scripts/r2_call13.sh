OUT=gpurun_out
PICLAS_GPU_DEBUG=1 timeout 1200 python bench.py --steps 6 --warmup 3 --no-cpu > $OUT/c13_full.json 2> $OUT/c13_full.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/c13_full.json').read().strip().splitlines()[-1])
print(d['ms_per_step'], d['event_ms_per_step'], d['value'], d['roofline']['phase_ms'], d['e2e'])
PY
grep "piclas_gpu" $OUT/c13_full.err | tail -16
