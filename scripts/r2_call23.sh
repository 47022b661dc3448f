OUT=gpurun_out
mkdir -p $OUT
timeout 1500 python -m pytest tests -m gpu -x -q > $OUT/c23_tests.log 2>&1; echo "tests rc=$?"; tail -5 $OUT/c23_tests.log
bash scripts/r2_call22.sh 2>&1 | grep -v '^{"metric"'
PICLAS_GPU_DEBUG=1 timeout 600 python bench.py --nelem 32 --particles 6.25e7 --steps 10 --warmup 4 --no-cpu --no-e2e > $OUT/c23_32.json 2> $OUT/c23_32.err
python - <<PY
import json
d=json.loads(open('gpurun_out/c23_32.json').read().strip().splitlines()[-1])
print('32^3', d['ms_per_step'], d['roofline']['phase_ms'])
PY
