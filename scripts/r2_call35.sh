OUT=gpurun_out
mkdir -p $OUT
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_errors.py -m gpu -x -q -k "partsource_async or unimplemented or cartesian_box_steps" > $OUT/c35_tests.log 2>&1; echo "tests rc=$?"; tail -3 $OUT/c35_tests.log
python __graft_entry__.py --smoke 2>&1 | tail -2
timeout 1500 python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu > $OUT/c35_bench.json 2> $OUT/c35_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/c35_bench.json').read().strip().splitlines()[-1])
print(round(d['ms_per_step'],3), d['value'], json.dumps(d['e2e']))
PY
