# single-GPU: GPU suite and the driver's bench command on the final build (asynchronous PartSource hand-off)
OUT=gpurun_out
mkdir -p $OUT
timeout 1500 python -m pytest tests -m gpu -q -rs > $OUT/r2final_gpu_tests.log 2>&1; echo "tests rc=$?"; tail -4 $OUT/r2final_gpu_tests.log
PICLAS_GPU_DEBUG=1 timeout 1500 python bench.py --gpus 1 --steps 20 --warmup 5 > $OUT/r2final_bench.json 2> $OUT/r2final_bench.err; echo "bench rc=$?"; grep -c "re-planning" $OUT/r2final_bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2final_bench.json').read().strip().splitlines()[-1])
print(round(d['ms_per_step'],3), d['value'], json.dumps(d['e2e']), d['checks']['ok'])
PY
