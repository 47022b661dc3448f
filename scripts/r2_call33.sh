OUT=gpurun_out
mkdir -p $OUT
timeout 600 python -m pytest tests/test_gpu_emission.py tests/test_gpu_parity.py -m gpu -x -q -k "emission or restart or full_regions or cartesian_box_steps" > $OUT/c33_tests.log 2>&1; echo "tests rc=$?"; tail -3 $OUT/c33_tests.log
# the driver's command, new default field (kick 0.5 % of v_th per step) and round 1's field (5 %)
PICLAS_GPU_DEBUG=1 timeout 1500 python bench.py --gpus 1 --steps 20 --warmup 5 > $OUT/c33_driver_cmd.json 2> $OUT/c33_driver_cmd.err; echo "bench rc=$?"
PICLAS_GPU_DEBUG=1 timeout 1500 python bench.py --gpus 1 --steps 20 --warmup 5 --kick 0.05 --no-cpu > $OUT/c33_driver_cmd_kick5.json 2> $OUT/c33_driver_cmd_kick5.err; echo "bench rc=$?"
python - <<'PY'
import json
for f in ('c33_driver_cmd','c33_driver_cmd_kick5'):
    d=json.loads(open('gpurun_out/%s.json'%f).read().strip().splitlines()[-1])
    print(f, round(d['ms_per_step'],3), d['value'], {k:round(v,2) for k,v in d['roofline']['phase_ms'].items()}, d['e2e'], d['checks']['ok'])
PY
for f in c33_driver_cmd c33_driver_cmd_kick5; do echo $f; grep -c "re-planning" $OUT/$f.err; grep "push_track" $OUT/$f.err | awk '{print $12, $20, $22}' | tr '\n' ';' | cut -c1-1200; echo; done
