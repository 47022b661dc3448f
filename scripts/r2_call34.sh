OUT=gpurun_out
mkdir -p $OUT
# per-rank load of the 8-GPU run on two GPUs: 40^3 elements, 1.22e8 particles
PICLAS_MULTI_TIMING=1 PICLAS_GPU_DEBUG=1 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29555 bench.py --gpus 2 --nelem 40 --particles 1.22e8 --steps 10 --warmup 4 --no-cpu --no-e2e --no-checks > $OUT/c34.json 2> $OUT/c34.err; echo "rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/c34.json').read().strip().splitlines()[-1])
print(d['n_gpus'], round(d['ms_per_step'],3), {k:round(v,2) for k,v in d['roofline']['phase_ms'].items()}, d.get('migrated_per_step'))
PY
grep "multi timing" $OUT/c34.err
grep "push_track (bins)" $OUT/c34.err | tail -2
