# usage: scripts/r2_final_multi.sh N -- final N-GPU run of round 2: (N = 2: the NCCL tests,) bench.py --gpus N at the flagship size
N=${1:-2}
OUT=gpurun_out
mkdir -p $OUT
if [ "$N" = 2 ]; then
  timeout 900 python -m pytest tests/test_multi_rank.py -m gpu -q -rs > $OUT/r2final_multi_rank_nccl_2gpu.log 2>&1; echo "pytest rc=$?"; tail -4 $OUT/r2final_multi_rank_nccl_2gpu.log
fi
PICLAS_MULTI_TIMING=${PICLAS_MULTI_TIMING:-} timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29555 bench.py --gpus $N --steps 20 --warmup 5 --no-cpu > $OUT/r2final_scale_${N}gpu.json 2> $OUT/r2final_scale_${N}gpu.err; echo "bench rc=$?"
python - <<PY
import json
d=json.loads(open('gpurun_out/r2final_scale_${N}gpu.json').read().strip().splitlines()[-1])
print(d['n_gpus'], round(d['ms_per_step'],3), d['value'], {k:round(v,2) for k,v in d['roofline']['phase_ms'].items()}, d['e2e'] and round(d['e2e']['ms_per_step'],2), d['checks']['ok'], d.get('migrated_per_step'))
PY
tail -2 $OUT/r2final_scale_${N}gpu.err
grep "host cpus\|NUMA" $OUT/r2final_scale_${N}gpu.err | head -8
