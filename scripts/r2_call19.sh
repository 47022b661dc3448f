OUT=gpurun_out
mkdir -p $OUT
timeout 900 python -m pytest tests/test_multi_rank.py -m gpu -q -rs > $OUT/c19_multi_rank_2gpu.log 2>&1; echo "pytest rc=$?"; tail -4 $OUT/c19_multi_rank_2gpu.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29555 bench.py --gpus 2 --nelem 32 --particles 6.25e7 --steps 10 --warmup 3 --no-cpu > $OUT/c19_bench_2gpu_32.json 2> $OUT/c19_bench_2gpu_32.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/c19_bench_2gpu_32.json').read().strip().splitlines()[-1])
print(d['n_gpus'], round(d['ms_per_step'],3), {k:round(v,2) for k,v in d['roofline']['phase_ms'].items()}, d['e2e']['ms_per_step'], d['checks']['ok'], d['checks'].get('parity_small_case'))
PY
tail -2 $OUT/c19_bench_2gpu_32.err
