# usage: scripts/run_variants.sh lib1.so lib2.so ...   ("default" = the in-tree library); 1/8-size bench per variant
for v in "$@"; do
  [ "$v" = default ] && lib="" || lib=$v
  echo "== variant: $v"
  PICLAS_GPU_LIB=$lib python bench.py --nelem 32 --particles 6.25e7 --steps 5 --warmup 2 --no-cpu --no-e2e 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(round(d['ms_per_step'],2), {k:round(v,2) for k,v in d['roofline']['phase_ms'].items()})"
done
