# quick 1-GPU bench at 1/8 of the flagship size (same particles per element), both arithmetic modes
for a in 1 0; do
  echo "== arithmetic=$a"
  python bench.py --nelem 32 --particles 6.25e7 --steps 5 --warmup 2 --no-cpu --no-e2e --arithmetic $a 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('%.3e p-steps/s'%d['value'], round(d['ms_per_step'],2),'ms', {k:round(v,2) for k,v in d['roofline']['phase_ms'].items()})"
done
