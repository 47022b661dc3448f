for v in "" piclas_b200/lib_ka5_dep4.so piclas_b200/lib_ka6_dep4.so piclas_b200/lib_ka4_dep6.so piclas_b200/lib_ka4_dep8.so; do
  echo "== variant: ${v:-default}"
  PICLAS_GPU_LIB=$v python bench.py --nelem 32 --particles 6.25e7 --steps 5 --warmup 2 --no-cpu --no-e2e 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(round(d['ms_per_step'],2), {k:round(v,2) for k,v in d['roofline']['phase_ms'].items()})"
done
