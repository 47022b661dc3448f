OUT=gpurun_out
mkdir -p $OUT
B="python bench.py --nelem 32 --particles 6.25e7 --steps 2 --warmup 3 --no-cpu --no-e2e --no-checks"
for v in build/variants/fh5.so build/variants/fh6.so ""; do
  PICLAS_GPU_LIB=$v timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'k_far_hint|k_far_walk' --launch-skip 4 -c 6 --csv --log-file $OUT/c24.csv $B > $OUT/c24.log 2>&1
  echo "== ${v:-default(minb 8)}"; grep -E "k_far_(hint|walk)" $OUT/c24.csv | awk -F'","' '{print $5, $NF}' | tr -d '"' | awk '{print $1, $NF}' | sort | uniq -c | head -12
done
