OUT=gpurun_out
mkdir -p $OUT
B="python bench.py --nelem 32 --particles 6.25e7 --steps 2 --warmup 3 --no-cpu --no-e2e --no-checks"
PICLAS_GPU_DEBUG=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'k_far|k_bin|radix|k_scan|k_hist|k_scatter|k_gather' --launch-skip 60 -c 60 --csv --log-file $OUT/c22_launches.csv $B > $OUT/c22.log 2>&1
grep "push_track" $OUT/c22.log | tail -2
python - <<'PY'
import csv,collections
rows=[r for r in csv.reader(open('gpurun_out/c22_launches.csv')) if len(r)>5]
hdr=[i for i,r in enumerate(rows) if 'Kernel Name' in r][0]
H=rows[hdr]; ki=H.index('Kernel Name'); vi=H.index('Metric Value')
d=collections.defaultdict(list)
for r in rows[hdr+1:]:
    try: d[r[ki].split('(')[0][:60]].append(float(r[vi].replace(',','')))
    except: pass
for k,v in sorted(d.items(), key=lambda kv:-sum(kv[1])): print(f"{k:60s} n={len(v):3d} avg={sum(v)/len(v)/1e3:9.1f} us")
PY
