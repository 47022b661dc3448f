OUT=gpurun_out
mkdir -p $OUT
B="python bench.py --nelem 32 --particles 6.25e7 --steps 2 --warmup 1 --no-cpu --no-e2e --no-checks"
timeout 800 ncu --set full --clock-control none --import-source on -k regex:'k_far_hint|k_far_walk' --launch-skip 2 -c 2 -f -o $OUT/c26_full $B > $OUT/c26_full.log 2>&1
ncu -i $OUT/c26_full.ncu-rep --page raw --csv > $OUT/c26_raw.csv 2>/dev/null
python scripts/summarize_raw.py $OUT/c26_raw.csv
for k in k_far_hint k_far_walk; do ncu -i $OUT/c26_full.ncu-rep --page source --csv --kernel-name regex:$k 2>/dev/null > $OUT/c26_src_$k.csv; done
python profiles/summarize_source.py $OUT/c26_src_k_far_hint.csv 2>&1 | head -60
