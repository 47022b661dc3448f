# usage: scripts/profile.sh TAG   -- launch list + one --set full capture of the step's kernels at 16^3 (same particles per element)
TAG=${1:-r1}
OUT=gpurun_out
B="python bench.py --nelem 16 --particles 7.8125e6 --steps 2 --warmup 1 --no-cpu --no-e2e"
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/${TAG}_launches.csv $B > $OUT/${TAG}_launches.log 2>&1
ncu --set full --clock-control none --import-source on --launch-skip 40 -c 24 -f -o $OUT/${TAG}_full $B > $OUT/${TAG}_full.log 2>&1
ncu -i $OUT/${TAG}_full.ncu-rep --page raw --csv > $OUT/${TAG}_raw.csv 2>/dev/null
for k in k_interp_push k_track_leavers k_deposit_cvwm k_gather_particles k_scatter; do
  ncu -i $OUT/${TAG}_full.ncu-rep --page source --csv --kernel-name regex:$k 2>/dev/null > $OUT/${TAG}_src_$k.csv
done
ls -la $OUT | tail -12
