# usage: scripts/r2_profile.sh TAG [extra bench args] -- launch list + one --set full capture of the step's kernels at 32^3 / 6.25e7 particles
TAG=${1:-r2}; shift
OUT=gpurun_out
B="python bench.py --nelem 32 --particles 6.25e7 --steps 2 --warmup 1 --no-cpu --no-e2e --no-checks $@"
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'k_bin_|k_far_|k_node|k_nodes|k_hist|k_col_|k_scatter|k_gather|k_segment|k_scan|k_halo' -c 400 --csv --log-file $OUT/${TAG}_launches.csv $B > $OUT/${TAG}_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'k_bin_push|k_bin_deposit_cvwm|k_far_walk|k_far_hint' --launch-skip 4 -c 4 -f -o $OUT/${TAG}_full $B > $OUT/${TAG}_full.log 2>&1
ncu -i $OUT/${TAG}_full.ncu-rep --page raw --csv > $OUT/${TAG}_raw.csv 2>/dev/null
for k in k_bin_push k_bin_deposit_cvwm k_far_walk k_far_hint; do
  ncu -i $OUT/${TAG}_full.ncu-rep --page source --csv --kernel-name regex:$k 2>/dev/null > $OUT/${TAG}_src_$k.csv
done
python scripts/summarize_raw.py $OUT/${TAG}_raw.csv
