OUT=gpurun_out
mkdir -p $OUT
# compute-sanitizer on a subset of the parity suite (memcheck + racecheck of the new kernels' shared-memory protocols)
SEL="cartesian_box_steps and restructured and 3 or full_regions or many_particles_per_element and restructured or open_boundaries and restructured or degenerate_flights and restructured"
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 7 --print-limit 5 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "$SEL" > $OUT/c20_memcheck.log 2>&1; echo "memcheck rc=$?"; grep -E "ERROR SUMMARY|passed|failed|Invalid|Error" $OUT/c20_memcheck.log | tail -6
timeout 900 compute-sanitizer --tool racecheck --racecheck-report analysis --error-exitcode 7 --print-limit 5 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "cartesian_box_steps and restructured and 3 or full_regions and replan" > $OUT/c20_racecheck.log 2>&1; echo "racecheck rc=$?"; grep -E "RACECHECK SUMMARY|passed|failed|hazard|Race" $OUT/c20_racecheck.log | tail -6
# secondary variant: launch list + full capture of its kernels at 5e6 particles
B="python bench.py --variant ref_sf --nelem 32 --particles 5e6 --steps 2 --warmup 1 --no-cpu --no-e2e"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'k_track_ref|k_interp_push|k_sf_|k_gather_particles|k_scatter|k_hist' -c 200 --csv --log-file $OUT/r2sf_launches.csv $B > $OUT/r2sf_launches.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_track_ref|k_interp_push|k_sf_gather|k_sf_prepare' --launch-skip 4 -c 4 -f -o $OUT/r2sf_full $B > $OUT/r2sf_full.log 2>&1
ncu -i $OUT/r2sf_full.ncu-rep --page raw --csv > $OUT/r2sf_raw.csv 2>/dev/null
ncu -i $OUT/r2sf_full.ncu-rep --page source --csv --kernel-name regex:k_track_ref 2>/dev/null > $OUT/r2sf_src_k_track_ref.csv
python scripts/summarize_raw.py $OUT/r2sf_raw.csv
