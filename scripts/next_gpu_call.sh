# usage (one gpurun call): gpurun --timeout 1500 -- 'bash scripts/next_gpu_call.sh'
# First GPU call after the round-1 budget ran out: the -m gpu halves of tests/test_reference_*.py have never run on a B200
# (DESIGN.md §2), the FP64 peak is unmeasured, and the whole suite should be re-confirmed after the ref.cuh change.
OUT=gpurun_out
mkdir -p $OUT
python __graft_entry__.py --smoke > $OUT/n_smoke.log 2>&1; tail -2 $OUT/n_smoke.log
python -m pytest tests/test_reference_deposition.py tests/test_reference_push.py tests/test_reference_shapefunction.py \
    tests/test_reference_tracking.py tests/test_zz_host_cpp.py tests/test_zz_gpu_properties.py -m gpu -q > $OUT/n_reference_tests.log 2>&1; tail -15 $OUT/n_reference_tests.log
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o $OUT/fp64_peak scripts/fp64_peak.cu && $OUT/fp64_peak | tee $OUT/n_fp64_peak.json
# -> copy gpurun_out/n_fp64_peak.json to profiles/fp64_peak.json: bench.py then reports roofline.fp64 beside the HBM roofline
python -m pytest tests -m gpu -x -q > $OUT/n_all_gpu_tests.log 2>&1; tail -5 $OUT/n_all_gpu_tests.log
python bench.py > $OUT/n_bench.json 2> $OUT/n_bench.err; tail -c 2500 $OUT/n_bench.json
# secondary variant of SURVEY.md 8d (RefMapping + shape_function), first at a tenth of the particles
python bench.py --variant ref_sf --particles 5e7 --steps 5 --warmup 3 --no-cpu > $OUT/n_bench_ref_sf.json 2> $OUT/n_bench_ref_sf.err; tail -c 1500 $OUT/n_bench_ref_sf.json
