# usage: scripts/final_round.sh TAG -- full-size bench (both arms) + ncu launch list / --set full summaries for profiles/
TAG=${1:-r1_final}
OUT=gpurun_out
python bench.py > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err; tail -c 2500 $OUT/${TAG}_bench.json
python bench.py --impl reference --steps 3 --warmup 1 > $OUT/${TAG}_bench_reference.json 2> $OUT/${TAG}_bench_reference.err; tail -c 1200 $OUT/${TAG}_bench_reference.json
bash scripts/profile.sh $TAG | tail -3
