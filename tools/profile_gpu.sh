#!/bin/bash
# Run under gpurun. Produces in gpurun_out/: launches.csv (every launch with its device time), prof_kw / prof_knn
# (.ncu-rep, --set full of the two dominant kernels), bench JSON + clocks.
set -x
OUT=gpurun_out
ARGS="--steps 2 --warmup 1 --no-cpu-baseline $EXTRA"
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'kw_|hnsw_|hybrid_|vec_|found_|bitmap_|flat_' -c 60 --csv --log-file $OUT/launches.csv python bench.py $ARGS > $OUT/bench_under_ncu.json 2> $OUT/ncu_launches.log
ncu --set full --clock-control none --import-source on -k regex:kw_search_kernel -s 1 -c 1 -o $OUT/prof_kw -f python bench.py $ARGS > /dev/null 2> $OUT/ncu_kw.log
ncu --set full --clock-control none --import-source on -k regex:hnsw_search_kernel -s 1 -c 1 -o $OUT/prof_knn -f python bench.py $ARGS > /dev/null 2> $OUT/ncu_knn.log
ls -la $OUT
