#!/bin/bash
OUT=gpurun_out
mkdir -p $OUT
timeout 1200 python -m pytest tests/test_hnsw_build_gpu.py tests/test_gpu_parity.py -q -m gpu -x -p no:cacheprovider -k "hnsw or knn or vector or hybrid" > $OUT/r2t_tests.log 2>&1
tail -6 $OUT/r2t_tests.log
ARGS="--steps 2 --warmup 1 --no-cpu-baseline --recall-queries 0 --no-other-configs"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:hnsw_walk_kernel -s 2 -c 1 -o $OUT/r2t_prof_walk -f python bench.py $ARGS > $OUT/r2t_ncu_walk.out 2> $OUT/r2t_ncu_walk.log
ls -la $OUT | grep r2t_
