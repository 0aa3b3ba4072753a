#!/bin/bash
OUT=gpurun_out
mkdir -p $OUT
timeout 900 python -m pytest tests/test_hnsw_build_gpu.py -q -m gpu -x -p no:cacheprovider > $OUT/r2d_tests.log 2>&1
tail -30 $OUT/r2d_tests.log
