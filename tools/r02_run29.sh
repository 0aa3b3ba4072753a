#!/bin/bash
# tensor-core flat scan v4 (accumulator pairs; <= 128 queries double-buffered, more: 256-wide tiles; 4 k-blocks of row loads in flight; output offsets in shared memory): check, pytest, ncu
OUT=gpurun_out
mkdir -p $OUT
timeout 200 ./tests/cpp/flat_tc_check 200000 > $OUT/flat_tc_check.log 2>&1; echo "check rc=$?"; grep -c "^ok" $OUT/flat_tc_check.log; grep "ids=200000\|FAIL\|PASSED\|FAILED" $OUT/flat_tc_check.log
timeout 600 python -m pytest tests/test_flat_tc_gpu.py -x -q -m gpu -p no:cacheprovider > $OUT/flat_tc_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 $OUT/flat_tc_pytest.log
timeout 400 ncu --set full --clock-control none --import-source on -k regex:flat_tc_kernel -s 12 -c 1 -o $OUT/r2fw_prof_flat_tc -f ./tests/cpp/flat_tc_check 200000 > $OUT/r2fw_ncu.log 2>&1; echo "ncu rc=$?"
