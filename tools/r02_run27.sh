#!/bin/bash
# tensor-core flat scan: stand-alone check, its pytest, and one ncu --set full capture of the 256-query x 200 K-candidate launch
OUT=gpurun_out
mkdir -p $OUT
timeout 200 ./tests/cpp/flat_tc_check 200000 > $OUT/flat_tc_check.log 2>&1; rc=$?; echo "check rc=$rc"; cat $OUT/flat_tc_check.log
if [ $rc -ne 0 ]; then
  timeout 300 compute-sanitizer --tool memcheck ./tests/cpp/flat_tc_check 20000 2>&1 | head -60 > $OUT/flat_tc_sanitizer.log; tail -40 $OUT/flat_tc_sanitizer.log
  exit 0
fi
timeout 900 python -m pytest tests/test_flat_tc_gpu.py -x -q -m gpu -p no:cacheprovider > $OUT/flat_tc_pytest.log 2>&1; echo "pytest rc=$?"; tail -15 $OUT/flat_tc_pytest.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:flat_tc_kernel -s 12 -c 1 -o $OUT/r2ft_prof_flat_tc -f ./tests/cpp/flat_tc_check 200000 > $OUT/r2ft_ncu.log 2>&1; echo "ncu rc=$?"; tail -3 $OUT/r2ft_ncu.log
ls -la $OUT | grep -i "flat\|r2ft"
