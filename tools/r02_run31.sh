#!/bin/bash
# f-4 (posting half) on the B200: patched mirror == fresh mirror == oracle, the C++ host scenarios (incremental_scenarios), memcheck of one patched search
OUT=gpurun_out
mkdir -p $OUT
timeout 900 python -m pytest tests/test_incremental_mirror.py tests/test_cpp_host.py -x -q -m gpu -p no:cacheprovider > $OUT/r2h_tests.log 2>&1; echo "tests rc=$?"; tail -6 $OUT/r2h_tests.log
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest "tests/test_incremental_mirror.py::test_patched_mirror_equals_fresh_mirror[False]" -x -q -m gpu -p no:cacheprovider > $OUT/r2h_memcheck.log 2>&1; echo "memcheck rc=$?"; grep -i "ERROR SUMMARY\|passed\|failed" $OUT/r2h_memcheck.log | tail -3
