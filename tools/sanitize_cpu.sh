#!/bin/bash
# AddressSanitizer + UBSan over everything of the product that can execute without a GPU: the __host__ __device__ functions
# (scoring, block probe, ART walk) compiled for the host, the ART mirror, and the C++ host layer on the test double.
# usage: tools/sanitize_cpu.sh   (from the repo root; ~3 min)
set -e
cd "$(dirname "$0")/.."
SAN="-std=c++17 -O1 -g -fsanitize=address,undefined -fno-omit-frame-pointer"
OUT=$(mktemp -d)
make -s -C oracle liboracle.so
echo "== ART mirror + art_walk() (random vocabularies)"
g++ $SAN tests/cpp/art_sanitize_driver.cpp tests/cpp/art_mirror_capi.cpp -o $OUT/art && ASAN_OPTIONS=detect_leaks=1 $OUT/art
echo "== C++ host layer on the oracle double (host walk, then device-walk marshalling)"
g++ $SAN -Wno-unused tests/cpp/host_scenarios.cpp tests/cpp/tsgpu_oracle_double.cpp -o $OUT/hs -L oracle -l:liboracle.so -Wl,-rpath,$PWD/oracle -pthread
TSGPU_HOST_HYBRID_KAT=1 ASAN_OPTIONS=detect_leaks=0 $OUT/hs tests/golden/documents.jsonl | tail -1
TSGPU_HOST_HYBRID_KAT=1 TSGPU_HOST_DEVICE_ART=1 ASAN_OPTIONS=detect_leaks=0 $OUT/hs tests/golden/documents.jsonl | tail -1
echo "== ThreadSanitizer: the lock-step multi_search of the C++ host layer"
g++ -std=c++17 -O1 -g -fsanitize=thread -fno-omit-frame-pointer -Wno-unused tests/cpp/host_scenarios.cpp tests/cpp/tsgpu_oracle_double.cpp -o $OUT/hs_tsan \
    -L oracle -l:liboracle.so -Wl,-rpath,$PWD/oracle -pthread
TSGPU_HOST_DEVICE_ART=1 $OUT/hs_tsan tests/golden/documents.jsonl 2>&1 | grep -E "WARNING: ThreadSanitizer|PASSED|FAIL" | sort | uniq -c
echo "== device scoring / probing functions through tests/test_hostsim.py"
cp tests/hostsim/libhostsim.so $OUT/keep.so 2>/dev/null || true
g++ $SAN -fPIC -shared -x c++ tests/hostsim/hostsim.cpp -o tests/hostsim/libhostsim.so
LD_PRELOAD="$(g++ -print-file-name=libasan.so) $(g++ -print-file-name=libubsan.so)" ASAN_OPTIONS=detect_leaks=0:verify_asan_link_order=0 \
    python -m pytest tests/test_hostsim.py -x -q -p no:cacheprovider | tail -1
[ -f $OUT/keep.so ] && cp $OUT/keep.so tests/hostsim/libhostsim.so && touch tests/hostsim/libhostsim.so
echo "== the CPU oracle itself (the checker) under the reference KATs and scenario replays"
cp oracle/liboracle.so $OUT/oracle_keep.so
g++ $SAN -fPIC -Wno-unused -pthread -mavx2 -mfma -shared oracle/ts_oracle.cpp -o oracle/liboracle.so
LD_PRELOAD="$(g++ -print-file-name=libasan.so) $(g++ -print-file-name=libubsan.so)" ASAN_OPTIONS=detect_leaks=0:verify_asan_link_order=0 \
    python -m pytest tests/test_oracle_ref.py tests/test_reference_scenarios.py tests/test_hybrid_reference_kat.py tests/test_filter_scenarios.py \
    tests/test_sorting_scenarios.py tests/test_typo_scenarios.py tests/test_synonym_scenarios.py -x -q -m "not gpu" -p no:cacheprovider | tail -1
cp $OUT/oracle_keep.so oracle/liboracle.so && touch oracle/liboracle.so
rm -rf $OUT
