#!/bin/bash
OUT=gpurun_out
mkdir -p $OUT
timeout 600 python -m pytest tests/test_facets.py tests/test_gpu_parity.py -q -m gpu -x -p no:cacheprovider -k "facet or knn or hybrid or keyword_search_random" > $OUT/r2h_tests.log 2>&1
tail -12 $OUT/r2h_tests.log
run() { # name env...
  name=$1; shift
  env "$@" python bench.py --steps 3 --warmup 3 --no-cpu-baseline --recall-queries 0 > $OUT/r2h_$name.json 2> $OUT/r2h_$name.err
  python - "$name" <<'PY'
import json, sys
n = sys.argv[1]
try:
    j = json.loads(open(f"gpurun_out/r2h_{n}.json").read().strip().splitlines()[-1])
    print(n, "value", round(j["value"]), "ms/step", round(j["ms_per_step"], 2), "knn", round(j["device_ms_isolated"]["knn"], 2), "small", round(j["latency_ms"]["small"]["p50"], 2))
except Exception as e:
    print(n, "unreadable", e)
PY
}
run base A=1
run rows4_cache1024_c6 TSGPU_WALK_CACHE=1024 TSGPU_WALK_CTAS=6
run rows4_cache512_c6 TSGPU_WALK_CACHE=512 TSGPU_WALK_CTAS=6
run rows2_cache2048_c7 TSGPU_WALK_ROWS=2 TSGPU_WALK_CACHE=2048 TSGPU_WALK_CTAS=7
run rows2_cache1024_c10 TSGPU_WALK_ROWS=2 TSGPU_WALK_CACHE=1024 TSGPU_WALK_CTAS=10
run rows2_cache512_c10 TSGPU_WALK_ROWS=2 TSGPU_WALK_CACHE=512 TSGPU_WALK_CTAS=10
run rows4_cache4096_c4 TSGPU_WALK_CACHE=4096 TSGPU_WALK_CTAS=4
