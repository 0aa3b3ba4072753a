#!/bin/bash
OUT=gpurun_out
mkdir -p $OUT
timeout 600 python -m pytest tests/test_filters_device.py tests/test_gpu_parity.py -q -m gpu -x -p no:cacheprovider -k "filter or knn" > $OUT/r2m_tests.log 2>&1
tail -4 $OUT/r2m_tests.log
run() { # name batch env...
  name=$1; batch=$2; shift; shift
  env "$@" python bench.py --batch $batch --steps 3 --warmup 2 --no-cpu-baseline --recall-queries 0 --no-other-configs > $OUT/r2m_$name.json 2> $OUT/r2m_$name.err
  python - "$name" <<'PY'
import json, sys
n = sys.argv[1]
try:
    j = json.loads(open(f"gpurun_out/r2m_{n}.json").read().strip().splitlines()[-1])
    w = j["work_per_step"]
    print(n, "value", round(j["value"]), "e2e", round(j["e2e"]["value"]), "ms/step", round(j["ms_per_step"], 2), "knn", round(j["device_ms_isolated"]["knn"], 2), "kw", round(j["device_ms_isolated"]["kw_search"], 2),
          "table_probes/dist", round(w["knn_table_probes"] / max(w["knn_dist"], 1), 2), "walks", j.get("knn_walks", {}).get("expanded_max"), "small p50", round(j["latency_ms"]["small"]["p50"], 2))
except Exception as e:
    print(n, "unreadable", e)
PY
}
run b4096 4096 A=1
run b512_base 512 A=1
run b512_c8192_ctas2 512 TSGPU_WALK_CACHE=8192 TSGPU_WALK_CTAS=2
run b512_rows8_c8192 512 TSGPU_WALK_ROWS=8 TSGPU_WALK_CACHE=8192 TSGPU_WALK_CTAS=1
run b512_rows8_c2048 512 TSGPU_WALK_ROWS=8 TSGPU_WALK_CACHE=2048 TSGPU_WALK_CTAS=2
tail -3 $OUT/r2m_b4096.err
