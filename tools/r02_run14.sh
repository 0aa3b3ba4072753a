#!/bin/bash
OUT=gpurun_out
mkdir -p $OUT
run() { # name batch env...
  name=$1; batch=$2; shift; shift
  env "$@" python bench.py --batch $batch --steps 3 --warmup 2 --no-cpu-baseline --recall-queries 0 --no-other-configs > $OUT/r2n_$name.json 2> $OUT/r2n_$name.err
  python - "$name" <<'PY'
import json, sys
n = sys.argv[1]
try:
    j = json.loads(open(f"gpurun_out/r2n_{n}.json").read().strip().splitlines()[-1])
    w = j["work_per_step"]
    print(n, "value", round(j["value"]), "e2e", round(j["e2e"]["value"]), "ms/step", round(j["ms_per_step"], 2), "knn", round(j["device_ms_isolated"]["knn"], 2), "kw", round(j["device_ms_isolated"]["kw_search"], 2),
          "probes/dist", round(w["knn_table_probes"] / max(w["knn_dist"], 1), 2), "waste/dist", round(w.get("knn_wasted_rows", 0) / max(w["knn_dist"], 1), 3), "walks", j.get("knn_walks", {}).get("expanded_max"), "small p50", round(j["latency_ms"]["small"]["p50"], 2))
except Exception as e:
    print(n, "unreadable", e)
PY
}
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_hnsw_build_gpu.py -q -m gpu -x -p no:cacheprovider -k "knn or hnsw or hybrid" > $OUT/r2n_tests.log 2>&1
tail -4 $OUT/r2n_tests.log
run spec0 4096 TSGPU_WALK_SPEC=0
run spec1 4096 TSGPU_WALK_SPEC=1
run spec1_rows8 4096 TSGPU_WALK_SPEC=1 TSGPU_WALK_ROWS=8
run spec1_c4096 4096 TSGPU_WALK_SPEC=1 TSGPU_WALK_CACHE=4096
run spec1_rows8_c4096 4096 TSGPU_WALK_SPEC=1 TSGPU_WALK_ROWS=8 TSGPU_WALK_CACHE=4096
run spec1_b512 512 TSGPU_WALK_SPEC=1
run spec1_b512_rows8 512 TSGPU_WALK_SPEC=1 TSGPU_WALK_ROWS=8 TSGPU_WALK_CACHE=8192
tail -3 $OUT/r2n_spec1.err
