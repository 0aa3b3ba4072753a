#!/bin/bash
OUT=gpurun_out
mkdir -p $OUT
run() { # name batch env...
  name=$1; batch=$2; shift; shift
  env "$@" python bench.py --batch $batch --steps 3 --warmup 2 --no-cpu-baseline --recall-queries 0 --no-other-configs > $OUT/r2o_$name.json 2> $OUT/r2o_$name.err
  python - "$name" <<'PY'
import json, sys
n = sys.argv[1]
try:
    j = json.loads(open(f"gpurun_out/r2o_{n}.json").read().strip().splitlines()[-1])
    w = j["work_per_step"]
    print(n, "value", round(j["value"]), "e2e", round(j["e2e"]["value"]), "ms/step", round(j["ms_per_step"], 2), "knn", round(j["device_ms_isolated"]["knn"], 2), "kw", round(j["device_ms_isolated"]["kw_search"], 2),
          "probes/dist", round(w["knn_table_probes"] / max(w["knn_dist"], 1), 2), "walks", j.get("knn_walks", {}).get("expanded_max"), "small p50", round(j["latency_ms"]["small"]["p50"], 2))
except Exception as e:
    print(n, "unreadable", e)
PY
}
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_hnsw_build_gpu.py tests/test_host_batched.py -q -m gpu -x -p no:cacheprovider -k "knn or hnsw or hybrid or batched" > $OUT/r2o_tests.log 2>&1
tail -4 $OUT/r2o_tests.log
run pf0 4096 TSGPU_WALK_PREFETCH=0
run pf1 4096 TSGPU_WALK_PREFETCH=1
run pf3 4096 TSGPU_WALK_PREFETCH=3
run pf7 4096 TSGPU_WALK_PREFETCH=7
run pf5 4096 TSGPU_WALK_PREFETCH=5
run pf7_b512 512 TSGPU_WALK_PREFETCH=7
python bench.py --batch 4096 --steps 4 --warmup 2 --no-cpu-baseline --recall-queries 0 --no-other-configs --e2e-depth 4 > $OUT/r2o_depth4.json 2> $OUT/r2o_depth4.err
python - <<'PY'
import json
j = json.loads(open("gpurun_out/r2o_depth4.json").read().strip().splitlines()[-1])
print("depth4 e2e", round(j["e2e"]["value"]), j["e2e"]["ms_per_step"], j["host_rounds_per_step"])
PY
tail -3 $OUT/r2o_pf0.err
