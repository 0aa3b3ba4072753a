#!/bin/bash
OUT=gpurun_out
mkdir -p $OUT
run() { # name batch env...
  name=$1; batch=$2; shift; shift
  env "$@" python bench.py --batch $batch --steps 3 --warmup 2 --no-cpu-baseline --recall-queries 0 --no-other-configs > $OUT/r2u_$name.json 2> $OUT/r2u_$name.err
  python - "$name" <<'PY'
import json, sys
n = sys.argv[1]
try:
    j = json.loads(open(f"gpurun_out/r2u_{n}.json").read().strip().splitlines()[-1])
    w = j["work_per_step"]
    print(n, "value", round(j["value"]), "e2e", round(j["e2e"]["value"]), "ms/step", round(j["ms_per_step"], 2), "knn", round(j["device_ms_isolated"]["knn"], 2), "kw", round(j["device_ms_isolated"]["kw_search"], 2),
          "probes/dist", round(w["knn_table_probes"] / max(w["knn_dist"], 1), 2), "walks", j.get("knn_walks", {}).get("expanded_max"), "small p50", round(j["latency_ms"]["small"]["p50"], 2))
except Exception as e:
    print(n, "unreadable", e)
PY
}
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_hnsw_build_gpu.py tests/test_host_batched.py -q -m gpu -x -p no:cacheprovider -k "knn or hnsw or hybrid or batched" > $OUT/r2u_tests.log 2>&1
tail -4 $OUT/r2u_tests.log
run fw 4096 A=1
run fw_b512 512 A=1
tail -3 $OUT/r2u_fw.err
